// K5: per-pixel front-to-back alpha compositing -> RGB u8 (or RGBA) + planar depth f32.
//
// Replaces cam.get_rgba() (simple_env.py:1380; generate_images.py:428) and
// frame['distance_to_image_plane'] (simple_env.py:1425-1432).
//
// One WARP = one 8 x 8 quadrant of a 32 x 16 tile: lane -> row lane>>2, two horizontally adjacent
// pixels (lane&3)*2, so the blend arithmetic runs on Blackwell's packed binary32 pipe
// (FFMA2/FMUL2/FADD2: fma.rn.f32x2), which is IEEE per lane and therefore bit-identical to the
// scalar statement in the oracle.  Warps are independent: CTA b holds four quadrants of the b/2-th
// heaviest tile (k_bin_order's length classes; CTAs are dispatched in index order, so the longest
// lists start first) and a warp never synchronises with another warp.  (Round 1 blended a tile with a
// CTA of four warps that met at a barrier per batch — 24 % of its stall samples — and filtered every
// entry with a ballot/ffs loop.)
//
// A warp walks the tile's depth-sorted list in three decoupled steps: it SCANS the raw list 128 entries
// at a time and appends the ids whose quadrant bit is set to a small id ring in shared memory (ballot +
// popc), GATHERS the next 32 ids' splat records (one 48-byte block per Gaussian, L2-resident) into one of
// two private stages with cp.async (LDGSTS, 16 bytes x 3 per lane), and BLENDS the stage that has landed
// while the next gather is in flight.  Every stage therefore holds 32 entries of this quadrant, however
// sparse they are in the tile's list.  (Round 1 issued one 48-byte TMA bulk copy per record: UBLKCP is a
// uniform-datapath instruction, so a warp issued its 32 copies one lane at a time — ~8 % of the kernel's
// instructions.  A per-lane gather is what LDGSTS is for.)  The entry loop is warp-uniform and blends E entries
// per trip (their exponents and alphas side by side - the loop is bound by its own dependency chain whenever a
// warp has its scheduler to itself); small whole frames also get a lane-parallel box test per stage (PRE) that
// leaves only the entries which can still reach a live pixel for that loop.
//
// The blend: log2(alpha) = qa dx^2 + qb dx dy + qc dy^2 + lo with the opacity
// already in the exponent; skip below log2(1/255); cap at log2(0.99); alpha by the degree-3
// deterministic exp2.  A finished (or off-image) pixel carries x = NaN, so every comparison on it is
// false and it drops out without an "active" flag.  Finished pixels can be stored straight into
// several frame buffers (peer-mapped, NVLink) — the multi-GPU strip exchange fused into the epilogue.
#include "sage_internal.cuh"

namespace {

// A CTA = four warps = the four 8 x 8 quadrants of one 16 x 16 HALF of a 32 x 16 tile; two CTAs per tile.
// (One CTA of eight warps per tile was measured: +5 % kernel time — a CTA holds its registers and shared
// memory until its slowest warp is done, and eight warps wait longer for one than four do.)
constexpr int C_WARPS = 4;
constexpr int C_HALVES = SAGE_BIN_W / 16;
constexpr int C_THREADS = C_WARPS * 32;
#ifndef SAGE_COMPOSITE_CTAS
#define SAGE_COMPOSITE_CTAS 8
#endif
constexpr int C_CTAS_PER_SM = SAGE_COMPOSITE_CTAS;
#ifndef SAGE_COMPOSITE_SCAN
#define SAGE_COMPOSITE_SCAN 4
#endif
constexpr int C_SCAN = SAGE_COMPOSITE_SCAN;          // 32-entry chunks of the raw list per scan step
constexpr int C_RING = C_SCAN >= 2 ? 64 * C_SCAN : 128;   // ids (power of two >= 64 + 32 C_SCAN): < 32 waiting + one scan step, never
                                                          // overtaking the 32 a slower lane may still be reading for its gather
#ifndef SAGE_COMPOSITE_CTAS_WIDE
#define SAGE_COMPOSITE_CTAS_WIDE 6
#endif
constexpr int C_CTAS_PER_SM_WIDE = SAGE_COMPOSITE_CTAS_WIDE;   // register budget of the 4-entries-per-trip instantiation
constexpr int C_LIVE_CHECK = 8;                      // entries between two "is any pixel of the quadrant still live" votes

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

#ifdef SAGE_CUDA_EMU
inline void cp_async16(void* dst_smem, const void* src) { memcpy(dst_smem, src, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// packed deterministic degree-3 exp2 (mirrors so_exp2_fast); valid for x in [L255, L099]
__device__ __forceinline__ float2 sage_exp2_fast_x2(float2 x)
{
    const float magic = 12582912.0f;
    float2 t = __fadd2_rn(x, f2(magic));
    float2 n = __fadd2_rn(t, f2(-magic));
    float2 f = __ffma2_rn(n, f2(-1.0f), x);          // x - n, exact product => same rounding
    float2 p = f2(0x1.c3f760p-5f);
    p = __ffma2_rn(p, f, f2(0x1.f0de1ap-3f));
    p = __ffma2_rn(p, f, f2(0x1.62f31ap-1f));
    p = __ffma2_rn(p, f, f2(0x1.fff692p-1f));
    float2 r;
    r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return r;
}

#ifdef SAGE_COMPOSITE_PROFILE
// development build only (tools/build_variant.sh prof -DSAGE_COMPOSITE_PROFILE): per-warp walk statistics of the last launch
constexpr int CPROF_WORDS = 8, CPROF_WARPS = 1 << 16;
__device__ uint32_t g_cprof[CPROF_WARPS * CPROF_WORDS];
__device__ __forceinline__ uint32_t cprof_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return (uint32_t)t; }
#endif

// E = entries of the stage blended per trip of the inner loop (2: frames of many waves of CTAs; 4: small frames);
// PRE = once per stage, a lane-parallel test of every entry against the box of the quadrant's live pixels comes first
template <int E, bool PRE>
__global__ void __launch_bounds__(C_THREADS, E <= 2 ? C_CTAS_PER_SM : C_CTAS_PER_SM_WIDE)
k_composite(SageFrame fr, const uint32_t* __restrict__ vals)
{
    __shared__ __align__(16) float4 srec[C_WARPS][2][32 * 3];       // per warp: two stages of 32 x 48-byte records
    __shared__ uint32_t sring[C_WARPS][C_RING];                     // per warp: ring of Gaussian ids waiting to be gathered
    const SageFrameParams* __restrict__ fp = fr.fp;
    const int W = fp->cam.W, H = fp->cam.H, bins_x = fp->cam.bins_x, bin_row0 = fp->cam.bin_row0;
    const int n_bins = fp->cam.n_bins;
    const float depth_lo = fp->cam.depth_lo, depth_hi = fp->cam.depth_hi;
    const float4* __restrict__ rec = fr.rec;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // work queue: tiles by descending length class, 4 work items (quadrants) each; lane l holds the
    // running count of classes 32 .. 32-l
    const uint32_t cls_n = fr.class_count[32 - lane];           // lane 0 <-> class 32 ... lane 31 <-> class 1
    uint32_t cls_incl = cls_n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, cls_incl, o);
        if (lane >= o) cls_incl += t;
    }
    const uint32_t n_nonempty = __shfl_sync(0xffffffffu, cls_incl, 31);
    const float NAN_X = __uint_as_float(0x7fc00000u);

    // Not persistent on purpose: as CTAs retire, the SM's registers and shared memory go to the
    // (higher-priority) front-half kernels of the next frame instead of staying pinned to this kernel.
    for (uint32_t wi = blockIdx.x; wi < (uint32_t)n_bins * C_HALVES; wi += gridDim.x) {     // (one trip: the grid is n_bins x halves)
        const uint32_t tt = wi / C_HALVES;                       // tile rank in the heaviest-first order
        const int half = (int)(wi % C_HALVES);
        const int quad = (w >> 1) * SAGE_QUADS_X + half * 2 + (w & 1);   // quadrant: bit qy*4 + qx of the instance masks
        uint32_t cls, idx;
        if (tt < n_nonempty) {
            const unsigned m = __ballot_sync(0xffffffffu, cls_incl > tt);
            const int L = __ffs(m) - 1;
            cls = (uint32_t)(32 - L);
            idx = tt - (__shfl_sync(0xffffffffu, cls_incl, L) - __shfl_sync(0xffffffffu, cls_n, L));
        } else {
            cls = 0; idx = tt - n_nonempty;                      // the empty tiles come last
        }
        const int bin = (int)fr.class_seg[(size_t)cls * fr.max_bins + idx];
        uint2 rg = fr.ranges[bin];                               // (~begin, end); (0, 0) = empty
        rg.x = rg.y ? ~rg.x : 0u;
        const int bx = bin % bins_x, by = bin_row0 + bin / bins_x;
        const int qx0 = bx * SAGE_BIN_W + (quad % SAGE_QUADS_X) * 8, qy0 = by * SAGE_BIN_H + (quad / SAGE_QUADS_X) * 8;
        if (qx0 >= W || qy0 >= H) continue;                      // quadrant entirely outside the image
        const int px = qx0 + (lane & 3) * 2;
        const int py = qy0 + (lane >> 2);
        const bool in0 = px < W && py < H, in1 = (px + 1) < W && py < H;
        float2 pxc = make_float2(in0 ? (float)px + 0.5f : NAN_X, in1 ? (float)(px + 1) + 0.5f : NAN_X);
        const float pyc = (float)py + 0.5f;
        float2 T = f2(1.0f), Cr = f2(0.0f), Cg = f2(0.0f), Cb = f2(0.0f), Dz = f2(0.0f);
#define SAGE_LIVE() (pxc.x == pxc.x || pxc.y == pxc.y)

        const uint32_t n = rg.y - rg.x;
        const uint32_t* lst = vals + rg.x;
#ifdef SAGE_COMPOSITE_PROFILE
        const long long prof_c0 = clock64();
        const uint32_t prof_t0 = cprof_ns();
        uint32_t prof_visited = 0, prof_hits = 0, prof_cand = 0;
#endif
        // The walk, in three decoupled steps:
        //   scan    the tile's list C_SCAN x 32 entries at a time (loads issued one step ahead); the ids of the entries
        //           whose quadrant bit is set are appended to the warp's id ring (ballot + popc)
        //   gather  the next 32 ids of the ring -> their 48-byte records into the other stage (cp.async)
        //   blend   the stage that has landed
        // so every stage holds 32 entries of THIS quadrant however sparse they are in the tile's list.  (Until run W
        // a stage was filled from one 32-entry chunk of the raw list: in the longest lists only a quarter of a chunk
        // belongs to a given quadrant, the blend of those ~8 entries was shorter than the L2 round trip of the next
        // gather, and the warps that finish last - the kernel's tail at 640 x 480 - ran at memory latency:
        // 210-290 cycles per gathered entry, profiles/r02_variants.md.)
#if defined(SAGE_CUDA_EMU) && defined(SAGE_EMU_POISON)
        for (int k = lane; k < 2 * 32 * 3; k += 32) (&srec[w][0][0])[k] = make_float4(NAN_X, NAN_X, NAN_X, NAN_X);   // emulator test: stale stage contents
        __syncwarp();
#endif
        uint32_t* ring = sring[w];
        uint32_t q_head = 0, q_tail = 0, scan = 0;               // warp-uniform
        uint32_t vq[C_SCAN];
        auto prefetch = [&]() {
#pragma unroll
            for (int k = 0; k < C_SCAN; ++k) {
                const uint32_t idx = scan + 32u * k + lane;
                vq[k] = idx < n ? __ldg(&lst[idx]) : 0u;          // 0: no quadrant bit, never kept
            }
        };
        auto fill = [&]() {
            while (q_tail - q_head < 32u && scan < n) {
#pragma unroll
                for (int k = 0; k < C_SCAN; ++k) {
                    const bool rel = (vq[k] >> quad) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, rel);
                    if (rel) ring[(q_tail + __popc(m & lt_mask)) & (C_RING - 1)] = vq[k] >> SAGE_MASK_BITS;
                    q_tail += __popc(m);
                }
                scan += 32u * C_SCAN;
                prefetch();
            }
            __syncwarp();                                   // the ring's new ids are visible to every lane
        };
        auto gather = [&](int stage) -> int {
            const int c = (int)min(q_tail - q_head, 32u);
            if (lane < c) {
                const float4* src = rec + (size_t)ring[(q_head + lane) & (C_RING - 1)] * 3;
                float4* dst = &srec[w][stage][lane * 3];
                cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
            }
            // count not a multiple of E: the index walk (!PRE) reads the slots up to the next multiple as masked partners of
            // the last entries - whatever an earlier stage left there, their colours must be finite
            if (!PRE && E > 1 && lane >= c && lane < ((c + E - 1) & ~(E - 1))) srec[w][stage][lane * 3 + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            cp_async_commit();                             // (possibly empty: keeps the group count uniform)
            q_head += (uint32_t)c;
            return c;
        };
        int cnt_next = 0;
        if (n > 0) {
            prefetch();
            fill();
            cnt_next = gather(0);
        }
        for (int st = 0; cnt_next > 0; st ^= 1) {
            const int cnt = cnt_next;
            fill();
            cnt_next = gather(st ^ 1);                     // in flight while stage st is blended
            cp_async_wait<1>();
            __syncwarp();
            const float4* s_rec = srec[w][st];
#ifdef SAGE_COMPOSITE_PROFILE
            prof_visited += (uint32_t)cnt;
#endif
            // Which entries of the stage can reach a pixel that is still live?  PRE: one entry per lane against the
            // bounding box of the quadrant's live pixels - the exact maximum of the (concave) exponent over the box, on
            // its four edges, plus a rounding allowance - so the serial loop below only walks the survivors.  In the
            // walks that make the kernel's tail two thirds of the gathered entries touch nothing that is still live
            // (profiles/r02_variants.md) and each costs the serial loop a full dependency chain; here they cost one
            // lane a few dozen independent flops.  Skipping is exact: an entry is dropped only if no live pixel can
            // pass the alpha >= 1/255 test, and pixels only ever leave the live set during the stage.
            unsigned cand = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
            if (PRE) {
                const bool l0 = pxc.x == pxc.x, l1 = pxc.y == pxc.y;
                const uint32_t INF_U = 0x7f800000u;                            // pixel centres are positive: uint order = float order
                const uint32_t ux0 = l0 ? __float_as_uint(pxc.x) : (l1 ? __float_as_uint(pxc.y) : INF_U);
                const uint32_t ux1 = l1 ? __float_as_uint(pxc.y) : (l0 ? __float_as_uint(pxc.x) : 0u);
                const uint32_t uy0 = (l0 || l1) ? __float_as_uint(pyc) : INF_U, uy1 = (l0 || l1) ? __float_as_uint(pyc) : 0u;
                const float X0 = __uint_as_float(__reduce_min_sync(0xffffffffu, ux0)), X1 = __uint_as_float(__reduce_max_sync(0xffffffffu, ux1));
                const float Y0 = __uint_as_float(__reduce_min_sync(0xffffffffu, uy0)), Y1 = __uint_as_float(__reduce_max_sync(0xffffffffu, uy1));
                bool keep = false;
                if (lane < cnt) {
                    const float4 a = s_rec[3 * lane];                                           // gx gy qa lo
                    const float2 bq = *reinterpret_cast<const float2*>(&s_rec[3 * lane + 1]);     // qb qc
                    const float ax0 = X0 - a.x, ax1 = X1 - a.x, ay0 = Y0 - a.y, ay1 = Y1 - a.y;   // the box, relative to the splat's centre
                    const float qa = a.z, qb = bq.x, qc = bq.y;
                    auto form = [&](float dx, float dy) { return dx * (qa * dx + qb * dy) + (qc * dy) * dy; };
                    float fmax = 0.0f;                                                          // centre inside the box
                    if (!(ax0 <= 0.0f && ax1 >= 0.0f && ay0 <= 0.0f && ay1 >= 0.0f)) {
                        const float kc = (-0.5f * qb) / qc, ka = (-0.5f * qb) / qa;             // argmax of dy given dx: kc dx; of dx given dy: ka dy
                        fmax = form(ax0, fminf(fmaxf(kc * ax0, ay0), ay1));
                        fmax = fmaxf(fmax, form(ax1, fminf(fmaxf(kc * ax1, ay0), ay1)));
                        fmax = fmaxf(fmax, form(fminf(fmaxf(ka * ay0, ax0), ax1), ay0));
                        fmax = fmaxf(fmax, form(fminf(fmaxf(ka * ay1, ax0), ax1), ay1));
                    }
                    // allowance for the binary32 rounding of this bound and of the loop's own exponent (a few ulp of the
                    // largest term each; thin diagonal splats cancel terms of 1e5 - 1e6 down to ~10)
                    const float mx = fmaxf(fabsf(ax0), fabsf(ax1)), my = fmaxf(fabsf(ay0), fabsf(ay1));
                    const float mag = fabsf(qa) * mx * mx + fabsf(qb) * mx * my + fabsf(qc) * my * my;
                    keep = !(a.w + fmax + (0.02f + 2e-6f * mag) < SAGE_L255);                   // (NaN anywhere: keep)
                }
                cand = __ballot_sync(0xffffffffu, keep);
            }
#ifdef SAGE_COMPOSITE_PROFILE
            prof_cand += (uint32_t)__popc(cand);
#endif
            // E entries per trip: their exponents and alphas are evaluated side by side (E independent dependency
            // chains in one basic block), then the E blends run in order.  A warp is latency-bound on this loop
            // whenever it has its scheduler to itself - the tail of every frame, most of a small frame (run X: 130
            // cycles per skipped and 270 per blended entry, one at a time) - and E chains overlap; where the kernel is
            // issue-bound the price is the alpha path of an entry no lane needed whenever a partner needed it.
            // one trip: rp[u] -> the 48-byte record of entry u of the trip, has_u[u]: it exists (warp-uniform)
            auto blend_trip = [&](const float4* const (&rp)[E], const bool (&has_u)[E]) {
                float2 pw[E];
                bool v0[E], v1[E], any = false;
#pragma unroll
                for (int u = 0; u < E; ++u) {
                    const float4 a = rp[u][0];                                  // gx gy qa lo
                    const float2 bq = *reinterpret_cast<const float2*>(&rp[u][1]);   // qb qc
                    const float dy = pyc - a.y;
                    const float2 t2 = __fmul2_rn(bq, f2(dy));                             // bdy, cdy
                    const float c2 = fmaf(t2.y, dy, a.w);                                 // cdy dy + lo
                    const float2 dx = __fadd2_rn(pxc, f2(-a.x));
                    const float2 wv = __ffma2_rn(f2(a.z), dx, f2(t2.x));
                    pw[u] = __ffma2_rn(dx, wv, f2(c2));                                   // log2(alpha)
                    const bool has = (u == 0) || has_u[u];                                 // warp-uniform (a missing partner is masked)
                    v0[u] = has && pw[u].x >= SAGE_L255;                                  // false for NaN (finished pixels)
                    v1[u] = has && pw[u].y >= SAGE_L255;
                    any = any || v0[u] || v1[u];
                }
#ifdef SAGE_COMPOSITE_PROFILE
                if (__any_sync(0xffffffffu, any)) prof_hits++;
#endif
                if (!any) return;
                float2 al[E];
#pragma unroll
                for (int u = 0; u < E; ++u) {
                    pw[u].x = fminf(pw[u].x, SAGE_L099);
                    pw[u].y = fminf(pw[u].y, SAGE_L099);
                    al[u] = sage_exp2_fast_x2(pw[u]);
                    al[u].x = v0[u] ? al[u].x : 0.0f;          // alpha 0: T, C, Dz stay bit-identical
                    al[u].y = v1[u] ? al[u].y : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < E; ++u) {
                    const float2 om = __ffma2_rn(al[u], f2(-1.0f), f2(1.0f));             // 1 - alpha
                    float2 Tn = __fmul2_rn(T, om);
                    float2 wgt = __fmul2_rn(al[u], T);
                    // stop: the splat would push T below 1e-4 -> the pixel is finished, the splat not blended.
                    // (ptxas if-converts this block into predicated instructions that issue for every splat; forcing a
                    // real branch was measured 2 % slower, profiles/r02_variants.md.)
                    if (fminf(Tn.x, Tn.y) < SAGE_T_MIN) {
                        const bool s0 = Tn.x < SAGE_T_MIN, s1 = Tn.y < SAGE_T_MIN;
                        wgt.x = s0 ? 0.0f : wgt.x;
                        wgt.y = s1 ? 0.0f : wgt.y;
                        Tn.x = s0 ? T.x : Tn.x;
                        Tn.y = s1 ? T.y : Tn.y;
                        pxc.x = s0 ? NAN_X : pxc.x;
                        pxc.y = s1 ? NAN_X : pxc.y;
#pragma unroll
                        for (int u2 = u + 1; u2 < E; ++u2) {                              // finished: the later entries of the trip no longer see the pixel
                            al[u2].x = s0 ? 0.0f : al[u2].x;
                            al[u2].y = s1 ? 0.0f : al[u2].y;
                        }
                    }
                    T = Tn;
                    const float4 c = rp[u][2];                              // r g b z
                    Cr = __ffma2_rn(f2(c.x), wgt, Cr);
                    Cg = __ffma2_rn(f2(c.y), wgt, Cg);
                    Cb = __ffma2_rn(f2(c.z), wgt, Cb);
                    Dz = __ffma2_rn(f2(c.w), wgt, Dz);
                }
            };
            if (PRE) {
                for (int trips = 0; cand != 0u; ) {                // the surviving entries, in list order
                    const float4* rp[E];
                    bool has_u[E];
#pragma unroll
                    for (int u = 0; u < E; ++u) {
                        has_u[u] = cand != 0u;
                        rp[u] = has_u[u] ? s_rec + 3 * (__ffs((int)cand) - 1) : rp[0];
                        cand &= cand - 1u;                         // (0 & 0xffffffff = 0)
                    }
                    blend_trip(rp, has_u);
                    if (((++trips) & (C_LIVE_CHECK / E - 1)) == 0 && cand != 0u && !__any_sync(0xffffffffu, SAGE_LIVE())) break;
                }
            } else {
                for (int j0 = 0; j0 < cnt; j0 += C_LIVE_CHECK) {   // every entry of the stage, by index
                    const int j1 = min(j0 + C_LIVE_CHECK, cnt);
                    for (int j = j0; j < j1; j += E) {
                        const float4* rp[E];
                        bool has_u[E];
#pragma unroll
                        for (int u = 0; u < E; ++u) {
                            has_u[u] = j + u < cnt;                // (a slot past cnt is stale, masked; its colour was zeroed by the gather)
                            rp[u] = s_rec + 3 * (j + u);
                        }
                        blend_trip(rp, has_u);
                    }
                    if (j1 < cnt && !__any_sync(0xffffffffu, SAGE_LIVE())) break;
                }
            }
            if (!__any_sync(0xffffffffu, SAGE_LIVE())) cnt_next = 0;       // every pixel of the quadrant is finished: the walk ends here
            __syncwarp();                                  // stage st is free for the gather after next
        }
        cp_async_wait<0>();                                // nothing may land in the stages after this bin
        __syncwarp();
#undef SAGE_LIVE
#ifdef SAGE_COMPOSITE_PROFILE
        if (lane == 0) {
            const uint32_t wid = blockIdx.x * C_WARPS + w;
            if (wid < (uint32_t)CPROF_WARPS) {
                uint32_t* o = g_cprof + (size_t)wid * CPROF_WORDS;
                o[0] = (uint32_t)(clock64() - prof_c0); o[1] = prof_visited; o[2] = min(scan, n); o[3] = prof_hits;
                o[4] = prof_t0; o[5] = cprof_ns(); o[6] = n; o[7] = prof_cand;
            }
        }
#endif

        // epilogue: quantise and store to every destination frame.  Wide path (frame width a multiple
        // of 8, 8-byte aligned buffers, quadrant inside the image): the 24 RGB bytes of a bin row (4 lanes x
        // 2 px) leave as three 8-byte words assembled with shuffles, RGBA as one 8-byte word per lane,
        // depth as one float2 per lane.
        uint32_t q[6];
        {
            const float vals_c[6] = {Cr.x, Cg.x, Cb.x, Cr.y, Cg.y, Cb.y};
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                float v = fminf(fmaxf(vals_c[k], 0.0f), 1.0f);
                q[k] = (uint32_t)(uint8_t)(v * 255.0f + 0.5f);
            }
        }
        const float cov0 = 1.0f - T.x, cov1 = 1.0f - T.y;
        const float INF_F = __uint_as_float(0x7f800000u);
        float d0 = cov0 > 1e-6f ? Dz.x / cov0 : INF_F;       // nothing hit: +inf, as Isaac's distance_to_image_plane
        float d1 = cov1 > 1e-6f ? Dz.y / cov1 : INF_F;
        if (depth_lo <= depth_hi) {                          // optional output clamp (façade: clip(0.1, 6.5))
            d0 = fminf(fmaxf(d0, depth_lo), depth_hi);
            d1 = fminf(fmaxf(d1, depth_lo), depth_hi);
        }
        const size_t p = (size_t)py * W + px;
        const int n_out = fp->out.n;
        const bool rgba = fp->out.rgba != 0;
        const bool wide = fp->out.wide && (qx0 + 8 <= W) && (qy0 + 8 <= H);
        if (wide) {
            const int sub = lane & 3;
            if (rgba) {
                const unsigned long long word = (unsigned long long)(q[0] | (q[1] << 8) | (q[2] << 16) | 0xFF000000u) |
                                                ((unsigned long long)(q[3] | (q[4] << 8) | (q[5] << 16) | 0xFF000000u) << 32);
                for (int k = 0; k < n_out; ++k) {
                    *reinterpret_cast<unsigned long long*>(fp->out.rgb[k] + p * 4) = word;
                    float* dep = fp->out.depth[k];
                    if (dep) *reinterpret_cast<float2*>(dep + p) = make_float2(d0, d1);
                }
            } else {
                // this lane's 6 bytes as a 64-bit value; lanes 4r..4r+3 form one row
                const unsigned long long mine = (unsigned long long)q[0] | ((unsigned long long)q[1] << 8) |
                                                ((unsigned long long)q[2] << 16) | ((unsigned long long)q[3] << 24) |
                                                ((unsigned long long)q[4] << 32) | ((unsigned long long)q[5] << 40);
                const unsigned long long nxt = __shfl_down_sync(0xffffffffu, mine, 1);
                // word k of the row (k = 0,1,2) starts at byte 8k = lane k's byte (8k - 6k) = 2k
                unsigned long long word = 0;
                if (sub == 0) word = mine | (nxt << 48);                       // b0..5 of lane0 + b0..1 of lane1
                else if (sub == 1) word = (mine >> 16) | (nxt << 32);          // b2..5 of lane1 + b0..3 of lane2
                else if (sub == 2) word = (mine >> 32) | (nxt << 16);          // b4..5 of lane2 + b0..5 of lane3
                const size_t row_byte = ((size_t)py * W + qx0) * 3;
                for (int k = 0; k < n_out; ++k) {
                    if (sub < 3) *reinterpret_cast<unsigned long long*>(fp->out.rgb[k] + row_byte + 8 * sub) = word;
                    float* dep = fp->out.depth[k];
                    if (dep) *reinterpret_cast<float2*>(dep + p) = make_float2(d0, d1);
                }
            }
        } else if (in0 || in1) {
            const int bpp = rgba ? 4 : 3;
            for (int k = 0; k < n_out; ++k) {
                uint8_t* rgb = fp->out.rgb[k];
                float* dep = fp->out.depth[k];
                if (in0) {
                    uint8_t* o = rgb + p * bpp;
                    o[0] = (uint8_t)q[0]; o[1] = (uint8_t)q[1]; o[2] = (uint8_t)q[2];
                    if (rgba) o[3] = 255;
                    if (dep) dep[p] = d0;
                }
                if (in1) {
                    uint8_t* o = rgb + (p + 1) * bpp;
                    o[0] = (uint8_t)q[3]; o[1] = (uint8_t)q[4]; o[2] = (uint8_t)q[5];
                    if (rgba) o[3] = 255;
                    if (dep) dep[p + 1] = d1;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace

#ifdef SAGE_COMPOSITE_PROFILE
extern "C" int sage_debug_composite_profile(uint32_t* host, int n_warps)
{
    if (n_warps > CPROF_WARPS) n_warps = CPROF_WARPS;
    return cudaMemcpyFromSymbol(host, g_cprof, sizeof(uint32_t) * CPROF_WORDS * (size_t)n_warps) == cudaSuccess ? 0 : 1;
}
#endif

// Entries per trip: 2 where the frame is many waves of CTAs (the SMs stay full, the kernel is issue-bound and a wider
// trip only adds alpha evaluations nobody needed), 4 for WHOLE frames of at most SAGE_COMPOSITE_WIDE_MAX_CTAS CTAs (640 x 480:
// 1200 CTAs, a single wave - the kernel's time is its longest walk, latency-bound, and the caller of such a frame is the
// synchronous facade: run CC, 1449 against 1366 agent steps/s).  A tile-row block of a multi-GPU split is as small, but
// its rank keeps six frames in flight and lives on throughput: E = 2 turns a block over 3 - 5 % faster (run DD).
static int g_wide_max_ctas = 2400;
// The per-stage box test (PRE) pays where walks are long and mostly miss: together with E = 4 on whole frames of a single wave
// of CTAs (640 x 480: composite 0.216 -> 0.188 ms, frames/s +2 %, frame latency 0.51 -> 0.49 ms, run II).  With E = 2 its
// candidate walk (a find-first-set per entry instead of an index) is not won back: 1024 x 768 -3 % frames/s, 1080p -6 %.
// SAGE_COMPOSITE_PREFILTER: 1 always, 0 never, unset: exactly the launches that get E = 4.
static int g_prefilter = -1;

void sage_launch_composite(const SageFrame& fr, const uint32_t* sorted_vals, int n_bins, bool whole_frame, cudaStream_t st)
{
    if (n_bins <= 0) return;
    const int ctas = n_bins * C_HALVES;
    const bool wide = whole_frame && ctas <= g_wide_max_ctas;
    const bool pre = g_prefilter >= 0 ? g_prefilter == 1 : wide;
    if (wide) {
        if (pre) SAGE_KLAUNCH((k_composite<4, true>), ctas, C_THREADS, st, fr, sorted_vals);
        else SAGE_KLAUNCH((k_composite<4, false>), ctas, C_THREADS, st, fr, sorted_vals);
    } else {
        if (pre) SAGE_KLAUNCH((k_composite<2, true>), ctas, C_THREADS, st, fr, sorted_vals);
        else SAGE_KLAUNCH((k_composite<2, false>), ctas, C_THREADS, st, fr, sorted_vals);
    }
}

void sage_composite_configure(void)
{
#ifndef SAGE_CUDA_EMU
    sage_prefer_smem(k_composite<2, false>);
    sage_prefer_smem(k_composite<2, true>);
    sage_prefer_smem(k_composite<4, false>);
    sage_prefer_smem(k_composite<4, true>);
#endif
    if (const char* v = getenv("SAGE_COMPOSITE_WIDE_MAX_CTAS")) g_wide_max_ctas = atoi(v);
    if (const char* v = getenv("SAGE_COMPOSITE_PREFILTER")) g_prefilter = atoi(v);

}
