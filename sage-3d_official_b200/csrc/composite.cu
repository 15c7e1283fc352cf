// K5: per-pixel front-to-back alpha compositing -> RGB u8 + planar depth f32.
//
// Replaces cam.get_rgba() (simple_env.py:1380; generate_images.py:428) and
// frame['distance_to_image_plane'] (simple_env.py:1425-1432).
//
// One CTA = one 16x16 tile, 128 threads, two horizontally adjacent pixels per thread so the
// blend arithmetic runs on Blackwell's packed binary32 pipe (FFMA2/FMUL2/FADD2: fma.rn.f32x2),
// which is IEEE per lane and therefore bit-identical to the scalar statement in the oracle.
// Splat records (one 48-byte block per Gaussian, L2-resident) of the tile's depth-sorted list are
// staged into shared memory 128 at a time by the TMA engine: every thread issues one
// cp.async.bulk (global -> shared, 48 B) that completes on an mbarrier; two stages, so the
// gather of batch b+1 is in flight while batch b is blended.  Finished pixels can be stored
// straight into several frame buffers (peer-mapped, NVLink) — the multi-GPU strip exchange
// fused into the epilogue.
#include "sage_internal.cuh"

namespace {

constexpr int C_THREADS = 128;
constexpr int MAX_PEERS = 8;

struct PeerOut {
    uint8_t* rgb[MAX_PEERS];
    float* depth[MAX_PEERS];
    int n;
    int wide;     // 8-byte stores allowed: W % 8 == 0 and every buffer 8-byte aligned
};

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

// ---- mbarrier + bulk-copy (TMA) primitives --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 22)) __trap();      // never hang the GPU
    }
}

// packed deterministic exp2; valid for lanes with x in [-125, 0] (callers mask the others)
__device__ __forceinline__ float2 sage_exp2_x2(float2 x)
{
    const float magic = 12582912.0f;
    float2 t = __fadd2_rn(x, f2(magic));
    float2 n = __fadd2_rn(t, f2(-magic));
    float2 f = __ffma2_rn(n, f2(-1.0f), x);          // x - n, exact product => same rounding
    float2 p = f2(0x1.5c08e4p-10f);
    p = __ffma2_rn(p, f, f2(0x1.3d0c52p-7f));
    p = __ffma2_rn(p, f, f2(0x1.c6b6e4p-5f));
    p = __ffma2_rn(p, f, f2(0x1.ebf918p-3f));
    p = __ffma2_rn(p, f, f2(0x1.62e428p-1f));
    p = __ffma2_rn(p, f, f2(0x1.000002p+0f));
    float2 r;
    r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return r;
}

// 12 CTAs per SM (40 registers, no spills): measured 2075 -> 2157 frames/s over the uncapped
// 50-register build; 16 CTAs (32 registers, 40 B of spill in the blend loop) drops to 2015.
__global__ void __launch_bounds__(C_THREADS, 12)
k_composite(SageCamDev cam, const float4* __restrict__ rec, const uint32_t* __restrict__ vals,
            const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, PeerOut out)
{
    __shared__ __align__(16) float4 srec[2][C_THREADS * 3];   // two stages of 128 x 48-byte records
    __shared__ uint32_t smask_s[2][C_THREADS];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int tile = (int)tile_order[blockIdx.x];
    const int tx = tile % cam.tiles_x, ty = cam.row0 + tile / cam.tiles_x;
    const uint2 rg = ranges[tile];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    // warp w owns the 8x8 quadrant (w&1, w>>1); lane -> row lane>>2, pixel pair lane&3
    const int px = tx * SAGE_TILE + (w & 1) * 8 + (lane & 3) * 2;
    const int py = ty * SAGE_TILE + (w >> 1) * 8 + (lane >> 2);
    const bool in0 = px < cam.W && py < cam.H, in1 = (px + 1) < cam.W && py < cam.H;
    // A finished (or off-image) pixel gets x = NaN: its power2 is NaN, every comparison on it is
    // false, so it drops out of the blend without carrying an "active" flag through the loop.
    const float NAN_X = __int_as_float(0x7fc00000);
    float2 pxc = make_float2(in0 ? (float)px + 0.5f : NAN_X, in1 ? (float)(px + 1) + 0.5f : NAN_X);
    const float pyc = (float)py + 0.5f;
    float2 T = f2(1.0f), Cr = f2(0.0f), Cg = f2(0.0f), Cb = f2(0.0f), Dz = f2(0.0f);
#define SAGE_LIVE() (pxc.x == pxc.x || pxc.y == pxc.y)

    if (tid == 0) {
        mbar_init(&sbar[0], 1);
        mbar_init(&sbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // issue the TMA gather of one batch into `stage`; v = this thread's instance value of the batch
    auto issue = [&](int stage, uint32_t batch_base, uint32_t v) {
        const uint32_t c = min((uint32_t)C_THREADS, rg.y - batch_base);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the stage
        if (tid == 0) mbar_expect_tx(&sbar[stage], c * 48u);
        smask_s[stage][tid] = (uint32_t)tid < c ? (v & 15u) : 0u;
        if ((uint32_t)tid < c) bulk_g2s(&srec[stage][tid * 3], rec + (size_t)(v >> 4) * 3, 48u, &sbar[stage]);
    };
    uint32_t v_pref = (rg.x + tid < rg.y) ? __ldg(&vals[rg.x + tid]) : 0u;
    bool pending[2] = {false, false};
    uint32_t parity[2] = {0u, 0u};
    if (rg.x < rg.y) {
        issue(0, rg.x, v_pref);
        pending[0] = true;
        v_pref = (rg.x + C_THREADS + tid < rg.y) ? __ldg(&vals[rg.x + C_THREADS + tid]) : 0u;
    }
    __syncthreads();                                   // stage-0 masks visible to every warp

    int b = 0;
    for (uint32_t base = rg.x; base < rg.y; base += C_THREADS, ++b) {
        const int cnt = min((uint32_t)C_THREADS, rg.y - base);
        const int st = b & 1;
        if (base + C_THREADS < rg.y) {                 // next batch: gather in flight while this one is blended
            issue(st ^ 1, base + C_THREADS, v_pref);
            pending[st ^ 1] = true;
            v_pref = (base + 2 * C_THREADS + tid < rg.y) ? __ldg(&vals[base + 2 * C_THREADS + tid]) : 0u;
        }
        mbar_wait(&sbar[st], parity[st]);
        parity[st] ^= 1u;
        pending[st] = false;
        const float4* s_rec = srec[st];
        const uint32_t* smask = smask_s[st];
        // warp-uniform liveness: every lane takes part in the ballots below
        if (__any_sync(0xffffffffu, SAGE_LIVE())) {
#pragma unroll 1
            for (int k = 0; k < C_THREADS / 32; ++k) {
                if (k * 32 >= cnt) break;
                // entries of this 32-chunk whose bounding box reaches this warp's quadrant
                unsigned m = __ballot_sync(0xffffffffu, (smask[k * 32 + lane] >> w) & 1u);
                while (m) {
                    const int j = k * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const float4 a = s_rec[3 * j];                            // gx gy qa thr
                    const float2 bq = *reinterpret_cast<const float2*>(&s_rec[3 * j + 1]);   // qb qc
                    const float dy = pyc - a.y;
                    const float2 t2 = __fmul2_rn(bq, f2(dy));                 // bdy, cdy
                    const float cdy2 = t2.y * dy;
                    const float2 dx = __fadd2_rn(pxc, f2(-a.x));
                    const float2 wv = __ffma2_rn(f2(a.z), dx, f2(t2.x));
                    const float2 pw = __ffma2_rn(dx, wv, f2(cdy2));
                    // a.w: conservative skip threshold (log2 units): below it alpha < 1/255 for sure
                    if (!(pw.x >= a.w || pw.y >= a.w)) continue;
                    bool v0 = pw.x >= a.w && pw.x <= 0.0f;
                    bool v1 = pw.y >= a.w && pw.y <= 0.0f;
                    if (!(v0 || v1)) continue;
                    const float2 oz = *reinterpret_cast<const float2*>(&s_rec[3 * j + 1].z);   // opacity, z
                    const float2 e = sage_exp2_x2(pw);
                    float2 al = __fmul2_rn(f2(oz.x), e);
                    al.x = fminf(SAGE_ALPHA_MAX, al.x);
                    al.y = fminf(SAGE_ALPHA_MAX, al.y);
                    v0 = v0 && al.x >= SAGE_ALPHA_MIN;
                    v1 = v1 && al.y >= SAGE_ALPHA_MIN;
                    const float2 om = __ffma2_rn(al, f2(-1.0f), f2(1.0f));   // 1 - alpha
                    const float2 Tn = __fmul2_rn(T, om);
                    // keep: the splat is blended; stop: it would push T below 1e-4 -> pixel done
                    const bool keep0 = v0 && !(Tn.x < SAGE_T_MIN);
                    const bool keep1 = v1 && !(Tn.y < SAGE_T_MIN);
                    pxc.x = (v0 && !keep0) ? NAN_X : pxc.x;
                    pxc.y = (v1 && !keep1) ? NAN_X : pxc.y;
                    float2 wgt = __fmul2_rn(al, T);
                    wgt.x = keep0 ? wgt.x : 0.0f;     // a zero weight leaves C, Dz bit-identical
                    wgt.y = keep1 ? wgt.y : 0.0f;
                    T.x = keep0 ? Tn.x : T.x;
                    T.y = keep1 ? Tn.y : T.y;
                    const float4 c = s_rec[3 * j + 2];
                    Cr = __ffma2_rn(f2(c.x), wgt, Cr);
                    Cg = __ffma2_rn(f2(c.y), wgt, Cg);
                    Cb = __ffma2_rn(f2(c.z), wgt, Cb);
                    Dz = __ffma2_rn(f2(oz.y), wgt, Dz);
                }
                if (!__any_sync(0xffffffffu, SAGE_LIVE())) break;
            }
        }
        if (__syncthreads_and(!SAGE_LIVE())) break;
    }
    // a gather issued for a batch nobody will read must still land before the CTA retires
    if (pending[0]) mbar_wait(&sbar[0], parity[0]);
    if (pending[1]) mbar_wait(&sbar[1], parity[1]);
#undef SAGE_LIVE

    // epilogue: quantise and store to every destination frame.  Wide path (frame width a multiple of
    // 8, 8-byte aligned buffers): the 24 RGB bytes of a quadrant row (4 lanes x 2 px) leave as three
    // 8-byte words assembled with shuffles, depth as one float2 per lane — 8x fewer store
    // instructions, which matters most when the destinations are peer GPUs behind NVLink.
    uint8_t q[6];
    {
        const float vals_c[6] = {Cr.x, Cg.x, Cb.x, Cr.y, Cg.y, Cb.y};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float v = fminf(fmaxf(vals_c[k], 0.0f), 1.0f);
            q[k] = (uint8_t)(v * 255.0f + 0.5f);
        }
    }
    const float cov0 = 1.0f - T.x, cov1 = 1.0f - T.y;
    float d0 = cov0 > 1e-6f ? Dz.x / cov0 : 0.0f;
    float d1 = cov1 > 1e-6f ? Dz.y / cov1 : 0.0f;
    if (cam.depth_lo <= cam.depth_hi) {                 // optional output clamp (façade: clip(0.1, 6.5))
        d0 = fminf(fmaxf(d0, cam.depth_lo), cam.depth_hi);
        d1 = fminf(fmaxf(d1, cam.depth_lo), cam.depth_hi);
    }
    const size_t p = (size_t)py * cam.W + px;
    // whole-warp decision: the quadrant lies fully inside the image and the layout is aligned
    const int qx0 = tx * SAGE_TILE + (w & 1) * 8, qy0 = ty * SAGE_TILE + (w >> 1) * 8;
    const bool wide = out.wide && (qx0 + 8 <= cam.W) && (qy0 + 8 <= cam.H);
    if (wide) {
        // this lane's 6 bytes as a 64-bit value; lanes 4r..4r+3 form one row
        const unsigned long long mine = (unsigned long long)q[0] | ((unsigned long long)q[1] << 8) |
                                        ((unsigned long long)q[2] << 16) | ((unsigned long long)q[3] << 24) |
                                        ((unsigned long long)q[4] << 32) | ((unsigned long long)q[5] << 40);
        const int sub = lane & 3;
        const unsigned long long nxt = __shfl_down_sync(0xffffffffu, mine, 1);
        // word k of the row (k = 0,1,2) starts at byte 8k = lane k's byte (8k - 6k) = 2k
        unsigned long long word = 0;
        if (sub == 0) word = mine | (nxt << 48);                       // b0..5 of lane0 + b0..1 of lane1
        else if (sub == 1) word = (mine >> 16) | (nxt << 32);          // b2..5 of lane1 + b0..3 of lane2
        else if (sub == 2) word = (mine >> 32) | (nxt << 16);          // b4..5 of lane2 + b0..5 of lane3
        const size_t row_byte = ((size_t)py * cam.W + qx0) * 3;
        for (int k = 0; k < out.n; ++k) {
            if (sub < 3) *reinterpret_cast<unsigned long long*>(out.rgb[k] + row_byte + 8 * sub) = word;
            if (out.depth[k]) *reinterpret_cast<float2*>(out.depth[k] + p) = make_float2(d0, d1);
        }
    } else if (in0 || in1) {
        for (int k = 0; k < out.n; ++k) {
            uint8_t* rgb = out.rgb[k];
            float* dep = out.depth[k];
            if (in0) {
                rgb[3 * p + 0] = q[0]; rgb[3 * p + 1] = q[1]; rgb[3 * p + 2] = q[2];
                if (dep) dep[p] = d0;
            }
            if (in1) {
                rgb[3 * p + 3] = q[3]; rgb[3 * p + 4] = q[4]; rgb[3 * p + 5] = q[5];
                if (dep) dep[p + 1] = d1;
            }
        }
    }
}

}  // namespace

void sage_launch_composite(const SageCamDev& cam, const SageFrame& fr, const uint32_t* sorted_vals,
                           int n_peers, uint8_t* const* rgb_peers, float* const* depth_peers,
                           cudaStream_t st)
{
    const int n_tiles = (cam.row1 - cam.row0) * cam.tiles_x;
    if (n_tiles <= 0) return;
    PeerOut out;
    bool aligned = (cam.W % 8) == 0;
    out.n = n_peers < MAX_PEERS ? n_peers : MAX_PEERS;
    for (int k = 0; k < MAX_PEERS; ++k) {
        out.rgb[k] = k < out.n ? rgb_peers[k] : nullptr;
        out.depth[k] = (k < out.n && depth_peers) ? depth_peers[k] : nullptr;
        if (k < out.n && (((uintptr_t)out.rgb[k] & 7u) || ((uintptr_t)out.depth[k] & 7u))) aligned = false;
    }
    out.wide = aligned ? 1 : 0;
    k_composite<<<n_tiles, C_THREADS, 0, st>>>(cam, fr.rec, sorted_vals, fr.ranges, fr.tile_order, out);
}
