// K2 (instance emission, fused with its offset scan) and the tiles' work queue (the tile ranges, K4,
// come out of the last pass of the by-tile sort: sort.cu).
//
// Binning is by 16 x 16 tiles, each instance carrying the mask of the tile's four 8 x 8 quadrants
// the splat really reaches (exact culling: its {alpha >= 1/255} ellipse against the two 8-row halves
// of every tile row).  The composite blends a quadrant per warp and drops the instances whose mask
// bit is clear while it gathers, so the lists stay tile-sized (D ~ 2.4 M at 1080p) while the blend
// works at quadrant granularity.  (Binning at 8 x 8 directly was measured: 2.2x the instances —
// 5.2 M — through emission and the by-bin sort for the same blend work.)
//
// Instances are emitted in depth order (order[] from the depth sort), so the stable by-tile radix
// sort that follows leaves every tile's list depth-sorted.  The exclusive scan of the per-Gaussian
// tile counts is fused into the emission kernel as a chained scan (decoupled look-back over one
// 64-bit word per batch of 256 Gaussians, read a warp wide).  All kernels here are persistent (fixed
// grids, ticket loops, element counts read on the device), so their launch shapes are frame-invariant.
#include "sage_internal.cuh"

namespace {

constexpr int B_THREADS = 256;
#ifdef SAGE_CUDA_EMU
constexpr int E_GRID = 2, O_GRID = 2;
#else
constexpr int E_GRID = 148 * 4;       // persistent emission blocks
constexpr int O_GRID = 148;
#endif
constexpr int ROW_POOL = 1536;        // tile-row entries of a batch's cullable splats (typically ~300 used)
constexpr int E_IPT = 8;              // consecutive instance slots one thread fills per round
#define ROWS_NONE 0xFFFFFFFFu

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifdef SAGE_CUDA_EMU
inline unsigned long long cs_load(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void cs_store(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ unsigned long long cs_load(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cs_store(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// Emission with exact tile culling.  A batch is 256 depth-sorted Gaussians, one per thread.  Setup: a
// splat whose pixel box spans >= 2 x 2 tiles intersects its {alpha >= 1/255} ellipse with the two
// 8-row halves of every tile row (sage_band_row) and keeps the column spans + running tile counts in
// a shared-memory pool; the others (one row or one column of tiles: nothing to cull) keep their box.
// The per-Gaussian totals are scanned (chained scan across batches).  Expansion is load-balanced over
// the OUTPUT: a thread takes E_IPT consecutive instance slots, finds the owner of the first one
// (binary search over the offsets, then over the owner's tile rows) and walks on from there — tile,
// next row, next Gaussian — so every thread writes one 32-byte run per array whatever the splat sizes.
__global__ void __launch_bounds__(B_THREADS, 4)
k_emit(SageFrame fr, const uint32_t* __restrict__ order, uint32_t* __restrict__ ikey,
       uint32_t* __restrict__ ival)
{
    const SageCamDev& cam = fr.fp->cam;
    const uint32_t epoch = fr.fp->epoch + 3u;
    const uint32_t cap = fr.fp->cap;
    const int bins_x = cam.bins_x, bin_row0 = cam.bin_row0;
    const uint32_t V = fr.counters[CNT_VISIBLE];
    const uint32_t nblk = (V + B_THREADS - 1) / B_THREADS;
    unsigned long long* chain = fr.chain;
    uint32_t* ticket = fr.tickets + SAGE_TICKET_EMIT;
    __shared__ uint32_t s_off[B_THREADS + 1];       // exclusive instance offsets of the batch's splats; [256] = total
    __shared__ uint32_t s_g[B_THREADS];
    __shared__ uint2 s_rc[B_THREADS];
    __shared__ uint32_t s_rows[B_THREADS];          // first pool entry of the splat, or ROWS_NONE (box mode)
    __shared__ uint2 s_pool[ROW_POOL];              // (upper, lower) column spans of one tile row
    __shared__ uint32_t s_cum[ROW_POOL];            // tiles of the splat in the rows before this one
    __shared__ uint32_t wsum[B_THREADS / 32];
    __shared__ uint32_t s_blk, s_base, s_pool_used;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_blk = atomicAdd(ticket, 1u); s_pool_used = 0; }
        __syncthreads();
        const uint32_t b = s_blk;
        if (b >= nblk) break;
        const uint32_t r = b * B_THREADS + threadIdx.x;
        uint32_t g = 0, c = 0, rows_at = ROWS_NONE;
        uint2 rc = make_uint2(0, 0);
        if (r < V) {
            g = __ldg(&order[r]);
            rc = __ldg(&fr.rect[g]);
            const int px0 = rc.x & 0xFFFF, px1 = rc.x >> 16, py0 = rc.y & 0xFFFF, py1 = rc.y >> 16;
            const int ty0 = py0 / SAGE_BIN, ty1 = py1 / SAGE_BIN;
            const uint32_t nrows = (uint32_t)(ty1 - ty0 + 1);
            c = (uint32_t)(px1 / SAGE_BIN - px0 / SAGE_BIN + 1) * nrows;
            if (sage_box_cullable(px0, px1, py0, py1)) {
                const uint32_t at = atomicAdd(&s_pool_used, nrows);
                if (at + nrows <= ROW_POOL) {            // (a full pool only costs the culling, never correctness)
                    const float4 r0 = __ldg(&fr.rec[3 * (size_t)g]);         // gx gy qa lo
                    const float4 r1 = __ldg(&fr.rec[3 * (size_t)g + 1]);     // qb qc
                    const SageBand band = sage_band_make(r0.x, r0.y, r0.z, r1.x, r1.y, r0.w - SAGE_L255);
                    uint32_t cum = 0;
                    for (uint32_t j = 0; j < nrows; ++j) {
                        const SageRowItem it = sage_band_row(band, px0, px1, py0, py1, ty0 + (int)j);
                        s_cum[at + j] = cum;
                        s_pool[at + j] = make_uint2(it.top, it.bot);
                        cum += (uint32_t)max(it.tx1 - it.tx0 + 1, 0);
                    }
                    c = cum;
                    rows_at = at;
                }
            }
        }
        uint32_t incl = warp_incl_scan_u32(c);
        if (lane == 31) wsum[w] = incl;
        s_g[threadIdx.x] = g;
        s_rc[threadIdx.x] = rc;
        s_rows[threadIdx.x] = rows_at;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < B_THREADS / 32; ++k) { woff += (k < w) ? wsum[k] : 0u; total += wsum[k]; }
        s_off[threadIdx.x] = woff + incl - c;
        if (threadIdx.x == 0) s_off[B_THREADS] = total;

        // chained scan across batches: chain[b] = epoch(30) | state(2) | value(32); warp 0 looks back
        // over 32 predecessors per round trip
        if (w == 0) {
            uint32_t excl = 0;
            if (b == 0) {
                if (lane == 0) cs_store(chain, ((unsigned long long)epoch << 34) | (2ull << 32) | total);
            } else {
                if (lane == 0) cs_store(chain + b, ((unsigned long long)epoch << 34) | (1ull << 32) | total);
                int t = (int)b - 1;
                uint32_t spins = 0;
                for (;;) {
                    const int idx = t - lane;
                    const unsigned long long word = idx >= 0 ? cs_load(chain + idx) : (((unsigned long long)epoch << 34) | (2ull << 32));
                    const bool ok = (uint32_t)(word >> 34) == epoch;
                    const bool isp = ok && ((word >> 32) & 3ull) == 2ull;
                    const unsigned m_bad = __ballot_sync(0xffffffffu, !ok);
                    const unsigned m_p = __ballot_sync(0xffffffffu, isp);
                    const int first_bad = m_bad ? __ffs(m_bad) - 1 : 32;
                    const int first_p = m_p ? __ffs(m_p) - 1 : 32;
                    const bool fin = first_p < first_bad;
                    const int take = fin ? first_p + 1 : first_bad;          // lanes [0, take) are consumed
                    excl += warp_sum_u32(lane < take ? (uint32_t)word : 0u);
                    if (fin) break;
                    t -= take;
                    if (take == 0) {                                         // predecessor has not published yet
                        if (++spins > (1u << 24)) __trap();                  // never hang the GPU
                        __nanosleep(32);
                    }
                }
                if (lane == 0) cs_store(chain + b, ((unsigned long long)epoch << 34) | (2ull << 32) | (excl + total));
            }
            if (lane == 0) {
                s_base = excl;
                if (b == nblk - 1) {                       // grand total D (clamped to the capacity)
                    unsigned long long D = (unsigned long long)excl + total;
                    fr.counters[CNT_WANTED] = (uint32_t)D;
                    if (D > cap) { fr.counters[CNT_OVERFLOW] = 1u; D = cap; }
                    fr.counters[CNT_INSTANCES] = (uint32_t)D;
                }
            }
        }
        __syncthreads();
        const uint32_t base = s_base;

        // load-balanced expansion: E_IPT consecutive instance slots per thread and round
        for (uint32_t k0 = threadIdx.x * E_IPT; k0 < total; k0 += B_THREADS * E_IPT) {
            const uint32_t kend = min(k0 + (uint32_t)E_IPT, total);
            int own = 0;
            {
                int lo = 0, hi = B_THREADS - 1;            // owner = last o with s_off[o] <= k0 (its count is > 0)
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_off[mid] <= k0) lo = mid; else hi = mid - 1;
                }
                own = lo;
            }
            // state of the walk: the owner's box / pool rows, tile row j of nr, tile tx of [t0, t1]
            uint2 bx2 = s_rc[own];
            uint32_t at = s_rows[own], gg = s_g[own];
            int px0 = (int)(bx2.x & 0xFFFF), px1 = (int)(bx2.x >> 16), py0 = (int)(bx2.y & 0xFFFF), py1 = (int)(bx2.y >> 16);
            int ty0 = py0 / SAGE_BIN, nr = py1 / SAGE_BIN - ty0 + 1;
            const uint32_t local = k0 - s_off[own];
            int j, t0, t1, tx;
            uint2 sp = make_uint2(0, 0);
            if (at == ROWS_NONE) {
                t0 = px0 / SAGE_BIN; t1 = px1 / SAGE_BIN;
                const uint32_t bw = (uint32_t)(t1 - t0 + 1);
                j = (int)(local / bw);
                tx = t0 + (int)(local - (uint32_t)j * bw);
            } else {
                int lo = 0, hi = nr - 1;                   // row = last j with cum[j] <= local (it is not empty)
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_cum[at + mid] <= local) lo = mid; else hi = mid - 1;
                }
                j = lo;
                sp = s_pool[at + j];
                const SageRowItem ri = sage_row_from_spans(sp.x, sp.y);
                t0 = ri.tx0; t1 = ri.tx1;
                tx = t0 + (int)(local - s_cum[at + j]);
            }
            for (uint32_t k = k0; k < kend; ++k) {
                const uint32_t slot = base + k;
                if (slot < cap) {
                    const int ty = ty0 + j;
                    const uint32_t mask = at == ROWS_NONE ? sage_box_quad_mask(px0, px1, py0, py1, tx, ty)
                                                          : sage_row_quad_mask(sp.x, sp.y, tx);
                    ikey[slot] = (uint32_t)((ty - bin_row0) * bins_x + tx);
                    ival[slot] = (gg << 4) | mask;
                }
                if (k + 1 == kend) break;
                if (++tx <= t1) continue;
                for (;;) {                                 // next non-empty tile row, of the next splat if need be
                    if (++j >= nr) {
                        do { ++own; } while (s_off[own + 1] == s_off[own]);     // k + 1 < total: one exists
                        bx2 = s_rc[own]; at = s_rows[own]; gg = s_g[own];
                        px0 = (int)(bx2.x & 0xFFFF); px1 = (int)(bx2.x >> 16); py0 = (int)(bx2.y & 0xFFFF); py1 = (int)(bx2.y >> 16);
                        ty0 = py0 / SAGE_BIN; nr = py1 / SAGE_BIN - ty0 + 1;
                        j = 0;
                    }
                    if (at == ROWS_NONE) { t0 = px0 / SAGE_BIN; t1 = px1 / SAGE_BIN; }
                    else {
                        sp = s_pool[at + j];
                        const SageRowItem ri = sage_row_from_spans(sp.x, sp.y);
                        t0 = ri.tx0; t1 = ri.tx1;
                    }
                    if (t0 <= t1) { tx = t0; break; }
                }
            }
        }
    }
}

// Heaviest tiles first: a quadrant is blended by one warp however long the tile's list is, so the
// longest lists must start early or they become the tail of the composite kernel.  Tiles are appended
// to one segment per power-of-two length class (order inside a class is irrelevant to the result);
// the composite's warps then draw tickets (four per tile) and walk the classes from the heaviest down.
__global__ void __launch_bounds__(B_THREADS)
k_bin_order(SageFrame fr)
{
    const int n_bins = fr.fp->cam.n_bins;
    const int lane = threadIdx.x & 31;
    const int n_round = (n_bins + 31) & ~31;
    for (int t = blockIdx.x * B_THREADS + threadIdx.x; t < n_round; t += gridDim.x * B_THREADS) {
        uint32_t cls = 0xFFFFFFFFu;
        if (t < n_bins) {
            const uint2 r = fr.ranges[t];                          // (~begin, end); (0, 0): empty
            const uint32_t n = r.y ? r.y - ~r.x : 0u;
            cls = n ? (uint32_t)(32 - __clz((int)n)) : 0u;      // class 0: empty, k: 2^(k-1) <= n < 2^k
        }
        const unsigned peers = __match_any_sync(0xffffffffu, cls);
        const int leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if (t < n_bins && lane == leader) pos = atomicAdd(&fr.class_count[cls], (uint32_t)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (t < n_bins) fr.class_seg[(size_t)cls * fr.max_bins + pos] = (uint32_t)t;
    }
}

}  // namespace

void sage_launch_emit(const SageFrame& fr, const uint32_t* order, cudaStream_t st)
{
    SAGE_KLAUNCH(k_emit, E_GRID, B_THREADS, st, fr, order, fr.ikey[0], fr.ival[0]);
}

void sage_launch_bin_order(const SageFrame& fr, cudaStream_t st)
{
    SAGE_KLAUNCH(k_bin_order, O_GRID, B_THREADS, st, fr);
}
