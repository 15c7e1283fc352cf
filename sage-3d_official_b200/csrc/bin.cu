// K2 (instance emission, fused with its offset scan) and the bins' work queue (the bin ranges, K4,
// come out of the last pass of the by-bin sort: sort.cu).
//
// The image is binned at 8 x 8 pixels — the footprint one warp of the composite kernel blends (two
// pixels per lane) — not at the 16 x 16 tiles of round 1: the splats of an interior scene are small
// (a visible Gaussian passes the alpha test on ~40 pixels of a 16 x 16 tile it touches), so a
// per-tile list was mostly filtered again per quadrant by ballots, and the four warps of a tile
// waited for each other at a barrier per batch.  With one list per warp-sized bin every entry a
// warp reads is its own, warps never synchronise with each other, and the work items are 4x
// smaller (better tails, also at 640 x 480).  Price: ~17 % more instances through the by-bin sort.
//
// Instances are emitted in depth order (order[] from the depth sort), so the stable by-bin radix
// sort that follows leaves every bin's list depth-sorted.  The exclusive scan of the per-Gaussian
// bin counts is fused into the emission kernel as a chained scan (decoupled look-back over one
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
constexpr int ROW_POOL = 2048;        // band entries of a batch's cullable splats (typically ~600 used)
constexpr int E_IPT = 8;              // consecutive instance slots one thread fills per round
#define ROWS_NONE 0xFFFFFFFFu
#define BAND_EMPTY 0x00000001u        /* b0 = 1 > b1 = 0 */

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

// Emission with exact bin culling.  A batch is 256 depth-sorted Gaussians, one per thread.  Setup: a
// splat whose pixel box spans >= 2 x 2 bins intersects its {alpha >= 1/255} ellipse with every band
// of 8 pixel rows (sage_band_bins) and keeps the bin spans + running counts in a shared-memory pool;
// the others (one row or one column of bins: nothing to cull) keep their box.  The per-Gaussian totals
// are scanned (chained scan across batches).  Expansion is load-balanced over the OUTPUT: a thread
// takes E_IPT consecutive instance slots, finds the owner of the first one (binary search over the
// offsets, then over the owner's bands) and walks on from there — bin, next band, next Gaussian — so
// every thread writes one 32-byte run per array whatever the splat sizes.
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
    __shared__ uint32_t s_pool[ROW_POOL];           // bin span of one band: b0 | b1 << 16 (BAND_EMPTY: none)
    __shared__ uint32_t s_cum[ROW_POOL];            // instances of the splat in the bands before this one
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
            const int by0 = py0 / SAGE_BIN, by1 = py1 / SAGE_BIN;
            const uint32_t nrows = (uint32_t)(by1 - by0 + 1);
            c = (uint32_t)(px1 / SAGE_BIN - px0 / SAGE_BIN + 1) * nrows;
            if (sage_box_cullable(px0, px1, py0, py1)) {
                const uint32_t at = atomicAdd(&s_pool_used, nrows);
                if (at + nrows <= ROW_POOL) {            // (a full pool only costs the culling, never correctness)
                    const float4 r0 = __ldg(&fr.rec[3 * (size_t)g]);         // gx gy qa lo
                    const float4 r1 = __ldg(&fr.rec[3 * (size_t)g + 1]);     // qb qc
                    const SageBand band = sage_band_make(r0.x, r0.y, r0.z, r1.x, r1.y, r0.w - SAGE_L255);
                    uint32_t cum = 0;
                    for (uint32_t j = 0; j < nrows; ++j) {
                        int b0, b1;
                        sage_band_bins(band, true, px0, px1, py0, py1, by0 + (int)j, &b0, &b1);
                        s_cum[at + j] = cum;
                        if (b0 <= b1) { cum += (uint32_t)(b1 - b0 + 1); s_pool[at + j] = (uint32_t)b0 | ((uint32_t)b1 << 16); }
                        else s_pool[at + j] = BAND_EMPTY;
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
            // state of the walk: owner's box / pool, band j of nb, bin bx of [b0, b1]
            uint2 bx2 = s_rc[own];
            uint32_t at = s_rows[own], gg = s_g[own];
            int bx0 = (int)(bx2.x & 0xFFFF) / SAGE_BIN, bx1 = (int)(bx2.x >> 16) / SAGE_BIN;
            int by0 = (int)(bx2.y & 0xFFFF) / SAGE_BIN, nb = (int)(bx2.y >> 16) / SAGE_BIN - by0 + 1;
            const uint32_t local = k0 - s_off[own];
            int j, b0, b1, bx;
            if (at == ROWS_NONE) {
                const uint32_t bw = (uint32_t)(bx1 - bx0 + 1);
                j = (int)(local / bw);
                b0 = bx0; b1 = bx1; bx = bx0 + (int)(local - (uint32_t)j * bw);
            } else {
                int lo = 0, hi = nb - 1;                   // band = last j with cum[j] <= local (it is not empty)
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_cum[at + mid] <= local) lo = mid; else hi = mid - 1;
                }
                j = lo;
                const uint32_t sp = s_pool[at + j];
                b0 = (int)(sp & 0xFFFFu); b1 = (int)(sp >> 16);
                bx = b0 + (int)(local - s_cum[at + j]);
            }
            for (uint32_t k = k0; k < kend; ++k) {
                const uint32_t slot = base + k;
                if (slot < cap) {
                    ikey[slot] = (uint32_t)((by0 + j - bin_row0) * bins_x + bx);
                    ival[slot] = gg;
                }
                if (k + 1 == kend) break;
                if (++bx <= b1) continue;
                for (;;) {                                 // next non-empty band, of the next splat if need be
                    if (++j >= nb) {
                        do { ++own; } while (s_off[own + 1] == s_off[own]);     // k + 1 < total: one exists
                        bx2 = s_rc[own]; at = s_rows[own]; gg = s_g[own];
                        bx0 = (int)(bx2.x & 0xFFFF) / SAGE_BIN; bx1 = (int)(bx2.x >> 16) / SAGE_BIN;
                        by0 = (int)(bx2.y & 0xFFFF) / SAGE_BIN; nb = (int)(bx2.y >> 16) / SAGE_BIN - by0 + 1;
                        j = 0;
                    }
                    if (at == ROWS_NONE) { b0 = bx0; b1 = bx1; }
                    else {
                        const uint32_t sp = s_pool[at + j];
                        b0 = (int)(sp & 0xFFFFu); b1 = (int)(sp >> 16);
                    }
                    if (b0 <= b1) { bx = b0; break; }
                }
            }
        }
    }
}

// Heaviest bins first: a bin is blended by one warp however long its list is, so the longest lists
// must start early or they become the tail of the composite kernel.  Bins are appended to one
// segment per power-of-two length class (order inside a class is irrelevant to the result); the
// composite's warps then draw tickets and walk the classes from the heaviest down.
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
