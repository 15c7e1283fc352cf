// K2 (instance emission, fused with its offset scan) and the tiles' work queue (the tile ranges, K4,
// come out of the last pass of the by-tile sort: sort.cu).
//
// Binning is by 32 x 16 tiles, each instance carrying the mask of the tile's eight 8 x 8 quadrants
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
#ifndef SAGE_EMIT_CTAS
#define SAGE_EMIT_CTAS 4
#endif
constexpr int E_GRID = 148 * SAGE_EMIT_CTAS;       // persistent emission blocks
constexpr int O_GRID = 148;
#endif
constexpr int ROW_POOL = 1536;        // tile-row entries of a batch's cullable splats (typically ~300 used)
constexpr int ITEM_POOL = 1024;       // (splat, tile row) work items of a batch (typically ~450)
#define ROWS_NONE 0xFFFFFFFFu
#define ITEM_NONE 0xFFFFFFFFu

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

// Emission with exact tile culling.  A batch is 256 depth-sorted Gaussians.  Every phase runs on a
// DENSE set of work items, so the lanes of a warp stay busy whatever the mix of splat sizes:
//   A1  thread / splat      box, tile rows; splats whose box spans >= 2 x 2 tiles are queued as "cullable"
//   A2  thread / cullable   the {alpha >= 1/255} ellipse against the two 8-row halves of each of its tile
//                           rows (sage_band_row): column spans + running tile counts into a row pool
//   B   block scan of the per-splat tile counts, chained scan across batches (decoupled look-back,
//       read a warp wide) -> first instance slot of every splat
//   C   thread / (splat, tile row) item: writes the row's tiles — tile id, Gaussian id << 4 | mask of
//       the 8 x 8 quadrants the splat reaches — to consecutive slots
__global__ void __launch_bounds__(B_THREADS, 4)
k_emit(SageFrame fr, const uint32_t* __restrict__ order, uint32_t* __restrict__ ikey,
       uint32_t* __restrict__ ival)
{
    const SageCamDev& cam = fr.fp->cam;
    const uint32_t epoch = fr.fp->epoch + 3u;
    const uint32_t cap = fr.fp->cap;
    const int bins_x = cam.bins_x, bin_row0 = cam.bin_row0;
    SAGE_DEP_LAUNCH_TOP();
    sage_dep_wait();
    const uint32_t V = fr.counters[CNT_VISIBLE];
    const uint32_t nblk = (V + B_THREADS - 1) / B_THREADS;
    unsigned long long* chain = fr.chain;
    uint32_t* ticket = fr.tickets + SAGE_TICKET_EMIT;
    __shared__ uint32_t s_off[B_THREADS];           // exclusive instance offsets of the batch's splats
    __shared__ uint32_t s_cnt[B_THREADS];           // tiles of the splat
    __shared__ uint32_t s_g[B_THREADS];
    __shared__ uint2 s_rc[B_THREADS];
    __shared__ uint32_t s_rows[B_THREADS];          // first pool entry of the splat, or ROWS_NONE (box mode)
    __shared__ uint32_t s_cull[B_THREADS];          // queue of cullable splats (thread ids)
    __shared__ uint2 s_pool[ROW_POOL];              // (upper, lower) column spans of one tile row
    __shared__ uint32_t s_cum[ROW_POOL];            // tiles of the splat in the rows before this one
    __shared__ uint32_t s_item[ITEM_POOL];          // (splat thread id) | (row j) << 8
    __shared__ uint32_t wsum[B_THREADS / 32];
    __shared__ uint32_t s_blk, s_base, s_pool_used, s_ncull, s_nitem;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    auto put = [&](uint32_t slot, uint32_t tile, uint32_t v) {
        if (slot < cap) { ikey[slot] = tile; ival[slot] = v; }
    };

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_blk = atomicAdd(ticket, 1u); s_pool_used = 0; s_ncull = 0; s_nitem = 0; }
        __syncthreads();
        const uint32_t b = s_blk;
        if (b >= nblk) break;
        // ---- A1: one splat per thread ----
        const uint32_t r = b * B_THREADS + threadIdx.x;
        uint32_t g = 0, c = 0, nrows = 0;
        uint2 rc = make_uint2(0, 0);
        bool overflow_rows = false;                  // its rows do not fit the item pool: expanded by its own thread
        if (r < V) {
            g = __ldg(&order[r]);
            rc = __ldg(&fr.rect[g]);
            const int px0 = rc.x & 0xFFFF, px1 = rc.x >> 16, py0 = rc.y & 0xFFFF, py1 = rc.y >> 16;
            nrows = (uint32_t)(py1 / SAGE_BIN_H - py0 / SAGE_BIN_H + 1);
            c = (uint32_t)(px1 / SAGE_BIN_W - px0 / SAGE_BIN_W + 1) * nrows;
            if (sage_box_cullable(px0, px1, py0, py1)) s_cull[atomicAdd(&s_ncull, 1u)] = threadIdx.x;
            const uint32_t it0 = atomicAdd(&s_nitem, nrows);
            if (it0 + nrows <= ITEM_POOL) {
                for (uint32_t j = 0; j < nrows; ++j) s_item[it0 + j] = threadIdx.x | (j << 8);
            } else {
                overflow_rows = true;
                for (uint32_t k = it0; k < ITEM_POOL; ++k) s_item[k] = ITEM_NONE;    // the tail it could not use
            }
        }
        s_g[threadIdx.x] = g;
        s_rc[threadIdx.x] = rc;
        s_rows[threadIdx.x] = ROWS_NONE;
        s_cnt[threadIdx.x] = c;
        __syncthreads();
        // ---- A2: one cullable splat per thread (dense) ----
        for (uint32_t k = threadIdx.x; k < s_ncull; k += B_THREADS) {
            const uint32_t t = s_cull[k];
            const uint2 q = s_rc[t];
            const int px0 = q.x & 0xFFFF, px1 = q.x >> 16, py0 = q.y & 0xFFFF, py1 = q.y >> 16;
            const int ty0 = py0 / SAGE_BIN_H;
            const uint32_t nr = (uint32_t)(py1 / SAGE_BIN_H - ty0 + 1);
            const uint32_t at = atomicAdd(&s_pool_used, nr);
            if (at + nr > ROW_POOL) continue;            // (a full pool only costs the culling, never correctness)
            const uint32_t gg = s_g[t];
            const float4 r0 = __ldg(&fr.rec[3 * (size_t)gg]);         // gx gy qa lo
            const float4 r1 = __ldg(&fr.rec[3 * (size_t)gg + 1]);     // qb qc
            const SageBand band = sage_band_make(r0.x, r0.y, r0.z, r1.x, r1.y, r0.w - SAGE_L255);
            uint32_t cum = 0;
            for (uint32_t j = 0; j < nr; ++j) {
                const SageRowItem it = sage_band_row(band, px0, px1, py0, py1, ty0 + (int)j);
                s_cum[at + j] = cum;
                s_pool[at + j] = make_uint2(it.top, it.bot);
                cum += (uint32_t)max(it.tx1 - it.tx0 + 1, 0);
            }
            s_cnt[t] = cum;
            s_rows[t] = at;
        }
        __syncthreads();
        // ---- B: offsets ----
        c = s_cnt[threadIdx.x];
        const uint32_t incl = warp_incl_scan_u32(c);
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < B_THREADS / 32; ++k) { woff += (k < w) ? wsum[k] : 0u; total += wsum[k]; }
        const uint32_t my_off = woff + incl - c;
        s_off[threadIdx.x] = my_off;
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
        // ---- C: one (splat, tile row) item per thread (dense) ----
        auto emit_row = [&](uint32_t t, uint32_t j) {
            const uint2 q = s_rc[t];
            const int px0 = q.x & 0xFFFF, px1 = q.x >> 16, py0 = q.y & 0xFFFF, py1 = q.y >> 16;
            const int ty = py0 / SAGE_BIN_H + (int)j;
            const uint32_t at = s_rows[t], gv = s_g[t] << SAGE_MASK_BITS;
            const uint32_t rowtile = (uint32_t)((ty - bin_row0) * bins_x);
            if (at == ROWS_NONE) {
                const int t0 = px0 / SAGE_BIN_W, t1 = px1 / SAGE_BIN_W;
                uint32_t slot = base + s_off[t] + j * (uint32_t)(t1 - t0 + 1);
                for (int tx = t0; tx <= t1; ++tx) put(slot++, rowtile + (uint32_t)tx, gv | sage_box_quad_mask(px0, px1, py0, py1, tx, ty));
            } else {
                const uint2 sp = s_pool[at + j];
                const SageRowItem ri = sage_row_from_spans(sp.x, sp.y);
                uint32_t slot = base + s_off[t] + s_cum[at + j];
                for (int tx = ri.tx0; tx <= ri.tx1; ++tx) put(slot++, rowtile + (uint32_t)tx, gv | sage_row_quad_mask(sp.x, sp.y, tx));
            }
        };
        const uint32_t n_items = min(s_nitem, (uint32_t)ITEM_POOL);
        for (uint32_t k = threadIdx.x; k < s_nitem; k += B_THREADS) {
            if (k >= n_items) break;
            const uint32_t it = s_item[k];
            if (it != ITEM_NONE) emit_row(it & 0xFFu, it >> 8);
        }
        if (overflow_rows)                                   // (splats taller than the item pool has room for)
            for (uint32_t j = 0; j < nrows; ++j) emit_row(threadIdx.x, j);
    }
    SAGE_DEP_LAUNCH_END();
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
    SAGE_DEP_LAUNCH_TOP();
    sage_dep_wait();
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
    SAGE_KLAUNCH_PDL(k_emit, E_GRID, B_THREADS, 0, st, fr, order, fr.ikey[0], fr.ival[0]);
}

void sage_launch_bin_order(const SageFrame& fr, cudaStream_t st)
{
    SAGE_KLAUNCH_PDL(k_bin_order, O_GRID, B_THREADS, 0, st, fr);
}

void sage_bin_configure(void)
{
#ifndef SAGE_CUDA_EMU
    sage_prefer_smem(k_emit);
    sage_prefer_smem(k_bin_order);
#endif
}
