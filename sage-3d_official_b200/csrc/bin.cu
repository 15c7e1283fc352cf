// K2 (instance emission, fused with its offset scan) and K4 (tile ranges).
//
// Instances are emitted in depth order (order[] from the depth sort), so the stable by-tile radix
// sort that follows leaves every tile's list depth-sorted.  The exclusive scan of the per-Gaussian
// tile counts is fused into the emission kernel as a chained scan (decoupled look-back over one
// 64-bit word per block), and the digit histograms the tile sort needs are accumulated on the
// fly — so the instance keys are written once and never re-read before the sort passes.
#include "sage_internal.cuh"

namespace {

constexpr int B_THREADS = 256;

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

// which 8x8 quadrants of tile (tx,ty) the inclusive pixel box [px0,px1]x[py0,py1] overlaps:
// bit (qy*2 + qx), matching the warp -> quadrant map of the composite kernel
__device__ __forceinline__ uint32_t box_quad_mask(int px0, int px1, int py0, int py1, int tx, int ty)
{
    const int bx = tx * SAGE_TILE, by = ty * SAGE_TILE;
    const uint32_t xl = (px0 <= bx + 7) ? 1u : 0u, xr = (px1 >= bx + 8) ? 1u : 0u;
    const uint32_t yt = (py0 <= by + 7) ? 1u : 0u, yb = (py1 >= by + 8) ? 1u : 0u;
    return (yt & xl) | ((yt & xr) << 1) | ((yb & xl) << 2) | ((yb & xr) << 3);
}

constexpr int ROW_POOL = 1536;      // tile-row entries of the block's cullable splats (typically ~300 used)
#define ROWS_NONE 0xFFFFFFFFu

// Load-balanced expansion with exact tile culling.  A block takes 256 depth-sorted Gaussians.  Setup
// (thread per Gaussian): a splat whose pixel box spans >= 2 x 2 tiles intersects its
// {alpha >= 1/255} ellipse with the two 8-row halves of every tile row (sage_band_row) and keeps the
// column spans + running tile counts in a shared-memory pool; the others (one row or one column of
// tiles: nothing to cull) keep their box.  The per-Gaussian totals are scanned (chained scan across
// blocks, decoupled look-back), then the threads walk the concatenated instance list: binary search
// for the owning Gaussian, a second short one for its tile row, so consecutive threads write
// consecutive instance slots (coalesced) whatever the splat sizes.
template <int TBITS_LO, int NP>
__global__ void __launch_bounds__(B_THREADS, 6)
k_emit(SageCamDev cam, SageFrame fr, const uint32_t* __restrict__ order, uint32_t* __restrict__ ikey,
       uint32_t* __restrict__ ival, uint32_t cap, unsigned long long* __restrict__ chain,
       uint32_t* __restrict__ ticket, uint32_t epoch, uint32_t* __restrict__ tile_hist)
{
    const uint32_t V = fr.counters[CNT_VISIBLE];
    const uint32_t nblk = (V + B_THREADS - 1) / B_THREADS;
    __shared__ uint32_t s_off[B_THREADS + 1];
    __shared__ uint32_t s_g[B_THREADS];
    __shared__ uint2 s_rc[B_THREADS];
    __shared__ uint32_t s_rows[B_THREADS];          // first pool entry of the splat, or ROWS_NONE (box mode)
    __shared__ uint2 s_pool_span[ROW_POOL];         // (upper, lower) column spans of one tile row
    __shared__ uint16_t s_pool_cum[ROW_POOL];       // tiles of the splat up to and including this row
    __shared__ uint32_t wsum[B_THREADS / 32];
    __shared__ uint32_t s_blk, s_base, s_pool_used;
    __shared__ uint32_t s_hist[NP][256];
    if (threadIdx.x == 0) { s_blk = atomicAdd(ticket, 1u); s_pool_used = 0; }
    for (int i = threadIdx.x; i < NP * 256; i += B_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t b = s_blk;
    if (b >= nblk) return;
    const uint32_t r = b * B_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t g = 0, c = 0, rows_at = ROWS_NONE;
    uint2 rc = make_uint2(0, 0);
    if (r < V) {
        g = __ldg(&order[r]);
        rc = __ldg(&fr.rect[g]);
        const int px0 = rc.x & 0xFFFF, px1 = rc.x >> 16, py0 = rc.y & 0xFFFF, py1 = rc.y >> 16;
        const int ty0 = py0 / SAGE_TILE, ty1 = py1 / SAGE_TILE;
        c = (uint32_t)(px1 / SAGE_TILE - px0 / SAGE_TILE + 1) * (uint32_t)(ty1 - ty0 + 1);
        if (sage_box_cullable(px0, px1, py0, py1)) {
            const uint32_t nrows = (uint32_t)(ty1 - ty0 + 1);
            const uint32_t at = atomicAdd(&s_pool_used, nrows);
            if (at + nrows <= ROW_POOL) {            // (a full pool only costs the culling, never correctness)
                const float4 r0 = __ldg(&fr.rec[3 * (size_t)g]);         // gx gy qa thr
                const float4 r1 = __ldg(&fr.rec[3 * (size_t)g + 1]);     // qb qc o z
                const SageBand band = sage_band_make(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                uint32_t cum = 0;
                for (uint32_t j = 0; j < nrows; ++j) {
                    const SageRowItem it = sage_band_row(band, px0, px1, py0, py1, ty0 + (int)j);
                    cum += (uint32_t)max(it.tx1 - it.tx0 + 1, 0);
                    s_pool_span[at + j] = make_uint2(it.top, it.bot);
                    s_pool_cum[at + j] = (uint16_t)cum;
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
    uint32_t woff = 0;
#pragma unroll
    for (int k = 0; k < B_THREADS / 32; ++k) woff += (k < w) ? wsum[k] : 0u;
    s_off[threadIdx.x] = woff + incl - c;
    if (threadIdx.x == B_THREADS - 1) s_off[B_THREADS] = woff + incl;
    __syncthreads();
    const uint32_t total = s_off[B_THREADS];

    // chained scan across blocks: chain[b] = epoch(30) | state(2) | value(32)
    if (threadIdx.x == 0) {
        uint32_t excl = 0;
        if (b == 0) {
            cs_store(chain, ((unsigned long long)epoch << 34) | (2ull << 32) | total);
        } else {
            cs_store(chain + b, ((unsigned long long)epoch << 34) | (1ull << 32) | total);
            int t = (int)b - 1;
            uint32_t spins = 0;
            bool done = false;
            while (!done) {                                    // batched look-back, as in sort.cu
                unsigned long long s[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] = (t - j >= 0) ? cs_load(chain + (t - j)) : 0ull;
                int used = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!done && used == j && t - j >= 0 && (uint32_t)(s[j] >> 34) == epoch) {
                        excl += (uint32_t)s[j];
                        used = j + 1;
                        done = ((s[j] >> 32) & 3ull) == 2ull;
                    }
                }
                t -= used;
                if (!done && used == 0) {
                    if (++spins > (1u << 24)) __trap();
                    __nanosleep(32);
                }
            }
            cs_store(chain + b, ((unsigned long long)epoch << 34) | (2ull << 32) | (excl + total));
        }
        s_base = excl;
        if (b == nblk - 1) {                       // grand total D (clamped to the capacity)
            unsigned long long D = (unsigned long long)excl + total;
            fr.counters[CNT_WANTED] = (uint32_t)D;
            if (D > cap) { fr.counters[CNT_OVERFLOW] = 1u; D = cap; }
            fr.counters[CNT_INSTANCES] = (uint32_t)D;
        }
    }
    __syncthreads();
    const uint32_t base = s_base;
    const int tiles_x = cam.tiles_x, row0 = cam.row0;
    for (uint32_t k = threadIdx.x; k < total; k += B_THREADS) {
        int lo = 0, hi = B_THREADS - 1;            // owner = last o with s_off[o] <= k
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] <= k) lo = mid; else hi = mid - 1;
        }
        const uint2 bx = s_rc[lo];
        const int px0 = bx.x & 0xFFFF, px1 = bx.x >> 16, py0 = bx.y & 0xFFFF, py1 = bx.y >> 16;
        const int tx0 = px0 / SAGE_TILE, ty0 = py0 / SAGE_TILE;
        const uint32_t local = k - s_off[lo];
        const uint32_t at = s_rows[lo];
        int tx, ty;
        uint32_t mask;
        if (at == ROWS_NONE) {                     // box mode
            const uint32_t bw = (uint32_t)(px1 / SAGE_TILE - tx0 + 1);
            ty = ty0 + (int)(local / bw); tx = tx0 + (int)(local % bw);
            mask = box_quad_mask(px0, px1, py0, py1, tx, ty);
        } else {                                   // row = first j with cum[j] > local
            int a = 0, z = py1 / SAGE_TILE - ty0;
            while (a < z) {
                const int mid = (a + z) >> 1;
                if ((uint32_t)s_pool_cum[at + mid] > local) z = mid; else a = mid + 1;
            }
            const uint32_t before = a > 0 ? (uint32_t)s_pool_cum[at + a - 1] : 0u;
            const uint2 sp = s_pool_span[at + a];
            const SageRowItem ri = sage_row_from_spans(sp.x, sp.y);
            ty = ty0 + a; tx = ri.tx0 + (int)(local - before);
            mask = sage_row_quad_mask(ri, tx);
        }
        const uint32_t tile = (uint32_t)((ty - row0) * tiles_x + tx);
        const uint32_t slot = base + k;
        if (slot < cap) {
            ikey[slot] = tile;
            ival[slot] = (s_g[lo] << 4) | mask;
#pragma unroll
            for (int p = 0; p < NP; ++p) atomicAdd(&s_hist[p][(tile >> (TBITS_LO * p)) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP * 256; i += B_THREADS) {
        const uint32_t cc = (&s_hist[0][0])[i];
        if (cc) atomicAdd(&tile_hist[i], cc);
    }
}

__global__ void __launch_bounds__(B_THREADS)
k_tile_ranges(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ counters,
              uint2* __restrict__ ranges)
{
    const uint32_t D = counters[CNT_INSTANCES];
    for (uint32_t j = blockIdx.x * B_THREADS + threadIdx.x; j < D; j += gridDim.x * B_THREADS) {
        const uint32_t k = __ldg(&keys[j]);
        if (j == 0 || __ldg(&keys[j - 1]) != k) ranges[k].x = j;
        if (j == D - 1 || __ldg(&keys[j + 1]) != k) ranges[k].y = j + 1;
    }
}

// Heaviest tiles first: a tile is blended by one CTA however long its list is, so the longest lists
// must start early or they become the tail of the composite kernel.  One block orders the tiles by
// descending power-of-two length class (order inside a class is irrelevant to the result).
__global__ void __launch_bounds__(1024)
k_tile_order(const uint2* __restrict__ ranges, int n_tiles, uint32_t* __restrict__ order)
{
    __shared__ uint32_t cls_count[33], cls_base[33];
    if (threadIdx.x < 33) cls_count[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += 1024) {
        const uint2 r = ranges[t];
        const uint32_t n = r.y - r.x;
        atomicAdd(&cls_count[n ? 32 - __clz(n) : 0], 1u);      // class 0: empty, k: 2^(k-1) <= n < 2^k
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int c = 32; c >= 0; --c) { cls_base[c] = run; run += cls_count[c]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += 1024) {
        const uint2 r = ranges[t];
        const uint32_t n = r.y - r.x;
        order[atomicAdd(&cls_base[n ? 32 - __clz(n) : 0], 1u)] = (uint32_t)t;
    }
}

}  // namespace

void sage_launch_emit(const SageCamDev& cam, const SageFrame& fr, const uint32_t* order, int64_t n,
                      int64_t cap, int tile_passes, uint32_t epoch, cudaStream_t st)
{
    if (n <= 0) return;
    const int blocks = (int)((n + B_THREADS - 1) / B_THREADS);
    uint32_t* ticket = fr.tickets + SAGE_TICKET_EMIT;
    switch (tile_passes) {
        case 1: k_emit<8, 1><<<blocks, B_THREADS, 0, st>>>(cam, fr, order, fr.ikey[0], fr.ival[0], (uint32_t)cap, fr.chain, ticket, epoch, fr.tile_hist); break;
        case 2: k_emit<8, 2><<<blocks, B_THREADS, 0, st>>>(cam, fr, order, fr.ikey[0], fr.ival[0], (uint32_t)cap, fr.chain, ticket, epoch, fr.tile_hist); break;
        default: k_emit<8, 3><<<blocks, B_THREADS, 0, st>>>(cam, fr, order, fr.ikey[0], fr.ival[0], (uint32_t)cap, fr.chain, ticket, epoch, fr.tile_hist); break;
    }
}

void sage_launch_ranges(const SageFrame& fr, const uint32_t* sorted_keys, int64_t cap, int n_tiles, cudaStream_t st)
{
    cudaMemsetAsync(fr.ranges, 0, sizeof(uint2) * (size_t)n_tiles, st);
    int blocks = (int)((cap + B_THREADS - 1) / B_THREADS);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;      // persistent grid-stride: D is only known on the device
    k_tile_ranges<<<blocks, B_THREADS, 0, st>>>(sorted_keys, fr.counters, fr.ranges);
    k_tile_order<<<1, 1024, 0, st>>>(fr.ranges, n_tiles, fr.tile_order);
}
