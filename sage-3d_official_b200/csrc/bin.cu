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

// Load-balanced two-level expansion.  A block takes 256 depth-sorted Gaussians; its unit of work is
// the ROW ITEM (Gaussian, tile row): one thread intersects the splat's {alpha >= 1/255} ellipse with
// the two 8-row halves of that tile row (sage_band_row, binary64) and gets the exact tile range and
// the column spans that later give every tile its 8x8-quadrant mask.  Items are processed 256 at a
// time: scan their tile counts, then let the threads walk the concatenated instance list (binary
// search for the owning item), so consecutive threads write consecutive instance slots (coalesced)
// whatever the splat sizes.  The per-Gaussian totals equal k_project's counts (same function, same
// inputs), which is what the chained scan across blocks is built on.
template <int TBITS_LO, int NP>
__global__ void __launch_bounds__(B_THREADS)
k_emit(SageCamDev cam, SageFrame fr, const uint32_t* __restrict__ order, uint32_t* __restrict__ ikey,
       uint32_t* __restrict__ ival, uint32_t cap, unsigned long long* __restrict__ chain,
       uint32_t* __restrict__ ticket, uint32_t epoch, uint32_t* __restrict__ tile_hist)
{
    const uint32_t V = fr.counters[CNT_VISIBLE];
    const uint32_t nblk = (V + B_THREADS - 1) / B_THREADS;
    __shared__ SageBand s_band[B_THREADS];
    __shared__ uint32_t s_roff[B_THREADS + 1];      // row items before Gaussian k of the block
    __shared__ uint32_t s_g[B_THREADS];
    __shared__ uint2 s_rc[B_THREADS];
    __shared__ uint32_t wsum[B_THREADS / 32], wsum2[B_THREADS / 32];
    __shared__ uint32_t s_blk, s_base, s_total;
    __shared__ uint32_t s_hist[NP][256];
    // one chunk of row items
    __shared__ uint32_t s_ioff[B_THREADS + 1];      // instances before item j of the chunk
    __shared__ uint32_t s_it_g[B_THREADS];          // Gaussian id << 4
    __shared__ uint32_t s_it_pos[B_THREADS];        // ty << 16 | first tile column
    __shared__ uint2 s_it_span[B_THREADS];          // column spans of the upper / lower half
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < NP * 256; i += B_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t b = s_blk;
    if (b >= nblk) return;
    const uint32_t r = b * B_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t g = 0, c = 0, nrows = 0;
    uint2 rc = make_uint2(0, 0);
    if (r < V) {
        g = __ldg(&order[r]);
        c = __ldg(&fr.ntiles[g]);
        rc = __ldg(&fr.rect[g]);
        if (sage_box_cullable(rc.x & 0xFFFF, rc.x >> 16, rc.y & 0xFFFF, rc.y >> 16)) {
            const float4 r0 = __ldg(&fr.rec[3 * (size_t)g]);         // gx gy qa thr
            const float4 r1 = __ldg(&fr.rec[3 * (size_t)g + 1]);     // qb qc o z
            s_band[threadIdx.x] = sage_band_make(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
        }
        nrows = (rc.y >> 16) / SAGE_TILE - (rc.y & 0xFFFFu) / SAGE_TILE + 1;
    }
    const uint32_t incl = warp_incl_scan_u32(c), incl2 = warp_incl_scan_u32(nrows);
    if (lane == 31) { wsum[w] = incl; wsum2[w] = incl2; }
    s_g[threadIdx.x] = g;
    s_rc[threadIdx.x] = rc;
    __syncthreads();
    uint32_t woff = 0, woff2 = 0;
#pragma unroll
    for (int k = 0; k < B_THREADS / 32; ++k) { woff += (k < w) ? wsum[k] : 0u; woff2 += (k < w) ? wsum2[k] : 0u; }
    s_roff[threadIdx.x] = woff2 + incl2 - nrows;
    if (threadIdx.x == B_THREADS - 1) { s_roff[B_THREADS] = woff2 + incl2; s_total = woff + incl; }
    __syncthreads();
    const uint32_t total = s_total, n_items = s_roff[B_THREADS];

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
    const int tiles_x = cam.tiles_x, row0 = cam.row0;
    uint32_t run = s_base;                          // instance slot of the chunk's first instance
    for (uint32_t it0 = 0; it0 < n_items; it0 += B_THREADS) {
        const uint32_t item = it0 + threadIdx.x;
        uint32_t cnt = 0;
        if (item < n_items) {
            int lo = 0, hi = B_THREADS - 1;         // owner = last o with s_roff[o] <= item
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_roff[mid] <= item) lo = mid; else hi = mid - 1;
            }
            const uint2 bx = s_rc[lo];
            const int px0 = bx.x & 0xFFFF, px1 = bx.x >> 16, py0 = bx.y & 0xFFFF, py1 = bx.y >> 16;
            const int ty = py0 / SAGE_TILE + (int)(item - s_roff[lo]);
            const SageRowItem ri = sage_band_row(s_band[lo], px0, px1, py0, py1, ty);
            cnt = (uint32_t)max(ri.tx1 - ri.tx0 + 1, 0);
            s_it_g[threadIdx.x] = s_g[lo] << 4;
            s_it_pos[threadIdx.x] = ((uint32_t)ty << 16) | (uint32_t)max(ri.tx0, 0);
            s_it_span[threadIdx.x] = make_uint2(ri.top, ri.bot);
        }
        const uint32_t ci = warp_incl_scan_u32(cnt);
        if (lane == 31) wsum[w] = ci;
        __syncthreads();
        uint32_t wo = 0;
#pragma unroll
        for (int k = 0; k < B_THREADS / 32; ++k) wo += (k < w) ? wsum[k] : 0u;
        s_ioff[threadIdx.x] = wo + ci - cnt;
        if (threadIdx.x == B_THREADS - 1) s_ioff[B_THREADS] = wo + ci;
        __syncthreads();
        const uint32_t chunk_total = s_ioff[B_THREADS];
        for (uint32_t k = threadIdx.x; k < chunk_total; k += B_THREADS) {
            int lo = 0, hi = B_THREADS - 1;         // owning item = last j with s_ioff[j] <= k
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_ioff[mid] <= k) lo = mid; else hi = mid - 1;
            }
            const uint32_t pos = s_it_pos[lo];
            const int ty = (int)(pos >> 16), tx = (int)(pos & 0xFFFFu) + (int)(k - s_ioff[lo]);
            SageRowItem ri;
            ri.top = s_it_span[lo].x; ri.bot = s_it_span[lo].y; ri.tx0 = 0; ri.tx1 = 0;
            const uint32_t tile = (uint32_t)((ty - row0) * tiles_x + tx);
            const uint32_t slot = run + k;
            if (slot < cap) {
                ikey[slot] = tile;
                ival[slot] = s_it_g[lo] | sage_row_quad_mask(ri, tx);
#pragma unroll
                for (int p = 0; p < NP; ++p) atomicAdd(&s_hist[p][(tile >> (TBITS_LO * p)) & 255u], 1u);   // (warp-aggregating these with match.any measured slower)
            }
        }
        run += chunk_total;
        __syncthreads();                            // the chunk arrays are rewritten by the next round
    }
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
