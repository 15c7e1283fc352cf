// K2 (offset scan + instance emission) and K4 (tile ranges).
//
// Instances are emitted in depth order (the order[] produced by the depth sort), so the stable
// by-tile radix sort that follows leaves each tile's list depth-sorted.
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

// partial[b] = sum of tiles touched by sorted Gaussians [256 b, 256 b + 256)
__global__ void __launch_bounds__(B_THREADS)
k_tiles_partial(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ntiles,
                const uint32_t* __restrict__ counters, uint32_t* __restrict__ partial)
{
    const uint32_t V = counters[CNT_VISIBLE];
    const uint32_t r = blockIdx.x * B_THREADS + threadIdx.x;
    if (blockIdx.x * B_THREADS >= V) return;
    uint32_t c = r < V ? __ldg(&ntiles[__ldg(&order[r])]) : 0u;
    __shared__ uint32_t sm[B_THREADS / 32];
    uint32_t s = c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < B_THREADS / 32; ++k) t += sm[k];
        partial[blockIdx.x] = t;
    }
}

// single block: exclusive scan of partial[0..nbv); total instance count -> counters
__global__ void __launch_bounds__(1024)
k_scan_partials(uint32_t* __restrict__ partial, uint32_t* __restrict__ counters, uint32_t cap)
{
    const uint32_t V = counters[CNT_VISIBLE];
    const uint32_t nbv = (V + B_THREADS - 1) / B_THREADS;
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t start = 0; start < nbv; start += 1024) {
        uint32_t idx = start + threadIdx.x;
        unsigned long long v = idx < nbv ? (unsigned long long)partial[idx] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            unsigned long long t = wsum[lane];
            unsigned long long ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            wsum[lane] = ti - t;
        }
        __syncthreads();
        unsigned long long carry = carry_s;
        unsigned long long ex = carry + wsum[w] + (incl - v);
        if (idx < nbv) partial[idx] = (uint32_t)(ex > 0xFFFFFFFFull ? 0xFFFFFFFFull : ex);
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ex + v;   // last thread's inclusive prefix == running total
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long total = carry_s;
        if (total > (unsigned long long)cap) {
            counters[CNT_OVERFLOW] = 1u;
            total = cap;
        }
        counters[CNT_INSTANCES] = (uint32_t)total;
    }
}

// which 8x8 quadrants of tile (tx,ty) the inclusive pixel box [px0,px1]x[py0,py1] overlaps:
// bit (qy*2 + qx), matching the warp -> quadrant map of the composite kernel
__device__ __forceinline__ uint32_t quad_mask(int px0, int px1, int py0, int py1, int tx, int ty)
{
    const int bx = tx * SAGE_TILE, by = ty * SAGE_TILE;
    const uint32_t xl = (px0 <= bx + 7) ? 1u : 0u, xr = (px1 >= bx + 8) ? 1u : 0u;
    const uint32_t yt = (py0 <= by + 7) ? 1u : 0u, yb = (py1 >= by + 8) ? 1u : 0u;
    return (yt & xl) | ((yt & xr) << 1) | ((yb & xl) << 2) | ((yb & xr) << 3);
}

// emit (tile, gaussian | quadrant mask) pairs in depth order
__global__ void __launch_bounds__(B_THREADS)
k_emit(SageCamDev cam, const uint32_t* __restrict__ order, const uint32_t* __restrict__ ntiles,
       const uint2* __restrict__ rect, const uint32_t* __restrict__ partial,
       const uint32_t* __restrict__ counters, uint32_t* __restrict__ ikey,
       uint32_t* __restrict__ ival, uint32_t cap)
{
    const uint32_t V = counters[CNT_VISIBLE];
    if (blockIdx.x * B_THREADS >= V) return;
    const uint32_t r = blockIdx.x * B_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t g = 0, c = 0;
    uint2 rc = make_uint2(0, 0);
    if (r < V) {
        g = __ldg(&order[r]);
        c = __ldg(&ntiles[g]);
        rc = __ldg(&rect[g]);
    }
    __shared__ uint32_t wsum[B_THREADS / 32];
    __shared__ uint32_t big_q[B_THREADS];
    __shared__ uint32_t big_off[B_THREADS];
    __shared__ uint32_t big_n;
    if (threadIdx.x == 0) big_n = 0;
    uint32_t incl = warp_incl_scan_u32(c);
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int k = 0; k < B_THREADS / 32; ++k) woff += (k < w) ? wsum[k] : 0u;
    uint32_t off = partial[blockIdx.x] + woff + (incl - c);

    const int tiles_x = cam.tiles_x, row0 = cam.row0;
    if (c > 32u) {                       // large footprints are spread over the whole block
        uint32_t q = atomicAdd(&big_n, 1u);
        big_q[q] = threadIdx.x;
        big_off[q] = off;
    } else if (c > 0u) {
        const int px0 = rc.x & 0xFFFF, px1 = rc.x >> 16, py0 = rc.y & 0xFFFF, py1 = rc.y >> 16;
        uint32_t o = off;
        for (int ty = py0 / SAGE_TILE; ty <= py1 / SAGE_TILE; ++ty)
            for (int tx = px0 / SAGE_TILE; tx <= px1 / SAGE_TILE; ++tx, ++o)
                if (o < cap) {
                    ikey[o] = (uint32_t)((ty - row0) * tiles_x + tx);
                    ival[o] = g | (quad_mask(px0, px1, py0, py1, tx, ty) << SAGE_ID_BITS);
                }
    }
    // broadcast the big ones through shared memory
    __shared__ uint32_t s_g[B_THREADS];
    __shared__ uint2 s_rc[B_THREADS];
    s_g[threadIdx.x] = g;
    s_rc[threadIdx.x] = rc;
    __syncthreads();
    const uint32_t nbig = big_n;
    for (uint32_t q = 0; q < nbig; ++q) {
        const uint32_t t = big_q[q];
        const uint32_t bo = big_off[q];
        const uint2 brc = s_rc[t];
        const uint32_t bg = s_g[t];
        const int px0 = brc.x & 0xFFFF, px1 = brc.x >> 16, py0 = brc.y & 0xFFFF, py1 = brc.y >> 16;
        const int bx0 = px0 / SAGE_TILE, by0 = py0 / SAGE_TILE;
        const int bw = px1 / SAGE_TILE - bx0 + 1;
        const uint32_t bc = (uint32_t)bw * (uint32_t)(py1 / SAGE_TILE - by0 + 1);
        for (uint32_t k = threadIdx.x; k < bc; k += B_THREADS) {
            uint32_t o = bo + k;
            if (o < cap) {
                int ty = by0 + (int)(k / (uint32_t)bw), tx = bx0 + (int)(k % (uint32_t)bw);
                ikey[o] = (uint32_t)((ty - row0) * tiles_x + tx);
                ival[o] = bg | (quad_mask(px0, px1, py0, py1, tx, ty) << SAGE_ID_BITS);
            }
        }
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

}  // namespace

void sage_launch_binning(const SageCamDev& cam, const SageFrame& fr, const uint32_t* order,
                         int64_t n, int64_t cap, int* launches, cudaStream_t st)
{
    if (n <= 0) return;
    int blocks = (int)((n + B_THREADS - 1) / B_THREADS);
    k_tiles_partial<<<blocks, B_THREADS, 0, st>>>(order, fr.ntiles, fr.counters, fr.partial);
    k_scan_partials<<<1, 1024, 0, st>>>(fr.partial, fr.counters, (uint32_t)cap);
    k_emit<<<blocks, B_THREADS, 0, st>>>(cam, order, fr.ntiles, fr.rect, fr.partial, fr.counters,
                                         fr.ikey[0], fr.ival[0], (uint32_t)cap);
    if (launches) *launches += 3;
}

void sage_launch_ranges(const SageFrame& fr, const uint32_t* sorted_keys, int64_t cap, int n_tiles,
                        int* launches, cudaStream_t st)
{
    cudaMemsetAsync(fr.ranges, 0, sizeof(uint2) * (size_t)n_tiles, st);
    int blocks = (int)((cap + B_THREADS - 1) / B_THREADS);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;      // persistent grid-stride: D is only known on the device
    k_tile_ranges<<<blocks, B_THREADS, 0, st>>>(sorted_keys, fr.counters, fr.ranges);
    if (launches) *launches += 1;
}
