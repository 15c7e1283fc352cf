// K2-K4: tile binning and the per-tile depth sort.
//
//   (k_project)       small footprints add 1 to the counter of each tile they touch
//   k_count_big       the same for footprints of more than SAGE_SMALL_FOOTPRINT tiles
//   k_scan_tiles      exclusive scan of the T tile counters -> per-tile ranges, D = total
//   k_expand<FILL>    every (tile, Gaussian) instance claims a slot in its tile's range (atomic
//                     cursor) and writes a 64-bit key  depth(27) << 32 | id(28) << 4 | quadrants(4)
//   k_tile_sort       one CTA per tile sorts its range by (depth, id) in shared memory
//
// The slot order produced by the atomics is arbitrary; the per-tile sort on the full (depth, id)
// key makes the final order — and with it every output bit — deterministic.  No global sort, no
// 64-bit global keys shuffled through HBM passes: each instance key is written once and read once.
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

// which 8x8 quadrants of tile (tx,ty) the inclusive pixel box [px0,px1]x[py0,py1] overlaps:
// bit (qy*2 + qx), matching the warp -> quadrant map of the composite kernel
__device__ __forceinline__ uint32_t quad_mask(int px0, int px1, int py0, int py1, int tx, int ty)
{
    const int bx = tx * SAGE_TILE, by = ty * SAGE_TILE;
    const uint32_t xl = (px0 <= bx + 7) ? 1u : 0u, xr = (px1 >= bx + 8) ? 1u : 0u;
    const uint32_t yt = (py0 <= by + 7) ? 1u : 0u, yb = (py1 >= by + 8) ? 1u : 0u;
    return (yt & xl) | ((yt & xr) << 1) | ((yb & xl) << 2) | ((yb & xr) << 3);
}

// Load-balanced expansion: a block takes 256 visible Gaussians, scans their tile counts and lets
// its threads walk the concatenated instance list (binary search for the owner), so a splat
// covering thousands of tiles costs the same per instance as one covering four.
template <bool FILL>
__global__ void __launch_bounds__(B_THREADS)
k_expand(SageCamDev cam, SageFrame fr, uint32_t cap)
{
    const uint32_t V = fr.counters[CNT_VISIBLE];
    if (blockIdx.x * B_THREADS >= V) return;
    const uint32_t r = blockIdx.x * B_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ uint32_t s_off[B_THREADS + 1];
    __shared__ uint32_t s_g[B_THREADS];
    __shared__ uint2 s_rc[B_THREADS];
    __shared__ uint32_t wsum[B_THREADS / 32];
    uint32_t g = 0, c = 0;
    uint2 rc = make_uint2(0, 0);
    if (r < V) {
        g = __ldg(&fr.vis_list[r]);
        c = __ldg(&fr.ntiles[g]);
        rc = __ldg(&fr.rect[g]);
    }
    uint32_t incl = warp_incl_scan_u32(c);
    if (lane == 31) wsum[w] = incl;
    s_g[threadIdx.x] = g;
    s_rc[threadIdx.x] = rc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int k = 0; k < B_THREADS / 32; ++k) woff += (k < w) ? wsum[k] : 0u;
    s_off[threadIdx.x] = woff + incl - c;
    if (threadIdx.x == B_THREADS - 1) s_off[B_THREADS] = woff + incl;
    __syncthreads();
    const uint32_t total = s_off[B_THREADS];
    const int tiles_x = cam.tiles_x, row0 = cam.row0;
    for (uint32_t k = threadIdx.x; k < total; k += B_THREADS) {
        // owner = last o with s_off[o] <= k
        int lo = 0, hi = B_THREADS - 1;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] <= k) lo = mid; else hi = mid - 1;
        }
        const uint2 b = s_rc[lo];
        const int px0 = b.x & 0xFFFF, px1 = b.x >> 16, py0 = b.y & 0xFFFF, py1 = b.y >> 16;
        const int bx0 = px0 / SAGE_TILE, by0 = py0 / SAGE_TILE;
        const uint32_t bw = (uint32_t)(px1 / SAGE_TILE - bx0 + 1);
        const uint32_t local = k - s_off[lo];
        const int ty = by0 + (int)(local / bw), tx = bx0 + (int)(local % bw);
        const uint32_t tile = (uint32_t)((ty - row0) * tiles_x + tx);
        if (!FILL) {
            atomicAdd(&fr.tile_count[tile], 1u);
        } else {
            const uint32_t gid = s_g[lo];
            const uint32_t slot = fr.ranges[tile].x + atomicAdd(&fr.tile_cursor[tile], 1u);
            if (slot < cap) {
                const unsigned long long key = ((unsigned long long)__ldg(&fr.dkey[gid]) << 32) |
                                               ((unsigned long long)gid << 4) |
                                               quad_mask(px0, px1, py0, py1, tx, ty);
                fr.inst_key[slot] = key;
            }
        }
    }
}

// Tile counts of the large footprints (K1a counted the small ones): a block per Gaussian, its
// threads striding over the tile rectangle.
__global__ void __launch_bounds__(B_THREADS)
k_count_big(SageCamDev cam, SageFrame fr)
{
    const uint32_t nbig = fr.counters[CNT_BIG];
    for (uint32_t q = blockIdx.x; q < nbig; q += gridDim.x) {
        const uint32_t g = fr.big_list[q];
        const uint2 b = fr.rect[g];
        const int bx0 = (b.x & 0xFFFF) / SAGE_TILE, by0 = (b.y & 0xFFFF) / SAGE_TILE;
        const uint32_t bw = (uint32_t)((b.x >> 16) / SAGE_TILE - bx0 + 1);
        const uint32_t bc = bw * (uint32_t)((b.y >> 16) / SAGE_TILE - by0 + 1);
        for (uint32_t k = threadIdx.x; k < bc; k += B_THREADS)
            atomicAdd(&fr.tile_count[(by0 + (int)(k / bw) - cam.row0) * cam.tiles_x + bx0 + (int)(k % bw)], 1u);
    }
}

// single block: exclusive scan of the tile counters -> ranges[t] = (begin, end), D -> counters.
// Each thread scans SCAN_IPT consecutive tiles in registers, so 8192 tiles take one block pass.
constexpr int SCAN_IPT = 8;
__global__ void __launch_bounds__(1024)
k_scan_tiles(SageFrame fr, int n_tiles, uint32_t cap)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int start = 0; start < n_tiles; start += 1024 * SCAN_IPT) {
        const int idx0 = start + threadIdx.x * SCAN_IPT;
        uint32_t c[SCAN_IPT];
        unsigned long long v = 0ull;
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k) {
            c[k] = (idx0 + k < n_tiles) ? fr.tile_count[idx0 + k] : 0u;
            v += c[k];
        }
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
        unsigned long long run = carry_s + wsum[w] + (incl - v);
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k) {
            if (idx0 + k < n_tiles) {
                unsigned long long b = run > cap ? cap : run, e = run + c[k] > cap ? cap : run + c[k];
                fr.ranges[idx0 + k] = make_uint2((uint32_t)b, (uint32_t)e);
            }
            run += c[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long total = carry_s;
        if (total > (unsigned long long)cap) {
            fr.counters[CNT_OVERFLOW] = 1u;
            total = cap;
        }
        fr.counters[CNT_INSTANCES] = (uint32_t)total;
    }
}

// ---------------------------------------------------------------------------------------------
// Per-tile sort.  LSD radix on the 27 depth bits (4 stable passes of 7 bits, 128 bins = one bin
// per thread), then equal-depth runs — rare: exact binary32 ties — are put in id order by
// odd-even transposition.  Lists up to TS_CAP keys are sorted in shared memory; longer ones
// ping-pong between the two global key buffers (they stay in L2).
// ---------------------------------------------------------------------------------------------
// Lists of up to WS_CAP keys (the common case) are sorted by ONE WARP each, warp-synchronously: no
// block barriers, keys ping-pong in the warp's slice of shared memory.  The key range of the list
// (max - min depth) decides how many 8-bit passes are needed (usually 3).
constexpr int WS_CAP = 512;
constexpr int WS_WARPS = 4;
constexpr int WS_BINS = 256;

__global__ void __launch_bounds__(WS_WARPS * 32)
k_tile_sort_warp(SageFrame fr, int n_tiles)
{
    __shared__ unsigned long long skeys[WS_WARPS][2][WS_CAP];
    __shared__ uint32_t shist[WS_WARPS][WS_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * WS_WARPS + w;
    if (tile >= n_tiles) return;
    const uint2 rg = fr.ranges[tile];
    const uint32_t n = rg.y - rg.x;
    if (n == 0 || n > WS_CAP) return;
    const unsigned long long* src = fr.inst_key + rg.x;
    uint32_t* out = fr.inst_val + rg.x;
    unsigned long long* buf0 = skeys[w][0];
    unsigned long long* buf1 = skeys[w][1];
    uint32_t* hist = shist[w];
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t dmin = 0xFFFFFFFFu, dmax = 0u;
    for (uint32_t e = lane; e < n; e += 32) {
        const unsigned long long k = src[e];
        buf0[e] = k;
        const uint32_t d = (uint32_t)(k >> 32);
        dmin = min(dmin, d);
        dmax = max(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    __syncwarp();
    const uint32_t span = dmax - dmin;
    const uint32_t nchunk = (n + 31) / 32;
    int cur = 0;
    for (int shift = 0; shift < 32 && (span >> shift) != 0u; shift += 8) {
        unsigned long long* a = cur ? buf1 : buf0;
        unsigned long long* b = cur ? buf0 : buf1;
#pragma unroll
        for (int k = 0; k < WS_BINS / 32; ++k) hist[k * 32 + lane] = 0;
        __syncwarp();
        for (uint32_t c = 0; c < nchunk; ++c) {
            const uint32_t e = c * 32 + lane;
            const bool valid = e < n;
            const uint32_t d = valid ? ((((uint32_t)(a[e] >> 32) - dmin) >> shift) & 255u) : 0xFFFFFFFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == __ffs(peers) - 1) hist[d] += __popc(peers);
            __syncwarp();
        }
        // exclusive scan of 256 bins: 8 consecutive bins per lane
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { loc[k] = hist[lane * 8 + k]; sum += loc[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { hist[lane * 8 + k] = run; run += loc[k]; }
        __syncwarp();
        for (uint32_t c = 0; c < nchunk; ++c) {
            const uint32_t e = c * 32 + lane;
            const bool valid = e < n;
            const unsigned long long key = valid ? a[e] : 0ull;
            const uint32_t d = valid ? ((((uint32_t)(key >> 32) - dmin) >> shift) & 255u) : 0xFFFFFFFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t prev = 0;
            if (valid && lane == leader) { prev = hist[d]; hist[d] = prev + __popc(peers); }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            if (valid) b[prev + __popc(peers & lt_mask)] = key;
            __syncwarp();
        }
        cur ^= 1;
    }
    unsigned long long* a = cur ? buf1 : buf0;
    // equal-depth runs -> ascending id (exact binary32 depth ties are rare)
    {
        int bad = 0;
        for (uint32_t e = lane; e + 1 < n; e += 32) {
            const unsigned long long x = a[e], y = a[e + 1];
            bad |= ((x >> 32) == (y >> 32)) && (x > y);
        }
        bad = __any_sync(0xffffffffu, bad);
        while (bad) {
            int swapped = 0;
            for (int phase = 0; phase < 2; ++phase) {
                for (uint32_t e = 2 * lane + phase; e + 1 < n; e += 64) {
                    const unsigned long long x = a[e], y = a[e + 1];
                    if ((x >> 32) == (y >> 32) && x > y) { a[e] = y; a[e + 1] = x; swapped = 1; }
                }
                __syncwarp();
            }
            bad = __any_sync(0xffffffffu, swapped);
        }
    }
    for (uint32_t e = lane; e < n; e += 32) out[e] = (uint32_t)a[e];   // id << 4 | quadrants
}

constexpr int TS_THREADS = 128;
constexpr int TS_WARPS = TS_THREADS / 32;
constexpr int TS_CAP = 2048;
constexpr int TS_BINS = 128;
constexpr int TS_DIGIT_BITS = 7;

__device__ __forceinline__ uint32_t ts_digit(unsigned long long k, int shift)
{
    return (uint32_t)(k >> shift) & (TS_BINS - 1);
}

// one stable pass src -> dst; returns false (and leaves dst untouched) when every key has the
// same digit.  whist: [TS_WARPS][TS_BINS] shared counters.
__device__ bool ts_pass(const unsigned long long* src, unsigned long long* dst, uint32_t n, int shift,
                        uint32_t (*whist)[TS_BINS])
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < TS_WARPS * TS_BINS; i += TS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the contiguous segment [seg0, seg1) so that (warp, position) order == list order
    const uint32_t per = (n + TS_WARPS - 1) / TS_WARPS;
    const uint32_t seg0 = min(n, w * per), seg1 = min(n, seg0 + per);
    for (uint32_t e = seg0 + lane; e < seg1; e += 32) atomicAdd(&whist[w][ts_digit(src[e], shift)], 1u);
    __syncthreads();
    // thread d: totals over warps, then exclusive scan over the 128 digits
    uint32_t cnt[TS_WARPS], tot = 0;
#pragma unroll
    for (int ww = 0; ww < TS_WARPS; ++ww) { cnt[ww] = whist[ww][tid]; tot += cnt[ww]; }
    const int uniform = __syncthreads_or(tot == n);
    if (uniform) return false;
    __shared__ uint32_t wtot[TS_WARPS];
    uint32_t incl = warp_incl_scan_u32(tot);
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    uint32_t base = incl - tot;
#pragma unroll
    for (int ww = 0; ww < TS_WARPS; ++ww) base += (ww < w) ? wtot[ww] : 0u;
#pragma unroll
    for (int ww = 0; ww < TS_WARPS; ++ww) { whist[ww][tid] = base; base += cnt[ww]; }
    __syncthreads();
    // ranked scatter of each warp's segment, 32 keys at a time
    const unsigned lt_mask = (1u << lane) - 1u;
    for (uint32_t e0 = seg0; e0 < seg1; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool valid = e < seg1;
        const unsigned long long k = valid ? src[e] : 0ull;
        const uint32_t d = valid ? ts_digit(k, shift) : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = whist[w][d];
            whist[w][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        if (valid) dst[prev + __popc(peers & lt_mask)] = k;
        __syncwarp();
    }
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(TS_THREADS)
k_tile_sort(SageFrame fr, int depth_bits)
{
    __shared__ unsigned long long sbuf[2][TS_CAP];
    __shared__ uint32_t whist[TS_WARPS][TS_BINS];
    const uint2 rg = fr.ranges[blockIdx.x];
    const uint32_t n = rg.y - rg.x;
    if (n <= WS_CAP) return;                 // short lists belong to k_tile_sort_warp
    const int tid = threadIdx.x;
    unsigned long long* g0 = fr.inst_key + rg.x;
    unsigned long long* g1 = fr.inst_key_alt + rg.x;
    uint32_t* out = fr.inst_val + rg.x;
    unsigned long long* buf[2];
    const bool in_smem = n <= TS_CAP;
    if (in_smem) {
        buf[0] = sbuf[0]; buf[1] = sbuf[1];
        for (uint32_t e = tid; e < n; e += TS_THREADS) sbuf[0][e] = g0[e];
        __syncthreads();
    } else {
        buf[0] = g0; buf[1] = g1;
        __threadfence_block();
    }
    int cur = 0;
    for (int shift = 32; shift < 32 + depth_bits; shift += TS_DIGIT_BITS)
        if (ts_pass(buf[cur], buf[cur ^ 1], n, shift, whist)) cur ^= 1;
    unsigned long long* a = buf[cur];
    // equal-depth runs -> ascending id (odd-even transposition restricted to equal depths)
    {
        int unordered = 0;
        for (uint32_t e = tid; e + 1 < n; e += TS_THREADS) {
            const unsigned long long x = a[e], y = a[e + 1];
            unordered |= ((x >> 32) == (y >> 32)) && (x > y);
        }
        unordered = __syncthreads_or(unordered);
        while (unordered) {
            int swapped = 0;
            for (int phase = 0; phase < 2; ++phase) {
                for (uint32_t e = 2 * tid + phase; e + 1 < n; e += 2 * TS_THREADS) {
                    const unsigned long long x = a[e], y = a[e + 1];
                    if ((x >> 32) == (y >> 32) && x > y) { a[e] = y; a[e + 1] = x; swapped = 1; }
                }
                __syncthreads();
            }
            unordered = __syncthreads_or(swapped);
        }
    }
    for (uint32_t e = tid; e < n; e += TS_THREADS) out[e] = (uint32_t)a[e];   // id << 4 | quadrants
}

}  // namespace

void sage_launch_binning(const SageCamDev& cam, const SageFrame& fr, int64_t n, int64_t cap,
                         int n_tiles, int* launches, cudaStream_t st)
{
    if (n <= 0 || n_tiles <= 0) return;
    const int blocks = (int)((n + B_THREADS - 1) / B_THREADS);
    k_count_big<<<148 * 4, B_THREADS, 0, st>>>(cam, fr);
    k_scan_tiles<<<1, 1024, 0, st>>>(fr, n_tiles, (uint32_t)cap);
    k_expand<true><<<blocks, B_THREADS, 0, st>>>(cam, fr, (uint32_t)cap);
    if (launches) *launches += 3;
}

void sage_launch_tile_sort(const SageFrame& fr, int n_tiles, int depth_bits, cudaStream_t st)
{
    if (n_tiles <= 0) return;
    k_tile_sort_warp<<<(n_tiles + WS_WARPS - 1) / WS_WARPS, WS_WARPS * 32, 0, st>>>(fr, n_tiles);
    k_tile_sort<<<n_tiles, TS_THREADS, 0, st>>>(fr, depth_bits);
}
