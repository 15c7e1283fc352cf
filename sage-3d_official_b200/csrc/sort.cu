// K3: stable LSD radix sort of (key32, val32) pairs, 8 bits per pass, element count read from
// device memory (no host round trip inside a frame).
//
// Used twice per frame: (a) depth sort of the Gaussians (key = binary32 bits of the view depth,
// ties broken by Gaussian index because the sort is stable and the input is in index order),
// (b) stable binning of the depth-ordered (tile, Gaussian) instances by tile id — which leaves
// every tile's list depth-sorted without ever forming 64-bit keys.
//
// Each pass = histogram -> per-digit scan -> ranked scatter.  Blocks own 4096 consecutive items;
// blocks past the live element count exit immediately, so grids are sized by capacity.
#include "sage_internal.cuh"

namespace {

constexpr int S_THREADS = 256;
constexpr int S_IPT = 16;
constexpr int S_TILE = S_THREADS * S_IPT;
constexpr int S_WARPS = S_THREADS / 32;

template <int RB, bool COMPACT>
__global__ void __launch_bounds__(S_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, int shift,
             uint32_t* __restrict__ block_hist, uint32_t* __restrict__ digit_total)
{
    constexpr int RADIX = 1 << RB;
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    const uint32_t b = blockIdx.x;
    if (b >= nb) return;
    __shared__ uint32_t h[RADIX];
    for (int d = threadIdx.x; d < RADIX; d += S_THREADS) h[d] = 0;
    __syncthreads();
    const uint32_t base = b * S_TILE;
#pragma unroll
    for (int i = 0; i < S_IPT; ++i) {
        uint32_t idx = base + i * S_THREADS + threadIdx.x;
        if (idx < n) {
            const uint32_t k = __ldg(&keys[idx]);
            if (!COMPACT || k != SAGE_KEY_CULLED) atomicAdd(&h[(k >> shift) & (RADIX - 1)], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += S_THREADS) {
        uint32_t c = h[d];
        block_hist[(size_t)d * nb + b] = c;
        if (c) atomicAdd(&digit_total[d], c);
    }
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan for 256 threads; returns exclusive prefix, writes block total.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* smem /*[9]*/, uint32_t& total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) smem[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < S_WARPS ? smem[lane] : 0;
        uint32_t ti = warp_incl_scan(t);
        if (lane < S_WARPS) smem[lane] = ti - t;
        if (lane == S_WARPS - 1) smem[S_WARPS] = ti;
    }
    __syncthreads();
    uint32_t res = incl - v + smem[w];
    total = smem[S_WARPS];
    __syncthreads();
    return res;
}

// One block per digit: turns block_hist[d][*] into global output offsets.
template <int RB>
__global__ void __launch_bounds__(S_THREADS)
k_radix_scan(const uint32_t* __restrict__ n_ptr, uint32_t* __restrict__ block_hist,
             const uint32_t* __restrict__ digit_total)
{
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    if (nb == 0) return;
    const uint32_t d = blockIdx.x;
    __shared__ uint32_t sm[S_WARPS + 1];
    uint32_t mine = 0;
    for (uint32_t k = threadIdx.x; k < d; k += S_THREADS) mine += digit_total[k];
    uint32_t base_total;
    (void)block_excl_scan_256(mine, sm, base_total);
    uint32_t carry = base_total;
    uint32_t* row = block_hist + (size_t)d * nb;
    for (uint32_t start = 0; start < nb; start += S_THREADS) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < nb ? row[idx] : 0;
        uint32_t tot;
        uint32_t ex = block_excl_scan_256(v, sm, tot);
        if (idx < nb) row[idx] = carry + ex;
        carry += tot;
    }
}

template <int RB, bool COMPACT>
__global__ void __launch_bounds__(S_THREADS)
k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                const uint32_t* __restrict__ n_ptr, int shift,
                const uint32_t* __restrict__ block_hist, int vals_are_iota)
{
    constexpr int RADIX = 1 << RB;
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    const uint32_t b = blockIdx.x;
    if (b >= nb) return;
    __shared__ uint32_t wh[S_WARPS][RADIX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < S_WARPS * RADIX; i += S_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();

    // warp-striped ownership: warp w holds items [w*512, (w+1)*512) of the block, item (i, lane)
    // at offset i*32 + lane, so (warp, i, lane) order == input order (stability).
    const uint32_t wbase = b * S_TILE + w * (32 * S_IPT);
    uint32_t key[S_IPT], val[S_IPT], rank[S_IPT];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < S_IPT; ++i) {
        uint32_t idx = wbase + i * 32 + lane;
        bool valid = idx < n;
        key[i] = valid ? __ldg(&keys_in[idx]) : 0xFFFFFFFFu;
        if (COMPACT) valid = valid && key[i] != SAGE_KEY_CULLED;   // pass 0 drops culled entries
        val[i] = valid ? (vals_are_iota ? idx : __ldg(&vals_in[idx])) : 0u;
        uint32_t d = valid ? ((key[i] >> shift) & (RADIX - 1)) : 0xFFFFFFFFu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = wh[w][d];
            wh[w][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[i] = prev + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += S_THREADS) {
        // global base of (block, digit), then per-warp exclusive offsets
        uint32_t run = block_hist[(size_t)d * nb + b];
#pragma unroll
        for (int ww = 0; ww < S_WARPS; ++ww) {
            uint32_t c = wh[ww][d];
            wh[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < S_IPT; ++i) {
        uint32_t idx = wbase + i * 32 + lane;
        if (idx < n && (!COMPACT || key[i] != SAGE_KEY_CULLED)) {
            uint32_t d = (key[i] >> shift) & (RADIX - 1);
            uint32_t pos = wh[w][d] + rank[i];
            keys_out[pos] = key[i];
            vals_out[pos] = val[i];
        }
    }
}

template <int RB, bool COMPACT>
void radix_pass(uint32_t* keys[2], uint32_t* vals[2], int cur, const uint32_t* n_dev, int grid, int shift,
                uint32_t* block_hist, uint32_t* tot, int iota, cudaStream_t st)
{
    k_radix_hist<RB, COMPACT><<<grid, S_THREADS, 0, st>>>(keys[cur], n_dev, shift, block_hist, tot);
    k_radix_scan<RB><<<1 << RB, S_THREADS, 0, st>>>(n_dev, block_hist, tot);
    k_radix_scatter<RB, COMPACT><<<grid, S_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                              n_dev, shift, block_hist, iota);
}

// digit width: 9 bits when that saves a pass (e.g. 27-bit depth keys: 3 passes), else 8
int digit_bits(int n_bits) { return ((n_bits + 8) / 9 < (n_bits + 7) / 8) ? 9 : 8; }

}  // namespace

int sage_radix_passes(int n_bits)
{
    const int rb = digit_bits(n_bits);
    return (n_bits + rb - 1) / rb;
}

size_t sage_radix_block_hist_elems(int64_t capacity)
{
    size_t nb = (size_t)((capacity + S_TILE - 1) / S_TILE);
    if (nb == 0) nb = 1;
    return nb * SAGE_RADIX_MAX_BINS;
}

int sage_radix_sort(uint32_t* keys[2], uint32_t* vals[2], const uint32_t* n_first_dev,
                    const uint32_t* n_rest_dev, int64_t capacity, int lo_bit, int n_bits,
                    bool vals_are_iota, bool compact_first, uint32_t* block_hist,
                    uint32_t* digit_total, int* launches, cudaStream_t st)
{
    const int rb = digit_bits(n_bits);
    const int passes = (n_bits + rb - 1) / rb;
    int grid = (int)((capacity + S_TILE - 1) / S_TILE);
    if (grid < 1) grid = 1;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = lo_bit + rb * p;
        uint32_t* tot = digit_total + (size_t)p * SAGE_RADIX_MAX_BINS;
        const int iota = (vals_are_iota && p == 0) ? 1 : 0;
        const uint32_t* n_dev = p == 0 ? n_first_dev : n_rest_dev;
        const bool compact = compact_first && p == 0;
        if (rb == 9) {
            if (compact) radix_pass<9, true>(keys, vals, cur, n_dev, grid, shift, block_hist, tot, iota, st);
            else radix_pass<9, false>(keys, vals, cur, n_dev, grid, shift, block_hist, tot, iota, st);
        } else {
            if (compact) radix_pass<8, true>(keys, vals, cur, n_dev, grid, shift, block_hist, tot, iota, st);
            else radix_pass<8, false>(keys, vals, cur, n_dev, grid, shift, block_hist, tot, iota, st);
        }
        if (launches) *launches += 3;
        cur ^= 1;
    }
    return cur;
}
