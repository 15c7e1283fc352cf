// K3: stable LSD radix sort of (key32, val32) pairs.
//
// Used twice per frame (no host round trip: element counts are read from device memory):
//  (a) depth sort of the Gaussians: key = bits(z) - bits(near) (<= 27 significant bits, 3 passes of
//      9 bits); pass 0 also compacts (drops SAGE_KEY_CULLED).  The sort is stable and its input
//      is in index order, so equal depths stay in Gaussian-index order — the spec's tie-break.
//  (b) stable binning of the depth-ordered (bin, Gaussian) instances by bin id (15 bits at
//      1080p, 2 passes of 8): every bin's list comes out depth-sorted, no 64-bit keys anywhere.
//      The last pass also writes every bin's [begin, end) range, so no separate ranges kernel runs.
//
// One pass = three small kernels, all with fixed grids (frame-invariant launch shapes, one CUDA graph):
//   k_sort_hist     per-tile digit histograms into a digit-major matrix H[digit][tile]
//   k_sort_scan     a warp per digit: exclusive scan of its row over the tiles (coalesced) + digit totals
//   k_sort_scatter  per tile: digit bases from the totals, local ranks (match.any + warp-private
//                   counters), staged in shared memory in digit order, coalesced scatter
// Round 1 (and the first cut of round 2) ran each pass as ONE kernel with decoupled look-back.  At
// these sizes (0.6 M keys, 2-5 M instances) a pass is a single wave of tiles: nobody ever finds an
// inclusive prefix, every tile sums every predecessor itself — O(tiles^2 x digits) loads, more than
// half of the kernel's instructions and most of its latency (ncu: 30 us for 0.7 M keys, 80-110 us
// for 5 M pairs).  With the counts known before the scatter starts there is no spinning, no epochs
// and no tickets; the price is two extra (tiny) launches per pass inside the graph.
#include "sage_internal.cuh"

namespace {

constexpr int S_THREADS = 512;
constexpr int S_WARPS = S_THREADS / 32;
#ifdef SAGE_CUDA_EMU
constexpr int S_GRID = 2;
#else
#ifndef SAGE_SORT_CTAS
#define SAGE_SORT_CTAS 2
#endif
constexpr int S_GRID = 148 * SAGE_SORT_CTAS;       // persistent: blocks per SM, grid-stride over the tiles
#endif

// lanes of the warp holding the same digit (invalid lanes: d = 0xFFFFFFFF).  The hardware MATCH
// instruction is one issue slot but has a long latency; the ballot form (one vote per digit bit) is
// ~3 RB simple instructions with none.  SAGE_SORT_MATCH=1 builds the MATCH form for comparison.
template <int RB>
__device__ __forceinline__ unsigned digit_peers(uint32_t d)
{
#if defined(SAGE_SORT_MATCH)
    return __match_any_sync(0xffffffffu, d);
#else
    unsigned m = __ballot_sync(0xffffffffu, d == 0xFFFFFFFFu);     // the invalid lanes
    m = (d == 0xFFFFFFFFu) ? m : ~m;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        const unsigned bit = (d >> b) & 1u;
        const unsigned v = __ballot_sync(0xffffffffu, bit);
        m &= bit ? v : ~v;
    }
    return m;
#endif
}

// ---- per-tile digit histograms --------------------------------------------------------------
template <int RB, bool COMPACT, int S_IPT>
__global__ void __launch_bounds__(S_THREADS)
k_sort_hist(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, int shift,
            uint32_t* __restrict__ H, int ld)
{
    constexpr int RADIX = 1 << RB;
    constexpr int S_TILE = S_THREADS * S_IPT;
    __shared__ uint32_t h[RADIX];
    SAGE_DEP_LAUNCH_TOP();
    sage_dep_wait();
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        for (int i = threadIdx.x; i < RADIX; i += S_THREADS) h[i] = 0;
        __syncthreads();
        const uint32_t base = b * S_TILE;
#pragma unroll
        for (int i = 0; i < S_IPT; ++i) {
            const uint32_t idx = base + i * S_THREADS + threadIdx.x;
            if (idx < n) {
                const uint32_t k = __ldg(&keys[idx]);
                if (!COMPACT || k != SAGE_KEY_CULLED) atomicAdd(&h[(k >> shift) & (RADIX - 1)], 1u);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < RADIX; i += S_THREADS) H[(size_t)i * ld + b] = h[i];
        __syncthreads();
    }
    SAGE_DEP_LAUNCH_END();
}

// ---- a warp per digit: exclusive scan over the tiles, digit totals ---------------------------
template <int S_IPT>
__global__ void __launch_bounds__(256)
k_sort_scan(const uint32_t* __restrict__ n_ptr, uint32_t* __restrict__ H, int ld, int radix,
            uint32_t* __restrict__ totals)
{
    constexpr int S_TILE = S_THREADS * S_IPT;
    SAGE_DEP_LAUNCH_TOP();
    sage_dep_wait();
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= radix) return;
    uint32_t* row = H + (size_t)d * ld;
    uint32_t run = 0;
    for (uint32_t t0 = 0; t0 < nb; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t v = t < nb ? row[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (t < nb) row[t] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) totals[d] = run;
    SAGE_DEP_LAUNCH_END();
}

// ---- rank + scatter -------------------------------------------------------------------------
// Register cap 64 (two blocks per SM): a resident sort block then leaves half the register file to
// the composite CTAs of the frame in flight on another stream.
// RANGES (last pass of the by-bin sort): every run of equal keys inside a tile's digit run reports
// its first / last output position with atomicMax into ranges[key] = (~begin, end) — runs of one
// key in different tiles combine to the bin's range, so no pass over the sorted keys is needed.
template <int RB, bool COMPACT, int S_IPT, bool RANGES>
__global__ void __launch_bounds__(S_THREADS, 2)
k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
               const uint32_t* __restrict__ n_ptr, int shift, const uint32_t* __restrict__ H, int ld,
               const uint32_t* __restrict__ totals, int vals_are_iota, uint2* __restrict__ ranges)
{
    constexpr int RADIX = 1 << RB;
    constexpr int S_TILE = S_THREADS * S_IPT;
    static_assert(RADIX <= S_THREADS, "one digit per thread");
    SAGE_DEP_LAUNCH_TOP();
    sage_dep_wait();
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    // dynamic shared memory: warp-private digit counters, then the staging area of the scatter
    SAGE_DYN_SMEM(uint32_t, smem_dyn);
    uint32_t (*wh)[RADIX] = reinterpret_cast<uint32_t (*)[RADIX]>(smem_dyn);      // [S_WARPS][RADIX]
    uint32_t* gdelta = smem_dyn + S_WARPS * RADIX;                                // [RADIX]
    uint32_t* skey = gdelta + RADIX;                                              // [S_TILE]
    uint32_t* sval = skey + S_TILE;                                               // [S_TILE]
    __shared__ uint32_t s_wtot[S_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (blockIdx.x >= nb) return;

    // where each digit's bucket starts in the output: exclusive scan of the digit totals
    uint32_t dbase = 0;
    {
        const uint32_t v = threadIdx.x < RADIX ? totals[threadIdx.x] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_wtot[w] = incl;
        __syncthreads();
        dbase = incl - v;
        for (int k = 0; k < w; ++k) dbase += s_wtot[k];
    }

    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        __syncthreads();                          // the previous tile's staging area has been drained
        for (int i = threadIdx.x; i < S_WARPS * RADIX; i += S_THREADS) (&wh[0][0])[i] = 0;
        // the tile's exclusive prefix of this digit over the earlier tiles (issued early)
        const uint32_t tpre = threadIdx.x < RADIX ? __ldg(&H[(size_t)threadIdx.x * ld + b]) : 0u;
        __syncthreads();

        // warp-striped ownership: warp w holds items [w*32*IPT, (w+1)*32*IPT) of the tile, item (i, lane)
        // at offset i*32 + lane, so (warp, i, lane) order == input order (stability).  All loads first.
        const uint32_t wbase = b * S_TILE + w * (32 * S_IPT);
        uint32_t key[S_IPT];
#pragma unroll
        for (int i = 0; i < S_IPT; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            key[i] = idx < n ? __ldg(&keys_in[idx]) : SAGE_KEY_CULLED;
        }
        // local ranks: warp-private digit counters, peers found with match.any
        uint32_t rank[S_IPT];
#pragma unroll
        for (int i = 0; i < S_IPT; ++i) {
            const uint32_t idx = wbase + i * 32 + lane;
            const bool valid = idx < n && (!COMPACT || key[i] != SAGE_KEY_CULLED);
            const uint32_t d = valid ? ((key[i] >> shift) & (RADIX - 1)) : 0xFFFFFFFFu;
            const unsigned peers = digit_peers<RB>(d);
            const int leader = __ffs(peers) - 1;
            uint32_t prev = 0;
            if (valid && lane == leader) {
                prev = wh[w][d];
                wh[w][d] = prev + __popc(peers);
            }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            rank[i] = valid ? prev + __popc(peers & lt_mask) : 0xFFFFFFFFu;
            __syncwarp();
        }
        __syncthreads();

        // per digit: block count -> local start in the staging area -> per-warp bases;
        // gdelta[d] = (global start of this tile's run of digit d) - (its local start)
        uint32_t cnt[S_WARPS], tot = 0;
        if (threadIdx.x < RADIX) {
            const int d = threadIdx.x;
#pragma unroll
            for (int ww = 0; ww < S_WARPS; ++ww) { cnt[ww] = wh[ww][d]; tot += cnt[ww]; }
        }
        {
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wtot[w] = incl;
            __syncthreads();
            uint32_t lstart = incl - tot;
            for (int k = 0; k < w; ++k) lstart += s_wtot[k];
            if (threadIdx.x < RADIX) {
                const int d = threadIdx.x;
                gdelta[d] = dbase + tpre - lstart;
                uint32_t run = lstart;
#pragma unroll
                for (int ww = 0; ww < S_WARPS; ++ww) { wh[ww][d] = run; run += cnt[ww]; }
            }
        }
        uint32_t n_valid = 0;
#pragma unroll
        for (int k = 0; k < S_WARPS; ++k) n_valid += s_wtot[k];       // s_wtot is final after the barrier above
        __syncthreads();
        // stage in digit order ...
#pragma unroll
        for (int i = 0; i < S_IPT; ++i) {
            if (rank[i] != 0xFFFFFFFFu) {
                const uint32_t d = (key[i] >> shift) & (RADIX - 1);
                const uint32_t lp = wh[w][d] + rank[i];
                const uint32_t idx = wbase + i * 32 + lane;
                skey[lp] = key[i];
                sval[lp] = vals_are_iota ? idx : __ldg(&vals_in[idx]);     // values are only read here
            }
        }
        __syncthreads();
        // ... and write out: consecutive threads -> consecutive addresses inside every digit run
        for (uint32_t j = threadIdx.x; j < n_valid; j += S_THREADS) {
            const uint32_t k = skey[j];
            const uint32_t pos = j + gdelta[(k >> shift) & (RADIX - 1)];
            if (!RANGES) keys_out[pos] = k;                      // after the last pass only the ranges are read
            vals_out[pos] = sval[j];
            if (RANGES) {
                if (j == 0 || skey[j - 1] != k) atomicMax(&ranges[k].x, ~pos);
                if (j == n_valid - 1 || skey[j + 1] != k) atomicMax(&ranges[k].y, pos + 1u);
            }
        }
    }
    SAGE_DEP_LAUNCH_END();
}

template <int RB, int S_IPT>
constexpr size_t scatter_smem() { return sizeof(uint32_t) * (S_WARPS * (1 << RB) + (1 << RB) + 2 * S_THREADS * S_IPT); }

template <int RB, int NP, bool COMPACT, int S_IPT>
void sort_impl(const SageSortPlan& pl, cudaStream_t st)
{
    constexpr int RADIX = 1 << RB;
    constexpr size_t SMEM = scatter_smem<RB, S_IPT>();
    auto k_hist_first = k_sort_hist<RB, COMPACT, S_IPT>;
    auto k_hist_rest = k_sort_hist<RB, false, S_IPT>;
    auto k_scan = k_sort_scan<S_IPT>;
    auto k_first = k_sort_scatter<RB, COMPACT, S_IPT, false>;
    auto k_rest = k_sort_scatter<RB, false, S_IPT, false>;
    auto k_last = k_sort_scatter<RB, false, S_IPT, true>;
    const int ld = pl.ld;
    int cur = 0;
    for (int p = 0; p < NP; ++p) {
        const uint32_t* n_dev = p == 0 ? pl.n_first : pl.n_rest;
        const uint32_t* kin = p == 0 ? pl.keys_src : pl.keys[cur];
        const int shift = pl.lo_bit + RB * p;
        const int iota = (pl.vals_are_iota && p == 0) ? 1 : 0;
        if (p == 0) SAGE_KLAUNCH_PDL(k_hist_first, S_GRID, S_THREADS, 0, st, kin, n_dev, shift, pl.H, ld);
        else SAGE_KLAUNCH_PDL(k_hist_rest, S_GRID, S_THREADS, 0, st, kin, n_dev, shift, pl.H, ld);
        SAGE_KLAUNCH_PDL(k_scan, (RADIX + 7) / 8, 256, 0, st, n_dev, pl.H, ld, RADIX, pl.totals);
        const bool last = pl.ranges != nullptr && p == NP - 1;
        if (last && !(COMPACT && p == 0))
            SAGE_KLAUNCH_PDL(k_last, S_GRID, S_THREADS, SMEM, st, kin, pl.vals[cur], pl.keys[cur ^ 1], pl.vals[cur ^ 1], n_dev,
                              shift, pl.H, ld, pl.totals, iota, pl.ranges);
        else if (p == 0)
            SAGE_KLAUNCH_PDL(k_first, S_GRID, S_THREADS, SMEM, st, kin, pl.vals[cur], pl.keys[cur ^ 1], pl.vals[cur ^ 1], n_dev,
                              shift, pl.H, ld, pl.totals, iota, (uint2*)nullptr);
        else
            SAGE_KLAUNCH_PDL(k_rest, S_GRID, S_THREADS, SMEM, st, kin, pl.vals[cur], pl.keys[cur ^ 1], pl.vals[cur ^ 1], n_dev,
                              shift, pl.H, ld, pl.totals, iota, (uint2*)nullptr);
        cur ^= 1;
    }
}

template <int RB, int S_IPT>
void sort_cfg()
{
#ifndef SAGE_CUDA_EMU
    constexpr size_t SMEM = scatter_smem<RB, S_IPT>();
    cudaFuncSetAttribute(k_sort_scatter<RB, true, S_IPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    cudaFuncSetAttribute(k_sort_scatter<RB, false, S_IPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    cudaFuncSetAttribute(k_sort_scatter<RB, false, S_IPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    sage_prefer_smem(k_sort_scatter<RB, true, S_IPT, false>);
    sage_prefer_smem(k_sort_scatter<RB, false, S_IPT, false>);
    sage_prefer_smem(k_sort_scatter<RB, false, S_IPT, true>);
    sage_prefer_smem(k_sort_hist<RB, true, S_IPT>);
    sage_prefer_smem(k_sort_hist<RB, false, S_IPT>);
    sage_prefer_smem(k_sort_scan<S_IPT>);
#endif
}

}  // namespace

// dynamic shared-memory opt-in of every scatter instantiation (per device; outside graph capture)
void sage_sort_configure(void)
{
    sort_cfg<5, 8>(); sort_cfg<5, 16>(); sort_cfg<6, 8>(); sort_cfg<6, 16>(); sort_cfg<7, 8>(); sort_cfg<7, 16>();
    sort_cfg<8, 8>(); sort_cfg<8, 16>(); sort_cfg<9, 8>(); sort_cfg<9, 16>();
}

void sage_sort_shape(int n_bits, int* digit_bits, int* passes)
{
    // 9-bit digits when that saves a pass (27-bit depth keys: 3 passes), else 8
    const int rb = ((n_bits + 8) / 9 < (n_bits + 7) / 8) ? 9 : 8;
    *digit_bits = rb;
    *passes = (n_bits + rb - 1) / rb;
}

// leading dimension (tiles) of the histogram matrix for `capacity` elements, and its size in words
int sage_sort_matrix_ld(int64_t capacity)
{
    const int tile_min = S_THREADS * 8;
    return (int)((capacity + tile_min - 1) / tile_min) + 32;
}

size_t sage_sort_matrix_words(int64_t capacity) { return (size_t)sage_sort_matrix_ld(capacity) * SAGE_RADIX_MAX_BINS; }

int sage_radix_sort(const SageSortPlan& pl, int* launches, cudaStream_t st)
{
    const int rb = pl.digit_bits, np = pl.passes;     // chosen by the caller (sage_sort_shape or forced)
    const bool c = pl.compact_first;
#define SAGE_SORT_CASE(RB_, NP_)                                                     \
    if (rb == RB_ && np == NP_) {                                                    \
        if (pl.items_per_thread == 16) {                                             \
            if (c) sort_impl<RB_, NP_, true, 16>(pl, st); else sort_impl<RB_, NP_, false, 16>(pl, st); \
        } else {                                                                     \
            if (c) sort_impl<RB_, NP_, true, 8>(pl, st); else sort_impl<RB_, NP_, false, 8>(pl, st);   \
        }                                                                            \
        if (launches) *launches += 3 * NP_;                                          \
        return np & 1;                                                               \
    }
    SAGE_SORT_CASE(5, 1) SAGE_SORT_CASE(5, 2) SAGE_SORT_CASE(6, 1) SAGE_SORT_CASE(6, 2) SAGE_SORT_CASE(7, 1) SAGE_SORT_CASE(7, 2)
    SAGE_SORT_CASE(8, 1) SAGE_SORT_CASE(8, 2) SAGE_SORT_CASE(8, 3) SAGE_SORT_CASE(8, 4)
    SAGE_SORT_CASE(9, 1) SAGE_SORT_CASE(9, 2) SAGE_SORT_CASE(9, 3)
#undef SAGE_SORT_CASE
    return -1;
}
