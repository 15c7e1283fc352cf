// K3: stable LSD radix sort of (key32, val32) pairs — ONE persistent kernel per pass ("onesweep"
// form): a block ranks a tile of its items locally, publishes its per-digit counts, and obtains the
// counts of all earlier tiles by decoupled look-back instead of a separate histogram + scan launch.
//
// Used twice per frame (no host round trip: element counts are read from device memory):
//  (a) depth sort of the Gaussians: key = bits(z) - bits(near) (<= 27 significant bits, 3 passes of
//      9 bits); pass 0 also compacts (drops SAGE_KEY_CULLED).  The sort is stable and its input
//      is in index order, so equal depths stay in Gaussian-index order — the spec's tie-break.
//  (b) stable binning of the depth-ordered (bin, Gaussian) instances by bin id (15 bits at
//      1080p, 2 passes of 8): every bin's list comes out depth-sorted, no 64-bit keys anywhere.
//
// Launch shape: the grid is a fixed 2 blocks per SM (296 on a B200), whatever the element count —
// blocks loop, taking tile indices from an atomic ticket until the tiles run out.  (Round 1 sized
// the grid by the buffer CAPACITY: 2048 blocks per pass of which ~290 had work; the other ~1750
// each paid a CTA launch with 80 KB of shared memory, and that — not the look-back — was most of a
// pass.)  The ticket also gives forward progress: every predecessor a tile waits for has already
// been taken by a resident block.  A fixed grid is what lets the whole frame be one CUDA graph.
//
// Look-back state is one 64-bit word per (tile, digit): epoch(30) | state(2) | value(32).  The
// epoch changes every pass of every frame (read from the frame's device-resident parameter block),
// so the state array is never cleared.  The spin is bounded and traps instead of hanging.
#include "sage_internal.cuh"

namespace {

constexpr int S_THREADS = 512;
constexpr int S_IPT_MIN = 8;                       // smallest items/thread used (sizes the look-back array)
constexpr int S_WARPS = S_THREADS / 32;
#ifdef SAGE_CUDA_EMU
constexpr int S_GRID = 2;
#else
constexpr int S_GRID = 148 * 2;                    // persistent: two blocks per SM
#endif

constexpr unsigned long long ST_AGG = 1ull, ST_PREFIX = 2ull;
constexpr int LB_BATCH = 8;

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, unsigned long long state, uint32_t v)
{
    return ((unsigned long long)epoch << 34) | (state << 32) | (unsigned long long)v;
}

#ifdef SAGE_CUDA_EMU
inline unsigned long long lb_load(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void lb_store(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// ---- global digit histograms for all passes of a sort, in one read of the keys ----------------
template <int RB, int NP, bool COMPACT>
__global__ void __launch_bounds__(S_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, int lo_bit,
             uint32_t* __restrict__ hist /*[NP][1<<RB]*/)
{
    constexpr int RADIX = 1 << RB;
    __shared__ uint32_t h[NP][RADIX];
    for (int i = threadIdx.x; i < NP * RADIX; i += S_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = *n_ptr;
    for (uint32_t idx = blockIdx.x * S_THREADS + threadIdx.x; idx < n; idx += gridDim.x * S_THREADS) {
        const uint32_t k = __ldg(&keys[idx]);
        if (COMPACT && k == SAGE_KEY_CULLED) continue;
#pragma unroll
        for (int p = 0; p < NP; ++p) atomicAdd(&h[p][(k >> (lo_bit + RB * p)) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP * RADIX; i += S_THREADS) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---- exclusive scan of each pass's histogram (one block per pass) -------------------------------
template <int RB>
__global__ void __launch_bounds__(S_THREADS)
k_radix_offsets(uint32_t* __restrict__ hist)
{
    constexpr int RADIX = 1 << RB;
    static_assert(RADIX <= S_THREADS, "one digit per thread");
    uint32_t* h = hist + (size_t)blockIdx.x * RADIX;
    __shared__ uint32_t wsum[S_WARPS];
    const uint32_t v = threadIdx.x < RADIX ? h[threadIdx.x] : 0u;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t base = incl - v;
    for (int k = 0; k < w; ++k) base += wsum[k];
    if (threadIdx.x < RADIX) h[threadIdx.x] = base;
}

// ---- one sort pass ------------------------------------------------------------------------------
// Register cap 64 (two blocks per SM): a resident sort block then leaves half the register file to
// the composite CTAs of the frame in flight on another stream.
template <int RB, bool COMPACT, int S_IPT>
__global__ void __launch_bounds__(S_THREADS, 2)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
           const uint32_t* __restrict__ n_ptr, int shift, const uint32_t* __restrict__ digit_base,
           unsigned long long* __restrict__ lookback, uint32_t* __restrict__ ticket,
           const SageFrameParams* __restrict__ fp, uint32_t epoch_off, int vals_are_iota)
{
    constexpr int RADIX = 1 << RB;
    constexpr int S_TILE = S_THREADS * S_IPT;
    static_assert(RADIX <= S_THREADS, "one digit per thread");
    const uint32_t n = *n_ptr;
    const uint32_t nb = (n + S_TILE - 1) / S_TILE;
    const uint32_t epoch = fp->epoch + epoch_off;
    // dynamic shared memory: warp-private digit counters, then the staging area of the scatter
    SAGE_DYN_SMEM(uint32_t, smem_dyn);
    uint32_t (*wh)[RADIX] = reinterpret_cast<uint32_t (*)[RADIX]>(smem_dyn);      // [S_WARPS][RADIX]
    uint32_t* gdelta = smem_dyn + S_WARPS * RADIX;                                // [RADIX]
    uint32_t* skey = gdelta + RADIX;                                              // [S_TILE]
    uint32_t* sval = skey + S_TILE;                                               // [S_TILE]
    __shared__ uint32_t s_tile, s_wtot[S_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t dbase = threadIdx.x < RADIX ? digit_base[threadIdx.x] : 0u;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (;;) {
        __syncthreads();                          // the previous tile's staging area has been drained
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        for (int i = threadIdx.x; i < S_WARPS * RADIX; i += S_THREADS) (&wh[0][0])[i] = 0;
        __syncthreads();
        const uint32_t b = s_tile;
        if (b >= nb) break;

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
            const unsigned peers = __match_any_sync(0xffffffffu, d);
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

        // per digit: block count -> publish -> look back -> publish inclusive -> per-warp bases
        uint32_t cnt[S_WARPS], tot = 0, excl = 0;
        if (threadIdx.x < RADIX) {
            const int d = threadIdx.x;
#pragma unroll
            for (int ww = 0; ww < S_WARPS; ++ww) { cnt[ww] = wh[ww][d]; tot += cnt[ww]; }
            unsigned long long* mine = lookback + (size_t)b * RADIX + d;
            if (b == 0) {
                lb_store(mine, lb_pack(epoch, ST_PREFIX, tot));
            } else {
                lb_store(mine, lb_pack(epoch, ST_AGG, tot));
                // Look back over the earlier tiles, LB_BATCH of them per round trip: the loads of a
                // batch are independent (memory-level parallelism), then consumed in order up to the
                // first tile that already holds an inclusive prefix or has not published yet.
                int t = (int)b - 1;
                uint32_t spins = 0;
                bool done = false;
                while (!done) {
                    unsigned long long s[LB_BATCH];
#pragma unroll
                    for (int j = 0; j < LB_BATCH; ++j)
                        s[j] = (t - j >= 0) ? lb_load(lookback + (size_t)(t - j) * RADIX + d) : 0ull;
                    int used = 0;
#pragma unroll
                    for (int j = 0; j < LB_BATCH; ++j) {
                        if (!done && used == j && t - j >= 0 && (uint32_t)(s[j] >> 34) == epoch) {
                            excl += (uint32_t)s[j];
                            used = j + 1;
                            done = ((s[j] >> 32) & 3ull) == ST_PREFIX;
                        }
                    }
                    t -= used;
                    if (!done && used == 0) {                      // predecessor has not published yet
                        if (++spins > (1u << 24)) __trap();        // never hang the GPU
                        __nanosleep(32);
                    }
                }
                lb_store(mine, lb_pack(epoch, ST_PREFIX, excl + tot));
            }
        }
        // block-local exclusive scan of the digit totals -> where each digit's run starts in the staging
        // area; gdelta[d] = (global start of this block's run of digit d) - (its local start)
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
                gdelta[d] = dbase + excl - lstart;
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
            keys_out[pos] = k;
            vals_out[pos] = sval[j];
        }
    }
}

template <int RB, int NP, bool COMPACT, int S_IPT>
void sort_impl(const SageSortPlan& pl, cudaStream_t st)
{
    constexpr int RADIX = 1 << RB;
    constexpr int S_TILE = S_THREADS * S_IPT;
    constexpr size_t SMEM = sizeof(uint32_t) * (S_WARPS * RADIX + RADIX + 2 * S_TILE);
    auto k_hist = k_radix_hist<RB, NP, COMPACT>;
    auto k_offs = k_radix_offsets<RB>;
    auto k_first = k_onesweep<RB, COMPACT, S_IPT>;
    auto k_rest = k_onesweep<RB, false, S_IPT>;
    if (!pl.hist_ready) SAGE_KLAUNCH(k_hist, S_GRID, S_THREADS, st, pl.keys_src, pl.n_first, pl.lo_bit, pl.hist);
    SAGE_KLAUNCH(k_offs, NP, S_THREADS, st, pl.hist);
    int cur = 0;
    for (int p = 0; p < NP; ++p) {
        const uint32_t* n_dev = p == 0 ? pl.n_first : pl.n_rest;
        const uint32_t* kin = p == 0 ? pl.keys_src : pl.keys[cur];
        const int iota = (pl.vals_are_iota && p == 0) ? 1 : 0;
        if (p == 0)
            SAGE_KLAUNCH_SMEM(k_first, S_GRID, S_THREADS, SMEM, st, kin, pl.vals[cur], pl.keys[cur ^ 1], pl.vals[cur ^ 1], n_dev,
                              pl.lo_bit + RB * p, pl.hist + (size_t)p * RADIX, pl.lookback, pl.tickets + p, pl.fp,
                              pl.epoch_off + (uint32_t)p, iota);
        else
            SAGE_KLAUNCH_SMEM(k_rest, S_GRID, S_THREADS, SMEM, st, kin, pl.vals[cur], pl.keys[cur ^ 1], pl.vals[cur ^ 1], n_dev,
                              pl.lo_bit + RB * p, pl.hist + (size_t)p * RADIX, pl.lookback, pl.tickets + p, pl.fp,
                              pl.epoch_off + (uint32_t)p, iota);
        cur ^= 1;
    }
}

template <int RB, int S_IPT>
void sort_cfg()
{
#ifndef SAGE_CUDA_EMU
    constexpr size_t SMEM = sizeof(uint32_t) * (S_WARPS * (1 << RB) + (1 << RB) + 2 * S_THREADS * S_IPT);
    cudaFuncSetAttribute(k_onesweep<RB, true, S_IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    cudaFuncSetAttribute(k_onesweep<RB, false, S_IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
#endif
}

}  // namespace

// dynamic shared-memory opt-in of every sort-pass instantiation (per device; outside graph capture)
void sage_sort_configure(void)
{
    sort_cfg<8, 8>(); sort_cfg<8, 16>(); sort_cfg<9, 8>(); sort_cfg<9, 16>();
}

void sage_sort_shape(int n_bits, int* digit_bits, int* passes)
{
    // 9-bit digits when that saves a pass (27-bit depth keys: 3 passes), else 8
    const int rb = ((n_bits + 8) / 9 < (n_bits + 7) / 8) ? 9 : 8;
    *digit_bits = rb;
    *passes = (n_bits + rb - 1) / rb;
}

size_t sage_sort_lookback_words(int64_t capacity)
{
    const int tile_min = S_THREADS * S_IPT_MIN;
    size_t nb = (size_t)((capacity + tile_min - 1) / tile_min) + 1;
    return nb * SAGE_RADIX_MAX_BINS;
}

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
        if (launches) *launches += NP_ + 1 + (pl.hist_ready ? 0 : 1);                \
        return np & 1;                                                               \
    }
    SAGE_SORT_CASE(8, 1) SAGE_SORT_CASE(8, 2) SAGE_SORT_CASE(8, 3) SAGE_SORT_CASE(8, 4)
    SAGE_SORT_CASE(9, 1) SAGE_SORT_CASE(9, 2) SAGE_SORT_CASE(9, 3)
#undef SAGE_SORT_CASE
    return -1;
}
