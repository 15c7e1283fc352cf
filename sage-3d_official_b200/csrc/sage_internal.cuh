// Internal declarations shared by the kernels and the C-ABI layer (sm_100a only).
//
// Numerics contract: every float operation on the render/collide path is an
// IEEE-754 binary32 operation in the order written.  The library is compiled with -fmad=false,
// so a multiply-add is fused ONLY where fmaf()/__ffma2_rn() is written.  That is what lets the
// parity tests demand bit-exact agreement with the CPU oracle.
#pragma once
#ifdef SAGE_CUDA_EMU            /* development-only lock-step emulator (tools/cuda_emu/), never shipped */
#include "../../tools/cuda_emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SAGE_KLAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define SAGE_KLAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define SAGE_DYN_SMEM(type, name) extern __shared__ type name[]
#endif
#include <stdlib.h>

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// The front half of a frame is a chain of ~18 small kernels on one stream; the gap between two of them
// (completion -> launch -> first CTA) is a few microseconds each.  A kernel launched with
// SAGE_KLAUNCH_PDL may start while its predecessor in the stream drains: it runs its prologue, then
// sage_dep_wait() (griddepcontrol.wait) blocks until the predecessor grid has completed and its
// writes are visible.  sage_dep_launch() (griddepcontrol.launch_dependents) is the predecessor's
// "my dependents may be scheduled now".  Both are no-ops in a kernel launched without the attribute.
// SAGE_NO_PDL=1 turns the attribute off (A/B).
#ifdef SAGE_CUDA_EMU
#define SAGE_KLAUNCH_PDL(kernel, grid, block, smem, stream, ...) SAGE_KLAUNCH(kernel, grid, block, stream, __VA_ARGS__)
inline void sage_dep_wait() {}
inline void sage_dep_launch() {}
#else
inline bool sage_pdl_enabled()
{
    static const bool on = getenv("SAGE_NO_PDL") == nullptr;
    return on;
}
template <typename... KArgs, typename... Args>
inline void sage_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = sage_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define SAGE_KLAUNCH_PDL(kernel, grid, block, smem, stream, ...) sage_launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)
__device__ __forceinline__ void sage_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sage_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
// where a kernel says "dependents may be scheduled": at its end (default: the dependent's launch overlaps
// only this kernel's drain, and no waiting CTA holds registers / shared memory the composite of the frame in
// flight on another stream could use) or at its top (-DSAGE_PDL_EARLY: experiment)
#ifdef SAGE_PDL_EARLY
#define SAGE_DEP_LAUNCH_TOP() sage_dep_launch()
#define SAGE_DEP_LAUNCH_END() do { } while (0)
#else
#define SAGE_DEP_LAUNCH_TOP() do { } while (0)
#define SAGE_DEP_LAUNCH_END() sage_dep_launch()
#endif
#include <stdint.h>
#include <stdlib.h>

#include "../../include/sage_splat.h"

#define SAGE_L255 (-7.99435344f)      /* log2(1/255): skip where log2(alpha) < this */
#define SAGE_L099 (-0.0144995697f)    /* log2(0.99): alpha cap, applied to the exponent */
#define SAGE_T_MIN 1e-4f
/* binning granularity: tiles of 32 x 16 pixels (W x H); a warp blends one of a tile's eight 8 x 8 quadrants.
 * (16 x 16 tiles, rounds 1-2a: 1.41x the (tile, Gaussian) instances through emission and the tile sort.) */
#define SAGE_BIN_W 32
#define SAGE_BIN_H 16
#define SAGE_QUADS_X (SAGE_BIN_W / 8)
#define SAGE_MASK_BITS 8              /* instance value = Gaussian id << 8 | quadrant mask (bit qy * 4 + qx) */
#define SAGE_BLUR 0.3f
#define SAGE_LOG2E 1.44269504f
#define SAGE_KEY_CULLED 0xFFFFFFFFu
// instance value = Gaussian id (24 bits) << 8 | 8x8-quadrant overlap mask of the tile (8 bits)
#define SAGE_ID_MASK 0x00FFFFFFu      /* scene size limit (2^24 - 1 = 16.7 M Gaussians) */
#define SAGE_MAX_PEERS 8

// device counters (uint32 each)
enum { CNT_VISIBLE = 0, CNT_INSTANCES = 1, CNT_OVERFLOW = 2, CNT_WANTED = 3, CNT_COUNT = 8 };

struct SageCamDev {
    float R[9];
    float t[3];
    float campos[3];
    float fx, fy, cx, cy;
    float limx, limy;
    float near_z, far_z;
    int W, H;
    int tiles_x, tiles_y;  // full-frame grid of 16 x 16 tiles (the public sharding unit)
    int row0, row1;        // tile rows rendered by this call
    float depth_lo, depth_hi;  // output clamp of the depth map (lo > hi: off)
    int bins_x;            // 8 x 8 bins per image row
    int bin_row0, bin_row1;    // bin rows rendered by this call
    int n_bins;            // bins_x * (bin_row1 - bin_row0)
    int tile_passes;       // radix passes of the by-tile sort
    int tile_digit_bits;   // their digit width: ceil(bits of the tile ids / passes), 5..8
};

// destination frame buffers of one render call (several: peer GPUs, NVLink)
struct SageOutDev {
    uint8_t* rgb[SAGE_MAX_PEERS];
    float* depth[SAGE_MAX_PEERS];
    int n;
    int wide;     // 8-byte stores allowed: W % 8 == 0 and every buffer 8-byte aligned
    int rgba;     // 4 bytes per pixel (alpha = 255) instead of 3
};

// Everything that changes from frame to frame.  Lives in DEVICE memory (one block per frame
// context, refreshed by a 400-byte host-to-device copy), so the kernels' launch parameters never
// change and a whole frame replays as one CUDA graph.
struct SageFrameParams {
    SageCamDev cam;
    SageOutDev out;
    uint32_t epoch;       // look-back epochs epoch .. epoch+7 are fresh for this frame
    uint32_t cap;         // instance capacity of the frame context
};

// ---- deterministic exp2 (mirrors the spec) --------------------------------
__device__ __forceinline__ float sage_exp2(float x)
{
    x = fminf(fmaxf(x, -125.0f), 125.0f);
    const float magic = 12582912.0f;
    float t = x + magic;
    float n = t - magic;
    float f = x - n;
    float p = 0x1.5c08e4p-10f;
    p = fmaf(p, f, 0x1.3d0c52p-7f);
    p = fmaf(p, f, 0x1.c6b6e4p-5f);
    p = fmaf(p, f, 0x1.ebf918p-3f);
    p = fmaf(p, f, 0x1.62e428p-1f);
    p = fmaf(p, f, 0x1.000002p+0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ float sage_exp(float x) { return sage_exp2(x * SAGE_LOG2E); }

// deterministic log2 for x in (0, 1] (mirrors so_log2); 0 -> -127
__device__ __forceinline__ float sage_log2(float x)
{
    const uint32_t u = __float_as_uint(x);
    float e = (float)((int)((u >> 23) & 255u) - 127);
    float m = __uint_as_float((u & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) { m = m * 0.5f; e = e + 1.0f; }
    const float s = (m - 1.0f) / (m + 1.0f);
    const float s2 = s * s;
    float p = 0x1.c71c72p-4f;
    p = fmaf(p, s2, 0x1.24924ap-3f);
    p = fmaf(p, s2, 0x1.99999ap-3f);
    p = fmaf(p, s2, 0x1.555556p-2f);
    p = fmaf(p, s2, 1.0f);
    const float ln_m = (2.0f * s) * p;
    return fmaf(ln_m, SAGE_LOG2E, e);
}

// ---- exact bin culling: the {alpha >= 1/255} ellipse against bands of 8 pixel rows -------------
// Binning is not part of the numerical spec: it only has to be conservative.
// The support of a splat is the ellipse  A dx^2 + 2B dx dy + C dy^2 <= t2  around (gx, gy), in the
// log2-scaled units of the conic the blend kernel uses (A = -qa, B = -qb/2, C = -qc; t2 = lo - L255).
// k_emit intersects it with the two 8-row halves of every tile row: a tile is emitted when one of
// its halves contains a pixel centre of the ellipse, with the mask of the 8 x 8 quadrants that do.  Binary32 arithmetic (binary64 sqrt/div sequences cost more than the instances
// they save) made conservative by slack: t2 * 1.003 + 0.02 (the bounding box uses 1.002) and every
// span is widened by 0.1 % of the half extent + 0.01 px; tools/cuda_emu/band_check.cpp brute-forces
// the guarantee on CPU.
// sqrt to ~2 ulp (one MUFU + one multiply); the spans carry 0.1 % of slack.  Negative -> 0.
__device__ __forceinline__ float sage_sqrt_fast(float x) { return x > 0.0f ? x * rsqrtf(x) : 0.0f; }

// true: the pixel box spans at least 2 x 2 blocks of 16 x 16 pixels — the splats worth intersecting with
// the row bands (the ellipse test then removes tiles AND tightens the quadrant masks; a 32-px-wide tile
// criterion would leave splats of 17-32 px with loose box masks: measured +6 % composite time)
__device__ __forceinline__ bool sage_box_cullable(int ix0, int ix1, int iy0, int iy1)
{
    return (ix0 / 16 != ix1 / 16) && (iy0 / 16 != iy1 / 16);
}

// The blend evaluates the exponent in binary32: at an offset d from the centre its rounding error is a few ulp of the
// LARGEST term (|qa| dx^2, |qb dx dy|, |qc| dy^2), which for needle-like splats is 1e5 - 1e6 times the sum - the rounded
// exponent of the spec can pass the 1/255 threshold up to ~10 % beyond the exact ellipse along the needle.  Culling has to
// bound where the ROUNDED exponent can pass: the ellipse of the conic deflated by that allowance, with
// |dx dy| <= (dx^2 + dy^2) / 2:  A' = A - eps (A + |B|),  C' = C - eps (C + |B|)  (eps = 1e-6, four times the worst case of
// the blend's operation sequence).  When the deflated form is no longer definite the caller covers everything.
#ifndef SAGE_CULL_EPS
#define SAGE_CULL_EPS 1e-6f        /* (overridable so that tools/cuda_emu/run_band_check.py can show what happens without it) */
#endif
__device__ __forceinline__ void sage_conic_deflate(float* A, float B, float* C)
{
    const float ab = fabsf(B);
    *A -= SAGE_CULL_EPS * (*A + ab);
    *C -= SAGE_CULL_EPS * (*C + ab);
}

// A C - B^2 without the cancellation of the plain form (Kahan): w = B B rounded, e = its rounding error (exact,
// by fma), f = A C - w fused; the sum is within ~2 ulp of the true determinant
__device__ __forceinline__ float sage_det_kahan(float A, float C, float B)
{
    const float w = B * B;
    const float e = fmaf(-B, B, w);
    const float f = fmaf(A, C, -w);
    return f + e;
}

struct SageBand {
    float gx, gy, A, B, C, t2, det2, invA;
    float hx, hy;       // half extents of the ellipse
    float dyr;          // dy of its right-most point (-dyr: left-most)
    float slack;        // outward widening of every span, pixels
    int unbounded;      // degenerate conic: treat as covering everything
};

// l2 = log2(255 * opacity) = lo - SAGE_L255
__device__ __forceinline__ SageBand sage_band_make(float gx, float gy, float qa, float qb, float qc, float l2)
{
    SageBand b;
    b.t2 = l2 * 1.003f + 0.02f;
    b.gx = gx; b.gy = gy;
    b.A = -qa; b.B = -0.5f * qb; b.C = -qc;
    sage_conic_deflate(&b.A, b.B, &b.C);
    b.det2 = sage_det_kahan(b.A, b.C, b.B);
    b.unbounded = (!(b.det2 > 0.0f) || !(b.A > 0.0f) || !(b.C > 0.0f) || !(b.t2 > 0.0f)) ? 1 : 0;
    b.hx = b.unbounded ? 1e30f : sqrtf(b.t2 * b.C / b.det2);
    b.hy = b.unbounded ? 1e30f : sqrtf(b.t2 * b.A / b.det2);
    b.dyr = b.unbounded ? 0.0f : -(b.B * b.hx) / b.C;
    b.invA = b.unbounded ? 0.0f : 1.0f / b.A;
    b.slack = b.unbounded ? 0.0f : 1e-3f * fmaxf(b.hx, b.hy) + 0.01f;
    if (!(b.hx < 1e9f) || !(b.hy < 1e9f)) b.unbounded = 1;       // overflow / NaN: cover everything
    return b;
}

// Inclusive pixel-column range [*xs, *xe] (clamped to [ix0, ix1]) that contains every pixel of rows
// y0..y1 (inclusive, y0 <= y1) whose centre lies in the ellipse.  Returns false if there is none.
__device__ __forceinline__ bool sage_band_span(const SageBand& b, int ix0, int ix1, int y0, int y1, int* xs, int* xe)
{
    if (b.unbounded) { *xs = ix0; *xe = ix1; return ix0 <= ix1; }
    float d0 = ((float)y0 + 0.5f) - b.gy, d1 = ((float)y1 + 0.5f) - b.gy;       // dy of the first / last row centre
    if (d0 > b.hy + b.slack || d1 < -(b.hy + b.slack)) return false;
    d0 = fmaxf(d0, -b.hy); d1 = fminf(d1, b.hy);
    // right edge u(dy) = (-B dy + sqrt(A t2 - det2 dy^2)) / A is concave with its maximum hx at dyr
    float xr;
    if (d0 <= b.dyr && b.dyr <= d1) xr = b.hx;
    else {
        const float dy = b.dyr < d0 ? d0 : d1;
        xr = (-(b.B * dy) + sage_sqrt_fast(b.A * b.t2 - b.det2 * dy * dy)) * b.invA;
    }
    // left edge l(dy) = (-B dy - sqrt(...)) / A is convex with its minimum -hx at -dyr
    float xl;
    if (d0 <= -b.dyr && -b.dyr <= d1) xl = -b.hx;
    else {
        const float dy = -b.dyr < d0 ? d0 : d1;
        xl = (-(b.B * dy) - sage_sqrt_fast(b.A * b.t2 - b.det2 * dy * dy)) * b.invA;
    }
    const float lo = floorf((b.gx + xl) - 0.5f - b.slack), hi = ceilf((b.gx + xr) - 0.5f + b.slack);
    if (hi < (float)ix0 || lo > (float)ix1) return false;
    *xs = (int)fmaxf(lo, (float)ix0);
    *xe = (int)fminf(hi, (float)ix1);
    return *xs <= *xe;
}

// One tile row ty of a splat with pixel box [ix0,ix1] x [iy0,iy1]: spans of its upper and lower
// 8-row halves (packed x0 | x1 << 16; SAGE_SPAN_EMPTY = x0 > x1) and the inclusive tile range.
struct SageRowItem {
    uint32_t top, bot;      // pixel spans of rows [16ty, 16ty+7] and [16ty+8, 16ty+15]; SAGE_SPAN_EMPTY if none
    int tx0, tx1;           // tiles [tx0, tx1]; tx0 > tx1: the row is empty
};
#define SAGE_SPAN_EMPTY 0x0000FFFFu     /* x0 = 0xFFFF > x1 = 0 */

// (cullable splats only: the box spans >= 2 x 2 tiles)
__device__ __forceinline__ SageRowItem sage_band_row(const SageBand& b, int ix0, int ix1, int iy0, int iy1, int ty)
{
    SageRowItem it;
    it.top = it.bot = SAGE_SPAN_EMPTY;
    int lo = 1 << 30, hi = -1;
    int xs, xe;
    const int ya = max(ty * SAGE_BIN_H, iy0), yb = min(ty * SAGE_BIN_H + 7, iy1);
    if (ya <= yb && sage_band_span(b, ix0, ix1, ya, yb, &xs, &xe)) {
        it.top = (uint32_t)xs | ((uint32_t)xe << 16);
        lo = min(lo, xs); hi = max(hi, xe);
    }
    const int yc = max(ty * SAGE_BIN_H + 8, iy0), yd = min(ty * SAGE_BIN_H + 15, iy1);
    if (yc <= yd && sage_band_span(b, ix0, ix1, yc, yd, &xs, &xe)) {
        it.bot = (uint32_t)xs | ((uint32_t)xe << 16);
        lo = min(lo, xs); hi = max(hi, xe);
    }
    if (hi < 0) { it.tx0 = 0; it.tx1 = -1; }
    else { it.tx0 = lo / SAGE_BIN_W; it.tx1 = hi / SAGE_BIN_W; }
    return it;
}

// rebuild a row item from its two stored spans
__device__ __forceinline__ SageRowItem sage_row_from_spans(uint32_t top, uint32_t bot)
{
    SageRowItem it;
    it.top = top; it.bot = bot;
    int lo = 1 << 30, hi = -1;
    if ((top & 0xFFFFu) <= (top >> 16)) { lo = min(lo, (int)(top & 0xFFFFu)); hi = max(hi, (int)(top >> 16)); }
    if ((bot & 0xFFFFu) <= (bot >> 16)) { lo = min(lo, (int)(bot & 0xFFFFu)); hi = max(hi, (int)(bot >> 16)); }
    if (hi < 0) { it.tx0 = 0; it.tx1 = -1; }
    else { it.tx0 = lo / SAGE_BIN_W; it.tx1 = hi / SAGE_BIN_W; }
    return it;
}

// quadrant columns (bits 0 .. SAGE_QUADS_X-1) of the tile starting at pixel bx that the inclusive pixel span
// [x0, x1] touches; 0 if the span is empty or misses the tile
__device__ __forceinline__ uint32_t sage_span_cols(int x0, int x1, int bx)
{
    const int lo = max(x0 - bx, 0) >> 3, hi = min(x1 - bx, SAGE_BIN_W - 1) >> 3;      // first / last quadrant column
    const uint32_t m = ((2u << hi) - 1u) & ~((1u << lo) - 1u);
    return (x0 <= x1 && x1 >= bx && x0 <= bx + SAGE_BIN_W - 1) ? m : 0u;
}

// 8x8-quadrant mask (bit qy*4+qx, the warp -> quadrant map of the composite kernel) of tile tx
__device__ __forceinline__ uint32_t sage_row_quad_mask(uint32_t top, uint32_t bot, int tx)
{
    const int bx = tx * SAGE_BIN_W;
    return sage_span_cols((int)(top & 0xFFFFu), (int)(top >> 16), bx) |
           (sage_span_cols((int)(bot & 0xFFFFu), (int)(bot >> 16), bx) << SAGE_QUADS_X);
}

// the same from the inclusive pixel box of a splat that is not culled per row
__device__ __forceinline__ uint32_t sage_box_quad_mask(int px0, int px1, int py0, int py1, int tx, int ty)
{
    const int by = ty * SAGE_BIN_H;
    const uint32_t cols = sage_span_cols(px0, px1, tx * SAGE_BIN_W);
    return (py0 <= by + 7 ? cols : 0u) | (py1 >= by + 8 ? cols << SAGE_QUADS_X : 0u);
}

// ---- frame workspace ---------------------------------------------------------------------
#define SAGE_RADIX_MAX_BINS 512
#define SAGE_MAX_PASSES 4
#define SAGE_N_CLASSES 33          /* list-length classes of the bins: 0 = empty, k: 2^(k-1) <= n < 2^k */
enum { SAGE_TICKET_EMIT = 0, SAGE_TICKET_COMPOSITE = 1, SAGE_TICKET_COUNT = 4 };

struct SageFrame {
    const SageFrameParams* fp;   // device copy of the frame parameters (camera + output buffers)
    // per-Gaussian (capacity n)
    float4* rec;         // 3 float4 per Gaussian (one 48-byte block): gx gy qa lo | qb qc - - | r g b z
    uint32_t* dkey_src;  // depth keys bits(z) - bits(near) (or CULLED) in index order; never overwritten by the sort
    uint32_t* dkey[2];   // sort ping-pong
    uint32_t* dval[2];   // Gaussian ids, ping-pong
    uint2* rect;         // inclusive PIXEL bounding box: x = x0 | x1<<16, y = y0 | y1<<16
    // per-instance (capacity cap)
    uint32_t* ikey[2];   // tile id, ping-pong
    uint32_t* ival[2];   // Gaussian id << 8 | quadrant mask, ping-pong
    // sort / scan state
    uint32_t* sort_h;       // [SAGE_RADIX_MAX_BINS][sort_ld] per-tile digit histograms / exclusive prefixes of a pass
    uint32_t* sort_totals;  // [SAGE_RADIX_MAX_BINS] digit totals of a pass
    int sort_ld;
    unsigned long long* chain;     // [n/256 + 1] chained-scan words of the emission kernel (epoch-tagged: never cleared)
    // one block cleared per frame: counters | tickets | class_count
    uint32_t* counters;     // CNT_COUNT
    uint32_t* tickets;      // SAGE_TICKET_COUNT
    uint32_t* class_count;  // [SAGE_N_CLASSES]
    // per-bin (capacity max_bins)
    uint2* ranges;          // (~begin, end) of the bin's list in the sorted instances; (0, 0) = empty (cleared per frame)
    uint32_t* class_seg;    // [SAGE_N_CLASSES][max_bins] bins of each length class, in arrival order
    uint32_t* bin_order;    // bins by descending length class: the composite's work queue
    int max_bins;
};
#define SAGE_ZERO_WORDS (CNT_COUNT + SAGE_TICKET_COUNT + SAGE_N_CLASSES + 3)

struct SageSortPlan {
    const uint32_t* keys_src;  // input keys of pass 0 (may alias keys[0])
    uint32_t* keys[2];
    uint32_t* vals[2];         // pass 0 reads vals[0] unless vals_are_iota
    const uint32_t* n_first;   // element count of pass 0 (device)
    const uint32_t* n_rest;    // element count of the later passes (device)
    int lo_bit, n_bits;
    int digit_bits, passes;    // from sage_sort_shape, or forced by the caller
    int items_per_thread;      // 8 or 16 (x 512 threads = items per tile)
    bool vals_are_iota, compact_first;
    uint32_t* H;               // [1 << digit_bits][ld] scratch (SageFrame::sort_h)
    int ld;
    uint32_t* totals;          // [1 << digit_bits] scratch
    uint2* ranges;             // optional: the last pass writes ranges[key] = (~begin, end) (cleared by the caller)
};

// ---- kernel launchers (defined in the .cu files) -------------------------------------------
struct SageSceneDev {
    int64_t n;
    int deg;
    int n_chunks;       // float4 chunks of SH per Gaussian
    float4* geo0;       // mu.xyz, log2(sigmoid(opacity))
    float4* geo1;       // S00 S01 S02 S11
    float2* geo2;       // S12 S22
    float4* shc;        // [n][n_chunks]: SH coefficients, contiguous per Gaussian
};

void sage_launch_precompute(const SageSceneDev& sc, const float* xyz, const float* log_scale,
                            const float* quat, const float* opacity, const float* sh, cudaStream_t st);
void sage_launch_project(const SageSceneDev& sc, const SageFrame& fr, cudaStream_t st);
void sage_launch_sh_color(const SageSceneDev& sc, const SageFrame& fr, cudaStream_t st);

// Stable LSD radix sort, three kernels per pass (tile histograms, scan, rank + scatter).  Returns the
// index (0/1) of the buffer pair that holds the result, or -1 for an unsupported shape.
int sage_radix_sort(const SageSortPlan& plan, int* launches, cudaStream_t st);
void sage_sort_shape(int n_bits, int* digit_bits, int* passes);
int sage_sort_matrix_ld(int64_t capacity);
size_t sage_sort_matrix_words(int64_t capacity);

void sage_launch_emit(const SageFrame& fr, const uint32_t* order, cudaStream_t st);
void sage_launch_bin_order(const SageFrame& fr, cudaStream_t st);
void sage_launch_composite(const SageFrame& fr, const uint32_t* sorted_vals, int n_bins, bool whole_frame, cudaStream_t st);

// One whole frame (every launch between the parameter upload and the counter read-back): the front
// half on stream `fs`, the view-dependent colours on `ss`, the composite on the caller's stream
// `cs` (forked from / joined to it with the four events).  Pure launch sequence: capturable into a
// CUDA graph.  ev: optional per-stage timing events (SAGE_N_STAGES + 1) — then everything runs in
// line on `cs`.  Returns the number of kernel launches, or -1 for an unsupported sort shape.
struct SageFrameShape {
    int depth_bits;      // significant bits of the depth keys
    int tile_passes;     // passes of the by-tile sort
    int tile_digit_bits; // digit width of those passes (5..8)
    int whole_frame;     // the rows are the whole frame (not a tile-row block of a multi-GPU split)
};
int sage_frame_enqueue(const SageSceneDev& sc, const SageFrame& fr, const SageFrameShape& shape,
                       const uint32_t* n_dev, cudaStream_t cs, cudaStream_t fs, cudaStream_t ss,
                       cudaEvent_t* fork_join, cudaEvent_t* ev, int n_bins);
void sage_sort_configure(void);
void sage_bin_configure(void);
void sage_composite_configure(void);
void sage_project_configure(void);

#ifndef SAGE_CUDA_EMU
// Experiment knob (off by default): ask for the maximum shared-memory carve-out on every kernel of a
// frame so an SM never changes configuration between them.  Measured (profiles/r02_variants.md):
// it costs k_sh_color its L1 (0.043 -> 0.069 ms) and the frame 5 %; left in for SAGE_CARVEOUT=1.
template <typename K>
inline void sage_prefer_smem(K kernel)
{
    static const bool on = getenv("SAGE_CARVEOUT") != nullptr;
    if (on) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
#endif

void sage_launch_collide_grid(const uint8_t* map, int H, int W, double min_x, double max_x,
                              double min_y, double max_y, double scale, const float* pos_f32,
                              const double* pos_f64, int P, int np2_promote, uint8_t* hit,
                              cudaStream_t st);
void sage_launch_first_hit(const uint8_t* hit, const int32_t* seg_len, int n_segments,
                           int32_t* first_hit, cudaStream_t st);
void sage_launch_collide_capsule(int64_t m, const float4* c0, const float4* c1, const float2* c2,
                                 const float* pos, int P, float radius, float z0, float z1,
                                 uint32_t* hit_bits, int32_t* first_body, cudaStream_t st);
void sage_launch_bodies_pack(int64_t m, const float* mu, const float* log_scale, const float* quat,
                             float4* c0, float4* c1, float2* c2, cudaStream_t st);
int sage_capsule_max_poses(void);

// ---- JPEG encoder internals shared with scene_api.cu (jpeg.cu) -------------------------------
struct sage_jpeg;
int sage_jpeg_enqueue(sage_jpeg* j, const uint8_t* rgb_dev, int bpp, int batch, int slot0, uint8_t* out_dev, int64_t out_stride,
                      long long* sizes_dev, cudaStream_t st);
uint8_t* sage_jpeg_stage_out(sage_jpeg* j);
long long* sage_jpeg_stage_sizes_dev(sage_jpeg* j);
long long* sage_jpeg_stage_sizes_host(sage_jpeg* j);
int64_t sage_jpeg_stage_stride(const sage_jpeg* j);
uint8_t* sage_jpeg_stage_out_host(sage_jpeg* j);
long long* sage_jpeg_stage_guess(sage_jpeg* j);
int sage_jpeg_width(const sage_jpeg* j);
int sage_jpeg_height(const sage_jpeg* j);
int sage_jpeg_device(const sage_jpeg* j);
int sage_jpeg_max_batch(const sage_jpeg* j);

// ---- depth PNG encoder internals shared with scene_api.cu (png.cu) ----------------------------
struct sage_png;
int sage_png_enqueue(sage_png* p, const float* depth_dev, float lo, float hi, float scale, int slot, uint8_t* out_dev,
                     int64_t out_cap, long long* size_dev, uint16_t* twin_dev, cudaStream_t st);
uint8_t* sage_png_stage_out(sage_png* p);
uint8_t* sage_png_stage_out_host(sage_png* p);
long long* sage_png_stage_sizes_dev(sage_png* p);
long long* sage_png_stage_sizes_host(sage_png* p);
int64_t sage_png_stage_stride(const sage_png* p);
int sage_png_width(const sage_png* p);
int sage_png_height(const sage_png* p);
int sage_png_device(const sage_png* p);
int sage_png_max_batch(const sage_png* p);
