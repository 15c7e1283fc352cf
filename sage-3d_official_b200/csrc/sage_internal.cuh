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
#endif
#include <stdint.h>

#include "../../include/sage_splat.h"

#define SAGE_ALPHA_MIN 0.003921569f
#define SAGE_ALPHA_MAX 0.99f
#define SAGE_T_MIN 1e-4f
#define SAGE_BLUR 0.3f
#define SAGE_LOG2E 1.44269504f
#define SAGE_KEY_CULLED 0xFFFFFFFFu
// instance value = Gaussian id (28 bits) << 4 | 8x8-quadrant overlap mask of the tile (4 bits)
#define SAGE_ID_MASK 0x0FFFFFFFu

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
    int tiles_x, tiles_y;  // full-frame tile grid
    int row0, row1;        // tile rows rendered by this call
    float depth_lo, depth_hi;  // output clamp of the depth map (lo > hi: off)
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

// ---- frame workspace ---------------------------------------------------------------------
#define SAGE_RADIX_MAX_BINS 512
#define SAGE_MAX_PASSES 4
enum { SAGE_TICKET_DEPTH = 0, SAGE_TICKET_EMIT = 4, SAGE_TICKET_TILE = 5, SAGE_TICKET_COUNT = 12 };

struct SageFrame {
    // per-Gaussian (capacity n)
    float4* rec;         // 3 float4 per Gaussian (one 48-byte TMA block): gx gy qa thr | qb qc o z | r g b 0
    uint32_t* dkey[2];   // depth keys bits(z) - bits(near) (or CULLED), ping-pong
    uint32_t* dval[2];   // Gaussian ids, ping-pong
    uint2* rect;         // inclusive PIXEL bounding box: x = x0 | x1<<16, y = y0 | y1<<16
    uint32_t* ntiles;    // tiles touched
    // per-instance (capacity cap)
    uint32_t* ikey[2];   // tile id, ping-pong
    uint32_t* ival[2];   // id << 4 | quadrant mask, ping-pong
    // sort / scan state
    unsigned long long* lookback;  // [blocks][SAGE_RADIX_MAX_BINS] epoch|state|value words (never cleared)
    unsigned long long* chain;     // [n/256 + 1] chained-scan words of the emission kernel
    // one block cleared per frame: counters | tickets | depth_hist | tile_hist
    uint32_t* counters;     // CNT_COUNT
    uint32_t* tickets;      // SAGE_TICKET_COUNT
    uint32_t* depth_hist;   // [SAGE_MAX_PASSES][SAGE_RADIX_MAX_BINS]
    uint32_t* tile_hist;    // [SAGE_MAX_PASSES][256]
    // per-tile
    uint2* ranges;
    uint32_t* tile_order;   // tiles by descending list length: CTA b of the composite takes tile_order[b]
};
#define SAGE_ZERO_WORDS (CNT_COUNT + SAGE_TICKET_COUNT + SAGE_MAX_PASSES * SAGE_RADIX_MAX_BINS + SAGE_MAX_PASSES * 256)

struct SageSortPlan {
    uint32_t* keys[2];
    uint32_t* vals[2];
    const uint32_t* n_first;   // element count of pass 0 (device)
    const uint32_t* n_rest;    // element count of the later passes (device)
    int64_t capacity;
    int lo_bit, n_bits;
    int digit_bits, passes;    // from sage_sort_shape, or forced by the caller
    int items_per_thread;      // 8, 16 or 32 (x 512 threads = items per block; fewer blocks = shorter look-back chains)
    bool vals_are_iota, compact_first, hist_ready;
    uint32_t* hist;            // [passes][1 << digit_bits], zeroed (or pre-accumulated if hist_ready)
    uint32_t* total_out;       // optional: total of pass 0's histogram
    unsigned long long* lookback;
    uint32_t* tickets;         // [passes], zeroed
    uint32_t epoch0;           // epochs epoch0 .. epoch0 + passes - 1 must be fresh
};

// ---- kernel launchers (defined in the .cu files) -------------------------------------------
struct SageSceneDev {
    int64_t n;
    int deg;
    int n_chunks;       // float4 chunks of SH per Gaussian
    float4* geo0;       // mu.xyz, sigmoid(opacity)
    float4* geo1;       // S00 S01 S02 S11
    float2* geo2;       // S12 S22
    float4* shc;        // [n][n_chunks]: SH coefficients, contiguous per Gaussian
};

void sage_launch_precompute(const SageSceneDev& sc, const float* xyz, const float* log_scale,
                            const float* quat, const float* opacity, const float* sh, cudaStream_t st);
void sage_launch_project(const SageSceneDev& sc, const SageCamDev& cam, const SageFrame& fr, cudaStream_t st);
void sage_launch_sh_color(const SageSceneDev& sc, const SageCamDev& cam, const SageFrame& fr,
                          const uint32_t* order, cudaStream_t st);

// Stable LSD radix sort, one kernel per pass (decoupled look-back).  Returns the index (0/1) of the
// buffer pair that holds the result, or -1 for an unsupported shape.
int sage_radix_sort(const SageSortPlan& plan, int* launches, cudaStream_t st);
void sage_sort_shape(int n_bits, int* digit_bits, int* passes);
size_t sage_sort_lookback_words(int64_t capacity);

void sage_launch_emit(const SageCamDev& cam, const SageFrame& fr, const uint32_t* order, int64_t n,
                      int64_t cap, int tile_passes, uint32_t epoch, cudaStream_t st);
void sage_launch_ranges(const SageFrame& fr, const uint32_t* sorted_keys, int64_t cap, int n_tiles, cudaStream_t st);
void sage_launch_composite(const SageCamDev& cam, const SageFrame& fr, const uint32_t* sorted_vals,
                           int n_peers, uint8_t* const* rgb_peers, float* const* depth_peers,
                           cudaStream_t st);

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
int sage_jpeg_enqueue(sage_jpeg* j, const uint8_t* rgb_dev, int batch, int slot0, uint8_t* out_dev, int64_t out_stride,
                      long long* sizes_dev, cudaStream_t st);
uint8_t* sage_jpeg_stage_out(sage_jpeg* j);
long long* sage_jpeg_stage_sizes_dev(sage_jpeg* j);
long long* sage_jpeg_stage_sizes_host(sage_jpeg* j);
int64_t sage_jpeg_stage_stride(const sage_jpeg* j);
int sage_jpeg_width(const sage_jpeg* j);
int sage_jpeg_height(const sage_jpeg* j);
int sage_jpeg_device(const sage_jpeg* j);
int sage_jpeg_max_batch(const sage_jpeg* j);
