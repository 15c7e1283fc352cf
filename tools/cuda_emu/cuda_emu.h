// cuda_emu.h — a tiny lock-step CUDA emulator for DEVELOPMENT ONLY.
//
// The authoring container has no GPU.  To shake out logic errors in warp-cooperative kernels
// (ballots, shuffles, shared-memory phases) before spending scarce GPU minutes, a kernel source
// can be compiled with g++ -DSAGE_CUDA_EMU: every CUDA thread of a CTA becomes an OS thread,
// __syncthreads/__syncwarp are barriers, shuffles and ballots exchange through per-warp slots,
// CTAs run one after another.  "Device memory" is host memory.  This is NOT a CPU fallback: it
// is never built into libsage_splat.so, never imported by the package, tests or bench — only
// tools/cuda_emu/run_*.py use it.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static const
#define __grid_constant__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
// packed binary32 pair operations: IEEE per lane, exactly as the hardware's *.f32x2 forms
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }

namespace cuda_emu {
struct Cta {
    int nthreads;
    std::unique_ptr<std::barrier<>> cta_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<unsigned long long> slots;   // [nthreads] exchange slots
};
inline thread_local Cta* t_cta = nullptr;
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;

template <typename F>
void launch(dim3 grid, dim3 block, F&& body)
{
    const int nt = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Cta cta;
                cta.nthreads = nt;
                cta.cta_bar = std::make_unique<std::barrier<>>(nt);
                for (int w = 0; w < (nt + 31) / 32; ++w) {
                    const int live = std::min(32, nt - 32 * w);
                    cta.warp_bar.push_back(std::make_unique<std::barrier<>>(live));
                }
                cta.slots.assign(nt, 0);
                std::vector<std::thread> th;
                th.reserve(nt);
                for (int t = 0; t < nt; ++t)
                    th.emplace_back([&, t] {
                        t_cta = &cta;
                        t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        t_blockIdx = dim3(bx, by, bz);
                        t_blockDim = block;
                        t_gridDim = grid;
                        body();
                        // a thread that returns early must not strand the others at a CTA barrier:
                        // kernels written for this tool only return warp- or CTA-uniformly and never
                        // hit __syncthreads afterwards, so dropping out of the barrier is correct
                        cta.cta_bar->arrive_and_drop();
                    });
                for (auto& x : th) x.join();
            }
}
inline int lane_id() { return (int)(t_threadIdx.x + t_threadIdx.y * t_blockDim.x) & 31; }
inline int lin_tid() { return (int)(t_threadIdx.x + t_threadIdx.y * t_blockDim.x + t_threadIdx.z * t_blockDim.x * t_blockDim.y); }
inline void warp_sync() { t_cta->warp_bar[lin_tid() >> 5]->arrive_and_wait(); }
inline unsigned long long exchange(unsigned long long v, int src_lane)
{
    const int tid = lin_tid(), w0 = tid & ~31;
    t_cta->slots[tid] = v;
    warp_sync();
    const unsigned long long r = t_cta->slots[w0 + src_lane];
    warp_sync();
    return r;
}
}  // namespace cuda_emu

#define threadIdx cuda_emu::t_threadIdx
#define blockIdx cuda_emu::t_blockIdx
#define blockDim cuda_emu::t_blockDim
#define gridDim cuda_emu::t_gridDim

inline void __syncthreads() { cuda_emu::t_cta->cta_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cuda_emu::warp_sync(); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    const int tid = cuda_emu::lin_tid(), w0 = tid & ~31;
    cuda_emu::t_cta->slots[tid] = pred ? 1 : 0;
    cuda_emu::warp_sync();
    unsigned m = 0;
    for (int l = 0; l < 32 && w0 + l < cuda_emu::t_cta->nthreads; ++l)
        if (cuda_emu::t_cta->slots[w0 + l]) m |= 1u << l;
    cuda_emu::warp_sync();
    return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
    const int tid = cuda_emu::lin_tid(), w0 = tid & ~31;
    cuda_emu::t_cta->slots[tid] = v;
    cuda_emu::warp_sync();
    unsigned r = 0xFFFFFFFFu;
    for (int l = 0; l < 32 && w0 + l < cuda_emu::t_cta->nthreads; ++l)
        if ((unsigned)cuda_emu::t_cta->slots[w0 + l] < r) r = (unsigned)cuda_emu::t_cta->slots[w0 + l];
    cuda_emu::warp_sync();
    return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    const int tid = cuda_emu::lin_tid(), w0 = tid & ~31;
    cuda_emu::t_cta->slots[tid] = v;
    cuda_emu::warp_sync();
    unsigned r = 0u;
    for (int l = 0; l < 32 && w0 + l < cuda_emu::t_cta->nthreads; ++l)
        if ((unsigned)cuda_emu::t_cta->slots[w0 + l] > r) r = (unsigned)cuda_emu::t_cta->slots[w0 + l];
    cuda_emu::warp_sync();
    return r;
}
inline int __all_sync(unsigned m, int pred)
{
    const int tid = cuda_emu::lin_tid(), w0 = tid & ~31;
    int live = cuda_emu::t_cta->nthreads - w0; if (live > 32) live = 32;
    return __ballot_sync(m, pred) == (live == 32 ? 0xFFFFFFFFu : ((1u << live) - 1u));
}
inline unsigned __match_any_sync(unsigned, unsigned v)
{
    const int tid = cuda_emu::lin_tid(), w0 = tid & ~31;
    cuda_emu::t_cta->slots[tid] = v;
    cuda_emu::warp_sync();
    unsigned m = 0;
    for (int l = 0; l < 32 && w0 + l < cuda_emu::t_cta->nthreads; ++l)
        if ((unsigned)cuda_emu::t_cta->slots[w0 + l] == v) m |= 1u << l;
    cuda_emu::warp_sync();
    return m;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline void __trap() { fprintf(stderr, "cuda_emu: __trap()\n"); abort(); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
template <typename T>
inline T __shfl_sync(unsigned, T v, int src)
{
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    b = cuda_emu::exchange(b, src & 31);
    T r;
    memcpy(&r, &b, sizeof(T));
    return r;
}
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, int d)
{
    const int l = cuda_emu::lane_id();
    return __shfl_sync(m, v, l - d >= 0 ? l - d : l);
}
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, int d)
{
    const int l = cuda_emu::lane_id();
    return __shfl_sync(m, v, l + d < 32 ? l + d : l);
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int x)
{
    return __shfl_sync(m, v, cuda_emu::lane_id() ^ x);
}
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline float __fmul_rn(float a, float b) { return a * b; }
inline int atomicCAS(int* p, int cmp, int val)
{
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline unsigned atomicMax(unsigned* p, unsigned v)
{
    unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
inline unsigned atomicMin(unsigned* p, unsigned v)
{
    unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

// ---- runtime subset -------------------------------------------------------------------------
typedef int cudaError_t;
#define cudaSuccess 0
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeRelaxed = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return 0; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (void*)1; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return 0; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return 1; }       // no graphs: direct launches
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return 1; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return 1; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return 0; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return 0; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return 1; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
#define cudaHostAllocDefault 0
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); memset(*p, 0xCD, n); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }

#define SAGE_KLAUNCH(kernel, grid, block, stream, ...) \
    cuda_emu::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#define SAGE_KLAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
    cuda_emu::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#define SAGE_DYN_SMEM(type, name) static type name[49152]      /* CTAs run one after another */
inline float __log2f(float x) { return log2f(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
