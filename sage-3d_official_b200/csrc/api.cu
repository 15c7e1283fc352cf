// C-ABI of libsage_splat.so (declared in include/sage_splat.h), part 1: library info, CUDA-IPC
// plumbing, 2-D grid collide and capsule collide.  The scene / render entry points live in
// scene_api.cu.
#include <cmath>
#include <cstring>
#include <vector>

#include "api_common.cuh"

std::string& sage_err_slot()
{
    static thread_local std::string err;
    return err;
}

struct sage_grid {
    int device = 0;
    uint8_t* map = nullptr;
    int H = 0, W = 0;
    double min_x = 0, max_x = 0, min_y = 0, max_y = 0, scale = 1;
    // staging for up to `cap` positions
    int cap = 0;
    float* pos32 = nullptr;
    double* pos64 = nullptr;
    uint8_t* hit = nullptr;
    int32_t* seg = nullptr;
    int32_t* first = nullptr;
};

struct sage_bodies {
    int device = 0;
    int64_t m = 0;
    float4* c0 = nullptr;
    float4* c1 = nullptr;
    float2* c2 = nullptr;
    float* pos = nullptr;
    uint32_t* hit_bits = nullptr;
    int32_t* first = nullptr;
    int maxp = 0;
};

extern "C" {

const char* sage_last_error(void) { return sage_err_slot().c_str(); }
int sage_abi_version(void) { return SAGE_ABI_VERSION; }

int sage_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sage_ipc_export(void* dev_ptr, uint8_t handle_out[64], uint64_t* offset_out)
{
    REQUIRE(dev_ptr && handle_out && offset_out, "sage_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    // base of the enclosing allocation (driver entry point fetched at run time: no link-time libcuda)
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    REQUIRE(fn != nullptr && qr == cudaDriverEntryPointSuccess, "sage_ipc_export: cuMemGetAddressRange unavailable");
    unsigned long long base = 0;
    size_t size = 0;
    REQUIRE(((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) == 0, "sage_ipc_export: cuMemGetAddressRange failed");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
    memcpy(handle_out, &h, 64);
    *offset_out = (uint64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return 0;
}

int sage_ipc_open(const uint8_t handle[64], uint64_t offset, void** dev_ptr_out)
{
    REQUIRE(handle && dev_ptr_out, "sage_ipc_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = (void*)((uint8_t*)base + offset);
    return 0;
}

int sage_ipc_close(void* dev_ptr, uint64_t offset)
{
    if (!dev_ptr) return 0;
    CK(cudaIpcCloseMemHandle((void*)((uint8_t*)dev_ptr - offset)));
    return 0;
}

// ---- device-side frame barrier between the ranks of a node (flags in peer memory) ----------------
// flag_peers[r] -> rank r's flag row (n_peers words, one per writer), mapped here with sage_ipc_open.
// One thread per peer: publish `value` in the peer's row (slot = this rank) with a system-scope release
// after this stream's earlier work (the composite's peer stores) has completed, then wait until this
// rank's own row shows `value` from that peer.  No rank waits on anything but the other ranks reaching
// the same point of their own streams; a peer that never arrives traps after 20 s instead of hanging.
struct SagePeerFlags {
    uint32_t* row[SAGE_MAX_PEERS];
};

static __global__ void __launch_bounds__(32) k_peer_barrier(SagePeerFlags f, int n, int me, uint32_t value)
{
    const int r = threadIdx.x;
    if (r >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.row[r] + me), "r"(value) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.row[me] + r) : "memory");
        if ((int32_t)(v - value) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) __trap();
        __nanosleep(200);
    }
}

int sage_peer_barrier(uint32_t* const* flag_peers, int n_peers, int my_rank, uint32_t value, sage_stream_t stream)
{
    REQUIRE(flag_peers != nullptr && n_peers >= 1 && n_peers <= SAGE_MAX_PEERS, "sage_peer_barrier: n_peers must be in [1, 8]");
    REQUIRE(my_rank >= 0 && my_rank < n_peers, "sage_peer_barrier: my_rank out of range");
    SagePeerFlags f{};
    for (int r = 0; r < n_peers; ++r) {
        REQUIRE(flag_peers[r] != nullptr, "sage_peer_barrier: NULL flag row");
        f.row[r] = flag_peers[r];
    }
    SAGE_KLAUNCH(k_peer_barrier, 1, 32, (cudaStream_t)stream, f, n_peers, my_rank, value);
    CK(cudaGetLastError());
    return 0;
}

// ---- grid --------------------------------------------------------------------------------

int sage_grid_create(sage_grid** out, int device, const uint8_t* map_host, int H, int W,
                     double min_x, double max_x, double min_y, double max_y, double scale)
{
    REQUIRE(out != nullptr, "sage_grid_create: out is NULL");
    *out = nullptr;
    REQUIRE(sage_device_count() > 0, "sage_grid_create: no CUDA device (this library has no CPU fallback)");
    REQUIRE((map_host == nullptr) || (H > 0 && W > 0), "sage_grid_create: bad grid shape");
    REQUIRE(scale != 0.0, "sage_grid_create: scale must be non-zero");
    CK(cudaSetDevice(device));
    sage_grid* g = new sage_grid();
    g->device = device;
    g->H = map_host ? H : 0; g->W = map_host ? W : 0;
    g->min_x = min_x; g->max_x = max_x; g->min_y = min_y; g->max_y = max_y; g->scale = scale;
    g->cap = 1024;
    cudaError_t e = cudaSuccess;
    if (map_host) {
        e = dalloc(&g->map, (size_t)H * W);
        if (e == cudaSuccess) e = cudaMemcpy(g->map, map_host, (size_t)H * W, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = dalloc(&g->pos32, (size_t)3 * g->cap);
    if (e == cudaSuccess) e = dalloc(&g->pos64, (size_t)3 * g->cap);
    if (e == cudaSuccess) e = dalloc(&g->hit, (size_t)g->cap);
    if (e == cudaSuccess) e = dalloc(&g->seg, (size_t)g->cap);
    if (e == cudaSuccess) e = dalloc(&g->first, (size_t)g->cap);
    if (e != cudaSuccess) { sage_grid_destroy(g); return sage_fail("sage_grid_create: allocation/upload", __FILE__, __LINE__, e); }
    *out = g;
    return 0;
}

int sage_grid_destroy(sage_grid* g)
{
    if (!g) return 0;
    cudaSetDevice(g->device);
    cudaFree(g->map); cudaFree(g->pos32); cudaFree(g->pos64); cudaFree(g->hit); cudaFree(g->seg); cudaFree(g->first);
    cudaGetLastError();
    delete g;
    return 0;
}

int sage_collide_grid_dev(sage_grid* g, const float* pos_dev, int P, int np2_promote,
                          uint8_t* hit_dev, sage_stream_t stream)
{
    REQUIRE(g != nullptr, "sage_collide_grid_dev: NULL grid");
    REQUIRE(P >= 0 && (P == 0 || (pos_dev && hit_dev)), "sage_collide_grid_dev: bad arguments");
    CK(cudaSetDevice(g->device));
    sage_launch_collide_grid(g->map, g->H, g->W, g->min_x, g->max_x, g->min_y, g->max_y, g->scale,
                             pos_dev, nullptr, P, np2_promote, hit_dev, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return 0;
}

int sage_collide_grid(sage_grid* g, const float* pos_host, int P, int np2_promote,
                      uint8_t* hit_host, const int32_t* seg_len_host, int n_segments,
                      int32_t* first_hit_host, sage_stream_t stream)
{
    REQUIRE(g != nullptr, "sage_collide_grid: NULL grid");
    REQUIRE(P >= 0 && P <= g->cap, "sage_collide_grid: P exceeds the per-call limit (1024)");
    REQUIRE(P == 0 || (pos_host && hit_host), "sage_collide_grid: NULL position/hit buffer");
    REQUIRE(n_segments >= 0 && n_segments <= g->cap, "sage_collide_grid: bad segment count");
    REQUIRE(n_segments == 0 || (seg_len_host && first_hit_host), "sage_collide_grid: NULL segment buffers");
    if (n_segments > 0) {
        long tot = 0;
        for (int k = 0; k < n_segments; ++k) { REQUIRE(seg_len_host[k] >= 0, "sage_collide_grid: negative segment"); tot += seg_len_host[k]; }
        REQUIRE(tot == P, "sage_collide_grid: segment lengths must sum to P");
    }
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(g->device));
    CK(cudaMemcpyAsync(g->pos32, pos_host, sizeof(float) * 3 * P, cudaMemcpyHostToDevice, st));
    sage_launch_collide_grid(g->map, g->H, g->W, g->min_x, g->max_x, g->min_y, g->max_y, g->scale,
                             g->pos32, nullptr, P, np2_promote, g->hit, st);
    if (n_segments > 0) {
        CK(cudaMemcpyAsync(g->seg, seg_len_host, sizeof(int32_t) * n_segments, cudaMemcpyHostToDevice, st));
        sage_launch_first_hit(g->hit, g->seg, n_segments, g->first, st);
        CK(cudaMemcpyAsync(first_hit_host, g->first, sizeof(int32_t) * n_segments, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hit_host, g->hit, (size_t)P, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int sage_collide_grid_f64(sage_grid* g, const double* pos_host, int P, uint8_t* hit_host,
                          sage_stream_t stream)
{
    REQUIRE(g != nullptr, "sage_collide_grid_f64: NULL grid");
    REQUIRE(P >= 0 && P <= g->cap, "sage_collide_grid_f64: P exceeds the per-call limit (1024)");
    REQUIRE(P == 0 || (pos_host && hit_host), "sage_collide_grid_f64: NULL buffer");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(g->device));
    CK(cudaMemcpyAsync(g->pos64, pos_host, sizeof(double) * 3 * P, cudaMemcpyHostToDevice, st));
    sage_launch_collide_grid(g->map, g->H, g->W, g->min_x, g->max_x, g->min_y, g->max_y, g->scale,
                             nullptr, g->pos64, P, 0, g->hit, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hit_host, g->hit, (size_t)P, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ---- capsule -----------------------------------------------------------------------------

int sage_bodies_create(sage_bodies** out, int device, int64_t m, const float* mu,
                       const float* log_scale, const float* quat_wxyz)
{
    REQUIRE(out != nullptr, "sage_bodies_create: out is NULL");
    *out = nullptr;
    REQUIRE(sage_device_count() > 0, "sage_bodies_create: no CUDA device (this library has no CPU fallback)");
    REQUIRE(m >= 0 && m < (int64_t)0x7FFFFFFF, "sage_bodies_create: m out of range");
    REQUIRE(m == 0 || (mu && log_scale && quat_wxyz), "sage_bodies_create: NULL attribute pointer");
    CK(cudaSetDevice(device));
    sage_bodies* b = new sage_bodies();
    b->device = device; b->m = m; b->maxp = sage_capsule_max_poses();
    cudaError_t e = dalloc(&b->c0, (size_t)m);
    if (e == cudaSuccess) e = dalloc(&b->c1, (size_t)m);
    if (e == cudaSuccess) e = dalloc(&b->c2, (size_t)m);
    if (e == cudaSuccess) e = dalloc(&b->pos, (size_t)3 * b->maxp);
    if (e == cudaSuccess) e = dalloc(&b->hit_bits, (size_t)(b->maxp / 32));
    if (e == cudaSuccess) e = dalloc(&b->first, (size_t)b->maxp);
    float *d_mu = nullptr, *d_ls = nullptr, *d_q = nullptr;
    if (e == cudaSuccess && m > 0) {
        e = dalloc(&d_mu, (size_t)3 * m);
        if (e == cudaSuccess) e = dalloc(&d_ls, (size_t)3 * m);
        if (e == cudaSuccess) e = dalloc(&d_q, (size_t)4 * m);
        if (e == cudaSuccess) e = cudaMemcpy(d_mu, mu, sizeof(float) * 3 * m, cudaMemcpyDefault);
        if (e == cudaSuccess) e = cudaMemcpy(d_ls, log_scale, sizeof(float) * 3 * m, cudaMemcpyDefault);
        if (e == cudaSuccess) e = cudaMemcpy(d_q, quat_wxyz, sizeof(float) * 4 * m, cudaMemcpyDefault);
        if (e == cudaSuccess) {
            sage_launch_bodies_pack(m, d_mu, d_ls, d_q, b->c0, b->c1, b->c2, nullptr);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
        }
    }
    cudaFree(d_mu); cudaFree(d_ls); cudaFree(d_q);
    if (e != cudaSuccess) { sage_bodies_destroy(b); return sage_fail("sage_bodies_create: allocation/upload", __FILE__, __LINE__, e); }
    *out = b;
    return 0;
}

int sage_bodies_destroy(sage_bodies* b)
{
    if (!b) return 0;
    cudaSetDevice(b->device);
    cudaFree(b->c0); cudaFree(b->c1); cudaFree(b->c2); cudaFree(b->pos); cudaFree(b->hit_bits); cudaFree(b->first);
    cudaGetLastError();
    delete b;
    return 0;
}

int sage_collide_capsule(sage_bodies* b, const float* pos_host, int P, float radius, float z0,
                         float z1, uint8_t* hit_host, int32_t* first_body_host, sage_stream_t stream)
{
    REQUIRE(b != nullptr, "sage_collide_capsule: NULL bodies");
    REQUIRE(P >= 0 && P <= b->maxp, "sage_collide_capsule: P exceeds the per-call limit (256)");
    REQUIRE(P == 0 || (pos_host && hit_host), "sage_collide_capsule: NULL buffer");
    REQUIRE(radius >= 0.0f, "sage_collide_capsule: negative radius");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpyAsync(b->pos, pos_host, sizeof(float) * 3 * P, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(b->hit_bits, 0, sizeof(uint32_t) * (b->maxp / 32), st));
    CK(cudaMemsetAsync(b->first, 0x7F, sizeof(int32_t) * b->maxp, st));   // 0x7F7F7F7F: > any body index
    sage_launch_collide_capsule(b->m, b->c0, b->c1, b->c2, b->pos, P, radius, z0, z1, b->hit_bits,
                                b->first, st);
    CK(cudaGetLastError());
    uint32_t bits[8];
    std::vector<int32_t> first((size_t)P);
    CK(cudaMemcpyAsync(bits, b->hit_bits, sizeof(uint32_t) * (b->maxp / 32), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(first.data(), b->first, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int p = 0; p < P; ++p) {
        hit_host[p] = (bits[p >> 5] >> (p & 31)) & 1u;
        if (first_body_host) first_body_host[p] = hit_host[p] ? first[p] : -1;
    }
    return 0;
}

}  // extern "C"
