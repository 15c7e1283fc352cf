// Map building and path planning on the obstacle grid — the callers either side
// of the collide path: the same inflated occupancy grid that check_collision_3d looks up is what
// the reference's trajectory generator plans on, thousands of A* queries per scene in Python.
//
//   K8  k_grid_inflate     robot-radius inflation,  dist_edt(grid == 0, sampling = s) <= r
//                          (collision_detector.py:119-127; vln_trajectory_generator.py:549-552)
//   K9  k_cc_init/k_cc_link/k_cc_flatten   4-connected components of the free cells (lock-free
//                          union-find, larger root linked under smaller): the partition
//                          build_connectivity_map computes (vln_trajectory_generator.py:639-666)
//   K10 k_astar            astar_pixel (vln_trajectory_generator.py:253-286), one query per thread,
//                          statement for statement: same neighbour order, no closed set, heap
//                          entries ordered by (f, x, y), binary64 g/f with
//                          hypot(dx, dy) = sqrt((double)(dx*dx + dy*dy)) — so the returned PATH
//                          (not just its cost) is the reference's.  The search is inherently
//                          sequential; the GPU runs thousands of them side by side.
// All integer / binary64 work: results are bit-exact against oracle/plan_oracle.c, which is pinned
// by golden vectors recorded from the unmodified reference functions.
#include <cmath>
#include <cstring>
#include <vector>

#include "api_common.cuh"

namespace {

// ---- K8 ------------------------------------------------------------------------------------------
// scipy evaluates the distance to the nearest obstacle cell as sqrt((dy*s)*(dy*s) + (dx*s)*(dx*s))
// in binary64; "<= r" holds for a cell iff it holds for some obstacle cell of the window.
__global__ void __launch_bounds__(256) k_grid_inflate(const uint8_t* __restrict__ raw, int H, int W, double s, double r,
                                                      int R, uint8_t* __restrict__ out)
{
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    int hit = 0;
    for (int dy = -R; dy <= R && !hit; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const double a = (double)dy * s;
        const double a2 = a * a;
        for (int dx = -R; dx <= R; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W || !raw[(size_t)yy * W + xx]) continue;
            const double b = (double)dx * s;
            if (sqrt(a2 + b * b) <= r) { hit = 1; break; }
        }
    }
    out[(size_t)y * W + x] = (uint8_t)hit;
}

// ---- K9 ------------------------------------------------------------------------------------------
__device__ __forceinline__ int cc_load(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__device__ __forceinline__ int cc_find(int* parent, int x)
{
    while (true) {
        const int px = cc_load(parent + x);
        if (px == x) return x;
        const int ppx = cc_load(parent + px);
        if (ppx != px) *reinterpret_cast<volatile int*>(parent + x) = ppx;   // path halving (x is not a root: benign)
        x = px;
    }
}

__device__ __forceinline__ void cc_unite(int* parent, int a, int b)
{
    while (true) {
        a = cc_find(parent, a);
        b = cc_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }         // link the larger root under the smaller
        if (atomicCAS(parent + a, a, b) == a) return;          // a stopped being a root meanwhile: retry
    }
}

__global__ void __launch_bounds__(256) k_cc_init(const uint8_t* __restrict__ grid, int n, int* __restrict__ parent)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) parent[i] = grid[i] == 0 ? i : -1;
}

__global__ void __launch_bounds__(256) k_cc_link(const uint8_t* __restrict__ grid, int H, int W, int* __restrict__ parent)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= H * W || grid[i] != 0) return;
    const int x = i % W, y = i / W;
    if (x + 1 < W && grid[i + 1] == 0) cc_unite(parent, i, i + 1);
    if (y + 1 < H && grid[i + W] == 0) cc_unite(parent, i, i + W);
}

__global__ void __launch_bounds__(256) k_cc_flatten(int n, int* __restrict__ parent, int* __restrict__ labels)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int p = cc_load(parent + i);
    labels[i] = p < 0 ? -1 : cc_find(parent, i);               // root = smallest index of the component
}

// ---- K10 -----------------------------------------------------------------------------------------
struct __align__(16) HeapEnt {
    double f;
    int x, y;
};

__device__ __forceinline__ bool heap_less(const HeapEnt& a, const HeapEnt& b)
{
    if (a.f < b.f) return true;
    if (a.f > b.f) return false;
    if (a.x != b.x) return a.x < b.x;
    return a.y < b.y;
}

// "not in g_score": the slot's g array is memset to 0x7F bytes = 1.38e306, above any real cost
#define ASTAR_G_UNSET_BYTE 0x7F

// One query per thread.  Slot t of the wave owns g[t], came[t], heap[t].
// path_len: >= 1 points written to path_xy (x,y pairs, start -> goal); 0: no path (the reference's
// None); -2: heap capacity exceeded; -3: path longer than max_path.
__global__ void __launch_bounds__(32) k_astar(const uint8_t* __restrict__ grid, int H, int W, const int* __restrict__ starts,
                                               const int* __restrict__ goals, int q0, int nq, double* __restrict__ g_all,
                                               int* __restrict__ came_all, HeapEnt* __restrict__ heap_all, long long heap_cap,
                                               int* __restrict__ path_len, int* __restrict__ path_xy, int max_path,
                                               long long* __restrict__ expansions)
{
    const int t = blockIdx.x * 32 + threadIdx.x;
    if (t >= nq) return;
    const int q = q0 + t;
    const size_t n = (size_t)H * W;
    double* g = g_all + (size_t)t * n;
    int* came = came_all + (size_t)t * n;
    HeapEnt* heap = heap_all + (size_t)t * (size_t)heap_cap;
    const int sx = starts[2 * q], sy = starts[2 * q + 1], gx = goals[2 * q], gy = goals[2 * q + 1];
    int* out = path_xy + (size_t)q * max_path * 2;
    long long nexp = 0;
    int result = 0;
    long long hn = 0;
    const int ddx[8] = {-1, 1, 0, 0, -1, -1, 1, 1}, ddy[8] = {0, 0, -1, 1, -1, 1, -1, 1};
    if (sx >= 0 && sx < W && sy >= 0 && sy < H) {
        g[(size_t)sy * W + sx] = 0.0;
        HeapEnt e0; e0.f = 0.0; e0.x = sx; e0.y = sy;
        heap[hn++] = e0;
    }
    while (hn > 0) {
        const HeapEnt cur = heap[0];
        const HeapEnt last = heap[--hn];
        long long i = 0;
        while (true) {                                         // sift down
            long long c = 2 * i + 1;
            if (c >= hn) break;
            HeapEnt ec = heap[c];
            if (c + 1 < hn) {
                const HeapEnt er = heap[c + 1];
                if (heap_less(er, ec)) { ec = er; ++c; }
            }
            if (!heap_less(ec, last)) break;
            heap[i] = ec;
            i = c;
        }
        if (hn > 0) heap[i] = last;
        ++nexp;
        const size_t ci = (size_t)cur.y * W + cur.x;
        if (cur.x == gx && cur.y == gy) {
            long long len = 1;
            size_t c = ci;
            while (came[c] >= 0) { c = (size_t)came[c]; ++len; }
            if (len > max_path) { result = -3; break; }
            c = ci;
            for (long long k = len - 1; k >= 0; --k) {
                out[2 * k] = (int)(c % W);
                out[2 * k + 1] = (int)(c / W);
                if (k > 0) c = (size_t)came[c];
            }
            result = (int)len;
            break;
        }
        const double gc = g[ci];
        bool overflow = false;
#pragma unroll 1
        for (int d = 0; d < 8; ++d) {
            const int nx = cur.x + ddx[d], ny = cur.y + ddy[d];
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            const size_t j = (size_t)ny * W + nx;
            if (grid[j] == 1) continue;
            const double step = sqrt((double)(ddx[d] * ddx[d] + ddy[d] * ddy[d]));
            const double tg = gc + step;
            if (tg < g[j]) {
                came[j] = (int)ci;
                g[j] = tg;
                const long long hx = nx - gx, hy = ny - gy;
                HeapEnt e;
                e.f = tg + sqrt((double)(hx * hx + hy * hy));
                e.x = nx; e.y = ny;
                if (hn >= heap_cap) { overflow = true; break; }
                long long k = hn++;
                while (k > 0) {                                // sift up
                    const long long p = (k - 1) / 2;
                    const HeapEnt ep = heap[p];
                    if (!heap_less(e, ep)) break;
                    heap[k] = ep;
                    k = p;
                }
                heap[k] = e;
            }
        }
        if (overflow) { result = -2; break; }
    }
    path_len[q] = result;
    if (expansions) expansions[q] = nexp;
}

}  // namespace

struct sage_planner {
    int device = 0;
    int H = 0, W = 0;
    int slots = 0, max_path = 0;
    long long heap_cap = 0;
    uint8_t* grid = nullptr;
    double* g = nullptr;
    int* came = nullptr;
    HeapEnt* heap = nullptr;
    // query staging (grown on demand)
    int qcap = 0;
    int *starts = nullptr, *goals = nullptr, *path_len = nullptr, *path_xy = nullptr;
    long long* expansions = nullptr;
};

extern "C" {

int sage_grid_inflate(int device, const uint8_t* raw_host, int H, int W, double scale, double radius,
                      uint8_t* out_host, sage_stream_t stream)
{
    REQUIRE(raw_host != nullptr && out_host != nullptr, "sage_grid_inflate: NULL argument");
    REQUIRE(H > 0 && W > 0 && (int64_t)H * W < ((int64_t)1 << 30), "sage_grid_inflate: bad grid size");
    REQUIRE(scale > 0.0, "sage_grid_inflate: scale must be positive");
    REQUIRE(sage_device_count() > 0, "sage_grid_inflate: no CUDA device (this library has no CPU fallback)");
    const size_t n = (size_t)H * W;
    if (!(radius > 0.0)) {                       // the reference inflates only for robot_radius_m > 0
        for (size_t i = 0; i < n; ++i) out_host[i] = raw_host[i] != 0;
        return 0;
    }
    const double Rf = floor(radius / scale) + 1.0;
    REQUIRE(Rf < 4096.0, "sage_grid_inflate: radius / scale too large");
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *d_raw = nullptr, *d_out = nullptr;
    CK(dalloc(&d_raw, n));
    CK(dalloc(&d_out, n));
    CK(cudaMemcpyAsync(d_raw, raw_host, n, cudaMemcpyHostToDevice, st));
    SAGE_KLAUNCH(k_grid_inflate, dim3((W + 15) / 16, (H + 15) / 16), 256, st, d_raw, H, W, scale, radius, (int)Rf, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, d_out, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_raw);
    cudaFree(d_out);
    return 0;
}

int sage_grid_components(int device, const uint8_t* grid_host, int H, int W, int32_t* labels_host, sage_stream_t stream)
{
    REQUIRE(grid_host != nullptr && labels_host != nullptr, "sage_grid_components: NULL argument");
    REQUIRE(H > 0 && W > 0 && (int64_t)H * W < ((int64_t)1 << 30), "sage_grid_components: bad grid size");
    REQUIRE(sage_device_count() > 0, "sage_grid_components: no CUDA device (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = H * W;
    uint8_t* d_grid = nullptr;
    int *d_parent = nullptr, *d_labels = nullptr;
    CK(dalloc(&d_grid, (size_t)n));
    CK(dalloc(&d_parent, (size_t)n));
    CK(dalloc(&d_labels, (size_t)n));
    CK(cudaMemcpyAsync(d_grid, grid_host, (size_t)n, cudaMemcpyHostToDevice, st));
    const int blocks = (n + 255) / 256;
    SAGE_KLAUNCH(k_cc_init, blocks, 256, st, d_grid, n, d_parent);
    SAGE_KLAUNCH(k_cc_link, blocks, 256, st, d_grid, H, W, d_parent);
    SAGE_KLAUNCH(k_cc_flatten, blocks, 256, st, n, d_parent, d_labels);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(labels_host, d_labels, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_grid); cudaFree(d_parent); cudaFree(d_labels);
    return 0;
}

int sage_planner_destroy(sage_planner* p)
{
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    cudaFree(p->grid); cudaFree(p->g); cudaFree(p->came); cudaFree(p->heap);
    cudaFree(p->starts); cudaFree(p->goals); cudaFree(p->path_len); cudaFree(p->path_xy); cudaFree(p->expansions);
    cudaGetLastError();
    delete p;
    return 0;
}

int sage_planner_create(sage_planner** out, int device, const uint8_t* grid_host, int H, int W, int slots,
                        int max_path, int64_t heap_capacity)
{
    REQUIRE(out != nullptr, "sage_planner_create: out is NULL");
    *out = nullptr;
    REQUIRE(sage_device_count() > 0, "sage_planner_create: no CUDA device (this library has no CPU fallback)");
    REQUIRE(grid_host != nullptr && H > 0 && W > 0 && (int64_t)H * W < ((int64_t)1 << 28), "sage_planner_create: bad grid");
    REQUIRE(slots >= 1 && slots <= 65536, "sage_planner_create: slots must be 1..65536");
    CK(cudaSetDevice(device));
    sage_planner* p = new sage_planner();
    p->device = device; p->H = H; p->W = W; p->slots = slots;
    const size_t n = (size_t)H * W;
    p->max_path = max_path > 0 ? max_path : 4 * (H + W);
    p->heap_cap = heap_capacity > 0 ? heap_capacity : (long long)(n / 2 + 1024);
    auto bail = [&](int rc) { sage_planner_destroy(p); return rc; };
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(sage_fail(#call, __FILE__, __LINE__, e__)); } while (0)
    CKB(dalloc(&p->grid, n));
    CKB(cudaMemcpy(p->grid, grid_host, n, cudaMemcpyHostToDevice));
    CKB(dalloc(&p->g, (size_t)slots * n));
    CKB(dalloc(&p->came, (size_t)slots * n));
    CKB(dalloc(&p->heap, (size_t)slots * (size_t)p->heap_cap));
#undef CKB
    *out = p;
    return 0;
}

int sage_plan_astar(sage_planner* p, const int32_t* starts_xy_host, const int32_t* goals_xy_host, int Q,
                    int32_t* path_len_host, int32_t* path_xy_host, int64_t* expansions_host, sage_stream_t stream)
{
    REQUIRE(p != nullptr && starts_xy_host && goals_xy_host && path_len_host && path_xy_host, "sage_plan_astar: NULL argument");
    REQUIRE(Q >= 0, "sage_plan_astar: negative query count");
    if (Q == 0) return 0;
    CK(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (Q > p->qcap) {
        CK(cudaStreamSynchronize(st));
        cudaFree(p->starts); cudaFree(p->goals); cudaFree(p->path_len); cudaFree(p->path_xy); cudaFree(p->expansions);
        p->starts = p->goals = p->path_len = p->path_xy = nullptr; p->expansions = nullptr; p->qcap = 0;
        CK(dalloc(&p->starts, (size_t)Q * 2));
        CK(dalloc(&p->goals, (size_t)Q * 2));
        CK(dalloc(&p->path_len, (size_t)Q));
        CK(dalloc(&p->path_xy, (size_t)Q * p->max_path * 2));
        CK(dalloc(&p->expansions, (size_t)Q));
        p->qcap = Q;
    }
    CK(cudaMemcpyAsync(p->starts, starts_xy_host, sizeof(int) * 2 * (size_t)Q, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(p->goals, goals_xy_host, sizeof(int) * 2 * (size_t)Q, cudaMemcpyHostToDevice, st));
    const size_t n = (size_t)p->H * p->W;
    for (int q0 = 0; q0 < Q; q0 += p->slots) {            // waves of `slots` concurrent searches
        const int nq = Q - q0 < p->slots ? Q - q0 : p->slots;
        CK(cudaMemsetAsync(p->g, ASTAR_G_UNSET_BYTE, sizeof(double) * n * (size_t)nq, st));
        CK(cudaMemsetAsync(p->came, 0xFF, sizeof(int) * n * (size_t)nq, st));
        SAGE_KLAUNCH(k_astar, (nq + 31) / 32, 32, st, p->grid, p->H, p->W, p->starts, p->goals, q0, nq, p->g, p->came, p->heap,
                     p->heap_cap, p->path_len, p->path_xy, p->max_path, p->expansions);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(path_len_host, p->path_len, sizeof(int) * (size_t)Q, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(path_xy_host, p->path_xy, sizeof(int) * 2 * (size_t)Q * p->max_path, cudaMemcpyDeviceToHost, st));
    if (expansions_host)
        CK(cudaMemcpyAsync(expansions_host, p->expansions, sizeof(long long) * (size_t)Q, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int sage_planner_limits(sage_planner* p, int* max_path_out, int64_t* heap_capacity_out)
{
    REQUIRE(p != nullptr, "sage_planner_limits: NULL planner");
    if (max_path_out) *max_path_out = p->max_path;
    if (heap_capacity_out) *heap_capacity_out = p->heap_cap;
    return 0;
}

}  // extern "C"
