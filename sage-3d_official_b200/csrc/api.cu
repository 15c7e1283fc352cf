// C-ABI of libsage_splat.so (declared in include/sage_splat.h).  Owns device memory, orders the
// kernels of a frame on the caller's stream, never synchronises inside sage_render.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sage_internal.cuh"

namespace {

thread_local std::string g_err;

int fail(const char* what, const char* file, int line, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s (%s:%d)", what, file, line);
    g_err = buf;
    return 1;
}

#define CK(call)                                                      \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return fail(#call, __FILE__, __LINE__, e__); \
    } while (0)
#define REQUIRE(cond, msg)                                            \
    do {                                                              \
        if (!(cond)) return fail(msg, __FILE__, __LINE__);            \
    } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t count)
{
    *p = nullptr;
    return cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
}


}  // namespace

struct sage_scene {
    int device = 0;
    SageSceneDev sc{};
    SageFrame fr{};
    int64_t cap = 0;           // instance capacity
    int max_tiles = 0;         // ranges capacity
    uint32_t* n_dev = nullptr; // device copy of n (element count for the depth sort)
    uint32_t* counters_host = nullptr;  // pinned
    uint8_t* frame_rgb = nullptr;       // internal frame (sage_render_host)
    float* frame_depth = nullptr;
    size_t frame_px = 0;
    cudaStream_t last_stream = nullptr;
    int last_W = 0, last_H = 0, last_depth = 0, last_launches = 0;
    bool rendered = false;
    uint32_t epoch = 8;        // look-back epochs: 8 fresh values per frame
    // optional per-stage timing: n_slots x (SAGE_N_STAGES + 1) events
    std::vector<cudaEvent_t> events;
    int n_slots = 0;
    long frame_index = 0;
};

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

static void free_instances(sage_scene* s)
{
    for (int k = 0; k < 2; ++k) {
        cudaFree(s->fr.ikey[k]); s->fr.ikey[k] = nullptr;
        cudaFree(s->fr.ival[k]); s->fr.ival[k] = nullptr;
    }
    cudaFree(s->fr.lookback); s->fr.lookback = nullptr;
}

static int alloc_instances(sage_scene* s, int64_t cap)
{
    free_instances(s);
    for (int k = 0; k < 2; ++k) {
        CK(dalloc(&s->fr.ikey[k], (size_t)cap));
        CK(dalloc(&s->fr.ival[k], (size_t)cap));
    }
    const int64_t mx = cap > s->sc.n ? cap : s->sc.n;
    const size_t words = sage_sort_lookback_words(mx);
    CK(dalloc(&s->fr.lookback, words));
    CK(cudaMemset(s->fr.lookback, 0, words * sizeof(unsigned long long)));   // epoch 0 = never published
    s->cap = cap;
    return 0;
}

extern "C" {

const char* sage_last_error(void) { return g_err.c_str(); }
int sage_abi_version(void) { return SAGE_ABI_VERSION; }

int sage_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sage_scene_create(sage_scene** out, int device, int64_t n, int sh_degree, const float* xyz,
                      const float* log_scale, const float* quat_wxyz, const float* opacity_raw,
                      const float* sh, int64_t max_instances)
{
    REQUIRE(out != nullptr, "sage_scene_create: out is NULL");
    *out = nullptr;
    REQUIRE(sage_device_count() > 0, "sage_scene_create: no CUDA device (this library has no CPU fallback)");
    REQUIRE(n >= 0 && n <= (int64_t)SAGE_ID_MASK, "sage_scene_create: n out of range (max 2^28-1 Gaussians)");
    REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sage_scene_create: sh_degree must be 0..3");
    REQUIRE(n == 0 || (xyz && log_scale && quat_wxyz && opacity_raw && sh), "sage_scene_create: NULL attribute pointer");
    CK(cudaSetDevice(device));
    sage_scene* s = new sage_scene();
    s->device = device;
    s->sc.n = n;
    s->sc.deg = sh_degree;
    const int nf = 3 * (sh_degree + 1) * (sh_degree + 1);
    s->sc.n_chunks = (nf + 3) / 4;
    auto bail = [&](int rc) { sage_scene_destroy(s); return rc; };
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(#call, __FILE__, __LINE__, e__)); } while (0)
    CKB(dalloc(&s->sc.geo0, (size_t)n));
    CKB(dalloc(&s->sc.geo1, (size_t)n));
    CKB(dalloc(&s->sc.geo2, (size_t)n));
    CKB(dalloc(&s->sc.shc, (size_t)n * s->sc.n_chunks));
    CKB(dalloc(&s->fr.rec0, (size_t)n));
    CKB(dalloc(&s->fr.rec1, (size_t)n));
    CKB(dalloc(&s->fr.rec2, (size_t)n));
    for (int k = 0; k < 2; ++k) {
        CKB(dalloc(&s->fr.dkey[k], (size_t)n));
        CKB(dalloc(&s->fr.dval[k], (size_t)n));
    }
    CKB(dalloc(&s->fr.chain, (size_t)(n / 256 + 2)));
    CKB(cudaMemset(s->fr.chain, 0, (size_t)(n / 256 + 2) * sizeof(unsigned long long)));
    CKB(dalloc(&s->fr.rect, (size_t)n));
    CKB(dalloc(&s->fr.ntiles, (size_t)n));
    CKB(dalloc(&s->fr.counters, (size_t)SAGE_ZERO_WORDS));
    s->fr.tickets = s->fr.counters + CNT_COUNT;
    s->fr.depth_hist = s->fr.tickets + SAGE_TICKET_COUNT;
    s->fr.tile_hist = s->fr.depth_hist + SAGE_MAX_PASSES * SAGE_RADIX_MAX_BINS;
    CKB(dalloc(&s->n_dev, 1));
    CKB(cudaHostAlloc((void**)&s->counters_host, sizeof(uint32_t) * CNT_COUNT, cudaHostAllocDefault));
    memset(s->counters_host, 0, sizeof(uint32_t) * CNT_COUNT);
    uint32_t n32 = (uint32_t)n;
    CKB(cudaMemcpy(s->n_dev, &n32, sizeof n32, cudaMemcpyHostToDevice));
    int64_t cap = max_instances > 0 ? max_instances : (8 * n > (1 << 24) ? 8 * n : (1 << 24));
    if (alloc_instances(s, cap)) return bail(1);

    if (n > 0) {
        // inputs may be host or device pointers: stage through device copies (UVA resolves the kind)
        float *d_xyz, *d_ls, *d_q, *d_op, *d_sh;
        CKB(dalloc(&d_xyz, (size_t)3 * n));
        CKB(dalloc(&d_ls, (size_t)3 * n));
        CKB(dalloc(&d_q, (size_t)4 * n));
        CKB(dalloc(&d_op, (size_t)n));
        CKB(dalloc(&d_sh, (size_t)nf * n));
        CKB(cudaMemcpy(d_xyz, xyz, sizeof(float) * 3 * n, cudaMemcpyDefault));
        CKB(cudaMemcpy(d_ls, log_scale, sizeof(float) * 3 * n, cudaMemcpyDefault));
        CKB(cudaMemcpy(d_q, quat_wxyz, sizeof(float) * 4 * n, cudaMemcpyDefault));
        CKB(cudaMemcpy(d_op, opacity_raw, sizeof(float) * n, cudaMemcpyDefault));
        CKB(cudaMemcpy(d_sh, sh, sizeof(float) * (size_t)nf * n, cudaMemcpyDefault));
        sage_launch_precompute(s->sc, d_xyz, d_ls, d_q, d_op, d_sh, nullptr);
        CKB(cudaGetLastError());
        CKB(cudaDeviceSynchronize());
        cudaFree(d_xyz); cudaFree(d_ls); cudaFree(d_q); cudaFree(d_op); cudaFree(d_sh);
    }
#undef CKB
    *out = s;
    return 0;
}

int sage_scene_reserve(sage_scene* s, int64_t max_instances)
{
    REQUIRE(s != nullptr, "sage_scene_reserve: NULL scene");
    REQUIRE(max_instances > 0 && max_instances < (int64_t)0xFFFFFFFF, "sage_scene_reserve: bad capacity");
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    return alloc_instances(s, max_instances);
}

int sage_scene_destroy(sage_scene* s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    cudaFree(s->sc.geo0); cudaFree(s->sc.geo1); cudaFree(s->sc.geo2); cudaFree(s->sc.shc);
    cudaFree(s->fr.rec0); cudaFree(s->fr.rec1); cudaFree(s->fr.rec2);
    for (int k = 0; k < 2; ++k) { cudaFree(s->fr.dkey[k]); cudaFree(s->fr.dval[k]); }
    cudaFree(s->fr.chain);
    cudaFree(s->fr.rect); cudaFree(s->fr.ntiles);
    free_instances(s);
    cudaFree(s->fr.ranges); cudaFree(s->fr.counters);
    cudaFree(s->n_dev);
    if (s->counters_host) cudaFreeHost(s->counters_host);
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    cudaFree(s->frame_rgb); cudaFree(s->frame_depth);
    cudaGetLastError();
    delete s;
    return 0;
}

static int make_cam(const sage_camera* cam, int row0, int row1, SageCamDev* d)
{
    REQUIRE(cam != nullptr, "camera is NULL");
    REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 16384 && cam->height <= 16384, "camera: bad resolution");
    REQUIRE(cam->fx > 0.0f && cam->fy > 0.0f, "camera: focal length must be positive");
    REQUIRE(cam->near_z > 0.0f && cam->far_z > cam->near_z, "camera: need 0 < near < far");
    memcpy(d->R, cam->R, sizeof d->R);
    memcpy(d->t, cam->t, sizeof d->t);
    memcpy(d->campos, cam->campos, sizeof d->campos);
    d->fx = cam->fx; d->fy = cam->fy; d->cx = cam->cx; d->cy = cam->cy;
    d->limx = 1.3f * ((0.5f * (float)cam->width) / cam->fx);
    d->limy = 1.3f * ((0.5f * (float)cam->height) / cam->fy);
    d->near_z = cam->near_z; d->far_z = cam->far_z;
    d->W = cam->width; d->H = cam->height;
    d->tiles_x = (cam->width + SAGE_TILE - 1) / SAGE_TILE;
    d->tiles_y = (cam->height + SAGE_TILE - 1) / SAGE_TILE;
    REQUIRE(row0 >= 0 && row0 <= row1 && row1 <= d->tiles_y, "tile-row range out of bounds");
    d->row0 = row0; d->row1 = row1;
    return 0;
}

// resolution-dependent scratch (tile ranges) grows on first use
static int ensure_tiles(sage_scene* s, int n_tiles, cudaStream_t st)
{
    if (n_tiles > s->max_tiles) {
        CK(cudaStreamSynchronize(st));
        cudaFree(s->fr.ranges);
        s->fr.ranges = nullptr; s->max_tiles = 0;
        CK(dalloc(&s->fr.ranges, (size_t)n_tiles));
        s->max_tiles = n_tiles;
    }
    return 0;
}

static int depth_key_bits(const SageCamDev& cd)
{
    // depth keys are bits(z) - bits(near) with near < z < far: only this many bits are significant
    uint32_t bn, bf;
    memcpy(&bn, &cd.near_z, 4);
    memcpy(&bf, &cd.far_z, 4);
    int dbits = 1;
    while (dbits < 32 && ((uint64_t)1 << dbits) <= (uint64_t)(bf - bn)) ++dbits;
    return dbits;
}

// project -> depth sort (+compaction) -> SH colour; *order_buf = buffer index holding order[]
static int front_half(sage_scene* s, const SageCamDev& cd, int* launches, cudaEvent_t* ev, cudaStream_t st, int* order_buf)
{
    if (ev) CK(cudaEventRecord(ev[0], st));
    sage_launch_project(s->sc, cd, s->fr, st);
    if (ev) CK(cudaEventRecord(ev[1], st));
    SageSortPlan dp{};
    dp.keys[0] = s->fr.dkey[0]; dp.keys[1] = s->fr.dkey[1];
    dp.vals[0] = s->fr.dval[0]; dp.vals[1] = s->fr.dval[1];
    dp.n_first = s->n_dev; dp.n_rest = s->fr.counters + CNT_VISIBLE;
    dp.capacity = s->sc.n; dp.lo_bit = 0; dp.n_bits = depth_key_bits(cd);
    sage_sort_shape(dp.n_bits, &dp.digit_bits, &dp.passes);
    dp.vals_are_iota = true; dp.compact_first = true; dp.hist_ready = false;
    dp.hist = s->fr.depth_hist; dp.total_out = nullptr;
    dp.lookback = s->fr.lookback; dp.tickets = s->fr.tickets + SAGE_TICKET_DEPTH; dp.epoch0 = s->epoch;
    const int res = sage_radix_sort(dp, launches, st);
    REQUIRE(res >= 0, "depth sort: unsupported key width");
    if (ev) CK(cudaEventRecord(ev[2], st));
    sage_launch_sh_color(s->sc, cd, s->fr, s->fr.dval[res], st);
    if (ev) CK(cudaEventRecord(ev[3], st));
    *launches += 2;
    *order_buf = res;
    return 0;
}

static int render_impl(sage_scene* s, const sage_camera* cam, int row0, int row1, int n_peers,
                       uint8_t* const* rgb_peers, float* const* depth_peers, cudaStream_t st)
{
    REQUIRE(s != nullptr, "sage_render: NULL scene");
    REQUIRE(n_peers >= 1 && n_peers <= 8, "sage_render: n_peers must be 1..8");
    for (int k = 0; k < n_peers; ++k) REQUIRE(rgb_peers[k] != nullptr, "sage_render: NULL rgb buffer");
    SageCamDev cd;
    if (make_cam(cam, row0, row1, &cd)) return 1;
    CK(cudaSetDevice(s->device));
    const int n_tiles = (row1 - row0) * cd.tiles_x;
    if (ensure_tiles(s, n_tiles, st)) return 1;
    int tbits = 1;
    while ((1 << tbits) < n_tiles) ++tbits;
    REQUIRE(tbits <= 24, "sage_render: too many tiles");
    const int tpasses = (tbits + 7) / 8;
    int launches = 0;
    CK(cudaMemsetAsync(s->fr.counters, 0, sizeof(uint32_t) * SAGE_ZERO_WORDS, st));
    s->epoch += 8;
    if (s->epoch > (1u << 30) - 16) s->epoch = 8;
    cudaEvent_t* ev = s->n_slots > 0
        ? &s->events[(size_t)(s->frame_index % s->n_slots) * (SAGE_N_STAGES + 1)] : nullptr;
    s->frame_index++;
#define STAGE_MARK(k) do { if (ev) CK(cudaEventRecord(ev[k], st)); } while (0)
    if (n_tiles > 0 && s->sc.n > 0) {
        int ob = 0;
        if (front_half(s, cd, &launches, ev, st, &ob)) return 1;     // stages 0..2
        const uint32_t* order = s->fr.dval[ob];
        sage_launch_emit(cd, s->fr, order, s->sc.n, s->cap, tpasses, s->epoch + 3, st);
        STAGE_MARK(4);
        SageSortPlan tp{};
        tp.keys[0] = s->fr.ikey[0]; tp.keys[1] = s->fr.ikey[1];
        tp.vals[0] = s->fr.ival[0]; tp.vals[1] = s->fr.ival[1];
        tp.n_first = tp.n_rest = s->fr.counters + CNT_INSTANCES;
        tp.capacity = s->cap; tp.lo_bit = 0; tp.n_bits = tbits; tp.digit_bits = 8; tp.passes = tpasses;
        tp.vals_are_iota = false; tp.compact_first = false; tp.hist_ready = true;
        tp.hist = s->fr.tile_hist; tp.total_out = nullptr;
        tp.lookback = s->fr.lookback; tp.tickets = s->fr.tickets + SAGE_TICKET_TILE; tp.epoch0 = s->epoch + 4;
        const int ires = sage_radix_sort(tp, &launches, st);
        REQUIRE(ires >= 0, "tile sort: unsupported key width");
        sage_launch_ranges(s->fr, s->fr.ikey[ires], s->cap, n_tiles, st);
        STAGE_MARK(5);
        sage_launch_composite(cd, s->fr, s->fr.ival[ires], n_peers, rgb_peers, depth_peers, st);
        launches += 3;
        STAGE_MARK(6);
    } else if (n_tiles > 0) {
        for (int k = 0; k <= SAGE_N_STAGES; ++k) STAGE_MARK(k);
        CK(cudaMemsetAsync(s->fr.ranges, 0, sizeof(uint2) * (size_t)n_tiles, st));
        sage_launch_composite(cd, s->fr, s->fr.ival[0], n_peers, rgb_peers, depth_peers, st);
        launches += 1;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s->counters_host, s->fr.counters, sizeof(uint32_t) * CNT_COUNT,
                       cudaMemcpyDeviceToHost, st));
    s->last_stream = st;
    s->last_W = cd.W; s->last_H = cd.H;
    s->last_depth = (depth_peers && depth_peers[0]) ? 1 : 0;
    s->last_launches = launches;
    s->rendered = true;
    return 0;
}

int sage_render(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                uint8_t* rgb_dev, float* depth_dev, sage_stream_t stream)
{
    uint8_t* rgbs[1] = {rgb_dev};
    float* deps[1] = {depth_dev};
    return render_impl(s, cam, row0_tiles, row1_tiles, 1, rgbs, depth_dev ? deps : nullptr,
                       (cudaStream_t)stream);
}

int sage_render_peers(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                      int n_peers, uint8_t* const* rgb_peers, float* const* depth_peers,
                      sage_stream_t stream)
{
    REQUIRE(rgb_peers != nullptr, "sage_render_peers: NULL peer table");
    return render_impl(s, cam, row0_tiles, row1_tiles, n_peers, rgb_peers, depth_peers,
                       (cudaStream_t)stream);
}

int sage_render_host(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                     uint8_t* rgb_host, float* depth_host, sage_stream_t stream)
{
    REQUIRE(s != nullptr, "sage_render_host: NULL scene");
    REQUIRE(cam != nullptr && rgb_host != nullptr, "sage_render_host: NULL camera or rgb buffer");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->device));
    const size_t px = (size_t)cam->width * (size_t)cam->height;
    if (px > s->frame_px) {
        CK(cudaStreamSynchronize(st));
        cudaFree(s->frame_rgb); cudaFree(s->frame_depth);
        s->frame_rgb = nullptr; s->frame_depth = nullptr; s->frame_px = 0;
        CK(dalloc(&s->frame_rgb, px * 3));
        CK(dalloc(&s->frame_depth, px));
        s->frame_px = px;
    }
    if (sage_render(s, cam, row0_tiles, row1_tiles, s->frame_rgb, depth_host ? s->frame_depth : nullptr, stream))
        return 1;
    const int y0 = row0_tiles * SAGE_TILE;
    int y1 = row1_tiles * SAGE_TILE;
    if (y1 > cam->height) y1 = cam->height;
    if (y1 > y0) {
        const size_t off = (size_t)y0 * cam->width, cnt = (size_t)(y1 - y0) * cam->width;
        CK(cudaMemcpyAsync(rgb_host + off * 3, s->frame_rgb + off * 3, cnt * 3, cudaMemcpyDeviceToHost, st));
        if (depth_host)
            CK(cudaMemcpyAsync(depth_host + off, s->frame_depth + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

int sage_frame_stats_get(sage_scene* s, sage_frame_stats* out)
{
    REQUIRE(s != nullptr && out != nullptr, "sage_frame_stats_get: NULL argument");
    REQUIRE(s->rendered, "sage_frame_stats_get: no frame rendered yet");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->last_stream));
    const uint32_t* c = s->counters_host;
    memset(out, 0, sizeof *out);
    out->n_gaussians = s->sc.n;
    out->n_visible = c[CNT_VISIBLE];
    out->n_instances = c[CNT_INSTANCES];
    out->instance_capacity = s->cap;
    out->overflow = (int32_t)c[CNT_OVERFLOW];
    out->kernel_launches = s->last_launches;
    const int64_t nb = (int64_t)(s->sc.deg + 1) * (s->sc.deg + 1);
    out->algorithmic_bytes = out->n_gaussians * 44 + out->n_visible * 12 * nb + out->n_visible * 32 +
                             out->n_instances * 24 + out->n_instances * 32 +
                             (int64_t)s->last_W * s->last_H * (3 + 4 * s->last_depth);
    return 0;
}

int sage_set_stage_timing(sage_scene* s, int n_slots)
{
    REQUIRE(s != nullptr, "sage_set_stage_timing: NULL scene");
    REQUIRE(n_slots >= 0 && n_slots <= 4096, "sage_set_stage_timing: n_slots must be 0..4096");
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    s->events.clear();
    s->n_slots = 0;
    s->frame_index = 0;
    s->events.resize((size_t)n_slots * (SAGE_N_STAGES + 1));
    for (auto& e : s->events) CK(cudaEventCreate(&e));
    s->n_slots = n_slots;
    return 0;
}

int sage_stage_times_get(sage_scene* s, int slot, float ms_out[SAGE_N_STAGES])
{
    REQUIRE(s != nullptr && ms_out != nullptr, "sage_stage_times_get: NULL argument");
    REQUIRE(s->n_slots > 0 && slot >= 0 && slot < s->n_slots, "sage_stage_times_get: bad slot / timing off");
    CK(cudaSetDevice(s->device));
    cudaEvent_t* ev = &s->events[(size_t)slot * (SAGE_N_STAGES + 1)];
    CK(cudaEventSynchronize(ev[SAGE_N_STAGES]));
    for (int k = 0; k < SAGE_N_STAGES; ++k) CK(cudaEventElapsedTime(&ms_out[k], ev[k], ev[k + 1]));
    return 0;
}

int sage_project(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                 sage_splat_record* records_host, uint8_t* visible_host)
{
    REQUIRE(s != nullptr && records_host != nullptr && visible_host != nullptr, "sage_project: NULL argument");
    SageCamDev cd;
    if (make_cam(cam, row0_tiles, row1_tiles, &cd)) return 1;
    CK(cudaSetDevice(s->device));
    CK(cudaMemsetAsync(s->fr.counters, 0, sizeof(uint32_t) * SAGE_ZERO_WORDS, nullptr));
    s->epoch += 8;
    int launches = 0, ob = 0;
    if (front_half(s, cd, &launches, nullptr, nullptr, &ob)) return 1;
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    const size_t n = (size_t)s->sc.n;
    uint32_t V = 0;
    CK(cudaMemcpy(&V, s->fr.counters + CNT_VISIBLE, sizeof V, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> vlist(V);
    if (V) CK(cudaMemcpy(vlist.data(), s->fr.dval[ob], V * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::vector<uint8_t> visv(n, 0);
    for (uint32_t g : vlist) visv[g] = 1;
    std::vector<float4> r0(n), r1(n), r2(n);
    std::vector<uint2> rc(n);
    if (n) {
        CK(cudaMemcpy(r0.data(), s->fr.rec0, n * sizeof(float4), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(r1.data(), s->fr.rec1, n * sizeof(float4), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(r2.data(), s->fr.rec2, n * sizeof(float4), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rc.data(), s->fr.rect, n * sizeof(uint2), cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < n; ++i) {
        sage_splat_record& o = records_host[i];
        memset(&o, 0, sizeof o);
        visible_host[i] = visv[i];
        if (!visible_host[i]) continue;
        o.gx = r0[i].x; o.gy = r0[i].y; o.qa = r0[i].z;
        o.qb = r1[i].x; o.qc = r1[i].y; o.opacity = r1[i].z; o.z = r1[i].w;
        o.r = r2[i].x; o.g = r2[i].y; o.b = r2[i].z;
        o.tile_x0 = (rc[i].x & 0xFFFF) / SAGE_TILE; o.tile_x1 = (rc[i].x >> 16) / SAGE_TILE;
        o.tile_y0 = (rc[i].y & 0xFFFF) / SAGE_TILE; o.tile_y1 = (rc[i].y >> 16) / SAGE_TILE;
    }
    return 0;
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
    if (e != cudaSuccess) { sage_grid_destroy(g); return fail("sage_grid_create: allocation/upload", __FILE__, __LINE__, e); }
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
    if (e != cudaSuccess) { sage_bodies_destroy(b); return fail("sage_bodies_create: allocation/upload", __FILE__, __LINE__, e); }
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
