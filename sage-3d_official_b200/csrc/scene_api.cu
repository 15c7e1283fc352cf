// C-ABI of libsage_splat.so, part 2: scene upload and rendering.
//
// A scene owns SAGE_N_CTX frame contexts (all per-frame device state, multi-buffered).  Consecutive
// frames rotate over the contexts, so a caller that rotates as many streams gets frame N+1's projection /
// sorts running while frame N is still compositing — the issue-bound blend kernel and the
// latency-bound sort kernels fill each other's gaps.  sage_render never synchronises the host
// (beyond waiting for the frame that used the same context SAGE_N_CTX frames ago).
//
// Everything that varies per frame (camera, output buffers, look-back epoch) is written into the
// context's pinned SageFrameParams block and reaches the kernels through a 400-byte host-to-device
// copy; all grids are fixed.  The launch sequence of a frame (frame.cu) is therefore captured ONCE
// per context into a CUDA graph and every later frame is a single cudaGraphLaunch.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "api_common.cuh"

#ifndef SAGE_N_CTX
#define SAGE_N_CTX 8      // most frame contexts a scene can own
#endif
// Contexts a scene is created with (frames in flight).  One GPU rendering whole frames: 5 (run V, same box:
// 2677 / 2766 / 2792 / 2800 frames/s device-resident and 2282 / 2539 / 2612 / 2651 end to end with 2 / 3 / 4 / 5;
// ~0.4 GB of frame state each at 2 M Gaussians).  A rank of a tile-row split renders a fraction of a frame per
// step: its kernels are smaller and more latency-bound, and deeper pipelining pays (4 GPUs: 5773 -> 6489
// frames/s with 6); the multi-GPU callers raise it with sage_set_pipeline_depth / SAGE_PIPELINE_DEPTH.
static int g_pipeline_depth = 0;
static int pipeline_depth_now()
{
    if (g_pipeline_depth == 0) {
        int d = 5;
        if (const char* v = getenv("SAGE_PIPELINE_DEPTH")) d = atoi(v);
        g_pipeline_depth = d < 1 ? 1 : (d > SAGE_N_CTX ? SAGE_N_CTX : d);
    }
    return g_pipeline_depth;
}

namespace {

struct GraphKey {
    int W = 0, H = 0, row0 = 0, row1 = 0, depth_bits = 0;
    long generation = -1;
    bool operator==(const GraphKey& o) const
    {
        return W == o.W && H == o.H && row0 == o.row0 && row1 == o.row1 && depth_bits == o.depth_bits &&
               generation == o.generation;
    }
};

struct GraphSlot {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    long last_use = 0;
};

constexpr int N_GRAPH_SLOTS = 4;

struct FrameCtx {
    SageFrame fr{};
    int64_t cap = 0;                 // instance capacity of this context
    uint32_t epoch = 8;              // look-back epochs: 8 fresh values per frame
    long generation = 0;             // bumped whenever a device buffer of the context is reallocated
    SageFrameParams* fp_host = nullptr;  // pinned staging copy of the frame parameters
    SageFrameParams* fp_dev = nullptr;
    uint32_t* counters_host = nullptr;   // pinned copy of the counters
    cudaEvent_t done = nullptr;      // end of the last frame that used this context
    cudaStream_t stream = nullptr;   // internal stream of the asynchronous host path
    cudaStream_t front = nullptr;    // highest-priority internal stream: every stage before the composite
    cudaStream_t side = nullptr;     // highest-priority internal stream: the SH-colour branch
    cudaStream_t low = nullptr;      // capture origin (stands in for the caller's stream)
    cudaEvent_t fork_join[4] = {nullptr, nullptr, nullptr, nullptr};
    GraphSlot graphs[N_GRAPH_SLOTS];
    uint8_t* out_rgb = nullptr;      // internal device frame (host paths)
    float* out_depth = nullptr;
    uint8_t* out_pack = nullptr;     // packed RGB copy of an RGBA frame (hand-off twin), allocated on first use
    uint16_t* out_u16 = nullptr;     // quantised depth (hand-off twin of the PNG), allocated on first use
    int frame_w = 0, frame_h = 0;    // resolution of the frame last rendered into out_rgb / out_depth
    bool frame_has_depth = false;
    size_t out_px = 0;
    cudaStream_t last_stream = nullptr;
    int last_W = 0, last_H = 0, last_depth = 0, last_launches = 0;
    int last_row0 = 0, last_row1 = 0, last_bins_x = 0;
    bool used = false;
};

}  // namespace

struct sage_scene {
    int device = 0;
    SageSceneDev sc{};
    FrameCtx ctx[SAGE_N_CTX];
    int n_ctx = 3;                   // contexts in use (pipeline depth at creation)
    uint32_t* n_dev = nullptr;       // device copy of n (element count of the first depth-sort pass)
    long frame_index = 0;
    int last_ctx = -1;
    // optional per-stage timing: n_slots x (SAGE_N_STAGES + 1) events
    std::vector<cudaEvent_t> events;
    int n_slots = 0;
    long timed_frames = 0;
    float depth_lo = 1.0f, depth_hi = 0.0f;   // depth output clamp (off while lo > hi)
    int rgba = 0;                    // colour output: 0 = RGB u8 x 3, 1 = RGBA u8 x 4 (alpha 255)
    bool use_graph = true;
    long graph_launches = 0, direct_frames = 0;
};

static void drop_graphs(FrameCtx& c)
{
    for (auto& g : c.graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = GraphSlot{};
    }
}

static void free_instances(FrameCtx& c)
{
    for (int k = 0; k < 2; ++k) {
        cudaFree(c.fr.ikey[k]); c.fr.ikey[k] = nullptr;
        cudaFree(c.fr.ival[k]); c.fr.ival[k] = nullptr;
    }
    cudaFree(c.fr.sort_h); c.fr.sort_h = nullptr;
    c.generation++;
}

static int alloc_instances(sage_scene* s, FrameCtx& c, int64_t cap)
{
    free_instances(c);
    for (int k = 0; k < 2; ++k) {
        CK(dalloc(&c.fr.ikey[k], (size_t)cap));
        CK(dalloc(&c.fr.ival[k], (size_t)cap));
    }
    const int64_t mx = cap > s->sc.n ? cap : s->sc.n;
    CK(dalloc(&c.fr.sort_h, sage_sort_matrix_words(mx)));
    c.fr.sort_ld = sage_sort_matrix_ld(mx);
    c.cap = cap;
    return 0;
}

static int alloc_ctx(sage_scene* s, FrameCtx& c, int64_t cap)
{
    const size_t n = (size_t)s->sc.n;
    CK(dalloc(&c.fr.rec, 3 * n));
    CK(dalloc(&c.fr.dkey_src, n));
    for (int k = 0; k < 2; ++k) {
        CK(dalloc(&c.fr.dkey[k], n));
        CK(dalloc(&c.fr.dval[k], n));
    }
    CK(dalloc(&c.fr.rect, n));
    CK(dalloc(&c.fr.chain, n / 256 + 2));
    CK(cudaMemset(c.fr.chain, 0, (n / 256 + 2) * sizeof(unsigned long long)));
    CK(dalloc(&c.fr.counters, (size_t)SAGE_ZERO_WORDS));
    c.fr.tickets = c.fr.counters + CNT_COUNT;
    c.fr.class_count = c.fr.tickets + SAGE_TICKET_COUNT;
    CK(dalloc(&c.fr.sort_totals, (size_t)SAGE_RADIX_MAX_BINS));
    CK(cudaHostAlloc((void**)&c.counters_host, sizeof(uint32_t) * CNT_COUNT, cudaHostAllocDefault));
    memset(c.counters_host, 0, sizeof(uint32_t) * CNT_COUNT);
    CK(cudaHostAlloc((void**)&c.fp_host, sizeof(SageFrameParams), cudaHostAllocDefault));
    memset(c.fp_host, 0, sizeof(SageFrameParams));
    CK(dalloc(&c.fp_dev, 1));
    c.fr.fp = c.fp_dev;
    CK(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    {
        // The stages before the composite are a chain of small latency-bound kernels; the composite
        // is one long issue-bound kernel.  Run the chain on a highest-priority stream: its CTAs are
        // then dispatched ahead of the queued composite CTAs of the previous frame (other context),
        // so frame i+1's projection/sorts run INSIDE frame i's composite instead of behind it.
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c.front, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&c.low, cudaStreamNonBlocking, lo));
        for (auto& e : c.fork_join) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return alloc_instances(s, c, cap);
}

static void free_ctx(FrameCtx& c)
{
    drop_graphs(c);
    cudaFree(c.fr.rec); cudaFree(c.fr.dkey_src);
    for (int k = 0; k < 2; ++k) { cudaFree(c.fr.dkey[k]); cudaFree(c.fr.dval[k]); }
    cudaFree(c.fr.rect); cudaFree(c.fr.chain); cudaFree(c.fr.sort_totals);
    free_instances(c);
    cudaFree(c.fr.ranges); cudaFree(c.fr.class_seg); cudaFree(c.fr.counters);
    cudaFree(c.fp_dev);
    if (c.fp_host) cudaFreeHost(c.fp_host);
    if (c.counters_host) cudaFreeHost(c.counters_host);
    if (c.done) cudaEventDestroy(c.done);
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.front) cudaStreamDestroy(c.front);
    if (c.side) cudaStreamDestroy(c.side);
    if (c.low) cudaStreamDestroy(c.low);
    for (auto& e : c.fork_join) if (e) cudaEventDestroy(e);
    cudaFree(c.out_rgb); cudaFree(c.out_depth); cudaFree(c.out_pack); cudaFree(c.out_u16);
    c = FrameCtx{};
}

static int make_cam(const sage_camera* cam, int row0, int row1, SageCamDev* d)
{
    REQUIRE(cam != nullptr, "camera is NULL");
    REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 16384 && cam->height <= 16384, "camera: bad resolution");
    REQUIRE(cam->fx > 0.0f && cam->fy > 0.0f, "camera: focal length must be positive");
    REQUIRE(cam->near_z > 0.0f && cam->far_z > cam->near_z, "camera: need 0 < near < far");
    memset(d, 0, sizeof *d);
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
    d->depth_lo = 1.0f; d->depth_hi = 0.0f;
    static_assert(SAGE_TILE % SAGE_BIN_H == 0, "public tile rows are whole bin rows");
    const int per = SAGE_TILE / SAGE_BIN_H;
    const int bins_y = (cam->height + SAGE_BIN_H - 1) / SAGE_BIN_H;
    d->bins_x = (cam->width + SAGE_BIN_W - 1) / SAGE_BIN_W;
    d->bin_row0 = row0 * per;
    d->bin_row1 = row1 * per < bins_y ? row1 * per : bins_y;
    if (d->bin_row1 < d->bin_row0) d->bin_row1 = d->bin_row0;
    d->n_bins = d->bins_x * (d->bin_row1 - d->bin_row0);
    int tbits = 1;
    while ((1 << tbits) < d->n_bins) ++tbits;
    d->tile_passes = (tbits + 7) / 8;
    REQUIRE(d->tile_passes <= 3, "sage_render: too many bins");
    d->tile_digit_bits = (tbits + d->tile_passes - 1) / d->tile_passes;
    if (d->tile_digit_bits < 5) d->tile_digit_bits = 5;
    if (d->tile_passes > 2) d->tile_digit_bits = 8;             // (> 65536 tiles: 8-bit digits, as before)
    return 0;
}

// resolution-dependent scratch (bin ranges, class segments) grows on first use
static int ensure_bins(FrameCtx& c, int n_bins)
{
    if (n_bins > c.fr.max_bins) {
        if (c.used) CK(cudaEventSynchronize(c.done));
        cudaFree(c.fr.ranges); cudaFree(c.fr.class_seg);
        c.fr.ranges = nullptr; c.fr.class_seg = nullptr; c.fr.max_bins = 0;
        CK(dalloc(&c.fr.ranges, (size_t)n_bins));
        CK(dalloc(&c.fr.class_seg, (size_t)n_bins * SAGE_N_CLASSES));
        c.fr.max_bins = n_bins;
        c.generation++;
    }
    return 0;
}

static int ensure_out(FrameCtx& c, size_t px)
{
    if (px > c.out_px) {
        if (c.used) CK(cudaEventSynchronize(c.done));
        cudaFree(c.out_rgb); cudaFree(c.out_depth); cudaFree(c.out_pack); cudaFree(c.out_u16);
        c.out_rgb = nullptr; c.out_depth = nullptr; c.out_pack = nullptr; c.out_u16 = nullptr; c.out_px = 0;
        CK(dalloc(&c.out_rgb, px * 4));          // room for RGBA
        CK(dalloc(&c.out_depth, px));
        c.out_px = px;
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

static void next_epoch(FrameCtx& c)
{
    c.epoch += 8;
    if (c.epoch > (1u << 30) - 16) c.epoch = 8;
}

// Build (capture + instantiate) the frame graph of context c for `key`; nullptr on any failure (the
// caller then launches directly).
static GraphSlot* get_graph(sage_scene* s, FrameCtx& c, const GraphKey& key, const SageFrameShape& shape, int n_bins)
{
    GraphSlot* victim = &c.graphs[0];
    for (auto& g : c.graphs) {
        if (g.exec && g.key == key) { g.last_use = s->frame_index; return &g; }
        if (!g.exec) { victim = &g; break; }
        if (g.last_use < victim->last_use) victim = &g;
    }
    if (victim->exec) { cudaGraphExecDestroy(victim->exec); *victim = GraphSlot{}; }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c.low, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemcpyAsync(c.fp_dev, c.fp_host, sizeof(SageFrameParams), cudaMemcpyHostToDevice, c.low);
    const int launches = sage_frame_enqueue(s->sc, c.fr, shape, s->n_dev, c.low, c.front, c.side, c.fork_join, nullptr, n_bins);
    cudaMemcpyAsync(c.counters_host, c.fr.counters, sizeof(uint32_t) * CNT_COUNT, cudaMemcpyDeviceToHost, c.low);
    const cudaError_t e_end = cudaStreamEndCapture(c.low, &graph);
    if (e_end != cudaSuccess || graph == nullptr || launches < 0) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return nullptr;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e_inst != cudaSuccess || exec == nullptr) { cudaGetLastError(); return nullptr; }
    victim->key = key;
    victim->exec = exec;
    victim->launches = launches;
    victim->last_use = s->frame_index;
    return victim;
}

// One frame on context `ci` (or the next context in rotation when ci < 0), on stream `st`.
static int render_impl(sage_scene* s, int ci, const sage_camera* cam, int row0, int row1, int n_peers,
                       uint8_t* const* rgb_peers, float* const* depth_peers, cudaStream_t st)
{
    REQUIRE(s != nullptr, "sage_render: NULL scene");
    REQUIRE(n_peers >= 1 && n_peers <= SAGE_MAX_PEERS, "sage_render: n_peers must be 1..8");
    for (int k = 0; k < n_peers; ++k) REQUIRE(rgb_peers[k] != nullptr, "sage_render: NULL rgb buffer");
    SageCamDev cd;
    if (make_cam(cam, row0, row1, &cd)) return 1;
    cd.depth_lo = s->depth_lo; cd.depth_hi = s->depth_hi;
    CK(cudaSetDevice(s->device));
    if (ci < 0) ci = (int)(s->frame_index % s->n_ctx);
    FrameCtx& c = s->ctx[ci];
    // the context's pinned parameter block and counters are reused: the frame that used them
    // SAGE_N_CTX frames ago must have drained (normally long done; bounds the host's run-ahead)
    if (c.used) CK(cudaEventSynchronize(c.done));
    if (ensure_bins(c, cd.n_bins)) return 1;
    next_epoch(c);
    SageFrameParams& P = *c.fp_host;
    P.cam = cd;
    P.epoch = c.epoch;
    P.cap = (uint32_t)c.cap;
    bool aligned = (cd.W % 8) == 0;
    P.out.n = n_peers;
    P.out.rgba = s->rgba;
    for (int k = 0; k < SAGE_MAX_PEERS; ++k) {
        P.out.rgb[k] = k < n_peers ? rgb_peers[k] : nullptr;
        P.out.depth[k] = (k < n_peers && depth_peers) ? depth_peers[k] : nullptr;
        if (k < n_peers && (((uintptr_t)P.out.rgb[k] & 7u) || ((uintptr_t)P.out.depth[k] & 7u))) aligned = false;
    }
    P.out.wide = aligned ? 1 : 0;
    SageFrameShape shape;
    shape.depth_bits = depth_key_bits(cd);
    shape.tile_passes = cd.tile_passes;
    shape.tile_digit_bits = cd.tile_digit_bits;
    shape.whole_frame = (row0 == 0 && row1 >= cd.tiles_y) ? 1 : 0;
    {
        int rb = 0, np = 0;
        sage_sort_shape(shape.depth_bits, &rb, &np);
        REQUIRE(np <= SAGE_MAX_PASSES && !(rb == 9 && np > 3), "depth sort: unsupported key width");
    }
    cudaEvent_t* ev = s->n_slots > 0
        ? &s->events[(size_t)(s->timed_frames % s->n_slots) * (SAGE_N_STAGES + 1)] : nullptr;
    if (ev) s->timed_frames++;
    s->frame_index++;
    int launches = -1;
    if (!ev && s->use_graph) {
        GraphKey key;
        key.W = cd.W; key.H = cd.H; key.row0 = row0; key.row1 = row1; key.depth_bits = shape.depth_bits;
        key.generation = c.generation;
        GraphSlot* g = get_graph(s, c, key, shape, cd.n_bins);
        if (g) {
            CK(cudaGraphLaunch(g->exec, st));
            launches = g->launches;
            s->graph_launches++;
        }
    }
    if (launches < 0) {                 // stage timing on, graphs off, or capture unavailable: direct launches
        CK(cudaMemcpyAsync(c.fp_dev, c.fp_host, sizeof(SageFrameParams), cudaMemcpyHostToDevice, st));
        launches = sage_frame_enqueue(s->sc, c.fr, shape, s->n_dev, st, c.front, c.side, c.fork_join, ev, cd.n_bins);
        REQUIRE(launches >= 0, "sage_render: unsupported sort shape");
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c.counters_host, c.fr.counters, sizeof(uint32_t) * CNT_COUNT, cudaMemcpyDeviceToHost, st));
        s->direct_frames++;
    }
    CK(cudaEventRecord(c.done, st));
    c.used = true;
    c.last_stream = st;
    c.last_W = cd.W; c.last_H = cd.H;
    c.last_depth = (depth_peers && depth_peers[0]) ? 1 : 0;
    c.last_launches = launches;
    c.last_row0 = row0; c.last_row1 = row1; c.last_bins_x = cd.bins_x;
    s->last_ctx = ci;
    return 0;
}

extern "C" {

int sage_scene_create(sage_scene** out, int device, int64_t n, int sh_degree, const float* xyz,
                      const float* log_scale, const float* quat_wxyz, const float* opacity_raw,
                      const float* sh, int64_t max_instances)
{
    REQUIRE(out != nullptr, "sage_scene_create: out is NULL");
    *out = nullptr;
    REQUIRE(sage_device_count() > 0, "sage_scene_create: no CUDA device (this library has no CPU fallback)");
    REQUIRE(n >= 0 && n <= (int64_t)SAGE_ID_MASK, "sage_scene_create: n out of range (max 2^24-1 = 16 777 215 Gaussians per scene)");
    REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sage_scene_create: sh_degree must be 0..3");
    REQUIRE(n == 0 || (xyz && log_scale && quat_wxyz && opacity_raw && sh), "sage_scene_create: NULL attribute pointer");
    CK(cudaSetDevice(device));
    sage_sort_configure();
    sage_bin_configure();
    sage_composite_configure();
    sage_project_configure();
    sage_scene* s = new sage_scene();
    s->device = device;
    s->sc.n = n;
    s->sc.deg = sh_degree;
    s->use_graph = getenv("SAGE_NO_GRAPH") == nullptr;
    const int nf = 3 * (sh_degree + 1) * (sh_degree + 1);
    s->sc.n_chunks = (nf + 3) / 4;
    auto bail = [&](int rc) { sage_scene_destroy(s); return rc; };
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(sage_fail(#call, __FILE__, __LINE__, e__)); } while (0)
    CKB(dalloc(&s->sc.geo0, (size_t)n));
    CKB(dalloc(&s->sc.geo1, (size_t)n));
    CKB(dalloc(&s->sc.geo2, (size_t)n));
    CKB(dalloc(&s->sc.shc, (size_t)n * s->sc.n_chunks));
    CKB(dalloc(&s->n_dev, 1));
    uint32_t n32 = (uint32_t)n;
    CKB(cudaMemcpy(s->n_dev, &n32, sizeof n32, cudaMemcpyHostToDevice));
    const int64_t cap = max_instances > 0 ? max_instances : (8 * n > (1 << 24) ? 8 * n : (1 << 24));
    s->n_ctx = pipeline_depth_now();
    for (int k = 0; k < s->n_ctx; ++k)
        if (alloc_ctx(s, s->ctx[k], cap)) return bail(1);

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
    for (int k = 0; k < s->n_ctx; ++k) {
        drop_graphs(s->ctx[k]);
        if (alloc_instances(s, s->ctx[k], max_instances)) return 1;
    }
    return 0;
}

int sage_scene_destroy(sage_scene* s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    cudaFree(s->sc.geo0); cudaFree(s->sc.geo1); cudaFree(s->sc.geo2); cudaFree(s->sc.shc);
    for (int k = 0; k < SAGE_N_CTX; ++k) free_ctx(s->ctx[k]);
    cudaFree(s->n_dev);
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    cudaGetLastError();
    delete s;
    return 0;
}

int sage_render(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                uint8_t* rgb_dev, float* depth_dev, sage_stream_t stream)
{
    uint8_t* rgbs[1] = {rgb_dev};
    float* deps[1] = {depth_dev};
    return render_impl(s, -1, cam, row0_tiles, row1_tiles, 1, rgbs, depth_dev ? deps : nullptr,
                       (cudaStream_t)stream);
}

int sage_render_peers(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                      int n_peers, uint8_t* const* rgb_peers, float* const* depth_peers,
                      sage_stream_t stream)
{
    REQUIRE(rgb_peers != nullptr, "sage_render_peers: NULL peer table");
    return render_impl(s, -1, cam, row0_tiles, row1_tiles, n_peers, rgb_peers, depth_peers,
                       (cudaStream_t)stream);
}

// render on context ci / stream st into the context's device frame, then copy the rows to the host
static int render_to_host(sage_scene* s, int ci, const sage_camera* cam, int row0, int row1,
                          uint8_t* rgb_host, float* depth_host, cudaStream_t st)
{
    REQUIRE(s != nullptr, "sage_render_host: NULL scene");
    REQUIRE(cam != nullptr && rgb_host != nullptr, "sage_render_host: NULL camera or rgb buffer");
    CK(cudaSetDevice(s->device));
    if (ci < 0) ci = (int)(s->frame_index % s->n_ctx);
    FrameCtx& c = s->ctx[ci];
    if (ensure_out(c, (size_t)cam->width * (size_t)cam->height)) return 1;
    uint8_t* rgbs[1] = {c.out_rgb};
    float* deps[1] = {c.out_depth};
    if (render_impl(s, ci, cam, row0, row1, 1, rgbs, depth_host ? deps : nullptr, st)) return 1;
    c.frame_w = cam->width; c.frame_h = cam->height; c.frame_has_depth = depth_host != nullptr;
    const int y0 = row0 * SAGE_TILE;
    int y1 = row1 * SAGE_TILE;
    if (y1 > cam->height) y1 = cam->height;
    if (y1 > y0) {
        const size_t bpp = s->rgba ? 4 : 3;
        const size_t off = (size_t)y0 * cam->width, cnt = (size_t)(y1 - y0) * cam->width;
        CK(cudaMemcpyAsync(rgb_host + off * bpp, c.out_rgb + off * bpp, cnt * bpp, cudaMemcpyDeviceToHost, st));
        if (depth_host)
            CK(cudaMemcpyAsync(depth_host + off, c.out_depth + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(c.done, st));     // the context is busy until its copies have drained
    return 0;
}

int sage_render_host(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                     uint8_t* rgb_host, float* depth_host, sage_stream_t stream)
{
    if (render_to_host(s, -1, cam, row0_tiles, row1_tiles, rgb_host, depth_host, (cudaStream_t)stream)) return 1;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int sage_render_host_async(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                           uint8_t* rgb_host, float* depth_host, int slot)
{
    REQUIRE(s != nullptr, "sage_render_host_async: NULL scene");
    REQUIRE(slot >= 0 && slot < s->n_ctx, "sage_render_host_async: slot must be in [0, sage_pipeline_depth())");
    return render_to_host(s, slot, cam, row0_tiles, row1_tiles, rgb_host, depth_host, s->ctx[slot].stream);
}

int sage_pipeline_depth(void) { return pipeline_depth_now(); }

int sage_set_pipeline_depth(int depth)
{
    REQUIRE(depth >= 1 && depth <= SAGE_N_CTX, "sage_set_pipeline_depth: depth must be in [1, 8]");
    g_pipeline_depth = depth;
    return 0;
}

int sage_host_wait(sage_scene* s, int slot)
{
    REQUIRE(s != nullptr, "sage_host_wait: NULL scene");
    REQUIRE(slot >= 0 && slot < s->n_ctx, "sage_host_wait: slot must be in [0, sage_pipeline_depth())");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->ctx[slot].stream));
    return 0;
}

// RGBA u8x4 -> packed RGB u8x3, four pixels (16 bytes in, 12 bytes out) per thread
static __global__ void __launch_bounds__(256) k_pack_rgb(const uint4* __restrict__ rgba, uint32_t* __restrict__ rgb, size_t n_quads)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = rgba[i];
        const uint32_t a = v.x & 0xFFFFFFu, b = v.y & 0xFFFFFFu, c = v.z & 0xFFFFFFu, d = v.w & 0xFFFFFFu;
        rgb[3 * i + 0] = a | (b << 24);
        rgb[3 * i + 1] = (b >> 8) | (c << 16);
        rgb[3 * i + 2] = (c >> 16) | (d << 8);
    }
}

int sage_render_jpeg_async(sage_scene* s, const sage_camera* cam, sage_jpeg* j, int slot,
                           uint8_t* rgb_host, float* depth_host)
{
    return sage_render_jpeg_async2(s, cam, j, slot, rgb_host, depth_host, nullptr);
}

int sage_render_jpeg_async2(sage_scene* s, const sage_camera* cam, sage_jpeg* j, int slot,
                            uint8_t* rgb_host, float* depth_host, uint8_t* rgb_twin_host)
{
    REQUIRE(s != nullptr && cam != nullptr && j != nullptr, "sage_render_jpeg_async: NULL argument");
    REQUIRE(slot >= 0 && slot < s->n_ctx, "sage_render_jpeg_async: slot must be in [0, sage_pipeline_depth())");
    REQUIRE(sage_jpeg_device(j) == s->device, "sage_render_jpeg_async: encoder lives on another device");
    REQUIRE(sage_jpeg_width(j) == cam->width && sage_jpeg_height(j) == cam->height,
            "sage_render_jpeg_async: encoder resolution differs from the camera's");
    REQUIRE(sage_jpeg_max_batch(j) > slot, "sage_render_jpeg_async: encoder needs max_batch > slot (one scratch slot per frame in flight)");
    CK(cudaSetDevice(s->device));
    FrameCtx& c = s->ctx[slot];
    cudaStream_t st = c.stream;
    const size_t px = (size_t)cam->width * (size_t)cam->height;
    if (ensure_out(c, px)) return 1;
    uint8_t* rgbs[1] = {c.out_rgb};
    float* deps[1] = {c.out_depth};
    const int rows = (cam->height + SAGE_TILE - 1) / SAGE_TILE;
    if (render_impl(s, slot, cam, 0, rows, 1, rgbs, depth_host ? deps : nullptr, st)) return 1;
    c.frame_w = cam->width; c.frame_h = cam->height; c.frame_has_depth = depth_host != nullptr;
    const int64_t stride = sage_jpeg_stage_stride(j);
    if (sage_jpeg_enqueue(j, c.out_rgb, s->rgba ? 4 : 3, 1, slot, sage_jpeg_stage_out(j) + (size_t)slot * stride, stride,
                          sage_jpeg_stage_sizes_dev(j) + slot, st)) return 1;
    CK(cudaMemcpyAsync(sage_jpeg_stage_sizes_host(j) + slot, sage_jpeg_stage_sizes_dev(j) + slot, sizeof(long long),
                       cudaMemcpyDeviceToHost, st));
    // the stream's size is unknown until the encode has run: copy last frame's size + slack along with the
    // frame, so the wait normally finds the whole file already on the host (no second round trip)
    {
        long long g = sage_jpeg_stage_guess(j)[slot];
        if (g > stride) g = stride;
        CK(cudaMemcpyAsync(sage_jpeg_stage_out_host(j) + (size_t)slot * stride, sage_jpeg_stage_out(j) + (size_t)slot * stride,
                           (size_t)g, cudaMemcpyDeviceToHost, st));
    }
    if (rgb_host) CK(cudaMemcpyAsync(rgb_host, c.out_rgb, px * (s->rgba ? 4 : 3), cudaMemcpyDeviceToHost, st));
    if (depth_host) CK(cudaMemcpyAsync(depth_host, c.out_depth, px * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (rgb_twin_host) {                              // the frame once more, always as packed RGB
        const uint8_t* src = c.out_rgb;
        if (s->rgba) {
            REQUIRE(px % 4 == 0, "sage_render_jpeg_async2: the packed twin of an RGBA frame needs width * height % 4 == 0");
            if (!c.out_pack) CK(dalloc(&c.out_pack, c.out_px * 3));
            const size_t quads = px / 4;
            SAGE_KLAUNCH(k_pack_rgb, 148 * 4, 256, st, reinterpret_cast<const uint4*>(c.out_rgb),
                         reinterpret_cast<uint32_t*>(c.out_pack), quads);
            src = c.out_pack;
        }
        CK(cudaMemcpyAsync(rgb_twin_host, src, px * 3, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(c.done, st));
    return 0;
}

int sage_render_jpeg_wait(sage_scene* s, sage_jpeg* j, int slot, uint8_t* jpeg_host, int64_t cap, int64_t* size_out)
{
    REQUIRE(s != nullptr && j != nullptr && jpeg_host != nullptr && size_out != nullptr, "sage_render_jpeg_wait: NULL argument");
    REQUIRE(slot >= 0 && slot < s->n_ctx, "sage_render_jpeg_wait: slot must be in [0, sage_pipeline_depth())");
    CK(cudaSetDevice(s->device));
    FrameCtx& c = s->ctx[slot];
    CK(cudaStreamSynchronize(c.stream));
    const long long sz = sage_jpeg_stage_sizes_host(j)[slot];
    *size_out = sz;
    REQUIRE(sz > 0, "sage_render_jpeg_wait: the encoded frame exceeds the encoder's max_bytes_per_frame");
    REQUIRE(sz <= cap, "sage_render_jpeg_wait: jpeg_host is too small (*size_out holds the size needed)");
    const int64_t stride = sage_jpeg_stage_stride(j);
    long long& guess = sage_jpeg_stage_guess(j)[slot];
    const long long have = guess < stride ? guess : stride;
    uint8_t* stage = sage_jpeg_stage_out_host(j) + (size_t)slot * stride;
    if (sz > have) {                                  // (a frame much larger than the previous one)
        CK(cudaMemcpyAsync(stage + have, sage_jpeg_stage_out(j) + (size_t)slot * stride + have, (size_t)(sz - have),
                           cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
    }
    memcpy(jpeg_host, stage, (size_t)sz);
    guess = sz + sz / 4 + 4096;
    return 0;
}

int sage_frame_depth_png_async(sage_scene* s, sage_png* p, int slot, float lo, float hi, float scale, uint16_t* twin_host)
{
    REQUIRE(s != nullptr && p != nullptr, "sage_frame_depth_png_async: NULL argument");
    REQUIRE(slot >= 0 && slot < s->n_ctx, "sage_frame_depth_png_async: slot must be in [0, sage_pipeline_depth())");
    REQUIRE(sage_png_device(p) == s->device, "sage_frame_depth_png_async: encoder lives on another device");
    REQUIRE(sage_png_max_batch(p) > slot, "sage_frame_depth_png_async: encoder needs max_batch > slot (one scratch slot per frame in flight)");
    CK(cudaSetDevice(s->device));
    FrameCtx& c = s->ctx[slot];
    REQUIRE(c.frame_has_depth && c.out_depth != nullptr, "sage_frame_depth_png_async: the slot's last frame was rendered without depth");
    REQUIRE(sage_png_width(p) == c.frame_w && sage_png_height(p) == c.frame_h,
            "sage_frame_depth_png_async: encoder resolution differs from the slot's frame");
    cudaStream_t st = c.stream;
    const size_t px = (size_t)c.frame_w * (size_t)c.frame_h;
    if (twin_host && !c.out_u16) CK(dalloc(&c.out_u16, c.out_px));
    const int64_t stride = sage_png_stage_stride(p);
    if (sage_png_enqueue(p, c.out_depth, lo, hi, scale, slot, sage_png_stage_out(p) + (size_t)slot * stride, stride,
                         sage_png_stage_sizes_dev(p) + slot, twin_host ? c.out_u16 : nullptr, st)) return 1;
    CK(cudaMemcpyAsync(sage_png_stage_sizes_host(p) + slot, sage_png_stage_sizes_dev(p) + slot, sizeof(long long),
                       cudaMemcpyDeviceToHost, st));
    if (twin_host) CK(cudaMemcpyAsync(twin_host, c.out_u16, px * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c.done, st));
    return 0;
}

int sage_frame_depth_png_wait(sage_scene* s, sage_png* p, int slot, uint8_t* png_host, int64_t cap, int64_t* size_out)
{
    REQUIRE(s != nullptr && p != nullptr && png_host != nullptr && size_out != nullptr, "sage_frame_depth_png_wait: NULL argument");
    REQUIRE(slot >= 0 && slot < s->n_ctx, "sage_frame_depth_png_wait: slot must be in [0, sage_pipeline_depth())");
    CK(cudaSetDevice(s->device));
    FrameCtx& c = s->ctx[slot];
    CK(cudaStreamSynchronize(c.stream));
    const long long sz = sage_png_stage_sizes_host(p)[slot];
    *size_out = sz;
    REQUIRE(sz > 0, "sage_frame_depth_png_wait: the encoder's stream buffer overflowed");
    REQUIRE(sz <= cap, "sage_frame_depth_png_wait: png_host is too small (*size_out holds the size needed)");
    CK(cudaMemcpyAsync(png_host, sage_png_stage_out(p) + (size_t)slot * sage_png_stage_stride(p), (size_t)sz,
                       cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return 0;
}

int sage_scene_set_depth_clip(sage_scene* s, float lo, float hi)
{
    REQUIRE(s != nullptr, "sage_scene_set_depth_clip: NULL scene");
    s->depth_lo = lo; s->depth_hi = hi;
    return 0;
}

int sage_scene_set_rgba(sage_scene* s, int on)
{
    REQUIRE(s != nullptr, "sage_scene_set_rgba: NULL scene");
    s->rgba = on ? 1 : 0;
    return 0;
}

int sage_scene_set_graph(sage_scene* s, int on)
{
    REQUIRE(s != nullptr, "sage_scene_set_graph: NULL scene");
    s->use_graph = on != 0;
    return 0;
}

int sage_scene_graph_stats(sage_scene* s, int64_t* graph_launches_out, int64_t* direct_frames_out)
{
    REQUIRE(s != nullptr, "sage_scene_graph_stats: NULL scene");
    if (graph_launches_out) *graph_launches_out = s->graph_launches;
    if (direct_frames_out) *direct_frames_out = s->direct_frames;
    return 0;
}

int sage_frame_stats_get(sage_scene* s, sage_frame_stats* out)
{
    REQUIRE(s != nullptr && out != nullptr, "sage_frame_stats_get: NULL argument");
    REQUIRE(s->last_ctx >= 0, "sage_frame_stats_get: no frame rendered yet");
    CK(cudaSetDevice(s->device));
    FrameCtx& c = s->ctx[s->last_ctx];
    CK(cudaEventSynchronize(c.done));
    const uint32_t* cn = c.counters_host;
    memset(out, 0, sizeof *out);
    out->n_gaussians = s->sc.n;
    out->n_visible = cn[CNT_VISIBLE];
    out->n_instances = cn[CNT_INSTANCES];
    out->n_instances_wanted = cn[CNT_WANTED];
    out->instance_capacity = c.cap;
    out->overflow = (int32_t)cn[CNT_OVERFLOW];
    out->kernel_launches = c.last_launches;
    const int64_t nb = (int64_t)(s->sc.deg + 1) * (s->sc.deg + 1);
    out->algorithmic_bytes = out->n_gaussians * 44 + out->n_visible * 12 * nb + out->n_visible * 32 +
                             out->n_instances * 24 + out->n_instances * 32 +
                             (int64_t)c.last_W * c.last_H * (3 + 4 * c.last_depth);
    return 0;
}

int sage_frame_row_instances(sage_scene* s, int64_t* rows_out_host, int n_rows)
{
    REQUIRE(s != nullptr && rows_out_host != nullptr, "sage_frame_row_instances: NULL argument");
    REQUIRE(s->last_ctx >= 0, "sage_frame_row_instances: no frame rendered yet");
    CK(cudaSetDevice(s->device));
    FrameCtx& c = s->ctx[s->last_ctx];
    CK(cudaEventSynchronize(c.done));
    const int rows = c.last_row1 - c.last_row0;
    REQUIRE(n_rows == rows, "sage_frame_row_instances: n_rows must equal the rows of the last frame");
    const int per = SAGE_TILE / SAGE_BIN_H;
    const int bins_y = (c.last_H + SAGE_BIN_H - 1) / SAGE_BIN_H;
    int bin_rows = rows * per;
    if (c.last_row0 * per + bin_rows > bins_y) bin_rows = bins_y - c.last_row0 * per;
    if (bin_rows < 0) bin_rows = 0;
    std::vector<uint2> rg((size_t)bin_rows * c.last_bins_x);
    if (!rg.empty()) CK(cudaMemcpy(rg.data(), c.fr.ranges, rg.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
    for (int r = 0; r < rows; ++r) rows_out_host[r] = 0;
    for (int br = 0; br < bin_rows; ++br) {
        int64_t acc = 0;
        for (int x = 0; x < c.last_bins_x; ++x) {
            const uint2 v = rg[(size_t)br * c.last_bins_x + x];      // (~begin, end); (0, 0) = empty
            if (v.y) acc += (int64_t)v.y - (int64_t)(~v.x);
        }
        rows_out_host[br / per] += acc;
    }
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
    s->timed_frames = 0;
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
    CK(cudaDeviceSynchronize());
    FrameCtx& c = s->ctx[0];
    c.fp_host->cam = cd;
    CK(cudaMemcpy(c.fp_dev, c.fp_host, sizeof(SageFrameParams), cudaMemcpyHostToDevice));
    CK(cudaMemsetAsync(c.fr.counters, 0, sizeof(uint32_t) * SAGE_ZERO_WORDS, nullptr));
    sage_launch_project(s->sc, c.fr, nullptr);
    sage_launch_sh_color(s->sc, c.fr, nullptr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    const size_t n = (size_t)s->sc.n;
    std::vector<uint32_t> keys(n);
    std::vector<float4> rr(3 * n);
    std::vector<uint2> rc(n);
    if (n) {
        CK(cudaMemcpy(keys.data(), c.fr.dkey_src, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rr.data(), c.fr.rec, 3 * n * sizeof(float4), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rc.data(), c.fr.rect, n * sizeof(uint2), cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < n; ++i) {
        sage_splat_record& o = records_host[i];
        memset(&o, 0, sizeof o);
        visible_host[i] = keys[i] != SAGE_KEY_CULLED ? 1 : 0;
        if (!visible_host[i]) continue;
        const float4 r0 = rr[3 * i], r1 = rr[3 * i + 1], r2 = rr[3 * i + 2];
        o.gx = r0.x; o.gy = r0.y; o.qa = r0.z; o.log2_opacity = r0.w;
        o.qb = r1.x; o.qc = r1.y;
        o.r = r2.x; o.g = r2.y; o.b = r2.z; o.z = r2.w;
        o.tile_x0 = (rc[i].x & 0xFFFF) / SAGE_TILE; o.tile_x1 = (rc[i].x >> 16) / SAGE_TILE;
        o.tile_y0 = (rc[i].y & 0xFFFF) / SAGE_TILE; o.tile_y1 = (rc[i].y >> 16) / SAGE_TILE;
    }
    return 0;
}

}  // extern "C"
