/*
 * sage_splat.h — C-ABI of libsage_splat.so, the B200-native (sm_100a) 3-D Gaussian-splatting
 * render-and-collide engine that stands behind SAGE-3D's Python render/collide surface.
 *
 * The reference has no FFI: its boundary is Python duck-typing at two seams.
 * Each entry point below names the reference call it replaces (paths relative to
 * Code/benchmark/environment_evaluation/ unless noted).  INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - every function returns 0 on success, non-zero on failure; sage_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors);
 *     "host" pointers are ordinary (ideally pinned) host memory;
 *   - calls on one handle are not re-entrant; different handles may be used from different
 *     threads;  no callbacks;  no allocation after create (sage_scene_reserve excepted);
 *   - there is NO CPU fallback: every entry point fails loudly without a CUDA device.
 */
#ifndef SAGE_SPLAT_H
#define SAGE_SPLAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAGE_ABI_VERSION 2
#define SAGE_TILE 16

typedef struct sage_scene sage_scene;   /* device-resident Gaussian scene + frame workspace */
typedef struct sage_grid sage_grid;     /* device-resident 2-D obstacle grid */
typedef struct sage_bodies sage_bodies; /* device-resident collision ellipsoids */
typedef void* sage_stream_t;

/* Pinhole camera, model -> camera: p_c = R p_m + t; camera axes +x right, +y down, +z forward.
 * Replaces the state Isaac's Camera holds after cam.set_world_pose(position, orientation)
 * (simple_env.py:1284; data_pipeline/training_data_construction/generate_images.py:419-421),
 * focal length 8.0 (simple_env.py:905) and clipping range (0.1, 50) (simple_env.py:899). */
typedef struct sage_camera {
    float R[9];      /* row-major */
    float t[3];
    float campos[3]; /* camera centre in model space (SH view direction) */
    float fx, fy, cx, cy;
    int32_t width, height;
    float near_z, far_z;
} sage_camera;

/* Device-side counters of the last frame rendered on a scene. */
typedef struct sage_frame_stats {
    int64_t n_gaussians;      /* N */
    int64_t n_visible;        /* V: survivors of the cull */
    int64_t n_instances;      /* D: (8x8 bin, Gaussian) pairs (clamped to the capacity) */
    int64_t n_instances_wanted; /* D before clamping: what sage_scene_reserve needs after an overflow */
    int64_t instance_capacity;
    int64_t algorithmic_bytes;/* N*44 + V*12*(deg+1)^2 + V*32 + D*24 + D*32 + W*H*(3+4*depth) */
    int32_t overflow;         /* 1: D exceeded instance_capacity; the frame is incomplete */
    int32_t kernel_launches;  /* kernels of this library launched for the frame */
} sage_frame_stats;

/* One row of sage_project's debug output (element-wise parity of the projection kernel). */
typedef struct sage_splat_record {
    float gx, gy, qa, qb, qc, log2_opacity, z, r, g, b;
    int32_t tile_x0, tile_y0, tile_x1, tile_y1; /* inclusive 16-px tile rectangle of the pixel box after clipping */
} sage_splat_record;

const char* sage_last_error(void);
int sage_abi_version(void);
/* Number of CUDA devices visible to the library (0 => nothing below can work). */
int sage_device_count(void);

/* ---- scene ------------------------------------------------------------------------------
 * Upload raw 3DGS attributes (host or device pointers) and repack them once into the engine's
 * HBM layout.  Stands in for `open_stage(usd_path)` + the USDZ Gaussian payload Isaac loads
 * (simple_env.py:219-230; generate_images.py:320-327; Data/template.usda:115-124).
 * sh is [n][(sh_degree+1)^2][3]; quat is (w,x,y,z), not necessarily normalised.
 * max_instances: capacity of the (tile, Gaussian) lists; <= 0 picks max(8 n, 2^24). */
int sage_scene_create(sage_scene** out, int device, int64_t n, int sh_degree, const float* xyz,
                      const float* log_scale, const float* quat_wxyz, const float* opacity_raw,
                      const float* sh, int64_t max_instances);
int sage_scene_reserve(sage_scene* s, int64_t max_instances);
int sage_scene_destroy(sage_scene* s);

/* ---- render -----------------------------------------------------------------------------
 * One frame: RGB u8 [H,W,3] and (optional) planar depth f32 [H,W] written to *device* buffers,
 * asynchronously on `stream`.  Replaces world.step(render=True) + cam.get_rgba()
 * (simple_env.py:1368-1380; generate_images.py:425-428) and
 * cam.get_current_frame()['distance_to_image_plane'] (simple_env.py:1425-1432).
 * Only tile rows [row0_tiles, row1_tiles) are produced (multi-GPU tile-row sharding; pass 0 and
 * ceil(H/16) for the whole frame); rows outside are left untouched.  rgb_dev/depth_dev always
 * address the full frame. */
int sage_render(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                uint8_t* rgb_dev, float* depth_dev, sage_stream_t stream);

/* Same frame, but into *host* buffers: render into the scene's internal device frame, copy the
 * produced rows device->host and synchronise.  This is the call behind env.get_rgb()/get_depth(). */
int sage_render_host(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                     uint8_t* rgb_host, float* depth_host, sage_stream_t stream);

/* Number of frame contexts of a scene = frames that can be in flight at once (3). */
int sage_pipeline_depth(void);
/* Frame contexts (frames in flight) of scenes created from now on: 1..8, default 3 (or the
 * SAGE_PIPELINE_DEPTH environment variable).  The ranks of a tile-row split render a fraction of a
 * frame per step and profit from a deeper pipeline (6); a context costs ~0.4 GB at 2 M Gaussians. */
int sage_set_pipeline_depth(int depth);

/* Pipelined host path: sage_pipeline_depth() frames can be in flight (slot 0 .. depth-1, each with
 * its own frame context and internal stream).  sage_render_host_async returns at once; sage_host_wait(slot) blocks until that
 * slot's RGB/depth have landed in the host buffers passed to it (keep them alive and distinct per
 * slot).  Submitting frame i+1 on the other slot before waiting for frame i overlaps frame i's
 * device->host copy and compositing with frame i+1's projection and sorts. */
int sage_render_host_async(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                           uint8_t* rgb_host, float* depth_host, int slot);
int sage_host_wait(sage_scene* s, int slot);

/* Multi-GPU fused composite+gather: like sage_render, but every finished pixel is stored
 * straight into each of n_peers frame buffers (peer-mapped device pointers, e.g. opened with
 * sage_ipc_open) over NVLink, so the strip exchange needs no separate collective copy. */
int sage_render_peers(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                      int n_peers, uint8_t* const* rgb_peers, float* const* depth_peers,
                      sage_stream_t stream);

/* Optional clamp applied to the depth output in the composite epilogue (the reference's
 * np.clip(depth, 0.1, 6.5), simple_env.py:1578, without a second pass over the frame on the host).
 * lo > hi disables it (default). */
int sage_scene_set_depth_clip(sage_scene* s, float lo, float hi);

/* Colour output format of every later render call on the scene: 0 (default) RGB u8 [H,W,3];
 * 1 RGBA u8 [H,W,4] with alpha = 255 — the array Isaac's Camera.get_rgba() returns
 * (simple_env.py:1380; generate_images.py:428), written as one 8-byte store per pixel pair, so the
 * Isaac shim needs no host-side expansion.  rgb buffers passed to render calls must then hold
 * 4 bytes per pixel. */
int sage_scene_set_rgba(sage_scene* s, int on);

/* Frames replay as one CUDA graph per frame context (default on; SAGE_NO_GRAPH=1 in the environment
 * or on = 0 launches every kernel directly).  sage_scene_graph_stats reports how many frames went
 * each way. */
int sage_scene_set_graph(sage_scene* s, int on);
int sage_scene_graph_stats(sage_scene* s, int64_t* graph_launches_out, int64_t* direct_frames_out);

/* Synchronises with the last frame on the scene and returns its counters. */
int sage_frame_stats_get(sage_scene* s, sage_frame_stats* out);

/* (bin, Gaussian) instances per 16-pixel tile row of the last frame (load-balancing the multi-GPU row split). */
int sage_frame_row_instances(sage_scene* s, int64_t* rows_out_host, int n_rows);

/* Per-stage device timing (bench/roofline): with n_slots > 0 every frame records CUDA events
 * between its stages on the caller's stream into slot (frame_index % n_slots); n_slots = 0 turns
 * it off.  sage_stage_times_get waits for that frame and returns milliseconds per stage. */
#define SAGE_N_STAGES 6 /* project, depth_sort, sh_color, emit, tile_sort (+ranges), composite */
int sage_set_stage_timing(sage_scene* s, int n_slots);
int sage_stage_times_get(sage_scene* s, int slot, float ms_out[SAGE_N_STAGES]);

/* Debug/parity: run only the projection kernel and return its per-Gaussian output to the host. */
int sage_project(sage_scene* s, const sage_camera* cam, int row0_tiles, int row1_tiles,
                 sage_splat_record* records_host, uint8_t* visible_host);

/* CUDA IPC plumbing for sage_render_peers.  A frame buffer may live inside a larger allocation
 * (e.g. a torch caching-allocator segment): export returns the 64-byte handle of the enclosing
 * allocation plus the byte offset of dev_ptr inside it; open maps the allocation in the calling
 * process and returns base + offset.  close takes the pointer returned by open and its offset. */
int sage_ipc_export(void* dev_ptr, uint8_t handle_out[64], uint64_t* offset_out);
int sage_ipc_open(const uint8_t handle[64], uint64_t offset, void** dev_ptr_out);
int sage_ipc_close(void* dev_ptr, uint64_t offset);
/* Device-side frame barrier between the ranks of one node, enqueued on `stream`: flag_peers[r] is rank
 * r's flag row (uint32 [n_peers], device memory, mapped with sage_ipc_open; flag_peers[my_rank] is this
 * rank's own).  After the stream's earlier work (the peer stores of sage_render_peers) it publishes
 * `value` to every peer and waits until every peer has published `value` (or later) here: the frame
 * is then complete in every rank's buffers.  Replaces a per-frame NCCL all-reduce;
 * `value` must increase from call to call on the same rows. */
int sage_peer_barrier(uint32_t* const* flag_peers, int n_peers, int my_rank, uint32_t value,
                      sage_stream_t stream);

/* ---- 2-D grid collide -------------------------------------------------------------------
 * Obstacle grid of SemanticMap2DCollisionDetector (collision_detector.py:95-132), uploaded once.
 * map_host is u8 [H][W] (1 = obstacle). */
int sage_grid_create(sage_grid** out, int device, const uint8_t* map_host, int H, int W,
                     double min_x, double max_x, double min_y, double max_y, double scale);
int sage_grid_destroy(sage_grid* g);

/* Batched check_collision_3d (collision_detector.py:191-237) over P float32 positions [P,3].
 * np2_promote != 0: NumPy >= 2 scalar promotion (binary32 flip step); 0: NumPy 1.26 (binary64).
 * pos/hit are HOST pointers (P <= 60 per agent step, simple_env.py:2116-2234); copies included.
 * first_hit_host (optional, int32[n_segments]): for consecutive segments of lengths
 * seg_len_host[0..n_segments) the index (within the segment) of the first colliding sample, or
 * -1 — the reduction _try_direct_movement/_try_short_movement perform by marching. */
int sage_collide_grid(sage_grid* g, const float* pos_host, int P, int np2_promote,
                      uint8_t* hit_host, const int32_t* seg_len_host, int n_segments,
                      int32_t* first_hit_host, sage_stream_t stream);
/* Double-precision positions (check_collision_at_position, collision_detector.py:263-272). */
int sage_collide_grid_f64(sage_grid* g, const double* pos_host, int P, uint8_t* hit_host,
                          sage_stream_t stream);
/* Device-resident variant: pos_dev f32 [P,3], hit_dev u8 [P]; asynchronous. */
int sage_collide_grid_dev(sage_grid* g, const float* pos_dev, int P, int np2_promote,
                          uint8_t* hit_dev, sage_stream_t stream);

/* ---- capsule collide --------------------------------------------------------------------
 * Collision bodies as ellipsoids (centre, log semi-axes, quaternion wxyz); stands in for the
 * convex collision payload (Data/template.usda:156-165) and the agent-vs-mesh test the reference
 * attempts through PhysX (simple_env.py:1878-1931; agent body :238,754-755). */
int sage_bodies_create(sage_bodies** out, int device, int64_t m, const float* mu,
                       const float* log_scale, const float* quat_wxyz);
int sage_bodies_destroy(sage_bodies* b);
/* pos_host f32 [P,3] -> hit_host u8 [P], first_body_host int32 [P] (optional; smallest index or -1). */
int sage_collide_capsule(sage_bodies* b, const float* pos_host, int P, float radius, float z0,
                         float z1, uint8_t* hit_host, int32_t* first_body_host,
                         sage_stream_t stream);

/* ---- frame hand-off: JPEG on the GPU --------------------------------
 * The reference turns every frame into a JPEG on the host with Pillow's defaults —
 *   Image.fromarray(img_array).save(traj_dir / filename)
 *       (data_pipeline/training_data_construction/generate_images.py:431-432)
 *   pil_image.save(buffer, format='JPEG')   (vlm_client_modular.py:351,402,453)
 * — after a 6.2 MB device->host copy of the 1080p frame.  sage_jpeg_* encode on the device;
 * the byte stream is IDENTICAL to Pillow/libjpeg-turbo's Image.save(format='JPEG',
 * quality=quality) of the same pixels (baseline, 4:2:0, Annex-K Huffman tables), so only the
 * encoded bytes need to cross PCIe.
 * max_batch: frames one call can encode (and the number of independent scratch slots);
 * max_bytes_per_frame <= 0 picks W*H*3/2 + 4096. */
typedef struct sage_jpeg sage_jpeg;
int sage_jpeg_create(sage_jpeg** out, int device, int width, int height, int quality,
                     int max_batch, int64_t max_bytes_per_frame);
int sage_jpeg_destroy(sage_jpeg* j);
int sage_jpeg_max_bytes(sage_jpeg* j, int64_t* out);
/* rgb_dev u8 [batch][H][W][3] -> out_dev[f * out_stride ...], sizes_dev[f] = bytes written, or
 * -(bytes needed) when out_stride is too small, or -1 when the scan exceeds the encoder's
 * max_bytes_per_frame.  Asynchronous on `stream`; all pointers are device pointers. */
int sage_jpeg_encode(sage_jpeg* j, const uint8_t* rgb_dev, int batch, uint8_t* out_dev,
                     int64_t out_stride, int64_t* sizes_dev, sage_stream_t stream);
/* Same, encoded bytes and sizes delivered to HOST memory; synchronises `stream`. */
int sage_jpeg_encode_host(sage_jpeg* j, const uint8_t* rgb_dev, int batch, uint8_t* out_host,
                          int64_t out_stride, int64_t* sizes_host, sage_stream_t stream);
/* Debug/parity: quantised coefficients (int16 [blocks][64], zig-zag, scan order), MCU bit
 * offsets and total scan bits of frame slot `frame` of the last encode. */
int sage_jpeg_debug(sage_jpeg* j, int frame, int16_t* coef_host, uint32_t* mcu_bitoff_host,
                    uint32_t* total_bits_host);

/* Fused hand-off behind env.get_rgb() -> VLM client / generate_images' save loop: render on
 * slot's frame context (as sage_render_host_async), encode on the same stream, and move only
 * what the caller asks for across PCIe.  rgb_host / depth_host may be NULL (then the raw
 * frame / depth never leave the device).  The encoder must match the camera's resolution and
 * have max_batch > slot (slot = scratch slot).  sage_render_jpeg_wait blocks until slot's frame
 * is done, copies exactly the encoded bytes into jpeg_host (capacity cap) and returns their
 * count in *size_out. */
int sage_render_jpeg_async(sage_scene* s, const sage_camera* cam, sage_jpeg* j, int slot,
                           uint8_t* rgb_host, float* depth_host);
/* The same with a second, private host copy of the frame, always packed RGB u8 [H,W,3] whatever the
 * scene's output mode: with sage_scene_set_rgba the caller gets the RGBA array Camera.get_rgba()
 * returns (generate_images.py:428) in rgb_host AND the exact pixels the JPEG was made from in
 * rgb_twin_host — what the hand-off compares a caller's array against before it lets
 * Image.fromarray(img[:, :, :3].copy()).save(...) (generate_images.py:431-432) reuse the device's bytes. */
int sage_render_jpeg_async2(sage_scene* s, const sage_camera* cam, sage_jpeg* j, int slot,
                            uint8_t* rgb_host, float* depth_host, uint8_t* rgb_twin_host);
int sage_render_jpeg_wait(sage_scene* s, sage_jpeg* j, int slot, uint8_t* jpeg_host,
                          int64_t cap, int64_t* size_out);

/* ---- frame hand-off: depth PNG on the GPU -----------------------------------
 * The RGB-D clients of the benchmark quantise and encode the depth map on the host for every query:
 *   depth_array_encoded = (np.clip(depth_image, 0.0, 6.5) * 10000.0).astype(np.uint16)
 *   cv2.imencode('.png', depth_array_encoded)                  (vlm_client_modular.py:354-364)
 * sage_depth_png_* do both on the device: the uint16 pixels are bit-identical to that NumPy sequence
 * (binary32 clip, binary32 multiply, truncation), the file is a valid 16-bit greyscale PNG of them
 * (Sub filter, one fixed-Huffman deflate block with in-group run-length matches, 32 KB IDAT chunks) —
 * NOT the bytes OpenCV's zlib build would produce (see csrc/png.cu); any reader decodes the same image.
 * twin (optional): the uint16 pixels themselves [H][W] — what a hand-off hook compares a caller's array
 * against before it substitutes the device's file. */
typedef struct sage_png sage_png;
int sage_png_create(sage_png** out, int device, int width, int height, int max_batch);
int sage_png_destroy(sage_png* p);
int sage_png_max_bytes(sage_png* p, int64_t* out);
/* depth_dev f32 [H][W] -> out_dev (capacity out_cap bytes), *size_dev = bytes written, -(bytes needed)
 * if out_cap is too small, -1 on internal overflow.  Asynchronous; device pointers. */
int sage_depth_png_encode(sage_png* p, const float* depth_dev, float lo, float hi, float scale,
                          uint8_t* out_dev, int64_t out_cap, int64_t* size_dev, uint16_t* twin_dev,
                          sage_stream_t stream);
/* Same, file bytes (and the optional uint16 twin) delivered to HOST memory; synchronises `stream`. */
int sage_depth_png_encode_host(sage_png* p, const float* depth_dev, float lo, float hi, float scale,
                               uint8_t* out_host, int64_t out_cap, int64_t* size_host,
                               uint16_t* twin_host, sage_stream_t stream);
/* Fused with the frame: encode the depth map of the frame last submitted on pipeline slot `slot`
 * (sage_render_host_async / sage_render_jpeg_async with a depth buffer) on that slot's stream;
 * sage_frame_depth_png_wait blocks until it is done and copies the file into png_host. */
int sage_frame_depth_png_async(sage_scene* s, sage_png* p, int slot, float lo, float hi, float scale,
                               uint16_t* twin_host);
int sage_frame_depth_png_wait(sage_scene* s, sage_png* p, int slot, uint8_t* png_host, int64_t cap,
                              int64_t* size_out);

/* ---- map building and path planning ---------------------------------------
 * The callers either side of the collide path work on the same occupancy grid
 * (data_pipeline/trajectory_generation/vln_trajectory_generator.py).
 * sage_grid_inflate: robot-radius inflation, `distance_transform_edt(raw == 0, sampling=scale)
 *   <= radius` (collision_detector.py:119-127; vln_trajectory_generator.py:549-552).  Host in,
 *   host out, u8 [H][W]; radius <= 0 copies (the reference inflates only for a positive radius).
 * sage_grid_components: 4-connected components of the free cells — the partition
 *   build_connectivity_map's union-find computes (vln_trajectory_generator.py:639-666).  labels
 *   int32 [H][W]: smallest linear index y*W+x of the cell's component, -1 for obstacle cells. */
int sage_grid_inflate(int device, const uint8_t* raw_host, int H, int W, double scale,
                      double radius, uint8_t* out_host, sage_stream_t stream);
int sage_grid_components(int device, const uint8_t* grid_host, int H, int W,
                         int32_t* labels_host, sage_stream_t stream);
/* Batched astar_pixel (vln_trajectory_generator.py:253-286): the PATH the reference returns
 * (same neighbour order, tie-breaking and binary64 arithmetic), one search per GPU thread,
 * `slots` searches side by side (each owns H*W*12 bytes of state + heap_capacity*16 bytes).
 * max_path <= 0 picks 4(H+W); heap_capacity <= 0 picks HW/2+1024 (the open list of a grid search
 * stays far below the cell count; -2 tells the caller to retry with more).
 * starts/goals int32 [Q][2] = (x, y).  path_len[q]: n >= 1 points written to
 * path_xy[q][0..n) (int32 [Q][max_path][2], start -> goal); 0: no path (the reference returns
 * None); -2: heap_capacity exceeded; -3: path longer than max_path (recreate with more room). */
typedef struct sage_planner sage_planner;
int sage_planner_create(sage_planner** out, int device, const uint8_t* grid_host, int H, int W,
                        int slots, int max_path, int64_t heap_capacity);
int sage_planner_destroy(sage_planner* p);
int sage_planner_limits(sage_planner* p, int* max_path_out, int64_t* heap_capacity_out);
int sage_plan_astar(sage_planner* p, const int32_t* starts_xy_host, const int32_t* goals_xy_host,
                    int Q, int32_t* path_len_host, int32_t* path_xy_host,
                    int64_t* expansions_host, sage_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAGE_SPLAT_H */
