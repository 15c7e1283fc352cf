"""ctypes binding of libsage_splat.so (C-ABI: include/sage_splat.h).

There is no CPU fallback: a missing library or a missing CUDA device raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsage_splat.so"
if os.environ.get("SAGE_SPLAT_LIB"):       # tuning experiments: another build of the same sources (tools/build_variant.sh)
    LIB_PATH = Path(os.environ["SAGE_SPLAT_LIB"]).resolve()

#: every symbol declared in include/sage_splat.h (tests check the .so exports all of them)
EXPORTS = [
    "sage_last_error", "sage_abi_version", "sage_device_count",
    "sage_scene_create", "sage_scene_reserve", "sage_scene_destroy", "sage_scene_set_depth_clip",
    "sage_scene_set_rgba", "sage_scene_set_graph", "sage_scene_graph_stats",
    "sage_render", "sage_render_host", "sage_render_host_async", "sage_host_wait", "sage_pipeline_depth", "sage_set_pipeline_depth", "sage_render_peers", "sage_frame_stats_get", "sage_project",
    "sage_frame_row_instances", "sage_set_stage_timing", "sage_stage_times_get",
    "sage_ipc_export", "sage_ipc_open", "sage_ipc_close", "sage_peer_barrier",
    "sage_grid_create", "sage_grid_destroy", "sage_collide_grid", "sage_collide_grid_f64",
    "sage_collide_grid_dev",
    "sage_bodies_create", "sage_bodies_destroy", "sage_collide_capsule",
    "sage_jpeg_create", "sage_jpeg_destroy", "sage_jpeg_max_bytes", "sage_jpeg_encode", "sage_jpeg_encode_host",
    "sage_png_create", "sage_png_destroy", "sage_png_max_bytes", "sage_depth_png_encode", "sage_depth_png_encode_host",
    "sage_frame_depth_png_async", "sage_frame_depth_png_wait",
    "sage_jpeg_debug", "sage_render_jpeg_async", "sage_render_jpeg_async2", "sage_render_jpeg_wait",
    "sage_grid_inflate", "sage_grid_components", "sage_planner_create", "sage_planner_destroy",
    "sage_planner_limits", "sage_plan_astar",
]


class SageError(RuntimeError):
    pass


class CCamera(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("campos", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32), ("near_z", C.c_float), ("far_z", C.c_float)]


class CFrameStats(C.Structure):
    _fields_ = [("n_gaussians", C.c_int64), ("n_visible", C.c_int64), ("n_instances", C.c_int64), ("n_instances_wanted", C.c_int64),
                ("instance_capacity", C.c_int64), ("algorithmic_bytes", C.c_int64),
                ("overflow", C.c_int32), ("kernel_launches", C.c_int32)]


class CSplatRecord(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("gx", "gy", "qa", "qb", "qc", "log2_opacity", "z", "r", "g", "b")] + \
               [(k, C.c_int32) for k in ("tile_x0", "tile_y0", "tile_x1", "tile_y1")]


_lib = None


def build(verbose: bool = False) -> Path:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    r = subprocess.run(["make", "-C", str(_HERE / "csrc"), "-j8"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise SageError("building libsage_splat.so failed")
    return LIB_PATH


def lib():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SageError(f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    L.sage_last_error.restype = C.c_char_p
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.sage_scene_create.argtypes = [C.POINTER(vp), i32, i64, i32, vp, vp, vp, vp, vp, i64]
    L.sage_scene_reserve.argtypes = [vp, i64]
    L.sage_scene_destroy.argtypes = [vp]
    L.sage_scene_set_depth_clip.argtypes = [vp, C.c_float, C.c_float]
    L.sage_scene_set_rgba.argtypes = [vp, i32]
    L.sage_scene_set_graph.argtypes = [vp, i32]
    L.sage_scene_graph_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.sage_render.argtypes = [vp, C.POINTER(CCamera), i32, i32, vp, vp, vp]
    L.sage_render_host.argtypes = [vp, C.POINTER(CCamera), i32, i32, vp, vp, vp]
    L.sage_render_host_async.argtypes = [vp, C.POINTER(CCamera), i32, i32, vp, vp, i32]
    L.sage_host_wait.argtypes = [vp, i32]
    L.sage_render_peers.argtypes = [vp, C.POINTER(CCamera), i32, i32, i32, C.POINTER(vp), C.POINTER(vp), vp]
    L.sage_frame_stats_get.argtypes = [vp, C.POINTER(CFrameStats)]
    L.sage_frame_row_instances.argtypes = [vp, vp, i32]
    L.sage_set_stage_timing.argtypes = [vp, i32]
    L.sage_stage_times_get.argtypes = [vp, i32, vp]
    L.sage_project.argtypes = [vp, C.POINTER(CCamera), i32, i32, vp, vp]
    L.sage_ipc_export.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.sage_ipc_open.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.sage_ipc_close.argtypes = [vp, C.c_uint64]
    L.sage_peer_barrier.argtypes = [vp, i32, i32, C.c_uint32, vp]
    dbl = C.c_double
    L.sage_grid_create.argtypes = [C.POINTER(vp), i32, vp, i32, i32, dbl, dbl, dbl, dbl, dbl]
    L.sage_grid_destroy.argtypes = [vp]
    L.sage_collide_grid.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp, vp]
    L.sage_collide_grid_f64.argtypes = [vp, vp, i32, vp, vp]
    L.sage_collide_grid_dev.argtypes = [vp, vp, i32, i32, vp, vp]
    L.sage_bodies_create.argtypes = [C.POINTER(vp), i32, i64, vp, vp, vp]
    L.sage_bodies_destroy.argtypes = [vp]
    L.sage_collide_capsule.argtypes = [vp, vp, i32, C.c_float, C.c_float, C.c_float, vp, vp, vp]
    L.sage_jpeg_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, i64]
    L.sage_jpeg_destroy.argtypes = [vp]
    L.sage_jpeg_max_bytes.argtypes = [vp, C.POINTER(i64)]
    L.sage_jpeg_encode.argtypes = [vp, vp, i32, vp, i64, vp, vp]
    L.sage_jpeg_encode_host.argtypes = [vp, vp, i32, vp, i64, vp, vp]
    L.sage_jpeg_debug.argtypes = [vp, i32, vp, vp, vp]
    L.sage_render_jpeg_async.argtypes = [vp, C.POINTER(CCamera), vp, i32, vp, vp]
    L.sage_render_jpeg_async2.argtypes = [vp, C.POINTER(CCamera), vp, i32, vp, vp, vp]
    L.sage_render_jpeg_wait.argtypes = [vp, vp, i32, vp, i64, C.POINTER(i64)]
    f32 = C.c_float
    L.sage_png_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    L.sage_png_destroy.argtypes = [vp]
    L.sage_png_max_bytes.argtypes = [vp, C.POINTER(i64)]
    L.sage_depth_png_encode.argtypes = [vp, vp, f32, f32, f32, vp, i64, vp, vp, vp]
    L.sage_depth_png_encode_host.argtypes = [vp, vp, f32, f32, f32, vp, i64, C.POINTER(i64), vp, vp]
    L.sage_frame_depth_png_async.argtypes = [vp, vp, i32, f32, f32, f32, vp]
    L.sage_frame_depth_png_wait.argtypes = [vp, vp, i32, vp, i64, C.POINTER(i64)]
    L.sage_grid_inflate.argtypes = [i32, vp, i32, i32, dbl, dbl, vp, vp]
    L.sage_grid_components.argtypes = [i32, vp, i32, i32, vp, vp]
    L.sage_planner_create.argtypes = [C.POINTER(vp), i32, vp, i32, i32, i32, i32, i64]
    L.sage_planner_destroy.argtypes = [vp]
    L.sage_planner_limits.argtypes = [vp, C.POINTER(i32), C.POINTER(i64)]
    L.sage_plan_astar.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name != "sage_last_error":
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise SageError(lib().sage_last_error().decode("utf-8", "replace"))


def require_gpu() -> None:
    if lib().sage_device_count() <= 0:
        raise SageError("no CUDA device visible: sage-3d_official_b200 has no CPU fallback")
