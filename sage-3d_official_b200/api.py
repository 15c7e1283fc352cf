"""Python surface of the engine: ``render(camera, gaussians) -> (rgb, depth)`` and
``collide(agent_pose, collision) -> hit`` (BASELINE.json north_star), torch tensors in/out,
every call funnelled through the C-ABI of ``libsage_splat.so``.

PyTorch is used for device memory and streams only; no tensor arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import CCamera, CFrameStats, CSplatRecord, SageError, check
from .camera import Camera
from .scenes import GaussianCloud

TILE = 16


def _torch():
    import torch
    return torch


def _ccam(cam: Camera) -> CCamera:
    """The C struct of a camera; cached on the object (a Camera is a value: callers that move the camera
    build a new one, so the cache cannot go stale unless fields are assigned in place — then `meta` is cleared)."""
    cached = cam.meta.get("_ccam")
    if cached is not None:
        return cached
    c = CCamera()
    c.R[:] = [float(v) for v in np.asarray(cam.R, np.float32).reshape(9)]
    c.t[:] = [float(v) for v in np.asarray(cam.t, np.float32).reshape(3)]
    c.campos[:] = [float(v) for v in np.asarray(cam.campos, np.float32).reshape(3)]
    c.fx, c.fy, c.cx, c.cy = float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy)
    c.width, c.height = int(cam.width), int(cam.height)
    c.near_z, c.far_z = float(cam.near), float(cam.far)
    cam.meta["_ccam"] = c
    return c


def _stream_ptr(stream=None):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _ptr(a) -> C.c_void_p:
    """Address of a numpy array or torch tensor (host or device)."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())


def _as_f32(a):
    """float32, contiguous; numpy stays numpy, torch stays torch (host or device)."""
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.float32)
    torch = _torch()
    if not isinstance(a, torch.Tensor):
        return np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    return a.detach().to(dtype=torch.float32).contiguous()


class GaussianScene:
    """A Gaussian scene resident in HBM (handle of ``sage_scene``)."""

    def __init__(self, gaussians, device: int = 0, max_instances: int = 0):
        _lib.require_gpu()
        if isinstance(gaussians, GaussianCloud):
            parts = gaussians.as_tuple()
        else:
            parts = tuple(gaussians)
        if len(parts) != 5:
            raise ValueError("gaussians = (xyz[N,3], log_scale[N,3], quat_wxyz[N,4], opacity_raw[N], sh[N,K,3])")
        xyz, ls, q, op, sh = (_as_f32(p) for p in parts)
        n = int(xyz.shape[0])
        if tuple(xyz.shape) != (n, 3) or tuple(ls.shape) != (n, 3) or tuple(q.shape) != (n, 4) \
                or int(np.prod(op.shape)) != n or sh.ndim != 3 or sh.shape[0] != n or sh.shape[2] != 3:
            raise ValueError("inconsistent Gaussian attribute shapes")
        nb = int(sh.shape[1])
        deg = int(round(np.sqrt(nb))) - 1
        if (deg + 1) ** 2 != nb or not 0 <= deg <= 3:
            raise ValueError(f"sh has {nb} coefficients per channel; expected 1, 4, 9 or 16")
        self.n, self.sh_degree, self.device = n, deg, int(device)
        self._h = C.c_void_p()
        check(_lib.lib().sage_scene_create(C.byref(self._h), self.device, n, deg, _ptr(xyz), _ptr(ls),
                                           _ptr(q), _ptr(op), _ptr(sh), int(max_instances)))
        self._pinned = {}
        #: frames that can be in flight at once (slots of render_host_async / render_jpeg_async)
        self.pipeline_depth = int(_lib.lib().sage_pipeline_depth())

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sage_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_depth_clip(self, lo: float = 1.0, hi: float = 0.0):
        """Clamp the depth output on the device (lo > hi switches it off)."""
        check(_lib.lib().sage_scene_set_depth_clip(self._h, float(lo), float(hi)))

    def reserve(self, max_instances: int):
        check(_lib.lib().sage_scene_reserve(self._h, int(max_instances)))

    def set_rgba(self, on: bool = True):
        """Colour output as RGBA u8 [H,W,4] (alpha 255) — what Isaac's ``Camera.get_rgba()`` returns —
        instead of RGB u8 [H,W,3].  Applies to every later render call on the scene."""
        check(_lib.lib().sage_scene_set_rgba(self._h, 1 if on else 0))
        self._rgba = bool(on)
        self._pinned = {}

    def set_graph(self, on: bool = True):
        """Replay frames as one CUDA graph per frame context (default) or launch every kernel directly."""
        check(_lib.lib().sage_scene_set_graph(self._h, 1 if on else 0))

    def graph_stats(self) -> dict:
        a, b = C.c_int64(0), C.c_int64(0)
        check(_lib.lib().sage_scene_graph_stats(self._h, C.byref(a), C.byref(b)))
        return {"graph_launches": int(a.value), "direct_frames": int(b.value)}

    @property
    def channels(self) -> int:
        return 4 if getattr(self, "_rgba", False) else 3

    # -- render --
    @staticmethod
    def _rows(cam: Camera, rows):
        ty = (cam.height + TILE - 1) // TILE
        if rows is None:
            return 0, ty
        r0, r1 = int(rows[0]), int(rows[1])
        return r0, r1

    def render(self, cam: Camera, rows=None, want_depth: bool = True, out=None, stream=None):
        """Asynchronous frame into CUDA tensors: (rgb u8 [H,W,3], depth f32 [H,W] | None)."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        if out is None:
            rgb = torch.empty((cam.height, cam.width, self.channels), dtype=torch.uint8, device=dev)
            depth = torch.empty((cam.height, cam.width), dtype=torch.float32, device=dev) if want_depth else None
            if rows is not None:
                rgb.zero_()
                if depth is not None:
                    depth.zero_()
        else:
            rgb, depth = out
        r0, r1 = self._rows(cam, rows)
        cc = _ccam(cam)
        check(_lib.lib().sage_render(self._h, C.byref(cc), r0, r1, _ptr(rgb), _ptr(depth), _stream_ptr(stream)))
        return rgb, depth

    def render_peers(self, cam: Camera, rows, rgb_ptrs: Sequence[int], depth_ptrs: Optional[Sequence[int]], stream=None):
        """Composite rows straight into several (peer-mapped) frame buffers given as raw addresses."""
        n = len(rgb_ptrs)
        # (multigpu.PeerFrames hands in ready-made ctypes arrays: nothing is rebuilt per frame)
        RA = rgb_ptrs if isinstance(rgb_ptrs, C.Array) else (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in rgb_ptrs])
        DA = None
        if depth_ptrs is not None and len(depth_ptrs):
            DA = depth_ptrs if isinstance(depth_ptrs, C.Array) else (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in depth_ptrs])
        cc = _ccam(cam)
        check(_lib.lib().sage_render_peers(self._h, C.byref(cc), int(rows[0]), int(rows[1]), n, RA, DA,
                                           _stream_ptr(stream)))

    def render_host(self, cam: Camera, rows=None, want_depth: bool = True, stream=None):
        """Synchronous frame into pinned host memory: (rgb u8 ndarray, depth f32 ndarray | None)."""
        torch = _torch()
        key = (cam.height, cam.width)
        if key not in self._pinned:
            self._pinned[key] = (torch.zeros((cam.height, cam.width, self.channels), dtype=torch.uint8).pin_memory(),
                                 torch.zeros((cam.height, cam.width), dtype=torch.float32).pin_memory())
        rgb_t, dep_t = self._pinned[key]
        r0, r1 = self._rows(cam, rows)
        cc = _ccam(cam)
        with torch.cuda.device(self.device):
            for attempt in range(2):
                check(_lib.lib().sage_render_host(self._h, C.byref(cc), r0, r1, _ptr(rgb_t),
                                                  _ptr(dep_t) if want_depth else C.c_void_p(0), _stream_ptr(stream)))
                st = self.stats()                  # the call above synchronised: no extra wait
                if not st["overflow"]:
                    break
                # the (tile, Gaussian) lists did not fit (camera inside geometry): grow and redo
                self.reserve(int(st["n_instances_wanted"] * 1.25) + 1024)
        return rgb_t.numpy(), (dep_t.numpy() if want_depth else None)

    def alloc_host_frame(self, height: int, width: int):
        """A pinned host (rgb u8 [H,W,C], depth f32 [H,W]) tensor pair for :meth:`render_host_async`'s ``out``."""
        torch = _torch()
        return (torch.zeros((height, width, self.channels), dtype=torch.uint8).pin_memory(),
                torch.zeros((height, width), dtype=torch.float32).pin_memory())

    def render_host_async(self, cam: Camera, slot: int, rows=None, want_depth: bool = True, out=None):
        """Submit a frame on pipeline slot 0 .. pipeline_depth-1; returns the slot's pinned host arrays
        (valid after :meth:`host_wait`).  Rotate the slots to keep pipeline_depth frames in flight.
        ``out``: a caller-owned pinned tensor pair (see :meth:`alloc_host_frame`) to land the frame in."""
        torch = _torch()
        if out is not None:
            rgb_t, dep_t = out
        else:
            key = (cam.height, cam.width, "async", int(slot))
            if key not in self._pinned:
                self._pinned[key] = self.alloc_host_frame(cam.height, cam.width)
            rgb_t, dep_t = self._pinned[key]
        r0, r1 = self._rows(cam, rows)
        cc = _ccam(cam)
        with torch.cuda.device(self.device):
            check(_lib.lib().sage_render_host_async(self._h, C.byref(cc), r0, r1, _ptr(rgb_t),
                                                    _ptr(dep_t) if want_depth else C.c_void_p(0), int(slot)))
        as_np = lambda t: t if isinstance(t, np.ndarray) else t.numpy()
        return as_np(rgb_t), (as_np(dep_t) if want_depth else None)

    def render_batch_host(self, cams, want_depth: bool = True):
        """Render a sequence of camera poses (e.g. every sampled waypoint of a trajectory, as
        generate_images.py does one by one) keeping pipeline_depth frames in flight.  Yields
        ``(index, rgb, depth)`` in order; the arrays are views of pinned buffers that are reused a few
        frames later — copy or encode them before advancing the generator ``pipeline_depth`` times."""
        cams = list(cams)
        bufs = {}
        d = self.pipeline_depth
        for i, cam in enumerate(cams):
            slot = i % d
            if i >= d:
                self.host_wait(slot)
                yield (i - d,) + bufs[slot]
            bufs[slot] = self.render_host_async(cam, slot, want_depth=want_depth)
        for i in range(max(0, len(cams) - d), len(cams)):
            self.host_wait(i % d)
            yield (i,) + bufs[i % d]

    def host_wait(self, slot: int):
        check(_lib.lib().sage_host_wait(self._h, int(slot)))

    # -- frame hand-off: render -> JPEG on the device --
    def render_jpeg_async(self, cam: Camera, encoder: "JpegEncoder", slot: int, want_rgb: bool = False,
                          want_depth: bool = False, out=None, twin=None):
        """Submit render + JPEG encode on a pipeline slot (< min(pipeline_depth, encoder.max_batch)).  Only what is asked for crosses PCIe:
        the encoded bytes (always, fetched by :meth:`jpeg_wait`), the raw frame (RGB, or RGBA after
        :meth:`set_rgba`) and/or depth (optional; returned as pinned host arrays valid after :meth:`jpeg_wait`).
        ``out``: a caller-owned pinned (rgb, depth) tensor pair (see :meth:`alloc_host_frame`) instead of the
        slot's own; ``twin``: a pinned u8 [H,W,3] tensor that receives the frame once more, always as packed
        RGB — the exact pixels the JPEG was made from (the hand-off's reference copy)."""
        torch = _torch()
        rgb_t = dep_t = None
        if out is not None:
            rgb_t, dep_t = out
        elif want_rgb or want_depth:
            key = (cam.height, cam.width, "async", int(slot))
            if key not in self._pinned:
                self._pinned[key] = self.alloc_host_frame(cam.height, cam.width)
            rgb_t, dep_t = self._pinned[key]
        cc = _ccam(cam)
        with torch.cuda.device(self.device):
            check(_lib.lib().sage_render_jpeg_async2(self._h, C.byref(cc), encoder._h, int(slot),
                                                     _ptr(rgb_t) if want_rgb else C.c_void_p(0),
                                                     _ptr(dep_t) if want_depth else C.c_void_p(0),
                                                     _ptr(twin) if twin is not None else C.c_void_p(0)))
        as_np = lambda t: t if isinstance(t, np.ndarray) else t.numpy()
        return (as_np(rgb_t) if want_rgb else None), (as_np(dep_t) if want_depth else None)

    def jpeg_wait(self, encoder: "JpegEncoder", slot: int) -> bytes:
        """Block until slot's frame is rendered and encoded; returns the JPEG file bytes."""
        buf = encoder._host_buf(slot)
        n = C.c_int64(0)
        check(_lib.lib().sage_render_jpeg_wait(self._h, encoder._h, int(slot), _ptr(buf), int(buf.numel()), C.byref(n)))
        return buf.numpy()[:n.value].tobytes()

    # -- frame hand-off: depth -> uint16 PNG on the device --
    def depth_png_async(self, encoder: "DepthPngEncoder", slot: int, lo: float = 0.0, hi: float = 6.5,
                        scale: float = 10000.0, twin=None):
        """Encode the depth map of the frame last submitted on ``slot`` (with depth) as the 16-bit PNG of
        ``(clip(depth, lo, hi) * scale).astype(uint16)`` (vlm_client_modular.py:354-364) on the slot's stream.
        ``twin``: optional pinned uint16 [H,W] tensor receiving the quantised pixels themselves."""
        check(_lib.lib().sage_frame_depth_png_async(self._h, encoder._h, int(slot), float(lo), float(hi), float(scale),
                                                    _ptr(twin) if twin is not None else C.c_void_p(0)))

    def depth_png_wait(self, encoder: "DepthPngEncoder", slot: int) -> bytes:
        buf = encoder._host_buf(slot)
        n = C.c_int64(0)
        check(_lib.lib().sage_frame_depth_png_wait(self._h, encoder._h, int(slot), _ptr(buf), int(buf.numel()), C.byref(n)))
        return buf.numpy()[:n.value].tobytes()

    def render_batch_jpeg(self, cams, encoder: "JpegEncoder"):
        """Every pose of a trajectory -> JPEG bytes, several frames in flight (what
        generate_images.py:408-456 does one waypoint at a time with 3 RTX frames, a 3 MB
        read-back and a host-side Pillow encode each).  Yields ``(index, jpeg_bytes)`` in order."""
        cams = list(cams)
        d = min(self.pipeline_depth, encoder.max_batch)
        for i, cam in enumerate(cams):
            if i >= d:
                yield i - d, self.jpeg_wait(encoder, i % d)
            self.render_jpeg_async(cam, encoder, i % d)
        for i in range(max(0, len(cams) - d), len(cams)):
            yield i, self.jpeg_wait(encoder, i % d)

    def host_wait_all(self):
        for k in range(self.pipeline_depth):
            self.host_wait(k)


    def stats(self) -> dict:
        st = CFrameStats()
        check(_lib.lib().sage_frame_stats_get(self._h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in CFrameStats._fields_}

    def row_instances(self, n_rows: int) -> np.ndarray:
        """(tile, Gaussian) instances per tile row of the last frame."""
        out = np.zeros(int(n_rows), np.int64)
        check(_lib.lib().sage_frame_row_instances(self._h, _ptr(out), int(n_rows)))
        return out

    STAGES = ("project", "depth_sort", "sh_color", "emit", "tile_sort", "composite")

    def set_stage_timing(self, n_slots: int):
        check(_lib.lib().sage_set_stage_timing(self._h, int(n_slots)))

    def stage_times(self, slot: int) -> dict:
        ms = (C.c_float * len(self.STAGES))()
        check(_lib.lib().sage_stage_times_get(self._h, int(slot), C.cast(ms, C.c_void_p)))
        return dict(zip(self.STAGES, (float(v) for v in ms)))

    def project(self, cam: Camera, rows=None):
        """Projection kernel only (parity/debug): structured records + visibility mask."""
        rec = (CSplatRecord * max(self.n, 1))()
        vis = np.zeros(max(self.n, 1), np.uint8)
        r0, r1 = self._rows(cam, rows)
        cc = _ccam(cam)
        check(_lib.lib().sage_project(self._h, C.byref(cc), r0, r1, C.cast(rec, C.c_void_p), _ptr(vis)))
        dt = np.dtype([(k, np.float32) for k in ("gx", "gy", "qa", "qb", "qc", "log2_opacity", "z", "r", "g", "b")] +
                      [(k, np.int32) for k in ("tile_x0", "tile_y0", "tile_x1", "tile_y1")])
        arr = np.frombuffer(rec, dtype=dt, count=self.n).copy()
        return arr, vis[:self.n].astype(bool)


class ObstacleGrid:
    """Device-resident 2-D obstacle grid (handle of ``sage_grid``)."""

    def __init__(self, obstacle_map, min_x, max_x, min_y, max_y, scale, device: int = 0):
        _lib.require_gpu()
        self._h = C.c_void_p()
        m = None if obstacle_map is None else np.ascontiguousarray(obstacle_map, dtype=np.uint8)
        H, W = (0, 0) if m is None else m.shape
        check(_lib.lib().sage_grid_create(C.byref(self._h), int(device), _ptr(m), H, W, float(min_x), float(max_x),
                                          float(min_y), float(max_y), float(scale)))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sage_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, positions, np2_promote: Optional[bool] = None, segments: Optional[Sequence[int]] = None):
        """Batched check_collision_3d.  float32 positions follow the NumPy promotion mode
        (default: the running interpreter's); float64 positions take the all-double path.
        Returns hit mask (bool [P]) and, if ``segments`` is given, first-hit index per segment."""
        pos = np.asarray(positions)
        if pos.dtype == np.float64:
            pos = np.ascontiguousarray(pos.reshape(-1, 3))
            hit = np.zeros(pos.shape[0], np.uint8)
            check(_lib.lib().sage_collide_grid_f64(self._h, _ptr(pos), pos.shape[0], _ptr(hit), C.c_void_p(0)))
            first = None
            if segments is not None:
                first = _first_hits(hit, segments)
            return (hit.astype(bool), first) if segments is not None else hit.astype(bool)
        if np2_promote is None:
            np2_promote = int(np.__version__.split(".")[0]) >= 2
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        P = pos.shape[0]
        hit = np.zeros(P, np.uint8)
        if segments is None:
            check(_lib.lib().sage_collide_grid(self._h, _ptr(pos), P, int(bool(np2_promote)), _ptr(hit),
                                               C.c_void_p(0), 0, C.c_void_p(0), C.c_void_p(0)))
            return hit.astype(bool)
        seg = np.ascontiguousarray(segments, dtype=np.int32)
        first = np.full(seg.shape[0], -1, np.int32)
        check(_lib.lib().sage_collide_grid(self._h, _ptr(pos), P, int(bool(np2_promote)), _ptr(hit), _ptr(seg),
                                           seg.shape[0], _ptr(first), C.c_void_p(0)))
        return hit.astype(bool), first

    def check_device(self, pos_dev, np2_promote: bool = True, stream=None):
        """positions: CUDA float32 tensor [P,3] -> CUDA uint8 tensor [P] (asynchronous)."""
        torch = _torch()
        pos_dev = pos_dev.contiguous()
        hit = torch.empty(pos_dev.shape[0], dtype=torch.uint8, device=pos_dev.device)
        check(_lib.lib().sage_collide_grid_dev(self._h, _ptr(pos_dev), int(pos_dev.shape[0]), int(bool(np2_promote)),
                                               _ptr(hit), _stream_ptr(stream)))
        return hit


def _first_hits(hit, segments):
    out, s = [], 0
    for n in segments:
        idx = np.nonzero(hit[s:s + n])[0]
        out.append(int(idx[0]) if idx.size else -1)
        s += n
    return np.asarray(out, np.int32)


class CollisionBodies:
    """Device-resident collision ellipsoids (handle of ``sage_bodies``)."""

    def __init__(self, mu, log_scale, quat_wxyz, device: int = 0):
        _lib.require_gpu()
        mu, ls, q = (_as_f32(a) for a in (mu, log_scale, quat_wxyz))
        self.m = int(mu.shape[0])
        self._h = C.c_void_p()
        check(_lib.lib().sage_bodies_create(C.byref(self._h), int(device), self.m, _ptr(mu), _ptr(ls), _ptr(q)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sage_bodies_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, positions, radius: float = 0.1, z0: float = 0.6, z1: float = 1.1):
        """Agent capsule (simple_env.py:238,754-755: r = 0.1, z in [0.6, 1.1]) at each position.
        Returns (hit bool [P], first_body int32 [P])."""
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        P = pos.shape[0]
        hit = np.zeros(P, np.uint8)
        first = np.full(P, -1, np.int32)
        for s in range(0, P, 256):
            e = min(P, s + 256)
            check(_lib.lib().sage_collide_capsule(self._h, _ptr(pos[s:e]), e - s, float(radius), float(z0), float(z1),
                                                  _ptr(hit[s:e]), _ptr(first[s:e]), C.c_void_p(0)))
        return hit.astype(bool), first


class JpegEncoder:
    """Baseline JPEG encoder on the GPU (handle of ``sage_jpeg``).  Output bytes are identical to
    Pillow's ``Image.fromarray(rgb).save(buf, format='JPEG', quality=quality)`` — the call the
    reference makes on the host for every frame (generate_images.py:431-432,
    vlm_client_modular.py:351,402,453)."""

    def __init__(self, width: int, height: int, quality: int = 75, max_batch: int = 4, device: int = 0,
                 max_bytes_per_frame: int = 0):
        _lib.require_gpu()
        self.width, self.height, self.quality = int(width), int(height), int(quality)
        self.max_batch, self.device = int(max_batch), int(device)
        self._h = C.c_void_p()
        check(_lib.lib().sage_jpeg_create(C.byref(self._h), self.device, self.width, self.height, self.quality,
                                          self.max_batch, int(max_bytes_per_frame)))
        mb = C.c_int64(0)
        check(_lib.lib().sage_jpeg_max_bytes(self._h, C.byref(mb)))
        self.max_bytes = int(mb.value)
        self._bufs = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sage_jpeg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _host_buf(self, slot):
        torch = _torch()
        if slot not in self._bufs:
            self._bufs[slot] = torch.zeros(self.max_bytes, dtype=torch.uint8).pin_memory()
        return self._bufs[slot]

    def encode(self, rgb, stream=None):
        """rgb: CUDA uint8 tensor [H,W,3] or [B,H,W,3] -> list of ``bytes`` (one JPEG file each)."""
        torch = _torch()
        if rgb.dim() == 3:
            rgb = rgb.unsqueeze(0)
        if tuple(rgb.shape[1:]) != (self.height, self.width, 3) or rgb.dtype != torch.uint8 or not rgb.is_cuda:
            raise ValueError(f"expected a CUDA uint8 tensor [B,{self.height},{self.width},3]")
        rgb = rgb.contiguous()
        B = int(rgb.shape[0])
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds max_batch {self.max_batch}")
        key = ("batch", B)
        if key not in self._bufs:
            self._bufs[key] = (torch.zeros((B, self.max_bytes), dtype=torch.uint8).pin_memory(),
                               torch.zeros(B, dtype=torch.int64).pin_memory())
        out, sizes = self._bufs[key]
        with torch.cuda.device(self.device):
            check(_lib.lib().sage_jpeg_encode_host(self._h, _ptr(rgb), B, _ptr(out), self.max_bytes, _ptr(sizes),
                                                   _stream_ptr(stream)))
        o = out.numpy()
        return [o[f, :int(sizes[f])].tobytes() for f in range(B)]

    def encode_device(self, rgb, out=None, sizes=None, stream=None):
        """Asynchronous, device to device: returns (out u8 [B,max_bytes], sizes i64 [B]) CUDA tensors."""
        torch = _torch()
        if rgb.dim() == 3:
            rgb = rgb.unsqueeze(0)
        rgb = rgb.contiguous()
        B = int(rgb.shape[0])
        if out is None:
            out = torch.empty((B, self.max_bytes), dtype=torch.uint8, device=rgb.device)
        if sizes is None:
            sizes = torch.empty(B, dtype=torch.int64, device=rgb.device)
        with torch.cuda.device(self.device):
            check(_lib.lib().sage_jpeg_encode(self._h, _ptr(rgb), B, _ptr(out), int(out.shape[1]), _ptr(sizes),
                                              _stream_ptr(stream)))
        return out, sizes

    def debug(self, frame: int = 0):
        """Stage-level parity: (coefficients int16 [blocks,64], MCU bit offsets u32, total bits)."""
        nmcu = ((self.width + 15) // 16) * ((self.height + 15) // 16)
        coef = np.zeros((nmcu * 6, 64), np.int16)
        off = np.zeros(nmcu, np.uint32)
        tot = np.zeros(1, np.uint32)
        check(_lib.lib().sage_jpeg_debug(self._h, int(frame), _ptr(coef), _ptr(off), _ptr(tot)))
        return coef, off, int(tot[0])


class DepthPngEncoder:
    """Depth map -> 16-bit greyscale PNG on the GPU (handle of ``sage_png``): the pixels are bit-identical to
    the reference client's ``(np.clip(depth, 0.0, 6.5) * 10000.0).astype(np.uint16)``
    (vlm_client_modular.py:354-355); the file is this encoder's own valid PNG of them, not OpenCV's bytes
    (csrc/png.cu explains why)."""

    def __init__(self, width: int, height: int, max_batch: int = 4, device: int = 0):
        _lib.require_gpu()
        self.width, self.height, self.max_batch, self.device = int(width), int(height), int(max_batch), int(device)
        self._h = C.c_void_p()
        check(_lib.lib().sage_png_create(C.byref(self._h), self.device, self.width, self.height, self.max_batch))
        mb = C.c_int64(0)
        check(_lib.lib().sage_png_max_bytes(self._h, C.byref(mb)))
        self.max_bytes = int(mb.value)
        self._bufs = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sage_png_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _host_buf(self, slot):
        torch = _torch()
        if slot not in self._bufs:
            self._bufs[slot] = torch.zeros(self.max_bytes, dtype=torch.uint8).pin_memory()
        return self._bufs[slot]

    def encode(self, depth, lo: float = 0.0, hi: float = 6.5, scale: float = 10000.0, want_pixels: bool = False, stream=None):
        """depth: CUDA float32 tensor [H,W] -> PNG file ``bytes`` (and the uint16 pixels if asked)."""
        torch = _torch()
        if tuple(depth.shape) != (self.height, self.width) or depth.dtype != torch.float32 or not depth.is_cuda:
            raise ValueError(f"expected a CUDA float32 tensor [{self.height},{self.width}]")
        depth = depth.contiguous()
        out = self._host_buf("encode")
        n = C.c_int64(0)
        twin = np.zeros((self.height, self.width), np.uint16) if want_pixels else None
        with torch.cuda.device(self.device):
            check(_lib.lib().sage_depth_png_encode_host(self._h, _ptr(depth), float(lo), float(hi), float(scale), _ptr(out),
                                                        self.max_bytes, C.byref(n), _ptr(twin), _stream_ptr(stream)))
        data = out.numpy()[:n.value].tobytes()
        return (data, twin) if want_pixels else data


def set_pipeline_depth(depth: int) -> None:
    """Frame contexts (frames in flight) of scenes created from now on (1..8; default 3)."""
    check(_lib.lib().sage_set_pipeline_depth(int(depth)))


# ---- functional surface (north_star) ---------------------------------------------------------

_scene_cache: dict = {}


def render(camera: Camera, gaussians, want_depth: bool = True, rows=None):
    """``render(camera, gaussians) -> (rgb u8 [H,W,3], depth f32 [H,W])`` as CUDA tensors.

    ``gaussians`` is a resident :class:`GaussianScene`, or raw attributes (GaussianCloud / 5-tuple
    of arrays or tensors) which are uploaded once and cached by identity.
    """
    if isinstance(gaussians, GaussianScene):
        scene = gaussians
    else:
        key = id(gaussians)
        scene = _scene_cache.get(key)
        if scene is None:
            scene = GaussianScene(gaussians)
            _scene_cache.clear()           # keep at most one implicit scene alive
            _scene_cache[key] = scene
            scene._keepalive = gaussians
    return scene.render(camera, rows=rows, want_depth=want_depth)


def collide(agent_pose, collision, **kw):
    """``collide(agent_pose, collision) -> hit mask``.

    ``agent_pose``: position(s) [P,3]; ``collision``: :class:`ObstacleGrid` (2-D semantic-map rule)
    or :class:`CollisionBodies` (ellipsoid-vs-capsule)."""
    if isinstance(collision, ObstacleGrid):
        return collision.check(agent_pose, **kw)
    if isinstance(collision, CollisionBodies):
        return collision.check(agent_pose, **kw)[0]
    raise TypeError("collision must be an ObstacleGrid or CollisionBodies")


__all__ = ["GaussianScene", "ObstacleGrid", "CollisionBodies", "JpegEncoder", "DepthPngEncoder", "set_pipeline_depth", "render", "collide", "SageError", "Camera"]
