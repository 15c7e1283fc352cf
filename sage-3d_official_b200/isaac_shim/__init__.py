"""Seam 2 — the handful of Isaac Sim symbols ``generate_images.py`` imports
(Code/data_pipeline/training_data_construction/generate_images.py:25-33), backed by the engine,
so that script runs unmodified:

    import sage3d_official_b200.isaac_shim as shim; shim.install()
    runpy.run_path(".../generate_images.py", run_name="__main__")

Provided (and nothing more): ``omni.isaac.kit.SimulationApp`` (+ ``isaacsim.simulation_app``),
``omni.usd.get_context().{open/close_stage,get_stage}``, ``omni.isaac.core.World{reset,step,clear}``,
``omni.isaac.core.utils.stage.open_stage(usd_path=)``, ``omni.isaac.sensor.Camera(prim_path,
frequency, resolution){initialize,set_world_pose,get_world_pose,get_rgba,get_current_frame,
add_distance_to_image_plane_to_frame}``, ``pxr.UsdGeom.Camera(prim).GetFocalLengthAttr().Set``,
``pxr.UsdLux.DomeLight.Define(stage, path).Create{Intensity,Color}Attr``, ``pxr.Gf.Vec3f``,
``stage.GetPrimAtPath``.

``world.step(render=True)`` is a no-op: a frame is rendered when pixels are asked for (the
reference steps 3 RTX frames per image, generate_images.py:49,425).  Lighting prims are accepted and
ignored (Gaussians carry their radiance).
"""
from __future__ import annotations

import sys
import types

import numpy as np

from ..camera import Camera as _RasterCamera, FOCAL_LENGTH_MM, USD_GAUSS_MODEL_TO_WORLD
from ..framepool import FramePool
from ..handoff import PrivateTwin

_state = {"stage": None, "jpeg_handoff": False}


class _Attr:
    def __init__(self, owner=None, name=None, value=None):
        self._owner, self._name, self._value = owner, name, value

    def Set(self, v):
        self._value = v
        if self._owner is not None and self._name is not None:
            setattr(self._owner, self._name, v)
        return True

    def Get(self):
        return self._value


class _Prim:
    def __init__(self, path, obj=None):
        self.path, self.obj = path, obj

    def IsValid(self):
        return True

    def __bool__(self):
        return True


class Stage:
    """The opened scene: Gaussians resident on the GPU + a prim table."""

    def __init__(self, usd_path, device=0):
        from ..api import GaussianScene
        from ..scene_io import load_cloud, resolve_stage_assets
        self.usd_path = str(usd_path)
        gauss, _, is_stage = resolve_stage_assets(self.usd_path)
        if gauss is None:
            raise FileNotFoundError(f"no Gaussian payload (.ply/.npz) found for {usd_path}")
        self.model_to_world = USD_GAUSS_MODEL_TO_WORLD if is_stage else None
        self.scene = GaussianScene(load_cloud(gauss), device=device)
        self.scene.set_rgba(True)                  # Camera.get_rgba() frames come out of the composite as RGBA
        self._pool = self._pool_shape = self._scratch = None
        self.prims = {"/World": _Prim("/World")}

    def GetPrimAtPath(self, path):
        return self.prims.get(str(path))          # None (falsy) when absent, like an invalid prim

    def close(self):
        if self.scene is not None:
            self.scene.close()
            self.scene = None


class _UsdContext:
    def get_stage(self):
        return _state["stage"]

    def close_stage(self):
        if _state["stage"] is not None:
            _state["stage"].close()
        _state["stage"] = None
        return True

    def open_stage(self, path):
        return open_stage(path)


_context = _UsdContext()


def get_context():
    return _context


def open_stage(usd_path=None, **kw):
    try:
        _context.close_stage()
        _state["stage"] = Stage(usd_path)
        return True
    except Exception as e:
        print(f"[isaac_shim] open_stage failed: {e}")
        _state["stage"] = None
        return False


class SimulationApp:
    def __init__(self, config=None, *a, **kw):
        self.config = dict(config or {})

    def update(self):
        pass

    def is_running(self):
        return True

    def close(self):
        _context.close_stage()


class World:
    def __init__(self, *a, **kw):
        self.steps = 0

    def reset(self):
        pass

    def step(self, render=True):
        self.steps += 1

    def clear(self):
        pass


class Camera:
    """``omni.isaac.sensor.Camera`` stand-in: pose in, frame out."""

    def __init__(self, prim_path, frequency=30, resolution=(640, 480), **kw):
        self.prim_path, self.frequency = prim_path, frequency
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self.focal_length = FOCAL_LENGTH_MM
        self.clipping_range = (0.1, 50.0)
        self._position = np.zeros(3, np.float32)
        self._orientation = np.array([1.0, 0.0, 0.0, 0.0], np.float32)
        self._depth_enabled = False
        self._frame_key = None
        self._frame = None
        self._served = False
        st = _state["stage"]
        if st is not None:
            st.prims[str(prim_path)] = _Prim(str(prim_path), self)

    def initialize(self, *a, **kw):
        pass

    def add_distance_to_image_plane_to_frame(self):
        self._depth_enabled = True

    def set_world_pose(self, position=None, orientation=None, **kw):
        if position is not None:
            self._position = np.asarray(position, np.float32).reshape(3).copy()
        if orientation is not None:
            self._orientation = np.asarray(orientation, np.float32).reshape(4).copy()

    def get_world_pose(self):
        return self._position.copy(), self._orientation.copy()

    def set_focal_length(self, v):
        self.focal_length = float(v)

    def set_clipping_range(self, near, far):
        self.clipping_range = (float(near), float(far))

    def _buffers(self, st):
        """A pinned (rgba, depth, twin) buffer triple for this resolution: lent to the caller zero-copy and
        taken back when the caller's arrays are garbage-collected; a scratch triple (and copies) when the
        caller keeps every frame."""
        W, H = self.resolution
        pool = getattr(st, "_pool", None)
        if pool is None or st._pool_shape != (H, W):
            import torch
            scene = st.scene

            def make():
                rgba, depth = scene.alloc_host_frame(H, W)
                return rgba, depth, torch.zeros((H, W, 3), dtype=torch.uint8).pin_memory()

            pool = st._pool = FramePool(make)
            st._pool_shape, st._scratch = (H, W), None
        triple = pool.take()
        if triple is None:
            if st._scratch is None:
                st._scratch = pool._make()
            return st._scratch, False
        return triple, True

    def _render(self):
        st = _state["stage"]
        if st is None or st.scene is None:
            return None
        key = (self._position.tobytes(), self._orientation.tobytes(), self.resolution, self.focal_length,
               self._depth_enabled)
        if key != self._frame_key:
            cam = _RasterCamera.from_isaac_pose(self._position, self._orientation, self.resolution[0],
                                                self.resolution[1], model_to_world=st.model_to_world,
                                                focal_mm=self.focal_length, near=self.clipping_range[0],
                                                far=self.clipping_range[1])
            scene = st.scene
            (rgba_t, dep_t, twin_t), lent = self._buffers(st)
            handoff = _state["jpeg_handoff"] and (self.resolution[0] * self.resolution[1]) % 4 == 0
            jpeg = None
            for _ in range(2):
                if handoff:
                    # frame hand-off: the JPEG generate_images.py:431-432 is about to make, made on the device
                    from ..api import JpegEncoder
                    enc = getattr(st, "_encoder", None)
                    if enc is None or (enc.width, enc.height) != self.resolution:
                        enc = st._encoder = JpegEncoder(self.resolution[0], self.resolution[1], max_batch=2)
                    scene.render_jpeg_async(cam, enc, 0, want_rgb=True, want_depth=self._depth_enabled,
                                            out=(rgba_t, dep_t), twin=twin_t)
                    jpeg = scene.jpeg_wait(enc, 0)
                else:
                    scene.render_host_async(cam, 0, want_depth=self._depth_enabled, out=(rgba_t, dep_t))
                    scene.host_wait(0)
                stats = scene.stats()
                if not stats["overflow"]:
                    break
                # the (tile, Gaussian) lists did not fit (camera inside geometry): grow and redo
                scene.reserve(int(stats["n_instances_wanted"] * 1.25) + 1024)
            rgba, depth, twin = rgba_t.numpy(), (dep_t.numpy() if self._depth_enabled else None), None
            if jpeg is not None:
                twin = twin_t.numpy()
            if not lent:
                rgba, depth, twin = rgba.copy(), (None if depth is None else depth.copy()), (None if twin is None else twin.copy())
            else:
                st._pool.lend((rgba_t, dep_t, twin_t), tuple(a for a in (rgba, depth, twin) if a is not None))
            if twin is not None:
                twin = twin.view(PrivateTwin)
            self._frame = (rgba, depth, twin, jpeg)
            self._frame_key = key
            self._served = False
        return self._frame

    def get_rgba(self):
        """RGBA u8 [H,W,4] straight from the composite's epilogue (alpha 255): no host-side expansion."""
        if self._served:
            self._frame_key = None            # the caller may have drawn on the array it was given: render afresh
        f = self._render()
        if f is None:
            return None
        self._served = True
        rgba, _, twin, jpeg = f
        from ..handoff import DeviceFrame
        return DeviceFrame(rgba, jpeg, twin, lambda: True, rgba_root=True)

    def get_current_frame(self):
        if not self._depth_enabled:
            self._depth_enabled = True        # Isaac delivers the annotators that were added; here depth is one flag away
        if self._served:
            self._frame_key = None
        f = self._render()
        out = {"rendering_time": 0.0, "rendering_frame": 0}
        if f is not None:
            self._served = True
            out["rgba"] = f[0]
            out["distance_to_image_plane"] = f[1]
        return out


class _UsdGeomCamera:
    def __init__(self, prim):
        self._cam = getattr(prim, "obj", None)

    def GetFocalLengthAttr(self):
        return _Attr(self._cam, "focal_length", getattr(self._cam, "focal_length", FOCAL_LENGTH_MM))

    def GetClippingRangeAttr(self):
        return _Attr(self._cam, "clipping_range", getattr(self._cam, "clipping_range", (0.1, 50.0)))


class _DomeLight:
    def __init__(self, path):
        self.path = path

    @classmethod
    def Define(cls, stage, path):
        light = cls(str(path))
        if stage is not None:
            stage.prims[str(path)] = _Prim(str(path), light)
        return light

    def CreateIntensityAttr(self, v=None):
        return _Attr(value=v)

    def CreateColorAttr(self, v=None):
        return _Attr(value=v)


class Vec3f(tuple):
    def __new__(cls, *a):
        return super().__new__(cls, a[0] if len(a) == 1 else a)


def install(jpeg_handoff=None):
    """Register the stand-in modules under the names the reference imports.

    jpeg_handoff (default: the SAGE_JPEG_HANDOFF environment variable): frames come back as
    handoff.DeviceFrame and ``Image.fromarray(frame[:, :, :3].copy()).save('x.jpg')``
    (generate_images.py:431-432) writes the JPEG the device already encoded."""
    from ..handoff import handoff_requested, install_pillow_hook
    _state["jpeg_handoff"] = handoff_requested() if jpeg_handoff is None else bool(jpeg_handoff)
    if _state["jpeg_handoff"]:
        install_pillow_hook()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    omni = mod("omni")
    omni.usd = mod("omni.usd", get_context=get_context)
    isaac = mod("omni.isaac")
    omni.isaac = isaac
    isaac.kit = mod("omni.isaac.kit", SimulationApp=SimulationApp)
    core = mod("omni.isaac.core", World=World)
    isaac.core = core
    core.utils = mod("omni.isaac.core.utils")
    core.utils.stage = mod("omni.isaac.core.utils.stage", open_stage=open_stage)
    isaac.sensor = mod("omni.isaac.sensor", Camera=Camera)
    mod("isaacsim")
    mod("isaacsim.simulation_app", SimulationApp=SimulationApp)
    pxr = mod("pxr")
    pxr.Gf = mod("pxr.Gf", Vec3f=Vec3f)
    pxr.UsdGeom = mod("pxr.UsdGeom", Camera=_UsdGeomCamera)
    pxr.UsdLux = mod("pxr.UsdLux", DomeLight=_DomeLight)
    return sys.modules["omni"]
