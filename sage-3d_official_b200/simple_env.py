"""``SimpleVLNEnv`` — drop-in for the reference's Isaac-Sim environment façade
(Code/benchmark/environment_evaluation/simple_env.py:43), re-backed by the B200 engine.

What ``run_benchmark.py`` / ``measures.py`` touch (SURVEY.md §8b seam 1) is reproduced member for
member; what Isaac did behind it is replaced:

=====================================  ==========================================================
reference (file:line)                   here
=====================================  ==========================================================
``get_rgb`` :1356 (5 RTX frames)        one ``sage_render_host`` frame, RGB u8 (H,W,3)
``get_depth`` :1395 (21 RTX frames)     same frame's planar depth, ``clip(0.1, 6.5)`` (:1578)
``_update_camera_position`` :1196       same pose arithmetic -> ``Camera.from_isaac_pose``
``apply_cmd_for`` :1987                 same march (20 cm cap, 1 cm steps, 4-way 5 mm exploration,
 + ``_try_*_movement`` :2116-2234       efficiency thresholds) but all <= 60 candidate positions are
 + ``_is_position_safe`` :2236          checked in ONE batched grid-collide launch instead of a
                                        Python call + RTX frame + physics step per sample
``_init_collision_detector`` :125       :class:`SemanticMap2DCollisionDetector` (GPU lookups),
                                        ``robot_radius_m=0.08, scale=0.05`` (:136-140)
=====================================  ==========================================================

Shipped-but-dead behaviour is kept dead on purpose: ``_check_collision`` is never
called by the motion code, so ``_total_collision_count`` / ``_collision_detected`` only change when a
caller invokes it explicitly.

Error convention.  Per call, as the reference: render returns ``None`` on a failed frame, collide
``False``; a missing or invalid 2-D map means "no collisions" (collision_detector.py:200-202).
But a broken ENGINE is never papered over: no CUDA device, a missing ``libsage_splat.so``, a scene
path without a Gaussian payload or a failed upload raise from the constructor / ``load_scene`` /
``update_map`` — the reference fails just as loudly when Isaac cannot start, and a benchmark that
completes on black frames and walks through walls is worse than one that stops.

Frames.  ``get_rgb`` / ``get_depth`` hand out views of pinned host buffers from a pool; a buffer goes
back to the pool only when the caller (and every view, PIL image ... derived from it) has dropped
it, so the arrays are as safe to keep as the reference's copies but cost no 14 MB memcpy per step.
After every pose change the next frame is submitted at once (asynchronously), so it renders and
crosses PCIe while the caller is busy with the VLM; ``get_rgb`` then only waits for it.

Candidate positions are generated with the same NumPy expressions as the reference so their dtypes
(float32 on the straight march, float64 on the exploration directions) and values are identical
under whichever NumPy runs; the batched kernel honours the per-dtype promotion rules.
"""
from __future__ import annotations

import atexit
import json
import math
import os
import time
import weakref
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np

from .camera import Camera, USD_GAUSS_MODEL_TO_WORLD
from .collision_detector import SemanticMap2DCollisionDetector


def _debug_print(msg: str) -> None:
    if not os.environ.get("SILENT_LOGGING_MODE", False):
        print(msg)


def _default_resolution(resolution):
    """The reference's callers never pass a resolution (run_benchmark.py:641,2188 -> 640 x 480); a
    launcher can ask for another default (BASELINE configs 2-5 run at 1920 x 1080) without touching
    them: SAGE_RESOLUTION=1920x1080 applies when the constructor is given the reference default."""
    env = os.environ.get("SAGE_RESOLUTION", "").lower().strip()
    if env and tuple(resolution) == (640, 480):
        w, h = env.split("x")
        return (int(w), int(h))
    return resolution


from .framepool import FramePool as _FramePool  # noqa: E402


class SimpleVLNEnv:
    """VLN environment with the reference's public surface, rendered and collided on the GPU."""

    #: test hooks (never set by the product): callables replacing the GPU scene / detector so the
    #: façade's host logic can be exercised without a device — tests back them with the oracle
    _scene_factory = None
    _detector_factory = None

    CAMERA_HEIGHT = 1.2            # simple_env.py:1204
    DEPTH_CLIP = (0.1, 6.5)        # simple_env.py:1578
    AGENT_CENTRE_Z = 0.85          # simple_env.py:238
    AGENT_RADIUS, AGENT_HEIGHT = 0.1, 0.5   # simple_env.py:754-755

    def __init__(self, scene_usd_path: str, headless: bool = True, hz: int = 30,
                 agent_prim_path: str = "/World/AgentCamera", resolution: Tuple[int, int] = (640, 480),
                 map_json_path: str = None, device: int = 0, gaussians=None, collision_bodies=None,
                 jpeg_handoff: Optional[bool] = None) -> None:
        self.scene_usd_path = str(Path(scene_usd_path).resolve())
        self.hz = hz
        self.agent_prim_path = agent_prim_path
        self.resolution = _default_resolution(resolution)
        self.headless = headless
        self.is_stop_called = False
        self.consecutive_collisions = 0
        self._total_collision_count = 0
        self._debug_disable_collision = False
        self._episode_start_time = time.time()
        self._current_time = time.time()
        self._collision_detected = False
        self._log_function = None
        self._map_json_path = map_json_path
        self.semantic_map_path = map_json_path
        self.collision_detector = None
        self._device = device
        self._preloaded = (gaussians, collision_bodies)     # optional in-memory assets (not in the reference)
        self._scene = None
        self._bodies = None
        self._model_to_world = None
        self._frame_key = None
        self._frame = None
        # frame hand-off (handoff.py): also encode every frame to JPEG on the device so an unmodified
        # caller's Image.fromarray(rgb).save(..., 'JPEG') costs a memcpy instead of a 13 ms encode
        from .handoff import handoff_requested, install_cv2_hook, install_pillow_hook
        self._jpeg_handoff = handoff_requested() if jpeg_handoff is None else bool(jpeg_handoff)
        if self._jpeg_handoff:
            install_pillow_hook()
            install_cv2_hook()
        self._jpeg = None
        self._encoder = None
        self._png, self._png_frame, self._png_encoder, self._png_twin = (None, None), None, None, None
        self.render_calls = 0          # engine frames actually rendered (1 per distinct pose)
        self.collide_launches = 0      # batched collide launches
        #: wall-clock accounting of the hot path (seconds / counts); dumped as JSON to $SAGE_PERF_JSON
        self.perf = {"render_s": 0.0, "render_wait_s": 0.0, "collide_s": 0.0, "get_rgb_calls": 0, "get_depth_calls": 0,
                     "apply_cmd_calls": 0, "prefetch_hits": 0, "prefetch_misses": 0, "first_call_t": None, "last_call_t": None}
        self._pos = np.array([0.0, 0.0, self.AGENT_CENTRE_Z], dtype=np.float32)
        self._yaw = 0.0
        self._pool = None
        self._scratch = None
        self._pending = None           # (key, slot, pair) of the frame submitted ahead of the caller
        self._slot = 0
        self._served = {"rgb": False, "depth": False}
        self._init_engine()
        if os.environ.get("SAGE_PERF_JSON"):
            atexit.register(self._dump_perf)

    # ---- logging / time (simple_env.py:96-123) -------------------------------------------------
    def _log(self, msg: str) -> None:
        _debug_print(msg)
        if self._log_function:
            try:
                self._log_function(msg)
            except Exception:
                pass

    def update_time_and_reset_collision(self) -> None:
        self._current_time = time.time()
        self._collision_detected = False

    def reset_episode_time(self) -> None:
        self._episode_start_time = time.time()
        self._current_time = time.time()
        self._collision_detected = False

    def set_log_function(self, log_func) -> None:
        self._log_function = log_func
        self._init_collision_detector(self._map_json_path)      # the reference re-inits here (:119-123)

    def set_collision_detection(self, enabled: bool) -> None:
        self._debug_disable_collision = not enabled

    # ---- engine ---------------------------------------------------------------------------------
    def _init_engine(self) -> None:
        """Upload the scene.  Raises when the engine cannot work (no device, no library, no payload)."""
        from .scene_io import load_cloud, resolve_stage_assets
        make_scene = type(self)._scene_factory
        if make_scene is None:
            from .api import GaussianScene
            make_scene = GaussianScene
        self._drop_pending()
        self._scene = None
        self._bodies = None
        self._pool = None
        self._scratch = None
        self._frame_key = None
        bodies_src = None
        if self._preloaded[0] is not None:
            self._model_to_world = None
            cloud, bodies_src = self._preloaded
        else:
            gauss, coll, is_stage = resolve_stage_assets(self.scene_usd_path)
            # a USDA stage carries rotateXYZ(-90,0,0) on the Gaussian prim (template.usda:120)
            self._model_to_world = USD_GAUSS_MODEL_TO_WORLD if is_stage else None
            if gauss is None:
                raise FileNotFoundError(f"[ENV_INIT] no Gaussian payload (.ply/.npz) found for {self.scene_usd_path}")
            cloud = load_cloud(gauss)
            if coll is not None:
                from .scene_io import load_collision_bodies
                bodies_src = load_collision_bodies(coll)
        self._scene = make_scene(cloud, device=self._device)
        self._scene.set_depth_clip(*self.DEPTH_CLIP)      # clip(0.1, 6.5) applied in the kernel epilogue
        self._log(f"[ENV_INIT] scene {self.scene_usd_path}: {self._scene.n} Gaussians resident on cuda:{self._device}")
        if bodies_src is not None and type(self)._scene_factory is None:
            from .api import CollisionBodies
            self._bodies = CollisionBodies(*bodies_src, device=self._device)

    def _init_collision_detector(self, map_json_path: str) -> None:
        self.collision_detector = None
        if map_json_path and os.path.exists(map_json_path):
            make = type(self)._detector_factory or SemanticMap2DCollisionDetector
            # an unreadable map leaves the detector without a grid (every query False, as the reference);
            # a broken device raises out of here
            self.collision_detector = make(map_json_path, robot_radius_m=0.08, scale=0.05, device=self._device)
            info = self.collision_detector.get_collision_info()
            self._log(f"[COLLISION_2D] map {map_json_path}: {info['obstacle_pixels']}/{info['total_pixels']} obstacle cells")
        else:
            self._log(f"[COLLISION_2D] no usable 2D semantic map ({map_json_path})")

    def load_scene(self, new_scene_usd_path: str) -> bool:
        """Switch scenes (run_benchmark.py:621).  A path without a payload keeps the current scene and
        returns False (the reference's switch is a no-op that returns False, simple_env.py:1104-1108);
        an engine failure raises."""
        from .scene_io import resolve_stage_assets
        path = str(Path(new_scene_usd_path).resolve())
        if self._preloaded[0] is None and resolve_stage_assets(path)[0] is None:
            self._log(f"[SCENE_SWITCH] no Gaussian payload for {path}; keeping {self.scene_usd_path}")
            return False
        old = (self._scene, self._bodies)
        self.scene_usd_path = path
        self._init_engine()
        for h in old:
            if h is not None:
                h.close()
        return True

    def update_map(self, new_map_path: str) -> bool:
        self._map_json_path = new_map_path
        self.semantic_map_path = new_map_path
        self._init_collision_detector(new_map_path)
        return True

    def close(self) -> None:
        self._dump_perf()
        self._drop_pending()
        for h in (self._scene, self._bodies):
            if h is not None:
                h.close()
        self._scene = self._bodies = None

    def _dump_perf(self) -> None:
        path = os.environ.get("SAGE_PERF_JSON")
        if not path:
            return
        try:
            out = dict(self.perf, render_calls=self.render_calls, collide_launches=self.collide_launches,
                       resolution=list(self.resolution))
            with open(path, "w") as f:
                json.dump(out, f)
        except Exception:
            pass

    # ---- pose (simple_env.py:1149-1353) ---------------------------------------------------------
    def set_start_pose(self, position: List[float], rotation_xyzw: List[float]) -> None:
        self.is_stop_called = False
        self.consecutive_collisions = 0
        self._total_collision_count = 0
        self._pos = np.array(position, dtype=np.float32)
        x, y, z, w = rotation_xyzw
        self._original_quaternion = list(rotation_xyzw).copy()
        yaw_new = 2 * math.atan2(-x, w)
        self._yaw = yaw_new - math.pi
        if self._yaw < -math.pi:
            self._yaw += 2 * math.pi
        elif self._yaw > math.pi:
            self._yaw -= 2 * math.pi
        self._initial_yaw = self._yaw
        self._update_camera_position()
        self._prefetch()

    def _update_camera_position(self) -> None:
        camera_pos = self._pos.copy()
        camera_pos[2] = self.CAMERA_HEIGHT
        if hasattr(self, "_original_quaternion"):
            qx_orig, qy_orig, qz_orig, qw_orig = self._original_quaternion
            qx_correction = math.sin(math.radians(-45) / 2)
            base_qx, base_qy, base_qz, base_qw = qx_orig + qx_correction, qy_orig, qz_orig, qw_orig
            yaw_delta = self._yaw - self._initial_yaw
            if abs(yaw_delta) > 0.01:
                qz_d, qw_d = math.sin(yaw_delta / 2.0), math.cos(yaw_delta / 2.0)
                qx = base_qx * qw_d + base_qw * (-qz_d)
                qy = base_qy * qw_d
                qz = base_qz * qw_d
                qw = base_qw * qw_d - base_qx * (-qz_d)
            else:
                qx, qy, qz, qw = base_qx, base_qy, base_qz, base_qw
        else:
            qx, qy, qz, qw = -math.sin(self._yaw / 2.0), 0.0, 0.0, math.cos(self._yaw / 2.0)
        # handed to Isaac as `orientation`, i.e. read scalar-first
        self._cam_position = camera_pos.astype(np.float32)
        self._cam_orientation = np.array([qx, qy, qz, qw], dtype=np.float32)

    def _apply_pose(self) -> None:
        self._update_camera_position()

    def get_agent_pos(self) -> np.ndarray:
        return self._pos.copy()

    def get_yaw(self) -> float:
        return self._yaw

    def get_collision_count(self) -> int:
        return self._total_collision_count

    # ---- render (simple_env.py:1356-1842) -------------------------------------------------------
    def _camera(self) -> Camera:
        W, H = int(self.resolution[0]), int(self.resolution[1])
        return Camera.from_isaac_pose(self._cam_position, self._cam_orientation, W, H,
                                      model_to_world=self._model_to_world)

    def _frame_pair(self):
        """A pinned (rgb, depth) buffer pair for the current resolution from the pool."""
        W, H = int(self.resolution[0]), int(self.resolution[1])
        if self._pool is None or self._pool_shape != (H, W):
            scene = self._scene
            self._pool = _FramePool(lambda: scene.alloc_host_frame(H, W))
            self._pool_shape = (H, W)
            self._scratch = None
        pair = self._pool.take()
        if pair is None:
            # the caller keeps every frame it was given (run_benchmark.py:1422 appends them all to a
            # list): stop lending pinned buffers, land frames in one scratch pair and hand out copies
            if self._scratch is None:
                self._scratch = self._scene.alloc_host_frame(H, W)
            return self._scratch
        return pair

    def _drop_pending(self) -> None:
        if self._pending is not None and self._scene is not None:
            _, slot, pair = self._pending
            try:
                self._scene.host_wait(slot)
            except Exception:
                pass
            if pair is not None and self._pool is not None and pair is not self._scratch:
                self._pool.give_back(pair)
        self._pending = None

    def _pose_key(self):
        return (self._cam_position.tobytes(), self._cam_orientation.tobytes(), tuple(self.resolution))

    def _submit(self, key):
        """Start rendering the frame of the current pose (asynchronous: returns at once)."""
        pair = self._frame_pair()
        slot = self._slot
        self._slot = (self._slot + 1) % max(1, int(getattr(self._scene, "pipeline_depth", 1)))
        self._scene.render_host_async(self._camera(), slot, want_depth=True, out=pair)
        return (key, slot, pair)

    def _prefetch(self) -> None:
        """After a pose change: submit the next frame now, so it renders and crosses PCIe while the
        caller is elsewhere (run_benchmark.py:1473-1640: measures, VLM query)."""
        if self._scene is None or self._jpeg_handoff or not hasattr(self._scene, "render_host_async"):
            return
        try:
            key = self._pose_key()
            if self._frame_key == key or (self._pending is not None and self._pending[0] == key):
                return
            self._drop_pending()
            self._pending = self._submit(key)
        except Exception as e:
            _debug_print(f"[WARN] frame prefetch failed: {e}")
            self._pending = None

    def _render_current(self):
        self._update_camera_position()
        key = self._pose_key()
        if self._frame_key != key:
            t0 = time.perf_counter()
            if self._jpeg_handoff:
                self._frame, self._jpeg = self._render_with_jpeg(self._camera())
            elif not hasattr(self._scene, "render_host_async"):
                self._frame = self._scene.render_host(self._camera(), want_depth=True)
            else:
                for attempt in range(2):
                    if self._pending is not None and self._pending[0] == key:
                        self.perf["prefetch_hits"] += 1
                    else:
                        self.perf["prefetch_misses"] += 1
                        self._drop_pending()
                        self._pending = self._submit(key)
                    _, slot, pair = self._pending
                    self._pending = None
                    self._scene.host_wait(slot)
                    st = self._scene.stats()
                    if not st["overflow"]:
                        break
                    # the (tile, Gaussian) lists did not fit (camera inside geometry): grow and redo
                    if pair is not self._scratch:
                        self._pool.give_back(pair)
                    self._scene.reserve(int(st["n_instances_wanted"] * 1.25) + 1024)
                if pair is self._scratch:
                    rgb, depth = pair[0].numpy().copy(), pair[1].numpy().copy()
                else:
                    rgb, depth = pair[0].numpy(), pair[1].numpy()
                    self._pool.lend(pair, (rgb, depth))
                self._frame = (rgb, depth)
            self.perf["render_s"] += time.perf_counter() - t0
            self._served = {"rgb": False, "depth": False}
            self._frame_key = key
            self.render_calls += 1
        return self._frame

    def _render_with_jpeg(self, cam: Camera):
        from .api import JpegEncoder
        if self._encoder is None or (self._encoder.width, self._encoder.height) != (cam.width, cam.height):
            self._encoder = JpegEncoder(cam.width, cam.height, quality=75, max_batch=2, device=self._device)
        for _ in range(2):
            frame = self._scene.render_jpeg_async(cam, self._encoder, 0, want_rgb=True, want_depth=True)
            jpeg = self._scene.jpeg_wait(self._encoder, 0)
            st = self._scene.stats()
            if not st["overflow"]:
                break
            self._scene.reserve(int(st["n_instances_wanted"] * 1.25) + 1024)    # as render_host does
        return frame, jpeg

    def _depth_png(self, frame):
        """Hand-off of the depth map (vlm_client_modular.py:354-364): the 16-bit PNG of
        ``(clip(depth, 0, 6.5) * 10000).astype(uint16)`` made on the device from the frame still resident in
        the slot, fetched with its uint16 pixels (the hook's reference copy).  Cached per frame."""
        if self._png_frame is frame:
            return self._png
        self._png, self._png_frame = (None, None), frame
        try:
            from .api import DepthPngEncoder
            import torch
            H, W = frame[1].shape
            if self._png_encoder is None or (self._png_encoder.width, self._png_encoder.height) != (W, H):
                self._png_encoder = DepthPngEncoder(W, H, max_batch=2, device=self._device)
                self._png_twin = torch.zeros((H, W), dtype=torch.uint16).pin_memory()
            self._scene.depth_png_async(self._png_encoder, 0, 0.0, 6.5, 10000.0, twin=self._png_twin)
            data = self._scene.depth_png_wait(self._png_encoder, 0)
            self._png = (data, self._png_twin.numpy().copy())
        except Exception as e:
            _debug_print(f"[WARN] depth PNG hand-off failed: {e}")
        return self._png

    def _tick(self, name: str) -> None:
        now = time.perf_counter()
        self.perf[name] += 1
        if self.perf["first_call_t"] is None:
            self.perf["first_call_t"] = now
        self.perf["last_call_t"] = now

    def _hand_out(self, which: int, name: str):
        """The frame's array: zero-copy the first time it is asked for, a private copy afterwards (a
        caller may have drawn on the first one)."""
        arr = self._frame[which]
        if self._jpeg_handoff or self._served[name] or not hasattr(self._scene, "render_host_async"):
            return np.array(arr, copy=True)
        self._served[name] = True
        return arr

    def get_rgb(self) -> Optional[np.ndarray]:
        if self._scene is None:
            return None
        try:
            self._tick("get_rgb_calls")
            frame = self._render_current()
            rgb = self._hand_out(0, "rgb")
            if self._jpeg_handoff and self._jpeg is not None:
                from .handoff import DeviceFrame
                # frame[0] is the engine's pinned copy: pristine until the next render replaces it
                return DeviceFrame(rgb, self._jpeg, frame[0], lambda f=frame: self._frame is f)
            return rgb
        except Exception as e:
            _debug_print(f"[WARN] get_rgb failed: {e}")
            return None

    def get_depth(self) -> Optional[np.ndarray]:
        if self._scene is None:
            return None
        try:
            self._tick("get_depth_calls")
            frame = self._render_current()
            depth = self._hand_out(1, "depth")                # already clipped to [0.1, 6.5] on the device
            if self._jpeg_handoff:
                png, pixels = self._depth_png(frame)
                if png is not None:
                    from .handoff import DeviceDepth
                    return DeviceDepth(depth, png, pixels, lambda f=frame: self._frame is f)
            return depth
        except Exception as e:
            _debug_print(f"[WARN] get_depth failed: {e}")
            return None

    def get_rgbd(self):
        return self.get_rgb(), self.get_depth()

    # ---- collide (simple_env.py:1844-2250) ------------------------------------------------------
    def _is_position_safe(self, pos: np.ndarray) -> bool:
        try:
            if self.collision_detector is not None and self.collision_detector.check_collision_3d(pos):
                return False
            return True
        except Exception:
            return False

    def _check_collision(self, new_pos) -> bool:
        """Kept for API parity; like the reference it is never called by the motion code."""
        try:
            if self._debug_disable_collision:
                return False
            if self.collision_detector is not None:
                if self.collision_detector.check_collision_3d(new_pos):
                    self._collision_detected = True
                    self._total_collision_count += 1
                    return True
                return False
            return self.check_capsule(new_pos)
        except Exception:
            return False

    def check_capsule(self, pos) -> bool:
        """3-D agent-capsule test against the scene's collision ellipsoids (K7), when loaded."""
        if self._bodies is None:
            return False
        z0 = self.AGENT_CENTRE_Z - self.AGENT_HEIGHT / 2
        return bool(self._bodies.check(np.asarray(pos, np.float32)[None, :3], radius=self.AGENT_RADIUS,
                                       z0=z0, z1=z0 + self.AGENT_HEIGHT)[0][0])

    def _safe_mask(self, marches: List[np.ndarray]) -> np.ndarray:
        """safe[k] for every candidate of the given marches ([K,3] arrays), one launch per dtype group
        (float32 / float64: the reference's positions drift between the two, and the grid rule differs)."""
        n = sum(len(m) for m in marches)
        safe = np.ones(n, dtype=bool)
        det = self.collision_detector
        if det is None or n == 0:
            return safe
        try:
            if len(marches) == 1 or all(m.dtype == marches[0].dtype for m in marches):
                arr = marches[0] if len(marches) == 1 else np.concatenate(marches)
                safe[:] = ~det.check_many(arr if arr.dtype == np.float32 else arr.astype(np.float64))
                self.collide_launches += 1
                return safe
            bounds = np.cumsum([0] + [len(m) for m in marches])
            for is32 in (True, False):
                sel = [k for k, m in enumerate(marches) if (m.dtype == np.float32) == is32]
                if not sel:
                    continue
                arr = np.concatenate([marches[k] for k in sel])
                res = ~det.check_many(arr if is32 else arr.astype(np.float64))
                o = 0
                for k in sel:
                    safe[bounds[k]:bounds[k + 1]] = res[o:o + len(marches[k])]
                    o += len(marches[k])
                self.collide_launches += 1
        except Exception:
            safe[:] = False
        return safe

    @staticmethod
    def _march(start_pos: np.ndarray, direction: np.ndarray, max_distance: float, step_size: float):
        """Candidate positions [K,3] and cumulative step lengths of a march: the arithmetic of
        _try_direct_movement / _try_short_movement (``next = current + direction * step`` sample after
        sample, simple_env.py:2150-2234) as one sequential accumulate in the dtypes NumPy gives those
        expressions, so every sample has the reference's bits."""
        total_moved = 0.0
        steps, moved = [], []
        while total_moved < max_distance:
            step_distance = min(step_size, max_distance - total_moved)
            steps.append(step_distance)
            total_moved += step_distance
            moved.append(total_moved)
        if not steps:
            return np.zeros((0, 3), np.result_type(start_pos.dtype, direction.dtype)), moved
        # direction * python_float keeps direction's dtype (the scalar is rounded to it first)
        delta = direction[None, :] * np.asarray(steps, dtype=direction.dtype)[:, None]
        out_dt = np.result_type(start_pos.dtype, direction.dtype)
        seq = np.empty((len(steps) + 1, 3), out_dt)
        seq[0] = start_pos
        seq[1:] = delta
        return np.add.accumulate(seq, axis=0)[1:], moved

    def apply_cmd_for(self, vx: float, vy: float, yaw_rate: float, duration_s: float) -> None:
        t_apply = time.perf_counter()
        self._tick("apply_cmd_calls")
        cos_yaw, sin_yaw = math.cos(self._yaw), math.sin(self._yaw)
        world_vx = vx * cos_yaw - vy * sin_yaw
        world_vy = vx * sin_yaw + vy * cos_yaw
        total_dx = world_vx * duration_s
        total_dy = world_vy * duration_s
        total_dyaw = yaw_rate * duration_s
        old_pos = self._pos.copy()
        intended_movement = np.linalg.norm([total_dx, total_dy])
        if intended_movement > 0.001:
            target_pos = old_pos.copy()
            target_pos[0] += total_dx
            target_pos[1] += total_dy
            actual_movement = self._safe_gradual_movement(old_pos, target_pos, intended_movement)
            movement_efficiency = actual_movement / intended_movement if intended_movement > 0 else 1.0
            if movement_efficiency < 0.3 and intended_movement > 0.05:
                self.consecutive_collisions += 1
            elif movement_efficiency > 0.6:
                self.consecutive_collisions = 0
        self._yaw += total_dyaw
        self._yaw = ((self._yaw + math.pi) % (2 * math.pi)) - math.pi
        self._update_camera_position()
        t_pre = time.perf_counter()
        self.perf["collide_s"] += t_pre - t_apply
        self._prefetch()
        self.perf["prefetch_s"] = self.perf.get("prefetch_s", 0.0) + (time.perf_counter() - t_pre)

    def _safe_gradual_movement(self, start_pos, target_pos, target_distance) -> float:
        current_pos = self._pos.copy()
        direction = target_pos - current_pos
        direction_norm = np.linalg.norm(direction[:2])
        if direction_norm < 0.001:
            return 0.0
        unit_direction = direction / direction_norm
        max_distance = min(0.20, target_distance)

        # --- direct march: <= 20 samples, one launch ---
        cands, _ = self._march(current_pos, unit_direction, max_distance, 0.01)
        safe = self._safe_mask([cands])
        n_ok = int(np.argmin(safe)) if not safe.all() else len(cands)
        reached = cands[n_ok - 1] if n_ok > 0 else current_pos
        if n_ok > 0:
            self._pos = reached
        direct_movement = np.linalg.norm(reached[:2] - current_pos[:2])
        if direct_movement > 0.01:
            return direct_movement

        # --- lateral exploration: 4 directions x <= 10 samples of 5 mm, one launch ---
        blocked = unit_direction
        perpendicular = np.array([-blocked[1], blocked[0], 0])
        exploration_directions = [perpendicular, -perpendicular,
                                  perpendicular * 0.707 + blocked * 0.707,
                                  -perpendicular * 0.707 + blocked * 0.707]
        marches = []
        for d in exploration_directions:
            dn = np.linalg.norm(d[:2])
            if dn > 0.001:
                marches.append(self._march(current_pos, d / dn, 0.05, 0.005))
        safe = self._safe_mask([cands_d for cands_d, _ in marches])
        best_movement, best_pos, k0 = 0.0, current_pos.copy(), 0
        for cands_d, moved_d in marches:
            s = safe[k0:k0 + len(cands_d)]
            k0 += len(cands_d)
            n_ok = int(np.argmin(s)) if not s.all() else len(cands_d)
            if n_ok > 0:
                self._pos = cands_d[n_ok - 1]          # the probing march leaves the agent there (:2206-2222)
            movement = moved_d[n_ok - 1] if n_ok > 0 else 0.0
            if movement > best_movement:
                best_movement = movement
                best_pos = self._pos.copy()
        if best_movement > 0.005:
            self._pos = best_pos
            return best_movement
        return 0.0

    # ---- misc -----------------------------------------------------------------------------------
    @staticmethod
    def write_video(frames, path: str, fps: int = 10) -> None:
        """MP4 (or GIF fallback) of RGB frames; the reference uses imageio (simple_env.py:2716)."""
        frames = [np.asarray(f, np.uint8) for f in frames if f is not None]
        if not frames:
            return
        try:
            import imageio
            imageio.mimsave(path, frames, fps=fps)
            return
        except Exception:
            pass
        try:
            import cv2
            h, w = frames[0].shape[:2]
            vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
            for f in frames:
                vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
            vw.release()
        except Exception as e:
            _debug_print(f"[WARN] write_video failed: {e}")
