"""Drop-in for the reference's ``SemanticMap2DCollisionDetector``
(Code/benchmark/environment_evaluation/collision_detector.py:26-288) whose lookups run on the GPU.

Same constructor, attributes and methods, same answers bit for bit:

* load time (host, once per map): bounds = min/max over every ``mask_coords_m`` entry (:69-77);
  grid ``H = ceil((max_y-min_y)/scale)+1`` (:82-83); cells of instances labelled 'wall' /
  'unable area' (case-insensitive, :104-114) set to 1; inflation ``EDT(free) <= robot_radius`` with
  ``scipy.ndimage.distance_transform_edt(sampling=scale)`` exactly as the reference (:122-125);
* query time (device): ``check_collision_3d`` and the batched ``check_many`` go through
  ``sage_collide_grid`` (K6).  ``check_many`` is what the environment façade uses to evaluate a
  whole 1 cm / 5 mm march (simple_env.py:2116-2234, <= 60 samples) in one launch.

Error convention of the reference is kept: a missing/invalid map leaves ``obstacle_map = None``
and every query answers ``False`` (:200-202); nothing raises out of a query (:235-237).  A grid that
cannot be uploaded (no CUDA device, library missing) raises from the constructor: the engine has no
CPU fallback and must not pretend the way is clear.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

OBSTACLE_LABELS = ("unable area", "wall")


def _quiet() -> bool:
    return bool(os.environ.get("SILENT_LOGGING_MODE", False))


def _say(msg: str) -> None:
    if not _quiet():
        print(msg)


def build_obstacle_grid(instances, robot_radius_m: float, scale: float):
    """Host-side grid construction.  Returns (grid u8 [H,W] | None, (min_x, max_x, min_y, max_y))."""
    ys = [float(c[0]) for inst in instances for c in inst.get("mask_coords_m", [])]
    xs = [float(c[1]) for inst in instances for c in inst.get("mask_coords_m", [])]
    if not xs or not ys:
        return None, (None, None, None, None)
    min_x, max_x, min_y, max_y = min(xs), max(xs), min(ys), max(ys)
    H = int(np.ceil((max_y - min_y) / scale)) + 1
    W = int(np.ceil((max_x - min_x) / scale)) + 1
    grid = np.zeros((H, W), dtype=np.uint8)
    for inst in instances:
        if str(inst.get("category_label", "")).lower() not in OBSTACLE_LABELS:
            continue
        for c in inst.get("mask_coords_m", []):
            px = int(round((float(c[1]) - min_x) / scale))
            py = int(round((float(c[0]) - min_y) / scale))
            if 0 <= py < H and 0 <= px < W:
                grid[py, px] = 1
    if robot_radius_m > 0:
        from scipy.ndimage import distance_transform_edt
        dist = distance_transform_edt(grid == 0, sampling=scale)
        grid = (dist <= robot_radius_m).astype(np.uint8)
    return grid, (min_x, max_x, min_y, max_y)


class SemanticMap2DCollisionDetector:
    """2-D semantic-map collision detector backed by ``sage_collide_grid``."""

    def __init__(self, map_json_path: str, robot_radius_m: float = 0.1, scale: float = 0.05,
                 device: int = 0, np2_promote: Optional[bool] = None):
        self.map_json_path = map_json_path
        self.robot_radius_m = robot_radius_m
        self.scale = scale
        self.min_x = self.max_x = self.min_y = self.max_y = None
        self.obstacle_map = None
        self.map_height = 0
        self.map_width = 0
        self._device = device
        #: float32 positions are combined with the Python-float bounds under the promotion rules
        #: of the interpreter's NumPy (>= 2: binary32, 1.x: binary64) unless forced here.
        self._np2 = (int(np.__version__.split(".")[0]) >= 2) if np2_promote is None else bool(np2_promote)
        self._grid = None
        self._load_map_data()

    # ---- load --------------------------------------------------------------------------------
    def _load_map_data(self) -> None:
        if not self.map_json_path or not os.path.exists(self.map_json_path):
            _say(f"[COLLISION_2D] Warning: Map file does not exist: {self.map_json_path}")
            return
        try:
            with open(self.map_json_path, "r") as f:
                instances = json.load(f)
            grid, (self.min_x, self.max_x, self.min_y, self.max_y) = build_obstacle_grid(
                instances, self.robot_radius_m, self.scale)
            if grid is None:
                _say("[COLLISION_2D] Error: No valid coordinates in map data")
                return
            self.map_height, self.map_width = grid.shape
            self.obstacle_map = grid
        except Exception as e:  # same blanket policy as the reference (:89-92): a bad MAP means no collisions
            _say(f"[COLLISION_2D] Error: Failed to load map data: {e}")
            self.obstacle_map = None
            return
        # a bad DEVICE is not a bad map: no CUDA / no library raises out of the constructor instead of
        # silently answering "no collision" for the rest of the run
        self._upload()

    def _upload(self) -> None:
        from .api import ObstacleGrid
        self._grid = ObstacleGrid(self.obstacle_map, self.min_x, self.max_x, self.min_y, self.max_y,
                                  self.scale, device=self._device)

    # ---- coordinate helpers (host; same formulas as :134-189) ----------------------------------
    def _world_to_pixel(self, x: float, y: float) -> Tuple[int, int]:
        px = int(round((float(x) - self.min_x) / self.scale))
        py = int(round((float(y) - self.min_y) / self.scale))
        return py, px

    def _pixel_to_world(self, px: int, py: int) -> Tuple[float, float]:
        return self.min_x + (px + 0.5) * self.scale, self.min_y + (py + 0.5) * self.scale

    def forward_position_mapping(self, px_3d, py_3d, flip_x: bool = True, flip_y: bool = True,
                                 negate_xy: bool = True):
        if self.min_x is None or self.max_x is None or self.min_y is None or self.max_y is None:
            return px_3d, py_3d
        px, py = px_3d, py_3d
        if negate_xy:
            px, py = -px, -py
        if flip_x:
            px = (self.min_x + self.max_x) - px
        if flip_y:
            py = (self.min_y + self.max_y) - py
        return px, py

    # ---- queries -----------------------------------------------------------------------------
    def check_many(self, positions, segments: Optional[Sequence[int]] = None):
        """Batched ``check_collision_3d`` in one launch.  positions [P,>=2 or 3]; float32 follows the
        NumPy promotion mode, float64 the all-double path.  With ``segments`` also returns the index
        of the first colliding sample of each consecutive segment (-1 if none)."""
        pos = np.asarray(positions)
        if pos.ndim == 1:
            pos = pos[None, :]
        P = pos.shape[0]
        if self.obstacle_map is None or self._grid is None:
            hit = np.zeros(P, dtype=bool)
            if segments is None:
                return hit
            return hit, np.full(len(segments), -1, np.int32)
        if pos.shape[1] < 3:
            pos = np.concatenate([pos, np.zeros((P, 3 - pos.shape[1]), pos.dtype)], axis=1)
        pos = pos[:, :3]
        if pos.dtype != np.float32:       # Python floats / ints / float64 take the all-double path
            pos = pos.astype(np.float64)
        out_hit = np.zeros(P, dtype=bool)
        firsts = None
        if segments is None:
            for s in range(0, P, 1024):
                out_hit[s:s + 1024] = self._grid.check(pos[s:s + 1024], np2_promote=self._np2)
            return out_hit
        if P > 1024:
            raise ValueError("check_many with segments supports at most 1024 positions per call")
        out_hit, firsts = self._grid.check(pos, np2_promote=self._np2, segments=list(segments))
        return out_hit, firsts

    def check_collision_3d(self, pos_3d) -> bool:
        if self.obstacle_map is None:
            return False
        try:
            return bool(self.check_many(np.asarray(pos_3d)[None, :])[0])
        except Exception as e:
            _say(f"[COLLISION_2D] Collision detection error: {e}")
            return False

    def check_path_collision_3d(self, start_pos_3d, end_pos_3d, num_samples: int = 10) -> bool:
        if num_samples <= 1:
            return self.check_collision_3d(end_pos_3d)
        a, b = np.asarray(start_pos_3d), np.asarray(end_pos_3d)
        samples = np.stack([a * (1 - i / float(num_samples)) + b * (i / float(num_samples))
                            for i in range(1, num_samples + 1)])
        try:
            return bool(np.any(self.check_many(samples)))
        except Exception:
            return False

    def check_collision_at_position(self, x: float, y: float) -> bool:
        return self.check_collision_3d(np.array([x, y, 0.0]))

    def get_collision_info(self) -> Dict:
        total = self.map_height * self.map_width
        obst = int(np.sum(self.obstacle_map)) if self.obstacle_map is not None else 0
        return {
            "map_path": self.map_json_path,
            "robot_radius_m": self.robot_radius_m,
            "scale": self.scale,
            "map_bounds": {"x": [self.min_x, self.max_x] if self.min_x is not None else None,
                           "y": [self.min_y, self.max_y] if self.min_y is not None else None},
            "map_size": [self.map_height, self.map_width],
            "obstacle_pixels": obst,
            "total_pixels": total,
            "obstacle_ratio": float(obst) / total if self.obstacle_map is not None and total else 0.0,
        }
