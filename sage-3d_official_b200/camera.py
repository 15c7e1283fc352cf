"""Camera model of the render path.

Mirrors what the reference hands to Isaac Sim's ``Camera``:

* pose = ``cam.set_world_pose(position, orientation)`` (simple_env.py:1284,
  generate_images.py:419-421); the 4 floats are consumed as scalar-first ``(w, x, y, z)``,
  normalised here, and rotate Isaac's "world" camera axes (+X forward, +Y left, +Z up);
* intrinsics: ``focalLength = 8.0`` (simple_env.py:905, generate_images.py:44,350) over the USD
  default horizontal aperture 20.955 with square pixels -> ``fx = fy = 8/20.955 * W``,
  ``cx = W/2``, ``cy = H/2``; clipping range (0.1, 50.0) (simple_env.py:899);
* the Gaussian prim carries ``rotateXYZ = (-90, 0, 0)`` (Data/template.usda:120), i.e.
  ``p_world = (x, z, -y)_model``; it is folded into the view matrix (``model_to_world``).

The raster camera frame is +x right, +y down, +z forward; all matrices are built in float64
and rounded once to float32, so the oracle and the CUDA path consume identical numbers.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FOCAL_LENGTH_MM = 8.0          # simple_env.py:905
HORIZONTAL_APERTURE_MM = 20.955  # USD default; vln_trajectory_generator.py:733-735
NEAR_FAR = (0.1, 50.0)          # simple_env.py:899

#: template.usda:120 — rotateXYZ(-90, 0, 0) applied to the Gaussian prim.
USD_GAUSS_MODEL_TO_WORLD = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]])


def quat_wxyz_to_matrix(q) -> np.ndarray:
    q = np.asarray(q, dtype=np.float64).reshape(4)
    n = np.linalg.norm(q)
    if not np.isfinite(n) or n == 0.0:
        raise ValueError(f"degenerate orientation {q!r}")
    w, x, y, z = q / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


@dataclass
class Camera:
    """Pinhole camera in the raster convention (model -> camera: ``p_c = R p_m + t``)."""
    R: np.ndarray          # (3,3) float32
    t: np.ndarray          # (3,)  float32
    campos: np.ndarray     # (3,)  float32, camera centre in *model* space
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int
    near: float = NEAR_FAR[0]
    far: float = NEAR_FAR[1]
    meta: dict = field(default_factory=dict)

    @staticmethod
    def intrinsics(width: int, height: int, focal_mm: float = FOCAL_LENGTH_MM,
                   aperture_mm: float = HORIZONTAL_APERTURE_MM):
        f = focal_mm / aperture_mm * width
        return float(np.float32(f)), float(np.float32(f)), width / 2.0, height / 2.0

    @classmethod
    def from_isaac_pose(cls, position, orientation_wxyz, width: int, height: int,
                        model_to_world=None, focal_mm: float = FOCAL_LENGTH_MM,
                        near: float = NEAR_FAR[0], far: float = NEAR_FAR[1]) -> "Camera":
        """Camera from the arguments of Isaac's ``Camera.set_world_pose`` ("world" axes)."""
        Q = quat_wxyz_to_matrix(orientation_wxyz)      # camera(+X fwd,+Y left,+Z up) -> world
        c = np.asarray(position, dtype=np.float64).reshape(3)
        fwd, left, up = Q[:, 0], Q[:, 1], Q[:, 2]
        R_cw = np.stack([-left, -up, fwd])             # world -> raster camera
        M = np.eye(3) if model_to_world is None else np.asarray(model_to_world, dtype=np.float64)
        R = R_cw @ M
        t = -R_cw @ c
        campos = M.T @ c
        fx, fy, cx, cy = cls.intrinsics(width, height, focal_mm)
        return cls(R.astype(np.float32), t.astype(np.float32), campos.astype(np.float32),
                   fx, fy, cx, cy, int(width), int(height), float(near), float(far))

    @classmethod
    def look_heading(cls, position, heading_rad: float, width: int, height: int, **kw) -> "Camera":
        """Level camera looking along ``heading`` (rotation about +Z), the 2-D map heading."""
        q = (np.cos(heading_rad / 2.0), 0.0, 0.0, np.sin(heading_rad / 2.0))
        return cls.from_isaac_pose(position, q, width, height, **kw)

    def tiles(self, tile: int = 16):
        return (self.width + tile - 1) // tile, (self.height + tile - 1) // tile
