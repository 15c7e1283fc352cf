"""sage-3d_official_b200 — B200-native 3-D Gaussian-splatting render-and-collide engine.

The hot path SAGE-3D delegates to Isaac Sim (render -> RGB/depth, collide -> hit) as hand-written
sm_100a CUDA behind a C-ABI shared library (include/sage_splat.h), with the reference's Python
surface on top:

* :func:`render` / :func:`collide`, :class:`GaussianScene`, :class:`ObstacleGrid`,
  :class:`CollisionBodies` (api.py);
* :class:`SimpleVLNEnv` — drop-in for the reference's environment façade (simple_env.py);
* :class:`SemanticMap2DCollisionDetector` — drop-in for the reference's detector
  (collision_detector.py).

The directory name contains a hyphen; import it as ``sage3d_official_b200`` (alias module at the
repository root).
"""
from ._lib import SageError, LIB_PATH, EXPORTS, build as build_library  # noqa: F401
from .camera import Camera  # noqa: F401
from .scenes import GaussianCloud, synthetic_room, synthetic_interior, camera_path  # noqa: F401
from .api import GaussianScene, ObstacleGrid, CollisionBodies, JpegEncoder, DepthPngEncoder, render, collide, set_pipeline_depth  # noqa: F401

__version__ = "0.1.0"
