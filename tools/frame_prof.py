"""Profiling driver: a few 1080p frames of the bench workload (run under ncu; see profiles/README.md)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import sage3d_official_b200 as sage  # noqa: E402
import torch  # noqa: E402
from sage3d_official_b200.scene_io import morton_order  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cloud, layout = sage.synthetic_interior(1, n)
cloud = cloud.subset(morton_order(cloud.xyz))
scene = sage.GaussianScene(cloud)
poses = sage.camera_path(layout, 256, seed=0)
rgb = dep = None
for i in range(frames):
    p, h = poses[(50 + i) % len(poses)]
    rgb, dep = scene.render(sage.Camera.look_heading(p, h, 1920, 1080), out=None if rgb is None else (rgb, dep))
    torch.cuda.synchronize()
print(scene.stats())
