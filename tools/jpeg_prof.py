"""Profiling driver for the JPEG hand-off kernels: one 1080p frame of the synthetic interior,
encoded a few times (run under ncu, see profiles/README.md)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import sage3d_official_b200 as sage  # noqa: E402
import torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
cloud, layout = sage.synthetic_interior(1, n)
scene = sage.GaussianScene(cloud)
pos, heading = sage.camera_path(layout, 8, seed=0)[3]
cam = sage.Camera.look_heading(pos, heading, 1920, 1080)
rgb, _ = scene.render(cam)
torch.cuda.synchronize()
enc = sage.JpegEncoder(1920, 1080, max_batch=8)
for _ in range(3):
    out, sizes = enc.encode_device(rgb)
torch.cuda.synchronize()
print("jpeg bytes", int(sizes[0]))
x8 = rgb.unsqueeze(0).repeat(8, 1, 1, 1).contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B, x in ((1, rgb.unsqueeze(0)), (8, x8)):
    for _ in range(3):
        enc.encode_device(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        enc.encode_device(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"batch {B}: {e0.elapsed_time(e1) / 20 / B * 1e3:.1f} us / frame")
