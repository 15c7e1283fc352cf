#!/usr/bin/env python
"""Depth hand-off timing: the device depth-PNG encoder on a 1080p frame of the bench scene against
the reference client's host path (vlm_client_modular.py:354-364: np.clip, * 10000, astype(uint16), cv2.imencode)."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import sage3d_official_b200 as sage  # noqa: E402
from sage3d_official_b200 import _lib  # noqa: E402

cloud, layout = sage.synthetic_interior(1, 2_000_000)
scene = sage.GaussianScene(cloud)
scene.set_depth_clip(0.1, 6.5)
p, h = sage.camera_path(layout, 8, seed=0)[3]
cam = sage.Camera.look_heading(p, h, 1920, 1080)
rgb, depth = scene.render(cam)
torch.cuda.synchronize()
enc = sage.DepthPngEncoder(1920, 1080)
out = torch.empty(enc.max_bytes, dtype=torch.uint8, device="cuda")
size = torch.zeros(1, dtype=torch.int64, device="cuda")
L = _lib.lib()


def go():
    L.sage_depth_png_encode(enc._h, C.c_void_p(depth.data_ptr()), C.c_float(0.0), C.c_float(6.5), C.c_float(10000.0),
                            C.c_void_p(out.data_ptr()), enc.max_bytes, C.c_void_p(size.data_ptr()), C.c_void_p(0),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))


for _ in range(5):
    go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    go()
e1.record()
torch.cuda.synchronize()
res = {"depth_png_1080p_device_us": e0.elapsed_time(e1) / 50 * 1e3, "png_bytes": int(size.item()), "raw_bytes": 1920 * 1080 * 2}
d = depth.cpu().numpy()
try:
    import cv2
    t0 = time.perf_counter()
    for _ in range(5):
        q = (np.clip(d, 0.0, 6.5) * 10000.0).astype(np.uint16)
        ok, buf = cv2.imencode('.png', q)
    res["host_numpy_cv2_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    res["cv2_png_bytes"] = int(buf.size)
    dev = cv2.imdecode(np.frombuffer(out[:int(size.item())].cpu().numpy(), np.uint8), cv2.IMREAD_UNCHANGED)
    res["decoded_pixels_equal"] = bool(np.array_equal(dev, q))
except ImportError:
    pass
print(json.dumps(res))
