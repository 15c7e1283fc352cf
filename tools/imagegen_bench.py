#!/usr/bin/env python
"""Batched image generation: the render-and-save loop of
generate_images.py:408-456 — set_world_pose, 3 x world.step(render=True), cam.get_rgba(),
Image.fromarray(img[:, :, :3].copy()).save(traj_dir / name.jpg) at 1024 x 768 — against the Isaac
stand-ins, with and without the device-JPEG hand-off, plus the engine's own batched form
(GaussianScene.render_batch_jpeg).  Prints one JSON line.

    python tools/imagegen_bench.py [--frames 300] [--n 2000000]
"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--n", type=int, default=2_000_000)
    args = ap.parse_args()
    from PIL import Image
    import sage3d_official_b200 as sage
    import sage3d_official_b200.isaac_shim as shim
    from sage3d_official_b200 import handoff
    from sage3d_official_b200.scene_io import save_npz, morton_order

    cloud, layout = sage.synthetic_interior(1, args.n)
    cloud = cloud.subset(morton_order(cloud.xyz))
    tmp = Path(tempfile.mkdtemp())
    save_npz(tmp / "scene.npz", cloud)
    poses = sage.camera_path(layout, args.frames, seed=5)
    out = {}
    for mode in ("pillow_on_host", "jpeg_handoff"):
        shim.install(jpeg_handoff=(mode == "jpeg_handoff"))
        from omni.isaac.core import World
        from omni.isaac.core.utils.stage import open_stage
        from omni.isaac.sensor import Camera
        assert open_stage(usd_path=str(tmp / "scene.npz"))
        world = World()
        cam = Camera(prim_path="/World/NaVILACamera", frequency=30, resolution=(1024, 768))
        cam.initialize()
        traj_dir = tmp / mode
        traj_dir.mkdir()
        t_render = t_save = 0.0
        for idx, (p, h) in enumerate(poses):
            a = time.perf_counter()
            position = np.array(p, dtype=np.float32)
            position[2] = 1.2
            cam.set_world_pose(position=position, orientation=np.array([np.cos(h / 2), 0.0, 0.0, np.sin(h / 2)], dtype=np.float32))
            for _ in range(3):
                world.step(render=True)
            img = cam.get_rgba()
            b = time.perf_counter()
            img_array = img[:, :, :3].copy()
            Image.fromarray(img_array).save(traj_dir / f"s_t_{idx:03d}.jpg")
            c = time.perf_counter()
            if idx > 0:                      # first frame: allocations
                t_render += b - a
                t_save += c - b
        n = len(poses) - 1
        out[mode] = {"images_per_s": n / (t_render + t_save), "get_rgba_ms": 1e3 * t_render / n, "save_ms": 1e3 * t_save / n}
        import omni.usd
        omni.usd.get_context().close_stage()
    shim.install(jpeg_handoff=False)
    handoff.uninstall_pillow_hook()
    same = all((tmp / "pillow_on_host" / f"s_t_{i:03d}.jpg").read_bytes() == (tmp / "jpeg_handoff" / f"s_t_{i:03d}.jpg").read_bytes()
               for i in range(len(poses)))
    # the engine's own batched form: every pose -> JPEG bytes, two frames in flight, files written by the host
    scene = sage.GaussianScene(cloud)
    enc = sage.JpegEncoder(1024, 768)
    cams = [sage.Camera.look_heading((p[0], p[1], 1.2), h, 1024, 768) for p, h in poses]
    list(scene.render_batch_jpeg(cams[:4], enc))
    t0 = time.perf_counter()
    d = tmp / "batched"
    d.mkdir()
    for i, data in scene.render_batch_jpeg(cams, enc):
        (d / f"s_t_{i:03d}.jpg").write_bytes(data)
    dt = time.perf_counter() - t0
    out["render_batch_jpeg"] = {"images_per_s": len(cams) / dt}
    out["files_identical_with_and_without_handoff"] = bool(same)
    out["workload"] = f"{len(poses)} waypoints, 1024x768 (generate_images.py:43), N={args.n}, JPEG files written to a tmp dir"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
