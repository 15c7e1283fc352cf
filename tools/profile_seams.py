#!/usr/bin/env python
"""Development diagnostic: cProfile of the two seams' per-frame host work (the Isaac stand-in loop of
generate_images.py with the JPEG hand-off at 1024x768; the SimpleVLNEnv get_rgb/get_depth/apply_cmd_for loop at
1080p), to see what of a step is Python and what is waiting for the device.  Prints the top functions by own time."""
import argparse
import cProfile
import io
import pstats
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def top(pr, n=28):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(n)
    return "\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--n", type=int, default=2_000_000)
    args = ap.parse_args()
    from PIL import Image
    import sage3d_official_b200 as sage
    import sage3d_official_b200.isaac_shim as shim
    from sage3d_official_b200.scene_io import save_npz, morton_order
    cloud, layout = sage.synthetic_interior(1, args.n)
    cloud = cloud.subset(morton_order(cloud.xyz))
    tmp = Path(tempfile.mkdtemp())
    save_npz(tmp / "scene.npz", cloud)
    poses = sage.camera_path(layout, args.frames, seed=5)
    shim.install(jpeg_handoff=True)
    from omni.isaac.core import World
    from omni.isaac.core.utils.stage import open_stage
    from omni.isaac.sensor import Camera
    assert open_stage(usd_path=str(tmp / "scene.npz"))
    world = World()
    cam = Camera(prim_path="/World/NaVILACamera", frequency=30, resolution=(1024, 768))
    cam.initialize()
    d = tmp / "out"
    d.mkdir()

    def loop(ps):
        for idx, (p, h) in enumerate(ps):
            position = np.array(p, dtype=np.float32)
            position[2] = 1.2
            cam.set_world_pose(position=position, orientation=np.array([np.cos(h / 2), 0.0, 0.0, np.sin(h / 2)], dtype=np.float32))
            for _ in range(3):
                world.step(render=True)
            img = cam.get_rgba()
            img_array = img[:, :, :3].copy()
            Image.fromarray(img_array).save(d / f"s_t_{idx:03d}.jpg")

    loop(poses[:8])
    t0 = time.perf_counter()
    loop(poses)
    dt = time.perf_counter() - t0
    pr = cProfile.Profile()
    pr.enable()
    loop(poses)
    pr.disable()
    print(f"=== imagegen with hand-off: {len(poses) / dt:.0f} images/s unprofiled ({1e3 * dt / len(poses):.3f} ms per image)")
    print(top(pr))


if __name__ == "__main__":
    main()
