"""Development harness (no GPU): the needle-splat scenes of tests/test_render_parity_gpu.py at 1920 x 1080, every sixth tile row
through the EMULATED render pipeline, against the oracle's BRUTE-FORCE mode (every splat against every pixel: the definition).
This is the configuration on which the oracle's tile mode once dropped a contribution (profiles/r02_variants.md).

    python tools/cuda_emu/run_needle_scan.py 2,5,6 [rebuild]
"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tools" / "cuda_emu", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import run_render_emu as R
import sage3d_official_b200 as sage
import oracle
oracle.build()
L = R.load(rebuild=(len(sys.argv) > 2) or not Path("/tmp/emu/librender_emu.so").exists())
from test_render_parity_gpu import needle_scene
tot = 0
for seed in [int(x) for x in sys.argv[1].split(",")]:
    cloud, cam = needle_scene(sage, seed, 1920, 1080)
    oc = oracle.make_camera(cam.R, cam.t, cam.campos, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
    b_rgb, b_dep, _ = oracle.render(oc, *cloud.as_tuple(), brute=True)
    sc = R.EmuScene(L, cloud)
    t0 = time.time()
    for r in range(0, 68, 6):
        rgb, dep = sc.render(cam, rows=(r, r + 1))
        y0, y1 = r * 16, min(r * 16 + 16, 1080)
        bad = (rgb[y0:y1] != b_rgb[y0:y1]).any(-1) | (dep[y0:y1].view(np.uint32) != b_dep[y0:y1].view(np.uint32))
        if bad.any():
            ys, xs = np.nonzero(bad)
            print(f"seed {seed} tile row {r}: {int(bad.sum())} px differ from BRUTE, e.g. ({xs[0]},{ys[0] + y0})", flush=True)
            tot += int(bad.sum())
    sc.close()
    print(f"seed {seed} scanned in {time.time() - t0:.0f}s", flush=True)
print("TOTAL differing px:", tot)
