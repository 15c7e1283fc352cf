"""Development harness (no GPU): the hostile-parameter scenes of tests/test_render_parity_gpu.py::test_randomised_stress_parity
through the EMULATED render pipeline (tools/cuda_emu), bit for bit against the CPU oracle.  Small whole frames: the composite's
E = 4 + box-test instantiation by default (SAGE_COMPOSITE_PREFILTER / SAGE_COMPOSITE_WIDE_MAX_CTAS select the others).

    python tools/cuda_emu/run_stress_emu.py 0,1,2,3
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tools" / "cuda_emu"):
    sys.path.insert(0, str(p))
import numpy as np
import run_render_emu as R
import sage3d_official_b200 as sage
import oracle
oracle.build()
L = R.load()
ok_all = True
for seed in [int(x) for x in sys.argv[1].split(",")]:
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 3000)); deg = int(rng.integers(0, 4))
    W, H = int(rng.integers(17, 400)), int(rng.integers(17, 300))
    cloud, _ = sage.synthetic_room(seed, n, sh_degree=deg)
    cloud.log_scale[:] = rng.uniform(-9.0, 1.5, size=(n, 3)).astype(np.float32)
    cloud.opacity[:] = rng.uniform(-8.0, 8.0, size=n).astype(np.float32)
    cloud.quat[:] = rng.normal(size=(n, 4)).astype(np.float32)
    cloud.sh[:] = rng.normal(0, 0.6, size=cloud.sh.shape).astype(np.float32)
    k = max(1, n // 50)
    bad = rng.integers(0, n, size=6 * k)
    cloud.quat[bad[:k]] = 0.0
    cloud.xyz[bad[k:2 * k], 0] = np.nan
    cloud.log_scale[bad[2 * k:3 * k], 1] = np.inf
    cloud.opacity[bad[3 * k:4 * k]] = -np.inf
    cloud.sh[bad[4 * k:5 * k], 0, 1] = 1e31
    cloud.sh[bad[5 * k:6 * k], -1, 2] = np.nan
    pos = rng.uniform([0.2, 0.2, 0.3], [5.8, 3.8, 2.7])
    q = rng.normal(size=4)
    cam = sage.Camera.from_isaac_pose(pos, q, W, H, focal_mm=float(rng.uniform(3.0, 30.0)), near=float(rng.uniform(0.01, 0.5)), far=float(rng.uniform(3.0, 80.0)))
    ok = R.check(L, sage, oracle, cloud, cam, f"stress seed {seed} n={n} {W}x{H}")
    ok_all &= ok
print("ALL OK" if ok_all else "FAILURES")
