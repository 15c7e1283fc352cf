"""CPU: the map/planning oracle (oracle/plan_oracle.c) is PINNED against golden vectors recorded
from the UNMODIFIED reference (vln_trajectory_generator.py build_2d_map / build_connectivity_map /
astar_pixel; tests/golden/make_plan_golden.py) and, where /root/reference exists, against the live
reference and scipy."""
import glob
import math
import sys
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(str(Path(__file__).parent / "golden" / "plan_*.npz")))
REF = Path("/root/reference/Code/data_pipeline/trajectory_generation")


@pytest.mark.parametrize("path", GOLDEN, ids=[Path(p).stem for p in GOLDEN])
def test_oracle_matches_reference_goldens(oracle, path):
    z = np.load(path)
    grid = z["grid"]
    assert np.array_equal(oracle.grid_inflate(z["raw"], float(z["scale"]), float(z["radius"])), grid)
    lab = oracle.grid_components(grid)
    assert np.array_equal(lab < 0, grid != 0)
    sl = lab[z["sample_yx"][:, 0], z["sample_yx"][:, 1]]
    g = z["sample_group"]
    for a in range(len(g)):                       # same partition of the sampled free cells
        assert np.array_equal(sl == sl[a], g == g[a])
    for k in range(len(z["starts"])):
        want = z["path_xy"][z["path_off"][k]:z["path_off"][k + 1]]
        got = oracle.astar(grid, z["starts"][k], z["goals"][k])
        if z["path_len"][k] == 0:
            assert got is None, k
        else:
            assert got is not None and np.array_equal(np.asarray(got, np.int32), want), k


def test_goldens_cover_the_cases():
    assert len(GOLDEN) >= 4
    n_path = n_none = 0
    for p in GOLDEN:
        z = np.load(p)
        n_path += int((z["path_len"] > 0).sum()); n_none += int((z["path_len"] == 0).sum())
        assert (z["path_len"] == 1).any()          # start == goal
    assert n_path >= 30 and n_none >= 10


def test_hypot_of_integers_is_correctly_rounded_sqrt():
    # the restatement (and the CUDA kernel) use sqrt((double)(dx*dx+dy*dy)) for math.hypot(dx, dy)
    for x in range(0, 1100, 7):
        for y in range(0, 1100, 3):
            assert math.hypot(x, y) == math.sqrt(float(x * x + y * y))
    assert math.hypot(1, 1) == math.sqrt(2.0) == math.hypot(-1, 1)


@pytest.mark.parametrize("rs", [(0.08, 0.05), (0.2, 0.05), (0.25, 0.05), (0.1, 0.05), (0.3, 0.1), (0.5, 0.05), (0.07, 0.07), (1.0, 0.05)])
def test_inflate_equals_scipy(oracle, rs):
    from scipy.ndimage import distance_transform_edt
    r, s = rs
    rng = np.random.default_rng(int(r * 1000 + s * 100))
    for density in (0.002, 0.02, 0.2):
        raw = (rng.uniform(size=(70, 93)) < density).astype(np.uint8)
        raw[5, 5] = 1
        want = (distance_transform_edt(raw == 0, sampling=s) <= r).astype(np.uint8)
        assert np.array_equal(oracle.grid_inflate(raw, s, r), want), (rs, density)


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present")
def test_oracle_matches_live_reference_astar(oracle):
    import contextlib, io
    sys.path.insert(0, str(REF))
    with contextlib.redirect_stdout(io.StringIO()):
        import vln_trajectory_generator as ref
    rng = np.random.default_rng(42)
    grid = (rng.uniform(size=(48, 64)) < 0.28).astype(np.uint8)       # dense clutter: many ties and detours
    free = np.argwhere(grid == 0)
    for _ in range(40):
        a, b = free[rng.integers(len(free))], free[rng.integers(len(free))]
        s, g = (int(a[1]), int(a[0])), (int(b[1]), int(b[0]))
        assert oracle.astar(grid, s, g) == ref.astar_pixel(grid, s, g), (s, g)
    empty = np.zeros((30, 40), np.uint8)                                # open floor: maximal tie-breaking
    for s, g in [((0, 0), (39, 29)), ((39, 0), (0, 29)), ((5, 20), (35, 3)), ((20, 29), (20, 0))]:
        assert oracle.astar(empty, s, g) == ref.astar_pixel(empty, s, g), (s, g)
