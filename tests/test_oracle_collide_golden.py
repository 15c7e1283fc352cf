"""CPU: the oracle's grid-collide restatement and the product's host-side grid builder against the
golden vectors produced by the UNMODIFIED reference class (tests/golden/make_collide_golden.py)."""
import glob
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(str(Path(__file__).parent / "golden" / "collide_grid_*.npz")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    b = z["bounds"]
    return z, dict(min_x=b[0], max_x=b[1], min_y=b[2], max_y=b[3], scale=b[4], radius=b[5])


def test_golden_files_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_oracle_matches_reference_float32_native(oracle, path):
    z, b = _load(path)
    np2 = int(str(z["numpy_version"]).split(".")[0]) >= 2
    hit = oracle.collide_grid(z["grid"], b["min_x"], b["max_x"], b["min_y"], b["max_y"], b["scale"],
                              z["pos32"], np2_promote=np2)
    assert np.array_equal(hit, z["hit_f32_native"])


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_oracle_matches_reference_float64_and_numpy1_mode(oracle, path):
    z, b = _load(path)
    args = (z["grid"], b["min_x"], b["max_x"], b["min_y"], b["max_y"], b["scale"])
    assert np.array_equal(oracle.collide_grid(*args, z["pos32"].astype(np.float64)), z["hit_f64"])
    # NumPy 1.26 promotion of float32 inputs == widening the inputs first
    assert np.array_equal(oracle.collide_grid(*args, z["pos32"], np2_promote=False), z["hit_f64"])


def test_promotion_modes_are_distinguishable():
    diff = 0
    for path in GOLDEN:
        z, _ = _load(path)
        diff += int(np.count_nonzero(z["hit_f32_native"] != z["hit_f64"]))
    assert diff > 0, "golden set no longer exercises the float32/float64 promotion difference"


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_host_grid_builder_matches_reference_grid(sage, path):
    from sage3d_official_b200.collision_detector import build_obstacle_grid
    z, b = _load(path)
    grid, (min_x, max_x, min_y, max_y) = build_obstacle_grid(json.loads(str(z["map_json"])), b["radius"], b["scale"])
    assert np.array_equal(grid, z["grid"])
    assert (min_x, max_x, min_y, max_y) == (b["min_x"], b["max_x"], b["min_y"], b["max_y"])


def test_missing_map_answers_false(oracle):
    hit = oracle.collide_grid(None, 0, 1, 0, 1, 0.05, np.zeros((3, 3), np.float32))
    assert not hit.any()
