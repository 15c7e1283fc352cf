"""CPU: host-side logic of planner.py (the reference's build_2d_map / nearest-free-cell search)
against goldens recorded from the unmodified reference, and the live reference where present."""
import copy
import glob
import sys
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(str(Path(__file__).parent / "golden" / "plan_*.npz")))
REF = Path("/root/reference/Code/data_pipeline/trajectory_generation")


@pytest.mark.parametrize("path", GOLDEN, ids=[Path(p).stem for p in GOLDEN])
def test_nearest_free_pixel_matches_reference_goldens(sage, path):
    from sage3d_official_b200 import planner
    z = np.load(path)
    grid = z["grid"]
    H, W = grid.shape
    for k in range(len(z["nf_plain"])):
        mask = [tuple(int(v) for v in yx) for yx in z["mask_yx"][z["mask_off"][k]:z["mask_off"][k + 1]]]
        a = planner.get_nearest_free_pixel_on_side(mask, grid)
        b = planner.get_nearest_free_pixel_on_side(mask, grid, towards_px=(W // 2, H // 2))
        assert ((-1, -1) if a is None else tuple(int(v) for v in a)) == tuple(z["nf_plain"][k]), k
        assert ((-1, -1) if b is None else tuple(int(v) for v in b)) == tuple(z["nf_side"][k]), k


def test_build_2d_map_without_inflation_needs_no_gpu(sage):
    from sage3d_official_b200 import planner
    sem = [{"category_label": "Wall", "mask_coords_m": [["0.00", "0.00"], ["1.00", "2.00"], ["bad", "1.00"]]},
           {"category_label": "chair", "mask_coords_m": [["0.50", "0.75"]]}]
    grid, s, min_x, min_y = planner.build_2d_map(sem, 0.05, 0.0)
    assert grid.shape == (21, 41) and grid.sum() == 2 and grid[0, 0] == 1 and grid[20, 40] == 1
    assert sem[1]["mask_coords"] == [(10, 15)] and (s, min_x, min_y) == (0.05, 0.0, 0.0)
    assert planner.build_2d_map([{"mask_coords_m": []}]) == (None, None, None, None)


def test_planner_fails_loudly_without_gpu(sage):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sage3d_official_b200 import planner
    with pytest.raises(sage.SageError, match="no CPU fallback"):
        planner.astar_pixel(np.zeros((4, 4), np.uint8), (0, 0), (3, 3))
    with pytest.raises(sage.SageError, match="no CPU fallback"):
        planner.inflate_grid(np.zeros((4, 4), np.uint8), 0.05, 0.2)


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present")
def test_build_2d_map_raw_matches_live_reference(sage):
    import contextlib, io
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(Path(__file__).parent / "golden"))
    with contextlib.redirect_stdout(io.StringIO()):
        import vln_trajectory_generator as ref
        import make_plan_golden as mk
    from sage3d_official_b200 import planner
    sem = mk.sem_map(9, 3.0, 2.5, 5)
    a, b = copy.deepcopy(sem), copy.deepcopy(sem)
    with contextlib.redirect_stdout(io.StringIO()):
        want = ref.build_2d_map(a, robot_radius_m=0.0)
    got = planner.build_2d_map(b, robot_radius_m=0.0)
    assert np.array_equal(got[0], want[0]) and got[1:] == want[1:]
    assert [i["mask_coords"] for i in a] == [i["mask_coords"] for i in b]


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present")
def test_trajectory_points_match_live_reference(sage):
    import contextlib, io
    sys.path.insert(0, str(REF))
    with contextlib.redirect_stdout(io.StringIO()):
        import vln_trajectory_generator as ref
    from sage3d_official_b200 import planner
    z = np.load(GOLDEN[1])
    k = int(np.argmax(z["path_len"]))
    path = [tuple(int(v) for v in p) for p in z["path_xy"][z["path_off"][k]:z["path_off"][k + 1]]]
    for step in (1, 3):
        want = ref.generate_trajectory_points(path, float(z["scale"]), float(z["min_x"]), float(z["min_y"]), sample_step=step)
        got = planner.generate_trajectory_points(path, float(z["scale"]), float(z["min_x"]), float(z["min_y"]), sample_step=step)
        assert got == want                       # dicts of Python floats: exact equality
    assert planner.generate_trajectory_points([(3, 4)], 0.05, 1.0, 2.0)[0]["rotation"] == [0.0, 0.0, 0.0, 1.0]


def test_trajectory_points_known_answer(sage):
    from sage3d_official_b200 import planner
    pts = planner.generate_trajectory_points([(0, 0), (1, 0), (1, 1)], 0.05, -1.0, 2.0)
    assert [p["point"] for p in pts] == ["0", "1", "2"]
    assert pts[0]["position"] == [-1.0 + 0.5 * 0.05, 2.0 + 0.5 * 0.05, 0.5]
    assert pts[0]["rotation"][2] == 0.0 and pts[0]["rotation"][3] == 1.0             # heading +x
    assert abs(pts[1]["rotation"][2] - np.sin(np.pi / 4)) < 1e-15                     # heading +y
    assert pts[2]["rotation"] == [0.0, 0.0, 0.0, 1.0] and pts[2]["focal_length"] == 7.0


def test_planner_slots_follow_the_batch_and_the_memory_budget(sage):
    """GridPlanner sizes its device arrays (12 B per cell + 16 B per heap entry, per search slot) by the largest batch it
    has seen, never beyond its slot limit or its memory budget - a single astar_pixel call must not allocate gigabytes
    (ADVICE round 1), a 16 384-query batch gets 16 384 searches in flight (throughput follows them: run Z)."""
    from sage3d_official_b200.planner import GridPlanner
    p = GridPlanner.__new__(GridPlanner)                 # the sizing rule alone: no device
    p.H, p.W, p.max_slots = 201, 241, 16384
    per_slot = 12 * 201 * 241 + 16 * (201 * 241 // 2 + 1024)
    assert p._fit_slots(0, 1) == 256                      # one query: the smallest array
    assert p._fit_slots(0, 300) == 512                    # next power of two
    assert p._fit_slots(0, 16384) == 16384                # fits the default 16 GiB budget on this map
    assert 16384 * per_slot < GridPlanner.MEMORY_BUDGET
    p.MEMORY_BUDGET = 1 << 30
    assert p._fit_slots(0, 16384) == (1 << 30) // per_slot
    p.H, p.W = 600, 600                                   # a large map: the budget, not the batch, decides
    p.MEMORY_BUDGET = 16 << 30
    big = 12 * 360000 + 16 * (180000 + 1024)
    assert p._fit_slots(0, 16384) == (16 << 30) // big
    assert p._fit_slots(8 * 360000 + 16, 16384) == (16 << 30) // (12 * 360000 + 16 * (8 * 360000 + 16))   # after capacity retries
