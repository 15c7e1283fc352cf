"""GPU: map building and planning kernels (csrc/plan.cu, through the C-ABI and the planner.py
mirror of the reference's names) against the goldens recorded from the unmodified reference and
against the pinned oracle.  Integer / binary64 work: everything is compared exactly."""
import glob
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(str(Path(__file__).parent / "golden" / "plan_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[Path(p).stem for p in GOLDEN])
def test_reference_goldens(sage, oracle, path):
    from sage3d_official_b200 import planner
    z = np.load(path)
    grid = z["grid"]
    assert np.array_equal(planner.inflate_grid(z["raw"], float(z["scale"]), float(z["radius"])), grid)
    lab = planner.connectivity_labels(grid)
    assert np.array_equal(lab, oracle.grid_components(grid))
    sl = lab[z["sample_yx"][:, 0], z["sample_yx"][:, 1]]
    g = z["sample_group"]
    for a in range(len(g)):
        assert np.array_equal(sl == sl[a], g == g[a])
    paths = planner.astar_pixel_batch(grid, z["starts"], z["goals"])
    for k, got in enumerate(paths):
        want = z["path_xy"][z["path_off"][k]:z["path_off"][k + 1]]
        if z["path_len"][k] == 0:
            assert got is None, k
        else:
            assert got is not None and np.array_equal(np.asarray(got, np.int32), want), k
    # the single-query drop-in
    k = int(np.argmax(z["path_len"]))
    got = planner.astar_pixel(grid, tuple(int(v) for v in z["starts"][k]), tuple(int(v) for v in z["goals"][k]))
    assert [tuple(p) for p in z["path_xy"][z["path_off"][k]:z["path_off"][k + 1]].tolist()] == got


def _interior_grid(sage, scale=0.05, radius=0.2):
    from sage3d_official_b200 import planner
    from sage3d_official_b200.scenes import layout_semantic_map
    _, layout = sage.synthetic_interior(1, 1000)
    sem = layout_semantic_map(layout)
    grid, s, min_x, min_y = planner.build_2d_map(sem, scale, radius)
    return grid, sem


def test_full_size_map_batch_vs_oracle(sage, oracle):
    """12 x 10 m floor at 5 cm (241 x 201): 300 searches in one call; paths and expansion counts
    identical to the oracle; inflation identical to scipy."""
    from scipy.ndimage import distance_transform_edt
    from sage3d_official_b200 import planner
    grid, sem = _interior_grid(sage)
    raw, _, _, _ = planner.build_2d_map(sem, 0.05, 0.0)
    assert np.array_equal(grid, (distance_transform_edt(raw == 0, sampling=0.05) <= 0.2).astype(np.uint8))
    assert np.array_equal(planner.inflate_grid(raw, 0.05, 0.08), oracle.grid_inflate(raw, 0.05, 0.08))
    rng = np.random.default_rng(0)
    free = np.argwhere(grid == 0)
    a, b = free[rng.integers(len(free), size=300)], free[rng.integers(len(free), size=300)]
    starts, goals = a[:, ::-1].copy(), b[:, ::-1].copy()
    pl = planner.GridPlanner(grid, slots=128)
    paths, nexp = pl.plan(starts, goals, want_expansions=True)
    lab = planner.connectivity_labels(grid)
    n_found = 0
    for k in range(300):
        want, ne = oracle.astar(grid, starts[k], goals[k], want_expansions=True)
        assert ne == nexp[k], k
        if want is None:
            assert paths[k] is None
            continue
        n_found += 1
        p = paths[k]
        assert np.array_equal(p, np.asarray(want, np.int32)), k
        # properties: endpoints, 8-connected unit steps, never on an obstacle after the start
        assert tuple(p[0]) == tuple(starts[k]) and tuple(p[-1]) == tuple(goals[k])
        if len(p) > 1:
            assert np.abs(np.diff(p, axis=0)).max() == 1 and (grid[p[1:, 1], p[1:, 0]] == 0).all()
        # 4-connected component equality implies reachability; 8-connected search may reach more
        if lab[starts[k][1], starts[k][0]] == lab[goals[k][1], goals[k][0]]:
            assert paths[k] is not None
    assert n_found > 100


def test_capacity_retry_and_edge_queries(sage, oracle):
    from sage3d_official_b200 import planner
    grid = np.zeros((40, 300), np.uint8)
    grid[5:35, 150] = 1
    pl = planner.GridPlanner(grid, slots=4, max_path=8, heap_capacity=1500)    # far too small: forces the re-run
    qs = [((0, 0), (299, 39)), ((10, 20), (10, 20)), ((0, 0), (150, 20)), ((-3, 0), (5, 5)), ((0, 0), (400, 5))]
    got = pl.plan([q[0] for q in qs], [q[1] for q in qs])
    assert np.array_equal(got[0], np.asarray(oracle.astar(grid, *qs[0]), np.int32))
    assert got[1].tolist() == [[10, 20]]
    assert got[2] is None                       # goal inside an obstacle
    assert got[3] is None                       # start outside the grid: nothing to expand
    assert got[4] is None                       # goal outside the grid: exhausts the floor, None
    assert pl.max_path > 8 and pl.heap_capacity > 1500


def test_connectivity_map_groups(sage, oracle):
    from sage3d_official_b200 import planner
    grid, sem = _interior_grid(sage)
    items = {str(i.get("item_id", k)): i for k, i in enumerate(sem)}
    groups = planner.build_connectivity_map(grid, items)
    lab = oracle.grid_components(grid)
    for key, members in groups.items():
        roots = set()
        for m in members:
            f = planner.get_nearest_free_pixel_on_side(items[m]["mask_coords"], grid)
            roots.add(int(lab[f[1], f[0]]))
        assert len(roots) == 1
    assert sum(len(v) for v in groups.values()) > 0
