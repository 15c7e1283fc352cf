"""GPU collide parity: K6 (grid) bit-exact against the reference's golden vectors and the oracle;
K7 (capsule) bit-exact against the oracle (parity unpinned: the reference has no such test)."""
import glob
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(str(Path(__file__).parent / "golden" / "collide_grid_*.npz")))


def _grid(sage, z):
    b = z["bounds"]
    return sage.ObstacleGrid(z["grid"], b[0], b[1], b[2], b[3], b[4]), b


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_grid_matches_reference_golden(sage, oracle, path):
    z = np.load(path)
    grid, b = _grid(sage, z)
    np2 = int(str(z["numpy_version"]).split(".")[0]) >= 2
    pos = z["pos32"]
    got32 = np.concatenate([grid.check(pos[s:s + 1000], np2_promote=np2) for s in range(0, len(pos), 1000)])
    assert np.array_equal(got32, z["hit_f32_native"])
    got_np1 = np.concatenate([grid.check(pos[s:s + 1000], np2_promote=False) for s in range(0, len(pos), 1000)])
    assert np.array_equal(got_np1, z["hit_f64"])
    p64 = pos.astype(np.float64)
    got64 = np.concatenate([grid.check(p64[s:s + 1000]) for s in range(0, len(pos), 1000)])
    assert np.array_equal(got64, z["hit_f64"])
    # and the oracle agrees on fresh random positions in both modes
    rng = np.random.default_rng(5)
    q = rng.uniform(-12, 12, size=(1000, 3)).astype(np.float32)
    for mode in (True, False):
        assert np.array_equal(grid.check(q, np2_promote=mode),
                              oracle.collide_grid(z["grid"], b[0], b[1], b[2], b[3], b[4], q, np2_promote=mode))


def test_detector_dropin_matches_golden(sage, tmp_path):
    from sage3d_official_b200.collision_detector import SemanticMap2DCollisionDetector
    z = np.load(GOLDEN[0])
    p = tmp_path / "map.json"
    p.write_text(str(z["map_json"]))
    radius = float(z["bounds"][5])
    det = SemanticMap2DCollisionDetector(str(p), robot_radius_m=radius, scale=0.05)
    assert np.array_equal(det.obstacle_map, z["grid"])
    info, ref_info = det.get_collision_info(), json.loads(str(z["info"]))
    for k in ("map_size", "obstacle_pixels", "total_pixels", "map_bounds"):
        assert info[k] == ref_info[k]
    pos = z["pos32"][:600]
    assert np.array_equal(det.check_many(pos), z["hit_f32_native"][:600])
    for i in (0, 5, 100, 599):
        assert det.check_collision_3d(pos[i]) == bool(z["hit_f32_native"][i])
        assert det.check_collision_3d(pos[i].astype(np.float64)) == bool(z["hit_f64"][i])
    assert det.check_collision_at_position(0.0, 0.0) in (True, False)
    missing = SemanticMap2DCollisionDetector(str(tmp_path / "nope.json"))
    assert missing.check_collision_3d(np.zeros(3, np.float32)) is False


def test_first_hit_segments(sage):
    z = np.load(GOLDEN[1])
    grid, _ = _grid(sage, z)
    pos = z["pos32"][:60]
    segs = [20, 10, 10, 10, 10]
    hit, first = grid.check(pos, np2_promote=True, segments=segs)
    s = 0
    for n, f in zip(segs, first):
        idx = np.nonzero(hit[s:s + n])[0]
        assert f == (idx[0] if idx.size else -1)
        s += n
    with pytest.raises(sage.SageError):
        grid.check(pos, np2_promote=True, segments=[1, 2, 3])


def test_device_resident_grid_query(sage, oracle):
    import torch
    z = np.load(GOLDEN[2])
    grid, b = _grid(sage, z)
    pos = torch.from_numpy(z["pos32"]).cuda()
    hit = grid.check_device(pos, np2_promote=True).cpu().numpy().astype(bool)
    assert np.array_equal(hit, oracle.collide_grid(z["grid"], b[0], b[1], b[2], b[3], b[4], z["pos32"], np2_promote=True))


@pytest.mark.parametrize("seed", [0, 1])
def test_capsule_matches_oracle(sage, oracle, seed):
    from sage3d_official_b200.scenes import collision_ellipsoids
    _, layout = sage.synthetic_interior(1, 10)
    mu, ls, q = collision_ellipsoids(layout, seed=seed, per_m2=200.0)
    assert mu.shape[0] > 10_000, mu.shape
    bodies = sage.CollisionBodies(mu, ls, q)
    rng = np.random.default_rng(seed)
    P = 300
    pos = np.stack([rng.uniform(-0.5, 12.5, P), rng.uniform(-0.5, 10.5, P), np.full(P, 0.85)], 1).astype(np.float32)
    hit, first = bodies.check(pos)
    o_hit, o_first = oracle.collide_capsule(mu, ls, q, pos)
    assert np.array_equal(hit, o_hit) and np.array_equal(first, o_first)
    assert 0.02 < hit.mean() < 0.98
    # other capsule parameters
    hit2, first2 = bodies.check(pos[:60], radius=0.25, z0=0.1, z1=1.8)
    o2 = oracle.collide_capsule(mu, ls, q, pos[:60], radius=0.25, z0=0.1, z1=1.8)
    assert np.array_equal(hit2, o2[0]) and np.array_equal(first2, o2[1])
    assert hit2.sum() >= hit[:60].sum()      # a fatter, taller capsule never hits less


def test_capsule_edge_cases(sage, oracle):
    mu = np.array([[0, 0, 0.85], [5, 5, 0.85]], np.float32)
    ls = np.log(np.array([[0.2, 0.1, 0.3], [0.1, 0.1, 0.1]], np.float32))
    q = np.array([[1, 0, 0, 0], [0, 0, 0, 0]], np.float32)      # second body has a degenerate quaternion
    bodies = sage.CollisionBodies(mu, ls, q)
    pos = np.array([[0, 0, 0], [0.29, 0, 0], [0.31, 0, 0], [5, 5, 0], [100, 100, 0]], np.float32)
    hit, first = bodies.check(pos)
    o_hit, o_first = oracle.collide_capsule(mu, ls, q, pos)
    assert np.array_equal(hit, o_hit) and np.array_equal(first, o_first)
    assert hit[0] and hit[1] and not hit[2] and not hit[4]
    empty = sage.CollisionBodies(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 4)))
    h, f = empty.check(pos)
    assert not h.any() and (f == -1).all()
