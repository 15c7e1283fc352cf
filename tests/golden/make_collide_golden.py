#!/usr/bin/env python
"""Generate golden vectors for the 2-D grid collide path by running the UNMODIFIED reference class
(/root/reference/Code/benchmark/environment_evaluation/collision_detector.py) in the authoring
container.  /root/reference does not exist on the GPU box, so the outputs are committed:

    tests/golden/collide_grid_<case>.npz   (map JSON text, grid, bounds, positions, expected hits)

Expected hits are recorded for
  * float32 positions under the running NumPy's promotion rules (recorded in `numpy_version`;
    NumPy >= 2 evaluates the flip step in binary32),
  * the same positions widened to float64 (what NumPy 1.26, the reference's pinned Isaac
    environment, computes for float32 inputs; also the check_collision_at_position path).
"""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

REF = "/root/reference/Code/benchmark/environment_evaluation"
HERE = Path(__file__).resolve().parent
os.environ["SILENT_LOGGING_MODE"] = "1"
sys.path.insert(0, REF)
from collision_detector import SemanticMap2DCollisionDetector  # noqa: E402  (the reference)


def cells(x0, x1, y0, y1, step=0.05):
    xs = np.arange(round(x0 / step), round(x1 / step) + 1) * step
    ys = np.arange(round(y0 / step), round(y1 / step) + 1) * step
    return [[f"{y:.2f}", f"{x:.2f}"] for y in ys for x in xs]


def inst(label, coords, iid):
    return {"category_id": iid, "category_label": label, "instance_id": iid, "item_id": iid, "mask_coords_m": coords}


def case_wall_ring():
    m = [inst("floor", cells(-2.0, 4.0, -1.5, 2.5), 0)]
    ring = cells(-2.0, 4.0, -1.5, -1.45) + cells(-2.0, 4.0, 2.45, 2.5) + cells(-2.0, -1.95, -1.5, 2.5) + cells(3.95, 4.0, -1.5, 2.5)
    m.append(inst("wall", ring, 1))
    return m


def case_l_corridor():
    m = [inst("floor", cells(0.0, 8.0, 0.0, 6.0), 0)]
    m.append(inst("Wall", cells(0.0, 8.0, 0.0, 0.05) + cells(0.0, 0.05, 0.0, 6.0), 1))
    m.append(inst("unable area", cells(2.0, 8.0, 2.0, 6.0), 2))      # fills the inside of the L
    return m


def case_pillar_and_casing():
    m = [inst("floor", cells(-3.0, 3.0, -3.0, 3.0), 0)]
    m.append(inst("Unable Area", cells(-0.2, 0.2, -0.2, 0.2), 1))    # label casing of semantic_map_builder.py:266
    m.append(inst("WALL", cells(1.5, 1.55, -3.0, 3.0), 2))
    m.append(inst("chair", cells(-2.0, -1.5, -2.0, -1.5), 3))         # not an obstacle class
    return m


def case_offset_bounds():
    # bounds that are not multiples of the cell size and far from the origin
    m = [inst("floor", [[f"{y:.2f}", f"{x:.2f}"] for y in np.arange(10.03, 14.0, 0.05) for x in np.arange(-7.51, -3.0, 0.05)], 0)]
    m.append(inst("wall", [[f"{y:.2f}", f"{x:.2f}"] for y in np.arange(11.03, 11.2, 0.05) for x in np.arange(-7.51, -3.0, 0.05)], 1))
    return m


def positions_for(det, rng, n_rand=4000):
    sx, sy = det.min_x + det.max_x, det.min_y + det.max_y
    W, H, s = det.map_width, det.map_height, det.scale
    P = []
    # random positions over a region larger than the map (3-D frame: x3 = x2 - (min+max))
    lo_x, hi_x = det.min_x - sx - 0.5, det.max_x - sx + 0.5
    lo_y, hi_y = det.min_y - sy - 0.5, det.max_y - sy + 0.5
    P.append(np.stack([rng.uniform(lo_x, hi_x, n_rand), rng.uniform(lo_y, hi_y, n_rand), rng.uniform(0, 2, n_rand)], 1))
    # exact half-cell boundaries (banker's rounding) for a band of cells, plus tiny offsets
    for k in range(-4, W + 4):
        for eps in (0.0, 1e-7, -1e-7, 3e-5, -3e-5):
            x2 = det.min_x + (k + 0.5) * s + eps
            P.append(np.array([[x2 - sx, (det.min_y + 0.5 * H * s) - sy, 0.5]]))
    for k in range(-4, H + 4):
        for eps in (0.0, 1e-7, -1e-7):
            y2 = det.min_y + (k + 0.5) * s + eps
            P.append(np.array([[(det.min_x + 0.5 * W * s) - sx, y2 - sy, 0.5]]))
    # margin rule: cells -3..+2 beyond every edge (collision_detector.py:213-225)
    for k in (-3, -2, -1, 0, W - 1, W, W + 1, W + 2):
        for j in (-3, -2, -1, 0, H - 1, H, H + 1, H + 2):
            P.append(np.array([[det.min_x + k * s - sx, det.min_y + j * s - sy, 0.0]]))
    # far away, and non-finite
    P.append(np.array([[1e6, 0.0, 0.0], [0.0, -1e6, 0.0], [1e30, 1e30, 0.0], [np.nan, 0.0, 0.0], [0.0, np.inf, 0.0], [-np.inf, 0.0, 0.0]]))
    return np.concatenate(P).astype(np.float32)


def main():
    rng = np.random.default_rng(2026)
    cases = {"wall_ring": (case_wall_ring(), 0.08), "l_corridor": (case_l_corridor(), 0.08),
             "pillar_casing": (case_pillar_and_casing(), 0.1), "offset_bounds": (case_offset_bounds(), 0.0)}
    for name, (instances, radius) in cases.items():
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(instances, f)
            path = f.name
        det = SemanticMap2DCollisionDetector(path, robot_radius_m=radius, scale=0.05)
        pos32 = positions_for(det, rng)
        hit32 = np.array([det.check_collision_3d(p) for p in pos32], dtype=bool)
        pos64 = pos32.astype(np.float64)
        hit64 = np.array([det.check_collision_3d(p) for p in pos64], dtype=bool)
        np.savez_compressed(HERE / f"collide_grid_{name}.npz", map_json=json.dumps(instances), grid=det.obstacle_map,
                            bounds=np.array([det.min_x, det.max_x, det.min_y, det.max_y, det.scale, radius]),
                            pos32=pos32, hit_f32_native=hit32, hit_f64=hit64, numpy_version=np.__version__,
                            info=json.dumps(det.get_collision_info()))
        os.unlink(path)
        print(name, det.obstacle_map.shape, "P", len(pos32), "hits32", int(hit32.sum()), "hits64", int(hit64.sum()),
              "differ", int((hit32 != hit64).sum()))
    # missing map: every query is False (collision_detector.py:200-202)
    det = SemanticMap2DCollisionDetector("/nonexistent/map.json")
    assert det.check_collision_3d(np.zeros(3, np.float32)) is False


if __name__ == "__main__":
    main()
