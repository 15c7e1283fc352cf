"""Golden vectors for the map/planning row, recorded from the UNMODIFIED
reference functions in /root/reference/Code/data_pipeline/trajectory_generation/
vln_trajectory_generator.py: build_2d_map (:487-559), build_connectivity_map (:639-689),
astar_pixel (:253-286).  Run in the authoring container (needs /root/reference):

    python tests/golden/make_plan_golden.py      # rewrites tests/golden/plan_*.npz
"""
import copy
import io
import contextlib
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/Code/data_pipeline/trajectory_generation")
sys.path.insert(0, str(REF))
with contextlib.redirect_stdout(io.StringIO()):
    import vln_trajectory_generator as ref  # noqa: E402

HERE = Path(__file__).resolve().parent


def sem_map(seed, w_m, h_m, n_boxes, scale=0.05):
    """Semantic-map instances in the on-disk format (2-decimal strings, [y, x] pairs,
    semantic_map_builder.py:176-197): a wall ring, interior partitions with doorways, boxes."""
    rng = np.random.default_rng(seed)
    nx, ny = int(round(w_m / scale)), int(round(h_m / scale))
    occ = np.zeros((ny + 1, nx + 1), bool)
    occ[0, :] = occ[-1, :] = occ[:, 0] = occ[:, -1] = True
    for _ in range(3):                                   # partitions with a door
        if rng.uniform() < 0.5:
            x = rng.integers(nx // 5, 4 * nx // 5); d = rng.integers(2, ny - 20)
            occ[:, x] = True; occ[d:d + 18, x] = False
        else:
            y = rng.integers(ny // 5, 4 * ny // 5); d = rng.integers(2, nx - 20)
            occ[y, :] = True; occ[y, d:d + 18] = False
    insts = [{"category_label": "Wall", "item_id": "wall_0",
              "mask_coords_m": [[f"{y * scale:.2f}", f"{x * scale:.2f}"] for y, x in np.argwhere(occ)]}]
    for k in range(n_boxes):
        bw, bh = rng.integers(3, 14), rng.integers(3, 14)
        x0, y0 = rng.integers(2, nx - bw - 2), rng.integers(2, ny - bh - 2)
        yy, xx = np.mgrid[y0:y0 + bh, x0:x0 + bw]
        label = ["Unable Area", "unable area", "chair", "table", "WALL"][k % 5]
        insts.append({"category_label": label, "item_id": f"{label}_{k}",
                      "mask_coords_m": [[f"{y * scale:.2f}", f"{x * scale:.2f}"] for y, x in zip(yy.ravel(), xx.ravel())]})
    return insts


def record(name, seed, w_m, h_m, n_boxes, radius, n_queries):
    rng = np.random.default_rng(seed + 77)
    insts = sem_map(seed, w_m, h_m, n_boxes)
    with contextlib.redirect_stdout(io.StringIO()):
        raw, scale, min_x, min_y = ref.build_2d_map(copy.deepcopy(insts), robot_radius_m=0.0)
        filled = copy.deepcopy(insts)
        grid, _, _, _ = ref.build_2d_map(filled, robot_radius_m=radius)
    H, W = grid.shape
    free = np.argwhere(grid == 0)
    # connectivity: one single-pixel "item" per sampled free cell -> groups = partition of the sample
    sample = free[rng.choice(len(free), min(150, len(free)), replace=False)]
    items = {f"it{k}": {"mask_coords": [(int(y), int(x))]} for k, (y, x) in enumerate(sample)}
    with contextlib.redirect_stdout(io.StringIO()):
        groups = ref.build_connectivity_map(grid, items)
    group_of = np.full(len(sample), -1, np.int32)
    for gi, members in enumerate(sorted(groups.values(), key=lambda s: min(int(m[2:]) for m in s))):
        for m in members:
            group_of[int(m[2:])] = gi
    # nearest free cell of every instance, plain and "on the side facing" the map centre
    nf_plain, nf_side, mask_off, mask_yx = [], [], [0], []
    for inst in filled:
        m = inst["mask_coords"]
        a = ref.get_nearest_free_pixel_on_side(m, grid)
        b = ref.get_nearest_free_pixel_on_side(m, grid, towards_px=(W // 2, H // 2))
        nf_plain.append((-1, -1) if a is None else a)
        nf_side.append((-1, -1) if b is None else b)
        mask_yx.extend(m)
        mask_off.append(len(mask_yx))
    # A*: free->free pairs, plus the awkward ones
    starts, goals = [], []
    for _ in range(n_queries):
        a, b = free[rng.integers(len(free))], free[rng.integers(len(free))]
        starts.append((int(a[1]), int(a[0]))); goals.append((int(b[1]), int(b[0])))
    occ = np.argwhere(grid == 1)
    o = occ[rng.integers(len(occ))]
    starts.append(starts[0]); goals.append(starts[0])                      # start == goal
    starts.append(starts[1]); goals.append((int(o[1]), int(o[0])))          # goal inside an obstacle -> None
    starts.append((int(o[1]), int(o[0]))); goals.append(goals[2])           # start inside an obstacle (not checked by the reference)
    lens, paths = [], []
    for s, g in zip(starts, goals):
        p = ref.astar_pixel(grid, s, g)
        lens.append(0 if p is None else len(p))
        paths.append(np.zeros((0, 2), np.int32) if p is None else np.asarray(p, np.int32))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    np.savez_compressed(HERE / f"plan_{name}.npz", raw=raw, grid=grid, scale=scale, radius=radius,
                        min_x=min_x, min_y=min_y, nf_plain=np.asarray(nf_plain, np.int32), nf_side=np.asarray(nf_side, np.int32),
                        mask_off=np.asarray(mask_off, np.int64), mask_yx=np.asarray(mask_yx, np.int32), sample_yx=sample.astype(np.int32), sample_group=group_of,
                        starts=np.asarray(starts, np.int32), goals=np.asarray(goals, np.int32),
                        path_len=np.asarray(lens, np.int32), path_off=off,
                        path_xy=np.concatenate(paths).astype(np.int32) if paths else np.zeros((0, 2), np.int32))
    print(name, grid.shape, "obstacles raw/inflated", int(raw.sum()), int(grid.sum()), "groups", len(groups),
          "paths", sum(1 for n in lens if n), "none", sum(1 for n in lens if not n))


if __name__ == "__main__":
    record("rooms_a", 1, 4.0, 3.0, 6, 0.2, 40)      # ROBOT_RADIUS_M of the trajectory generator (:64)
    record("rooms_b", 2, 6.0, 5.0, 10, 0.2, 40)
    record("rooms_c", 3, 5.0, 4.0, 8, 0.08, 32)     # the env's detector radius (simple_env.py:136-140)
    record("rooms_d", 4, 3.0, 3.0, 4, 0.25, 24)     # r/s = 5: the (5,0)/(3,4) tie radius
