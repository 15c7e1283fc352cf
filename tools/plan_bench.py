#!/usr/bin/env python
"""Map/planning row: batched A* on the 12 x 10 m synthetic floor at 5 cm
(241 x 201 cells, robot radius 0.2 m as vln_trajectory_generator.py:64) — GPU queries/s against the
CPU oracle (C port of astar_pixel, one core) on a bounded sample.  Prints one JSON line.

    python tools/plan_bench.py [--queries 4096] [--slots 1024] [--scale 0.05]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=16384)
    ap.add_argument("--slots", type=int, default=8192)
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--cpu-queries", type=int, default=64)
    ap.add_argument("--budget-gb", type=float, default=4.0, help="device memory the search slots may take")
    args = ap.parse_args()
    import sage3d_official_b200 as sage
    from sage3d_official_b200 import planner
    from sage3d_official_b200.scenes import layout_semantic_map
    import oracle
    _, layout = sage.synthetic_interior(1, 1000)
    sem = layout_semantic_map(layout)
    planner.inflate_grid(np.zeros((8, 8), np.uint8), args.scale, 0.2)        # CUDA context + first-launch costs are not the map's
    t0 = time.perf_counter()
    raw, _, _, _ = planner.build_2d_map(sem, args.scale, 0.0)                 # host: parse the 2-decimal strings, rasterise
    t_parse = time.perf_counter() - t0
    t0 = time.perf_counter()
    grid, _, _, _ = planner.build_2d_map(sem, args.scale, 0.2)
    t_map = time.perf_counter() - t0
    t0 = time.perf_counter()
    lab = planner.connectivity_labels(grid)
    t_cc = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    free = np.argwhere(grid == 0)
    # reachable pairs only (what the generator plans after its connectivity pre-filter)
    main_lab = np.bincount(lab[lab >= 0]).argmax()
    cells = free[lab[free[:, 0], free[:, 1]] == main_lab]
    a = cells[rng.integers(len(cells), size=args.queries)][:, ::-1].copy()
    b = cells[rng.integers(len(cells), size=args.queries)][:, ::-1].copy()
    pl = planner.GridPlanner(grid, slots=args.slots, memory_budget=int(args.budget_gb * (1 << 30)))
    pl.plan(a[:args.slots], b[:args.slots])                      # warm-up (allocations, capacity growth)
    t0 = time.perf_counter()
    paths, nexp = pl.plan(a, b, want_expansions=True)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    ok = 0
    for k in range(args.cpu_queries):
        want = oracle.astar(grid, a[k], b[k])
        ok += int(want is not None and np.array_equal(np.asarray(want, np.int32), paths[k]))
    dt_cpu = time.perf_counter() - t1
    print(json.dumps({
        "workload": f"A* on {grid.shape[0]}x{grid.shape[1]} grid ({args.scale} m cells, r=0.2 m), {args.queries} reachable start/goal pairs",
        "gpu_queries_per_s": args.queries / dt, "gpu_ms_total": 1e3 * dt, "slots": pl.slots, "slots_asked": args.slots, "budget_gb": args.budget_gb,
        "mean_path_len": float(np.mean([len(p) for p in paths if p is not None])),
        "mean_expansions": float(nexp.mean()), "gpu_expansions_per_s": float(nexp.sum() / dt),
        "cpu_oracle_queries_per_s_1core": args.cpu_queries / dt_cpu, "cpu_sample": args.cpu_queries,
        "cpu_paths_identical": ok == args.cpu_queries,
        "build_2d_map_ms": 1e3 * t_map, "build_2d_map_host_parse_ms": 1e3 * t_parse, "connectivity_labels_ms": 1e3 * t_cc,
    }))


if __name__ == "__main__":
    main()
