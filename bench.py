#!/usr/bin/env python
"""bench.py — frames/sec of the render hot path at 1920x1080 on a ~2M-Gaussian interior.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one frame (RGB u8 + planar depth f32) of BASELINE.json configs[1]:
InteriorGS assets are not in the reference tree and there is no network, so the workload is the
seeded synthetic multi-room interior of SURVEY.md §8d (12 x 10 x 3 m, N = 2,000,000, SH degree 3)
rendered along a seeded 256-pose walk.  Rank 0 prints ONE JSON line.

native arm  (this repository's CUDA path)
  value      frames/s, scene resident in HBM, outputs left on the device.  A block of exactly K frames
             is timed with CUDA events between barriers (max over ranks); the block is repeated until
             >= 0.5 s have been measured and the MEDIAN block is reported (min / max next to it).
  e2e        frames/s through the host-facing call (sage_render_host_async + sage_host_wait: host
             camera in, RGB+depth of EVERY frame copied device->pinned host inside the timed region,
             sage_pipeline_depth() = 5 frames in flight).  N > 1: every rank copies ITS strip over its
             own PCIe link into one shared pinned host frame (no NVLink hop on the host-facing path).
  roofline   dominant kernel by device time (per-stage CUDA events recorded on the launch stream):
             algorithmic bytes per launch / its mean duration, against MEASURED_PEAKS.json's hbm_gbs.
  collide    BASELINE config 3: the P = 60 sweep of one agent step through the grid detector, the
             capsule test against the scene's collision ellipsoids, and the reference's own
             check_collision_3d (collision_detector.py:191) looped on the host beside them.
  batch8     BASELINE config 4: an 8-scene batch, as tile-row shards (every frame split over the ranks)
             and as replicas (whole scenes per rank — the reference's only scale-out).
  cpu_baseline  the CPU oracle (a port of the frozen spec — the reference's own renderer is Isaac
             Sim's closed RTX path and cannot run) on all host cores, a few frames of the workload.

reference arm (--impl reference): the same frames rendered by the CPU oracle on the host cores.
Multi-GPU (torchrun, N ranks): every frame is split into N tile-row blocks; each rank composites its
rows straight into every rank's frame buffer over NVLink (peer stores fused into the composite
kernel's epilogue).  scaling = "strong" (one frame is split).  The assembled frame is checked bit for
bit against a full frame rendered by rank 0 alone ("multigpu_frame_equal").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "frames/sec @1920x1080 InteriorGS 2M Gaussians; HBM GB/s vs roofline"
WORKLOAD = "synthetic 12x10x3 m multi-room interior (seed 1), N=2,000,000, SH deg 3, 1920x1080 RGB+depth, 256-pose walk"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--min-time", type=float, default=0.5, help="repeat the K-step block until this many seconds are measured")
    ap.add_argument("--max-blocks", type=int, default=40)
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the bounded CPU-baseline sample")
    ap.add_argument("--batch-scenes", type=int, default=8, help="scenes of the config-4 batch leg (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-collide", action="store_true")
    ap.add_argument("--no-handoff", action="store_true")
    ap.add_argument("--no-morton", action="store_true", help="skip the scene-ingest Morton reordering")
    return ap.parse_args()


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons, sampled from before the warm-up to the end of the timed
    region (so that even a 10 ms region has samples taken under load)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Samples from here on count as 'under load'."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if self.t_mark is not None and t < self.t_mark:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args, seed: int = 1):
    import sage3d_official_b200 as sage
    t0 = time.time()
    cloud, layout = sage.synthetic_interior(seed=seed, n=args.n)
    if not getattr(args, "no_morton", False):
        # scene ingest: spatially coherent index order (what a loader does once per scene); the
        # oracle and the engine both see the reordered cloud, so parity is unaffected
        from sage3d_official_b200.scene_io import morton_order
        cloud = cloud.subset(morton_order(cloud.xyz))
    poses = sage.camera_path(layout, 256, seed=0)
    cams = [sage.Camera.look_heading(p, h, args.width, args.height) for p, h in poses]
    return sage, cloud, layout, cams, time.time() - t0


def oracle_fps(cloud, cams, frames, threads=None):
    import oracle
    if threads:
        oracle.set_threads(threads)
    idx = np.linspace(0, len(cams) - 1, frames).astype(int)
    t0 = time.perf_counter()
    st = None
    for i in idx:
        c = cams[i]
        oc = oracle.make_camera(c.R, c.t, c.campos, c.fx, c.fy, c.cx, c.cy, c.width, c.height, c.near, c.far)
        _, _, st = oracle.render(oc, *cloud.as_tuple())
    dt = time.perf_counter() - t0
    return frames / dt, st["threads"], dt


def run_reference(args):
    """Reference arm: the CPU oracle on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    sage, cloud, layout, cams, gen_s = build_workload(args)
    cores = os.cpu_count() or 1
    oracle.set_threads(cores)

    def frame(i):
        c = cams[i % len(cams)]
        oc = oracle.make_camera(c.R, c.t, c.campos, c.fx, c.fy, c.cx, c.cy, c.width, c.height, c.near, c.far)
        return oracle.render(oc, *cloud.as_tuple())[2]

    for i in range(args.warmup):
        frame(i)
    t0 = time.perf_counter()
    st = None
    for i in range(args.steps):
        st = frame(args.warmup + i)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_gaussians": args.n, "resolution": [args.width, args.height]},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": int(st["threads"]) if st else cores, "kind": "port",
                         "sample": f"{args.steps} full frames of the same pose walk; CPU oracle (frozen-spec port; "
                                   "the reference's own renderer is Isaac Sim RTX, not runnable)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def collide_leg(sage, layout, device):
    """BASELINE config 3: per-step collide — the P = 60 sample sweep of one agent step (20 x 1 cm forward
    + 4 x 10 x 5 mm lateral, simple_env.py:2116-2234) through the grid detector, the capsule test
    against the collision ellipsoids, and the reference's check_collision_3d looped on the host."""
    import tempfile
    import torch
    from sage3d_official_b200.scenes import write_semantic_map, collision_ellipsoids
    from sage3d_official_b200.collision_detector import SemanticMap2DCollisionDetector
    tmp = Path(tempfile.mkdtemp(prefix="sage_bench_"))
    map_path = write_semantic_map(layout, tmp / "map.json")
    det = SemanticMap2DCollisionDetector(str(map_path), robot_radius_m=0.08, scale=0.05, device=device)
    sx, sy = det.min_x + det.max_x, det.min_y + det.max_y
    rng = np.random.default_rng(0)
    # sweeps from 32 start poses of the walk (trajectory frame: x3 = x2 - (min+max))
    sweeps = []
    for p, h in sage.camera_path(layout, 32, seed=3):
        start = np.array([p[0] - sx, p[1] - sy, 0.85], np.float32)
        d = np.array([np.cos(h), np.sin(h), 0.0], np.float32)
        perp = np.array([-d[1], d[0], 0.0], np.float32)
        pts = [start + d * np.float32(0.01 * (k + 1)) for k in range(20)]
        for e in (perp, -perp, (perp + d) * np.float32(0.707), (d - perp) * np.float32(0.707)):
            pts += [start + e * np.float32(0.005 * (k + 1)) for k in range(10)]
        sweeps.append(np.stack(pts).astype(np.float32))
    P = sweeps[0].shape[0]
    seg = [20, 10, 10, 10, 10]
    for s in sweeps[:4]:
        det.check_many(s, segments=seg)
    reps = 200
    t0 = time.perf_counter()
    masks = []
    for k in range(reps):
        masks.append(det.check_many(sweeps[k % len(sweeps)], segments=seg)[0])
    grid_us = (time.perf_counter() - t0) / reps * 1e6
    # device-resident variant (positions already in HBM): kernel + launch only, CUDA events
    dev = torch.device("cuda", device)
    pos_dev = [torch.from_numpy(s).to(dev) for s in sweeps]
    for pd in pos_dev[:4]:
        det._grid.check_device(pd)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for k in range(reps):
        det._grid.check_device(pos_dev[k % len(pos_dev)])
    ev1.record()
    torch.cuda.synchronize()
    grid_dev_us = ev0.elapsed_time(ev1) / reps * 1e3
    # capsule vs the scene's collision ellipsoids
    mu, ls, q = collision_ellipsoids(layout, per_m2=400.0)
    bodies = sage.CollisionBodies(mu, ls, q, device=device)
    p3 = sweeps[0].copy()
    for _ in range(4):
        bodies.check(p3)
    t0 = time.perf_counter()
    for k in range(reps):
        bodies.check(sweeps[k % len(sweeps)])
    capsule_us = (time.perf_counter() - t0) / reps * 1e6
    # the reference's own class, looped on the host (staged copy: oracle/make_ref.py; else the C port)
    ref_us, ref_kind, equal = None, None, None
    try:
        sys.path.insert(0, str(ROOT / "oracle"))
        import make_ref
        d = make_ref.ref_dir()
        if d is not None:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ref_collision_detector", str(d / "collision_detector.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            ref = mod.SemanticMap2DCollisionDetector(str(map_path), robot_radius_m=0.08, scale=0.05)
            ref_reps = 40
            t0 = time.perf_counter()
            ref_masks = []
            for k in range(ref_reps):
                s = sweeps[k % len(sweeps)]
                ref_masks.append(np.array([ref.check_collision_3d(s[i]) for i in range(P)]))
            ref_us = (time.perf_counter() - t0) / ref_reps * 1e6
            ref_kind = "reference"
            equal = all(np.array_equal(ref_masks[k], masks[k]) for k in range(min(ref_reps, len(masks))))
    except Exception as e:
        ref_kind = f"unavailable ({e})"
    if ref_us is None:
        import oracle
        t0 = time.perf_counter()
        om = []
        for k in range(reps):
            s = sweeps[k % len(sweeps)]
            om.append(oracle.collide_grid(det.obstacle_map, det.min_x, det.max_x, det.min_y, det.max_y, det.scale, s))
        ref_us = (time.perf_counter() - t0) / reps * 1e6
        ref_kind = "port"
        equal = all(np.array_equal(om[k], masks[k]) for k in range(reps))
    return {"P": P, "grid_cells": [int(det.map_height), int(det.map_width)],
            "grid_sweep_us_per_step": grid_us, "grid_sweep_device_us": grid_dev_us,
            "capsule_us_per_step": capsule_us, "capsule_bodies": int(bodies.m),
            "reference_us_per_step": ref_us, "reference_us_per_query": ref_us / P, "reference_kind": ref_kind,
            "masks_equal": bool(equal), "hits_in_sample": int(sum(int(m.sum()) for m in masks[:len(sweeps)])),
            "algorithmic_bytes": {"grid": P * 13, "capsule": int(bodies.m) * 40 + P * 17},
            "note": "grid/capsule: host API incl. H2D/D2H copies and sync (what apply_cmd_for pays); "
                    "grid_sweep_device_us: positions resident, CUDA events; reference: collision_detector.py:191 on one host core"}


def batch_leg(args, sage, world, rank, local, n_scenes, steps=48):
    """BASELINE config 4: a batch of n_scenes scenes (seeds 1..n), one 1080p frame each per step.
    tilerow : every scene resident on every rank, every frame split into tile-row blocks (peer stores)
    replicas: scene s lives on rank s % world and is rendered whole there (the reference's scale-out:
              independent processes per scene, generate_images.py:136-139)."""
    import torch
    import torch.distributed as dist
    from sage3d_official_b200.multigpu import rows_of, PeerFrames
    from sage3d_official_b200.scene_io import morton_order
    W, H = args.width, args.height
    dev = torch.device("cuda", local)
    n_small = args.n                           # full-size scenes (seeds 1..n), generated on every rank
    clouds, cams = [], []
    for s in range(n_scenes):
        cloud, layout = sage.synthetic_interior(seed=1 + s, n=n_small)
        cloud = cloud.subset(morton_order(cloud.xyz))
        clouds.append(cloud)
        cams.append([sage.Camera.look_heading(p, h, W, H) for p, h in sage.camera_path(layout, 16, seed=s)])
    out = {"n_scenes": n_scenes, "n_gaussians_per_scene": n_small, "frames_per_mode": steps * n_scenes}

    from sage3d_official_b200 import _lib as _sage_lib
    depth = int(_sage_lib.lib().sage_pipeline_depth())
    streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]

    def timed(fn, frames):
        for i in range(2 * depth + 2):          # every frame context of every scene has captured its graph
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(2 * depth + 2 + i)
        for st_ in streams:
            torch.cuda.current_stream().wait_stream(st_)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return frames / (ms / 1000.0)

    # replicas: scene s lives on rank s % N and is rendered whole there (the reference's only scale-out,
    # generate_images.py:136-139); frames rotate over the frame contexts / streams as in the main loop
    mine = [s for s in range(n_scenes) if s % world == rank]
    scenes = {s: sage.GaussianScene(clouds[s], device=local) for s in mine}
    rgb = [torch.zeros((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(depth)]
    dep = [torch.zeros((H, W), dtype=torch.float32, device=dev) for _ in range(depth)]

    def step_rep(i):
        b = i % depth
        for s in mine:
            scenes[s].render(cams[s][i % 16], out=(rgb[b], dep[b]), stream=streams[b])
    out["replicas_fps"] = timed(step_rep, steps * n_scenes)
    for sc in scenes.values():
        sc.close()
    # tile rows: every scene resident on every rank, every frame split into N tile-row blocks
    if world > 1:
        scenes = [sage.GaussianScene(c, device=local) for c in clouds]
        rows = rows_of(rank, H, world)
        peers = PeerFrames(rgb, dep, rank, world)

        def step_rows(i):
            b = i % depth
            for s in range(n_scenes):
                scenes[s].render_peers(cams[s][i % 16], rows, peers.rgb[b], peers.depth[b], stream=streams[b])
            peers.close_frame(b, stream=streams[b])
        out["tilerow_fps"] = timed(step_rows, steps * n_scenes)
        peers.close()
        for sc in scenes:
            sc.close()
    else:
        out["tilerow_fps"] = out["replicas_fps"]
    out["note"] = ("frames/s over the whole batch, device-resident; replicas = scene s rendered whole on rank s % N; "
                   "tilerow = every frame split into N tile-row blocks with peer stores + one frame-close per step")
    return out


def run_native(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # before the warm-up: a 10 ms timed region still gets samples under load
    sage, cloud, layout, cams, gen_s = build_workload(args)
    W, H = args.width, args.height
    if world > 1 and "SAGE_PIPELINE_DEPTH" not in os.environ:
        # a rank of the tile-row split renders a fraction of a frame per step: smaller, more latency-bound
        # kernels, which a deeper pipeline hides (4 GPUs: 5773 -> 6489 frames/s with 6 frame contexts)
        sage.set_pipeline_depth(6)
    scene = sage.GaussianScene(cloud, device=local)
    dev = torch.device("cuda", local)
    from sage3d_official_b200.multigpu import rows_of, balanced_cuts
    rows = rows_of(rank, H, world)      # tile-row blocks: 68 rows at 1080p -> 9,9,9,9,8,8,8,8 for 8 ranks
    if world > 1:
        # load-balance the split once, from the per-row instance counts of calibration frames
        # (identical on every rank: same scene, same poses): cost = instances + 0.35 x pixels/16
        ty_all = (H + 15) // 16
        cost = np.zeros(ty_all)
        for ci in np.linspace(0, len(cams) - 1, 8).astype(int):
            scene.render(cams[ci])
            cost += scene.row_instances(ty_all)
        cost = cost / 8 + 0.35 * (W * 16 / 16.0)
        cuts = balanced_cuts(cost, world)
        rows = (cuts[rank], cuts[rank + 1])

    nbuf = int(os.environ.get("SAGE_BENCH_STREAMS", str(scene.pipeline_depth)))   # one stream per frame context
    rgb = [torch.zeros((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    dep = [torch.zeros((H, W), dtype=torch.float32, device=dev) for _ in range(nbuf)]

    peers = None
    if world > 1:
        from sage3d_official_b200.multigpu import PeerFrames
        peers = PeerFrames(rgb, dep, rank, world)

    # consecutive frames rotate over the scene's frame contexts, one stream each, so frame i+1's
    # projection/sorts overlap frame i's compositing
    streams = [torch.cuda.Stream(device=dev) for _ in range(nbuf)]

    def step(i):
        cam = cams[i % len(cams)]
        b = i % nbuf
        if world == 1:
            scene.render(cam, out=(rgb[b], dep[b]), stream=streams[b])
        else:
            scene.render_peers(cam, rows, peers.rgb[b], peers.depth[b], stream=streams[b])
            peers.close_frame(b, stream=streams[b])     # device-side: every rank's strips have landed everywhere

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()

    # ---- timed region: blocks of exactly K steps, repeated until >= min_time, median reported ----
    if rank == 0:
        sampler.mark()
    block_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = args.warmup
    t_wall = time.perf_counter()
    while True:
        sync_all()
        e0.record()
        for i in range(args.steps):
            step(it + i)
        for st_ in streams:
            torch.cuda.current_stream().wait_stream(st_)
        e1.record()
        sync_all()
        it += args.steps
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        block_ms.append(ms)
        done = (sum(block_ms) >= args.min_time * 1000.0) or len(block_ms) >= args.max_blocks
        if world > 1:                   # every rank must take the same decision
            flag = torch.tensor([1 if done else 0], device=dev)
            dist.broadcast(flag, 0)
            done = bool(flag.item())
        if done:
            break
    clocks = sampler.stop() if rank == 0 else None
    ms = float(np.median(block_ms))
    fps = args.steps / (ms / 1000.0)
    graph = scene.graph_stats()

    # ---- multi-GPU: the peer-assembled frame must equal a full frame rendered by one rank ----
    frame_equal = None
    if world > 1:
        cam = cams[7]
        sync_all()
        with torch.cuda.stream(streams[0]):
            scene.render_peers(cam, rows, peers.rgb[0], peers.depth[0])
            peers.close_frame(0)
        sync_all()
        ref_rgb, ref_dep = scene.render(cam)
        torch.cuda.synchronize()
        eq = torch.equal(ref_rgb, rgb[0]) and torch.equal(ref_dep.view(torch.int32), dep[0].view(torch.int32))
        t = torch.tensor([1 if eq else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)          # on EVERY rank: all buffers hold the same, right frame
        frame_equal = bool(t.item())

    # ---- per-stage device times without inter-frame overlap (one stream, direct launches) ----
    scene.set_stage_timing(48)
    for i in range(48):
        scene.render(cams[(args.warmup + i) % len(cams)], rows=rows if world > 1 else None, out=(rgb[0], dep[0]))
    stage_serial = {k: 0.0 for k in scene.STAGES}
    serial_frames = []
    for s_ in range(48):
        tot = 0.0
        for k, v in scene.stage_times(s_).items():
            stage_serial[k] += v / 48
            tot += v
        serial_frames.append(tot)
    scene.set_stage_timing(0)
    # latency of ONE frame (no frames in flight), replayed from the graph: what a synchronous caller sees
    lat = []
    for i in range(40):
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        scene.render(cams[(args.warmup + i) % len(cams)], rows=rows if world > 1 else None, out=(rgb[0], dep[0]))
        b_.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b_))
    frame_latency_ms = float(np.median(lat[8:]))

    # frame-averaged counters over a sample of the walk (V, D vary per pose)
    Vs, Ds, Bs = [], [], []
    st = None
    for i in np.linspace(0, len(cams) - 1, 16).astype(int):
        scene.render(cams[i], rows=rows if world > 1 else None, out=(rgb[0], dep[0]))
        st = scene.stats()
        Vs.append(st["n_visible"]); Ds.append(st["n_instances"]); Bs.append(st["algorithmic_bytes"])
        assert st["overflow"] == 0, "instance capacity overflow"
    V, D, B = float(np.mean(Vs)), float(np.mean(Ds)), float(np.mean(Bs))

    # ---- e2e: host-facing call, frames delivered to (pinned) host memory inside the timed region ----
    e2e_steps = max(30, min(args.steps, 200))
    d = scene.pipeline_depth
    e2e_equal = None
    if world == 1:
        for i in range(2 * d):
            scene.render_host_async(cams[i], i % d)
            scene.host_wait(i % d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):              # d frames in flight: wait for frame i-d, submit frame i
            if i >= d:
                scene.host_wait(i % d)
            scene.render_host_async(cams[(args.warmup + i) % len(cams)], i % d)
        scene.host_wait_all()
        e2e_dt = time.perf_counter() - t0
    else:
        from sage3d_official_b200.multigpu import SharedHostFrames
        shf = SharedHostFrames(H, W, d, rank, world)
        for i in range(2 * d):
            shf.submit(scene, cams[i], rows, i)
            shf.finish(scene, i)
        shf.drain(scene)
        sync_all()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            if i >= d:
                shf.finish(scene, i - d)        # own strip landed, published; rank 0: whole frame delivered
            shf.submit(scene, cams[(args.warmup + i) % len(cams)], rows, i)
        last = e2e_steps - 1
        for i in range(max(0, e2e_steps - d), e2e_steps):
            shf.finish(scene, i, release=(i != last))      # the last frame stays in its slot for the check below
        e2e_dt = time.perf_counter() - t0
        t = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
        # the last frame in the shared host buffers against rank 0's own full frame
        if rank == 0:
            ref_rgb, ref_dep = scene.render_host(cams[(args.warmup + last) % len(cams)])
            hr, hd = shf.frame(last)
            e2e_equal = bool(np.array_equal(ref_rgb, hr) and np.array_equal(ref_dep.view(np.uint32), hd.view(np.uint32)))
        sync_all()
        shf.close()
    e2e_fps = e2e_steps / e2e_dt

    # ---- frame hand-off: what crosses PCIe is the JPEG the reference would have made ----
    handoff = None
    if world == 1 and not args.no_handoff:
        try:
            enc = sage.JpegEncoder(W, H, quality=75, max_batch=d, device=local)      # one scratch slot per frame in flight
            sizes = []
            for i in range(2 * d):
                scene.render_jpeg_async(cams[i], enc, i % d)
                scene.jpeg_wait(enc, i % d)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                if i >= d:
                    sizes.append(len(scene.jpeg_wait(enc, i % d)))
                scene.render_jpeg_async(cams[(args.warmup + i) % len(cams)], enc, i % d)
            for k in range(min(d, e2e_steps)):
                sizes.append(len(scene.jpeg_wait(enc, (e2e_steps + k) % d)))
            ho_dt = time.perf_counter() - t0
            pil_ms = None
            try:
                import io
                from PIL import Image
                fr = rgb[0].cpu().numpy()
                t1 = time.perf_counter()
                for _ in range(3):
                    Image.fromarray(fr).save(io.BytesIO(), format="JPEG")
                pil_ms = (time.perf_counter() - t1) / 3 * 1e3
            except Exception:
                pass
            handoff = {"value": e2e_steps / ho_dt, "unit": "frames/s", "steps": e2e_steps,
                       "d2h_bytes_per_step": float(np.mean(sizes)) + 8, "h2d_bytes_per_step": 400,
                       "pillow_encode_ms_host_1thread": pil_ms,
                       "note": "sage_render_jpeg_async/sage_render_jpeg_wait: render + device JPEG (byte-identical to Pillow "
                               "quality 75), only the encoded bytes copied to the host, " + f"{d} frames in flight"}

        except Exception as e:        # an auxiliary leg never takes the headline line down with it
            handoff = {"error": repr(e)}
    collide = None
    if rank == 0 and not args.no_collide:
        try:
            collide = collide_leg(sage, layout, local)
        except Exception as e:
            collide = {"error": repr(e)}
    batch = None
    if args.batch_scenes > 0:
        try:
            batch = batch_leg(args, sage, world, rank, local, args.batch_scenes)
        except Exception as e:
            batch = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel = longest stage when frames do not overlap (under the multi-stream overlap a
        # stage's event bracket also contains time spent waiting behind other frames' kernels)
        dom = max(stage_serial, key=stage_serial.get)
        nb = 16
        px = W * H if world == 1 else W * (min(rows[1] * 16, H) - rows[0] * 16)
        alg = {   # algorithmic bytes per launch of each stage
            "project": args.n * 44 + V * 32,
            "depth_sort": args.n * 4 + V * 8 * 2 * 3,
            "sh_color": V * (12 * nb + 16 + 16),
            "emit": V * 16 + D * 8,
            "tile_sort": D * 8 * 2 * 2,
            "composite": D * 36 + px * 7,
        }
        dom_ms = stage_serial[dom]
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                if dom in tj:
                    traffic, traffic_src = float(tj[dom]["dram_bytes"]), tj[dom].get("source")
            except Exception:
                pass
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # the dominant kernel is bound by the SMs' issue rate, not by HBM: next to the HBM figure
        # the contract asks for, its warp-instructions per launch (committed ncu capture) over the time measured here,
        # against 148 SMs x 4 schedulers x the SM clock sampled during this run
        issue = None
        try:
            if not (world == 1 and (W, H) == (1920, 1080) and args.n == 2_000_000):
                raise KeyError("the committed instruction count is for the headline workload only")
            wi = float(tj[dom]["warp_instr"])
            mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
            peak_issue = 148 * 4 * mhz * 1e6
            issue = {"warp_instructions_per_launch": wi, "achieved_per_s": wi / (dom_ms * 1e-3), "peak_per_s": peak_issue,
                     "frac": wi / (dom_ms * 1e-3) / peak_issue, "sm_mhz": mhz,
                     "source": tj[dom].get("warp_instr_source", traffic_src),
                     "note": "instruction count of ONE profiled frame (pose 54 of the walk), time = mean over 48 poses of this run"}
        except Exception:
            issue = None
        frame_ms = ms / args.steps
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": frame_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_gaussians": args.n, "resolution": [W, H],
                       "timing": "scene (472 MB) + per-frame working set exceed the 126 MB L2 and every frame uses a new pose; no explicit L2 flush",
                       "parallelism": (f"single GPU, consecutive frames rotate over {nbuf} streams / frame contexts, one CUDA graph per context"
                                       if world == 1 else
                                       f"tile-row blocks (load-balanced) over {world} GPUs, peer-store composite epilogue + a flag barrier in peer memory per frame, {nbuf} frame contexts in flight per rank"),
                       "mean_visible": V, "mean_instances": D, "scene_gen_s": round(gen_s, 1)},
            "timed_blocks": {"blocks": len(block_ms), "steps_per_block": args.steps, "median_ms": ms,
                             "min_ms": float(np.min(block_ms)), "max_ms": float(np.max(block_ms)),
                             "frames_per_s_min_max": [args.steps / (float(np.max(block_ms)) / 1e3), args.steps / (float(np.min(block_ms)) / 1e3)],
                             "wall_s": time.perf_counter() - t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 400,
                    "d2h_bytes_per_step": W * H * 7, "steps": e2e_steps,
                    "note": ("sage_render_host_async/sage_host_wait: host camera in, RGB+depth copied to pinned host for every frame, "
                             f"frames in flight = pipeline depth ({d})") if world == 1 else
                            (f"every rank renders its tile rows and copies its strip over its own PCIe link into one shared pinned host "
                             f"frame; rank 0 waits for all {world} strips of every frame; {d} frames in flight"),
                    "frame_equal": e2e_equal},
            "gpu_launches": int(st["kernel_launches"]) * args.steps * len(block_ms),
            "graph": graph,
            "frame_latency_ms": frame_latency_ms,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": alg[dom],
                         "traffic_source": traffic_src, "issue": issue,
                         "frame": {"algorithmic_bytes": B, "stage_ms_serial": stage_serial,
                                   "stage_gbs_serial": {k: (alg[k] / (v * 1e-3) / 1e9 if v > 0 else 0.0) for k, v in stage_serial.items()},
                                   "serial_frame_ms": sum(stage_serial.values()),
                                   "serial_frame_ms_p50_p99_max": [float(np.percentile(serial_frames, 50)), float(np.percentile(serial_frames, 99)), float(np.max(serial_frames))],
                                   "frame_ms": frame_ms,
                                   "achieved_gbs": B / (frame_ms * 1e-3) / 1e9 if frame_ms > 0 else 0.0,
                                   "frac": B / (frame_ms * 1e-3) / 1e9 / peak if frame_ms > 0 else 0.0}},
        }
        if frame_equal is not None:
            line["multigpu_frame_equal"] = frame_equal
        if handoff is not None:
            line["handoff_jpeg"] = handoff
        if collide is not None:
            line["collide"] = collide
        if batch is not None:
            line["batch8"] = batch
        if not args.no_cpu_baseline and world == 1:      # CPU baseline: rank 0 at N=1 only
            import oracle
            oracle.build()
            cores = os.cpu_count() or 1
            cfps, thr, cdt = oracle_fps(cloud, cams, args.cpu_frames, cores)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": thr, "kind": "port",
                                    "sample": f"{args.cpu_frames} full 1920x1080 frames of the same walk, CPU oracle "
                                              f"(OpenMP, {thr} threads, {cdt:.1f} s); the reference's own renderer (Isaac Sim RTX) cannot run"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
