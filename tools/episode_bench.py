#!/usr/bin/env python
"""Configs 3 and 5 of BASELINE.json: an agent-step loop through the SimpleVLNEnv drop-in —
render (RGB + depth) + batched collide march per step — with a scripted action source in place of
the VLM (0 ms think time).  Prints one JSON line: steps/s and the per-step latency split.

    python tools/episode_bench.py [--steps 500] [--n 2000000] [--width 1920 --height 1080]

The reference's own loop for the same step issues 5 + 21 RTX frames for get_rgb/get_depth and up
to 60 more inside apply_cmd_for, plus <= 60 sequential Python collision lookups
(simple_env.py:1356-1589, 2116-2234); here it is one engine frame and <= 2 collide launches.
"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--capsule", action="store_true", help="also run the ellipsoid-vs-capsule test per step")
    ap.add_argument("--vlm-client", action="store_true",
                    help="also do, per step, the host work of the reference's RGB socket client "
                         "(run_benchmark.py:1421 + vlm_client_modular.py:445-455): Image.fromarray(rgb), 8-frame "
                         "history, convert('RGB') + JPEG + base64 of every history frame")
    ap.add_argument("--jpeg-handoff", action="store_true", help="device-side JPEG + Pillow hook (handoff.py)")
    ap.add_argument("--stage-timing", action="store_true", help="per-stage device times (serialises the frame: slower)")
    args = ap.parse_args()
    import sage3d_official_b200 as sage
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    from sage3d_official_b200.scenes import write_semantic_map, collision_ellipsoids, layout_semantic_map
    from sage3d_official_b200.scene_io import morton_order

    from sage3d_official_b200.scenes import to_trajectory_frame
    from sage3d_official_b200.collision_detector import build_obstacle_grid
    cloud, layout = sage.synthetic_interior(1, args.n)
    cloud = cloud.subset(morton_order(cloud.xyz))
    tmp = Path(tempfile.mkdtemp())
    map_path = write_semantic_map(layout, tmp / "map.json")
    # the stage (and the agent) live in the trajectory frame; the generator's layout is in map coordinates
    _, bounds = build_obstacle_grid(json.loads(Path(map_path).read_text()), 0.08, 0.05)
    cloud.xyz[:] = to_trajectory_frame(cloud.xyz, bounds)
    env = SimpleVLNEnv(str(tmp / "scene_in_memory.npz"), resolution=(args.width, args.height),
                       map_json_path=str(map_path), gaussians=cloud, jpeg_handoff=args.jpeg_handoff)
    history, t_client, payload = [], 0.0, 0
    env.set_log_function(lambda m: None)
    bodies = None
    if args.capsule:
        mu, ls, q = collision_ellipsoids(layout, per_m2=400.0)
        bodies = sage.CollisionBodies(to_trajectory_frame(mu, bounds), ls, q)
    # start pose: free space, trajectory-frame coordinates (x3 = x2 - (min+max), trajectory_2d_to_3d.py:56-63)
    det = env.collision_detector
    p0, h0 = sage.camera_path(layout, 1, seed=3)[0]
    sx, sy = det.min_x + det.max_x, det.min_y + det.max_y
    psi = (h0 + np.pi + np.pi) % (2 * np.pi) - np.pi
    env.set_start_pose([float(p0[0] - sx), float(p0[1] - sy), 0.85], [-float(np.sin(psi / 2)), 0.0, 0.0, float(np.cos(psi / 2))])
    rng = np.random.default_rng(0)
    t_render = t_collide = t_capsule = 0.0
    moved = 0
    frame_ms, inst = [], []
    env.get_rgb()                                                   # warm-up frame
    if args.stage_timing:
        env._scene.set_stage_timing(1)      # serialises the frame's stages on one stream: slower frames
    stage_acc = {k: 0.0 for k in env._scene.STAGES}
    t0 = time.perf_counter()
    timed = 0
    for step in range(args.steps):
        calls0 = env.render_calls
        a = time.perf_counter()
        rgb = env.get_rgb()
        depth = env.get_depth()
        b = time.perf_counter()
        frame_ms.append(1000 * (b - a))
        inst.append(env._scene.stats()["n_instances_wanted"])
        if env.render_calls > calls0 and args.stage_timing:
            timed += 1
            for k_, v_ in env._scene.stage_times(0).items():
                stage_acc[k_] += v_
        if args.vlm_client:
            import base64
            import io
            from PIL import Image
            e0 = time.perf_counter()
            history.append(Image.fromarray(rgb))
            history = history[-8:]
            seq = list(history)
            while len(seq) < 8:
                seq.insert(0, seq[0])
            enc_imgs = []
            for img in seq:
                img = img.convert("RGB")
                buf = io.BytesIO()
                img.save(buf, format="JPEG")
                enc_imgs.append(base64.b64encode(buf.getvalue()).decode())
            data = json.dumps({"images": enc_imgs, "query": "go"}).encode("utf-8")
            payload = len(data)
            e1 = time.perf_counter()
            t_client += e1 - e0
        r = rng.uniform()
        if r < 0.6:
            cmd = (0.25, 0.0, 0.0, 1.0)
        elif r < 0.8:
            cmd = (0.0, 0.0, np.radians(15.0), 1.0)
        else:
            cmd = (0.0, 0.0, -np.radians(15.0), 1.0)
        before = env.get_agent_pos()
        env.apply_cmd_for(*cmd)
        c = time.perf_counter()
        if bodies is not None:
            pos = env.get_agent_pos().astype(np.float32)[None, :]
            bodies.check(pos)
        d = time.perf_counter()
        moved += int(np.linalg.norm(env.get_agent_pos()[:2] - before[:2]) > 1e-6)
        t_render += b - a; t_collide += c - max(b, e1 if args.vlm_client else b); t_capsule += d - c
    total = time.perf_counter() - t0
    out = {
        "workload": f"{args.steps}-step episode, {args.width}x{args.height} RGB+depth, N={args.n}, grid collide march"
                    + (" + capsule vs %d ellipsoids" % bodies.m if bodies is not None else ""),
        "steps_per_s": args.steps / total,
        "ms_per_step": 1000 * total / args.steps,
        "render_ms": 1000 * t_render / args.steps,
        "collide_ms": 1000 * t_collide / args.steps,
        "capsule_ms": 1000 * t_capsule / args.steps if bodies is not None else None,
        "vlm_client_ms": 1000 * t_client / args.steps if args.vlm_client else None,
        "vlm_payload_bytes": payload if args.vlm_client else None,
        "jpeg_handoff": bool(args.jpeg_handoff),
        "python_other_ms": 1000 * (total - t_render - t_collide - t_capsule - t_client) / args.steps,
        "render_ms_p50_p90_max": [float(np.percentile(frame_ms, 50)), float(np.percentile(frame_ms, 90)), float(np.max(frame_ms))],
        "instances_p50_p90_max": [int(np.percentile(inst, 50)), int(np.percentile(inst, 90)), int(np.max(inst))],
        "stage_ms_mean": {k_: v_ / max(timed, 1) for k_, v_ in stage_acc.items()} if args.stage_timing else None,
        "depth_mean_m": float(depth.mean()),
        "engine_frames": env.render_calls, "collide_launches": env.collide_launches, "steps_that_moved": moved,
        "frame_shape": list(rgb.shape), "depth_range": [float(depth.min()), float(depth.max())],
        "facade_perf": {k: (round(v, 6) if isinstance(v, float) else v) for k, v in env.perf.items() if not k.endswith("_t")},
        "pipeline_depth": int(getattr(env._scene, "pipeline_depth", 0)),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
