#!/usr/bin/env python
"""BASELINE config 5: a SAGE-Bench episode through the UNMODIFIED reference ``run_benchmark.py``
(its step loop :1331-1698) on top of the drop-in environment, with a scripted socket VLM.

    python tools/run_benchmark_episode.py [--steps 500] [--width 1920 --height 1080] [--n 2000000]
                                          [--out profiles/episode_r02_run_benchmark.json]

What runs: ``tools/launch_sage_bench.py  <reference>/run_benchmark.py --scene_usd_path scene.npz
--traj_json_path episodes.json --map_path map.json --input-type rgb --output-type text --protocol socket
--max-steps N ...`` in a child process; ``tools/fake_vlm_server.py`` (length-prefixed JSON, deterministic
action script, 0 ms think time) in this one.  The scene is the seeded synthetic interior of the bench
(InteriorGS assets are not in the reference tree), written to disk as the loader's NPZ; the 2-D
semantic map and the GVLN episode JSON are generated from the same layout.

The reference script is taken from ``oracle/_ref/environment_evaluation/`` (staged by
``oracle/make_ref.py``; on the GPU box) or from ``/root/reference``.  Prints and writes ONE JSON
object: steps/s and the per-step latency split render / collide / VLM client + Python.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def reference_script() -> Path:
    for d in (ROOT / "oracle" / "_ref" / "environment_evaluation", Path("/root/reference/Code/benchmark/environment_evaluation")):
        if (d / "run_benchmark.py").exists():
            return d / "run_benchmark.py"
    raise SystemExit("run_benchmark.py not found: run `python oracle/make_ref.py` where /root/reference exists")


def write_assets(tmp: Path, n: int, seed: int = 1, n_points: int = 6):
    """scene.npz + map.json + episodes.json (GVLN format, episodes_adapter.py:40-89) for one episode."""
    import sage3d_official_b200 as sage
    from sage3d_official_b200.scene_io import morton_order, save_npz
    from sage3d_official_b200.scenes import write_semantic_map
    from sage3d_official_b200.collision_detector import build_obstacle_grid
    from sage3d_official_b200.scenes import to_trajectory_frame
    cloud, layout = sage.synthetic_interior(seed, n)
    cloud = cloud.subset(morton_order(cloud.xyz))
    map_path = write_semantic_map(layout, tmp / "map.json")
    inst = json.loads(Path(map_path).read_text())
    _, (min_x, max_x, min_y, max_y) = build_obstacle_grid(inst, 0.08, 0.05)
    sx, sy = min_x + max_x, min_y + max_y
    # the stage (and the agent) live in the trajectory frame; the generator's layout is in map coordinates
    cloud.xyz[:] = to_trajectory_frame(cloud.xyz, (min_x, max_x, min_y, max_y))
    save_npz(tmp / "scene_0001.npz", cloud)
    # trajectory frame: x3 = x2 - (min+max) (trajectory_2d_to_3d.py:56-63); rotation [-sin(psi'/2),0,0,cos(psi'/2)]
    path = sage.camera_path(layout, n_points, seed=3)
    pts = []
    for p, h in path:
        psi = (h + np.pi + np.pi) % (2 * np.pi) - np.pi
        pts.append({"position": [float(p[0] - sx), float(p[1] - sy), 0.85],
                    "rotation": [-float(np.sin(psi / 2)), 0.0, 0.0, float(np.cos(psi / 2))]})
    # the goal is put far from the scripted walk so the episode runs its full step budget
    pts[-1]["position"] = [pts[0]["position"][0] + 500.0, pts[0]["position"][1] + 500.0, 0.85]
    episodes = {"scenes": [{"scene_id": 1, "scene_name": "scene_0001",
                            "samples": [{"trajectory_id": "0", "points": pts,
                                         "instructions": ["walk forward through the flat and keep turning at the doors"]}]}]}
    (tmp / "episodes.json").write_text(json.dumps(episodes))
    return tmp / "scene_0001.npz", Path(map_path), tmp / "episodes.json"


def run(steps: int, width: int, height: int, n: int, launcher=None, extra_env=None, jpeg_handoff: bool = False,
        timeout: float = 3600.0):
    from fake_vlm_server import FakeVLMServer
    tmp = Path(tempfile.mkdtemp(prefix="sage_ep_"))
    scene, map_path, episodes = write_assets(tmp, n)
    srv = FakeVLMServer().start()
    perf_path = tmp / "perf.json"
    env = dict(os.environ, SILENT_LOGGING_MODE="1", SAGE_RESOLUTION=f"{width}x{height}", SAGE_PERF_JSON=str(perf_path),
               SAGE_JPEG_HANDOFF="1" if jpeg_handoff else "0", MPLBACKEND="Agg")
    env.update(extra_env or {})
    launcher = launcher or [sys.executable, str(ROOT / "tools" / "launch_sage_bench.py")]
    cmd = launcher + [str(reference_script()), "--scene_usd_path", str(scene), "--traj_json_path", str(episodes),
                      "--map_path", str(map_path), "--output_root", str(tmp / "out"), "--vlm-host", "127.0.0.1",
                      "--vlm-port", str(srv.port), "--input-type", "rgb", "--output-type", "text", "--protocol", "socket",
                      "--max-steps", str(steps), "--headless", "--no-skip-completed", "--goal-radius", "0.5",
                      "--no-progress", "--silent-logging"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=tmp, timeout=timeout)
    wall = time.perf_counter() - t0
    vlm = srv.stop()
    if not perf_path.exists():
        raise SystemExit("run_benchmark.py produced no perf record:\n" + r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    perf = json.loads(perf_path.read_text())
    meas = sorted((tmp / "out").rglob("measurements/*.json"))
    measurements = json.loads(meas[0].read_text()) if meas else None
    n_steps = int(perf["apply_cmd_calls"])
    loop_s = float(perf["last_call_t"] - perf["first_call_t"]) if perf["first_call_t"] is not None else float("nan")
    out = {
        "workload": f"{steps}-step VLN episode through the UNMODIFIED run_benchmark.py (step loop :1331-1698), "
                    f"{width}x{height} RGB, N={n}, socket VLM (scripted, 0 ms think time), grid collide march per step",
        "steps": n_steps, "vlm_requests": vlm["requests"], "vlm_images": vlm["images"], "vlm_bad_images": vlm["bad_images"],
        "vlm_payload_bytes_per_request": vlm["bytes_in"] / max(vlm["requests"], 1),
        "steps_per_s": n_steps / loop_s if loop_s > 0 else None,
        "ms_per_step": 1000.0 * loop_s / max(n_steps, 1),
        "render_ms": 1000.0 * perf["render_s"] / max(n_steps, 1),
        "collide_ms": 1000.0 * perf["collide_s"] / max(n_steps, 1),
        "vlm_server_ms": 1000.0 * vlm["serve_s"] / max(n_steps, 1),
        "client_and_python_ms": 1000.0 * (loop_s - perf["render_s"] - perf["collide_s"]) / max(n_steps, 1),
        "engine_frames": perf["render_calls"], "collide_launches": perf["collide_launches"],
        "prefetch_hits": perf["prefetch_hits"], "prefetch_misses": perf["prefetch_misses"],
        "resolution": perf["resolution"], "jpeg_handoff": bool(jpeg_handoff),
        "process_wall_s": wall, "returncode": r.returncode,
        "measurements_keys": sorted(measurements.keys())[:12] if isinstance(measurements, dict) else None,
        "reference_script": str(reference_script()),
    }
    if r.returncode != 0 or n_steps == 0:
        out["stdout_tail"] = r.stdout[-1500:]
        out["stderr_tail"] = r.stderr[-1500:]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--jpeg-handoff", action="store_true", help="device-side JPEG + Pillow hook (handoff.py)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    out = run(args.steps, args.width, args.height, args.n, jpeg_handoff=args.jpeg_handoff)
    line = json.dumps(out)
    if args.out:
        Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
    print(line)


if __name__ == "__main__":
    main()
