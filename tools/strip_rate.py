"""Diagnostic (1 GPU): how fast does ONE rank of an N-way tile-row split turn frames over, and is the limit the
host loop or the device?  Renders the rank-0 block of an N-way split (no peers, no barrier) on `ctx` frame contexts and
reports the host's submit time per frame (loop without waiting) next to the device's frames/s."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--split", type=int, default=8)
    ap.add_argument("--ctx", type=int, default=6)
    ap.add_argument("--frames", type=int, default=2000)
    args = ap.parse_args()
    sage, cloud, layout, cams, _ = bench.build_workload(args)
    sage.set_pipeline_depth(args.ctx)
    scene = sage.GaussianScene(cloud, device=0)
    dev = torch.device("cuda", 0)
    H, W = args.height, args.width
    ty = (H + 15) // 16
    out = {}
    for which in ("first", "middle"):
        per = ty // args.split
        r0 = 0 if which == "first" else (args.split // 2) * per
        rows = (r0, r0 + per)
        rgb = [torch.zeros((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(args.ctx)]
        dep = [torch.zeros((H, W), dtype=torch.float32, device=dev) for _ in range(args.ctx)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(args.ctx)]
        for i in range(64):
            b = i % args.ctx
            scene.render(cams[i % len(cams)], rows=rows, out=(rgb[b], dep[b]), stream=streams[b])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(args.frames):
            b = i % args.ctx
            scene.render(cams[i % len(cams)], rows=rows, out=(rgb[b], dep[b]), stream=streams[b])
        t_submit = time.perf_counter() - t0
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        out[which] = {"rows": rows, "host_submit_us_per_frame": 1e6 * t_submit / args.frames,
                      "device_fps": args.frames / (e0.elapsed_time(e1) / 1e3), "wall_fps": args.frames / t_all,
                      "stats": {k: scene.stats()[k] for k in ("n_visible", "n_instances", "kernel_launches")}}
    print(json.dumps({"split": args.split, "ctx": args.ctx, **out}))


if __name__ == "__main__":
    main()
