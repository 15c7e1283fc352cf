"""Development diagnostic (needs the -DSAGE_COMPOSITE_PROFILE variant: tools/build_variant.sh prof -DSAGE_COMPOSITE_PROFILE,
run with SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_prof.so): per-warp walk statistics of k_composite -
cycles, list entries scanned / gathered / hit, start and end time - for a few poses of the bench walk."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--poses", default="0,40,80,120,160,200,240")
    args = ap.parse_args()
    sage, cloud, layout, cams, _ = bench.build_workload(args)
    from sage3d_official_b200 import _lib
    L = _lib.lib()
    scene = sage.GaussianScene(cloud, device=0)
    W, H = args.width, args.height
    n_warps = ((W + 31) // 32) * ((H + 15) // 16) * 8
    buf = np.zeros((n_warps, 8), np.uint32)
    rows = []
    for p in [int(x) for x in args.poses.split(",")]:
        for _ in range(3):
            scene.render(cams[p])
        torch.cuda.synchronize()
        assert L.sage_debug_composite_profile(buf.ctypes.data_as(C.c_void_p), n_warps) == 0
        cyc, visited, scanned, hits, t0, t1, n, tile = [buf[:, k].astype(np.int64) for k in range(8)]
        live = n > 0
        start = t0[live].min()
        end_rel = ((t1 - start) & 0xFFFFFFFF)[live]
        beg_rel = ((t0 - start) & 0xFFFFFFFF)[live]
        order = np.argsort(-cyc)
        top = order[:5]
        rows.append({
            "pose": p, "warps": int(live.sum()), "kernel_span_us": float(end_rel.max() / 1e3),
            "last_start_us": float(beg_rel.max() / 1e3),
            "end_us_p50_p90_p99": [float(np.percentile(end_rel, q) / 1e3) for q in (50, 90, 99)],
            "visited_sum": int(visited.sum()), "hits_sum": int(hits.sum()), "scanned_sum": int(scanned.sum()),
            "cycles_sum": int(cyc.sum()),
            "top5": [{"cycles": int(cyc[i]), "visited": int(visited[i]), "hits": int(hits[i]), "scanned": int(scanned[i]), "list": int(n[i]),
                      "cyc_per_visited": float(cyc[i] / max(visited[i], 1)), "start_us": float(((t0[i] - start) & 0xFFFFFFFF) / 1e3),
                      "end_us": float(((t1[i] - start) & 0xFFFFFFFF) / 1e3)} for i in top],
        })
    print(json.dumps({"resolution": [W, H], "poses": rows}))


if __name__ == "__main__":
    main()
