"""Diagnostic: the needle-splat scene of tests/test_render_parity_gpu.py on the GPU against the oracle; where do the frames differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import sage3d_official_b200 as sage
import oracle
from test_render_parity_gpu import needle_scene
oracle.build()
for seed in (0, 1, 2, 3, 4):
    for res in ((640, 480), (1920, 1080)):
        cloud, cam = needle_scene(sage, seed, *res)
        sc = sage.GaussianScene(cloud)
        rgb, dep = sc.render(cam)
        rgb, dep = rgb.cpu().numpy(), dep.cpu().numpy()
        oc = oracle.make_camera(cam.R, cam.t, cam.campos, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
        o_rgb, o_dep, st = oracle.render(oc, *cloud.as_tuple())
        bad = (rgb != o_rgb).any(-1) | (dep.view(np.uint32) != o_dep.view(np.uint32))
        ys, xs = np.nonzero(bad)
        msg = f"seed {seed} {res}: differing px {bad.sum()}"
        if bad.any():
            msg += f" bbox x[{xs.min()},{xs.max()}] y[{ys.min()},{ys.max()}] quadrants {len(set(zip((ys // 8).tolist(), (xs // 8).tolist())))}"
            i = 0
            msg += f" first px ({xs[i]},{ys[i]}) gpu rgb {rgb[ys[i], xs[i]]} oracle {o_rgb[ys[i], xs[i]]} depth {dep[ys[i], xs[i]]} vs {o_dep[ys[i], xs[i]]}; stats {sc.stats()}"
        print(msg, flush=True)
