import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import sage3d_official_b200 as sage, oracle
def ocam(cam): return oracle.make_camera(cam.R, cam.t, cam.campos, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
for deg in (0,1,2,3):
    cloud,_ = sage.synthetic_room(3, 8000, sh_degree=deg)
    cam = sage.Camera.look_heading((2.0,2.0,1.2), np.radians(10.0), 320, 240)
    o_rgb, o_dep, o_st = oracle.render(ocam(cam), *cloud.as_tuple())
    scene = sage.GaussianScene(cloud)
    for rep in range(4):
        rgb, dep = scene.render(cam); rgb=rgb.cpu().numpy(); dep=dep.cpu().numpy()
        st = scene.stats()
        bad = (rgb!=o_rgb).any(-1) | (dep.view(np.uint32)!=o_dep.view(np.uint32))
        ys,xs = np.nonzero(bad)
        print('deg',deg,'rep',rep,'V',st['n_visible'],'oV',o_st['visible'],'D',st['n_instances'],'bad px',bad.sum(), 'tiles', sorted(set(zip((ys//16).tolist(),(xs//16).tolist())))[:6], 'maxdd', float(np.abs(dep-o_dep).max()))
