"""Development harness (no GPU): the whole render path — project, depth sort, SH colour, emit, by-bin
sort, ranges/order, composite — compiled with g++ under tools/cuda_emu/cuda_emu.h and compared bit
for bit with the CPU oracle on small scenes.  Not a product path: see cuda_emu.h."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def load(out="/tmp/emu/librender_emu.so", rebuild=True):
    if rebuild or not Path(out).exists():
        subprocess.run([str(ROOT / "tools/cuda_emu/build_render_emu.sh"), out], check=True)
    L = C.CDLL(out)
    L.sage_last_error.restype = C.c_char_p
    return L


class EmuScene:
    def __init__(self, L, cloud, max_instances=0):
        from sage3d_official_b200._lib import CCamera
        self.L, self.CCamera = L, CCamera
        self.parts = [np.ascontiguousarray(p, np.float32) for p in cloud.as_tuple()]
        xyz, ls, q, op, sh = self.parts
        self.n = xyz.shape[0]
        deg = int(round(np.sqrt(sh.shape[1]))) - 1
        self.h = C.c_void_p()
        vp = C.c_void_p
        L.sage_scene_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_int64]
        L.sage_render.argtypes = [vp, C.POINTER(CCamera), C.c_int, C.c_int, vp, vp, vp]
        L.sage_scene_set_rgba.argtypes = [vp, C.c_int]
        L.sage_scene_destroy.argtypes = [vp]
        rc = L.sage_scene_create(C.byref(self.h), 0, self.n, deg, *[p.ctypes.data for p in self.parts], max_instances)
        assert rc == 0, L.sage_last_error()

    def render(self, cam, rows=None, rgba=False, out=None):
        from sage3d_official_b200.api import _ccam
        cc = _ccam(cam)
        H, W = cam.height, cam.width
        rgb = np.zeros((H, W, 4 if rgba else 3), np.uint8) if out is None else out[0]
        dep = np.zeros((H, W), np.float32) if out is None else out[1]
        r0, r1 = (0, (H + 15) // 16) if rows is None else rows
        self.L.sage_scene_set_rgba(self.h, 1 if rgba else 0)
        rc = self.L.sage_render(self.h, C.byref(cc), r0, r1, rgb.ctypes.data, dep.ctypes.data, None)
        assert rc == 0, self.L.sage_last_error()
        return rgb, dep

    def close(self):
        self.L.sage_scene_destroy(self.h)


def check(L, sage, oracle, cloud, cam, label, rows_split=None):
    sc = EmuScene(L, cloud)
    rgb, dep = sc.render(cam)
    oc = oracle.make_camera(cam.R, cam.t, cam.campos, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
    o_rgb, o_dep, st = oracle.render(oc, *cloud.as_tuple())
    ok_rgb = np.array_equal(rgb, o_rgb)
    ok_dep = np.array_equal(dep.view(np.uint32), o_dep.view(np.uint32))
    print(f"{label}: rgb {'OK' if ok_rgb else 'DIFF %d' % np.count_nonzero(rgb != o_rgb)}  depth {'OK' if ok_dep else 'DIFF %d' % np.count_nonzero(dep.view(np.uint32) != o_dep.view(np.uint32))}  oracle V={st['visible']}", flush=True)
    ok = ok_rgb and ok_dep
    rgba, dep2 = sc.render(cam, rgba=True)
    ok_rgba = np.array_equal(rgba[..., :3], o_rgb) and bool((rgba[..., 3] == 255).all()) and np.array_equal(dep2.view(np.uint32), o_dep.view(np.uint32))
    print(f"   rgba {'OK' if ok_rgba else 'DIFF'}")
    ok = ok and ok_rgba
    if rows_split:
        a = np.zeros_like(o_rgb); d = np.zeros_like(o_dep)
        for r0, r1 in zip(rows_split[:-1], rows_split[1:]):
            sc.render(cam, rows=(r0, r1), out=(a, d))
        ok_rows = np.array_equal(a, o_rgb) and np.array_equal(d.view(np.uint32), o_dep.view(np.uint32))
        print(f"   row shards {'OK' if ok_rows else 'DIFF'}")
        ok = ok and ok_rows
    sc.close()
    return ok


def main():
    import sage3d_official_b200 as sage
    import oracle
    L = load()
    ok = True
    cloud, _ = sage.synthetic_room(0, 3000)
    cam = sage.Camera.look_heading((1.5, 1.0, 1.2), np.radians(30.0), 160, 120)
    ok &= check(L, sage, oracle, cloud, cam, "room 3000 @160x120", rows_split=[0, 3, 4, 8])
    cam = sage.Camera.look_heading((3.0, 2.0, 1.2), np.radians(200.0), 97, 61)
    ok &= check(L, sage, oracle, cloud, cam, "room 3000 @97x61 (ragged)")
    cloud, _ = sage.synthetic_room(8, 600)
    cam = sage.Camera.look_heading((0.2, 2.0, 1.2), np.radians(180.0), 192, 112)
    ok &= check(L, sage, oracle, cloud, cam, "wall close-up (big splats)")
    empty = cloud.subset(np.zeros(0, np.int64))
    ok &= check(L, sage, oracle, empty, cam, "empty scene")
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
