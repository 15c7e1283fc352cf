"""Development check (no GPU): csrc/plan.cu under the lock-step emulator vs the pinned oracle and
the reference goldens.  Build first (see plan_emu.cpp), or let this script do it."""
import ctypes as C, glob, subprocess, sys, os
from pathlib import Path
import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # dev tool: the checker

so = "/tmp/libplan_emu.so"
if not os.path.exists(so) or "--build" in sys.argv:
    subprocess.run(["g++", "-std=c++20", "-O1", "-x", "c++", "-DSAGE_CUDA_EMU", "-shared", "-fPIC", "-pthread", "-w",
                    "-o", so, str(HERE / "plan_emu.cpp")], check=True)
L = C.CDLL(so)
L.sage_last_error.restype = C.c_char_p
vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
L.sage_grid_inflate.argtypes = [i32, vp, i32, i32, dbl, dbl, vp, vp]
L.sage_grid_components.argtypes = [i32, vp, i32, i32, vp, vp]
L.sage_planner_create.argtypes = [C.POINTER(vp), i32, vp, i32, i32, i32, i32, i64]
L.sage_planner_destroy.argtypes = [vp]
L.sage_plan_astar.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]

bad = 0
for path in sorted(glob.glob(str(ROOT / "tests" / "golden" / "plan_*.npz"))):
    z = np.load(path)
    raw, grid = np.ascontiguousarray(z["raw"]), np.ascontiguousarray(z["grid"])
    H, W = grid.shape
    out = np.zeros_like(raw)
    assert L.sage_grid_inflate(0, raw.ctypes.data, H, W, float(z["scale"]), float(z["radius"]), out.ctypes.data, None) == 0
    ok_inf = np.array_equal(out, grid)
    lab = np.zeros((H, W), np.int32)
    assert L.sage_grid_components(0, grid.ctypes.data, H, W, lab.ctypes.data, None) == 0
    ok_cc = np.array_equal(lab, oracle.grid_components(grid))
    starts, goals = np.ascontiguousarray(z["starts"]), np.ascontiguousarray(z["goals"])
    Q = len(starts)
    h = vp()
    assert L.sage_planner_create(C.byref(h), 0, grid.ctypes.data, H, W, 8, 0, 0) == 0, L.sage_last_error()
    mp = 4 * (H + W)
    plen = np.zeros(Q, np.int32); pxy = np.zeros((Q, mp, 2), np.int32); nexp = np.zeros(Q, np.int64)
    assert L.sage_plan_astar(h, starts.ctypes.data, goals.ctypes.data, Q, plen.ctypes.data, pxy.ctypes.data, nexp.ctypes.data, None) == 0
    L.sage_planner_destroy(h)
    ok_a = True
    for k in range(Q):
        want = z["path_xy"][z["path_off"][k]:z["path_off"][k + 1]]
        if plen[k] != z["path_len"][k] or not np.array_equal(pxy[k, :plen[k]], want):
            ok_a = False
            print("  A* mismatch", Path(path).stem, k, plen[k], z["path_len"][k])
        _, ne = oracle.astar(grid, starts[k], goals[k], want_expansions=True)
        if ne != nexp[k]:
            ok_a = False
            print("  expansions differ", k, ne, nexp[k])
    print(Path(path).stem, "inflate", ok_inf, "components", ok_cc, "astar", ok_a, flush=True)
    bad += (not ok_inf) + (not ok_cc) + (not ok_a)
print("bad", bad)
sys.exit(1 if bad else 0)
