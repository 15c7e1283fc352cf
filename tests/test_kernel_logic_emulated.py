"""CPU: the SOURCE of two CUDA translation units (csrc/jpeg.cu, csrc/plan.cu) compiled with g++ under
the development-only lock-step emulator (tools/cuda_emu/cuda_emu.h: CUDA threads -> OS threads,
barriers, shuffles, ballots) and checked against the pinned oracle / Pillow / the reference goldens.

This is NOT a product path and not a fallback — the emulated build is never part of libsage_splat.so
and nothing in the package can reach it; it exists so that the no-GPU suite exercises the kernels'
logic (bit positions, scans, heap order) and not only their host-side wrappers.  The real parity
tests are the gpu-marked ones."""
import ctypes as C
import io
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
EMU = ROOT / "tools" / "cuda_emu"


def _build(name, tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("emu") / f"lib{name}.so"
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-x", "c++", "-DSAGE_CUDA_EMU", "-shared", "-fPIC", "-pthread", "-w",
                        "-o", str(out), str(EMU / f"{name}.cpp")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("emulated build failed: " + r.stderr[-400:])
    L = C.CDLL(str(out))
    L.sage_last_error.restype = C.c_char_p
    return L


@pytest.fixture(scope="module")
def jpeg_emu(tmp_path_factory):
    L = _build("jpeg_emu", tmp_path_factory)
    vp = C.c_void_p
    L.sage_jpeg_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
    L.sage_jpeg_encode_host.argtypes = [vp, vp, C.c_int, vp, C.c_int64, vp, vp]
    L.sage_jpeg_destroy.argtypes = [vp]
    return L


@pytest.fixture(scope="module")
def png_emu(tmp_path_factory):
    L = _build("png_emu", tmp_path_factory)
    vp, f32 = C.c_void_p, C.c_float
    L.sage_png_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]
    L.sage_png_max_bytes.argtypes = [vp, C.POINTER(C.c_int64)]
    L.sage_depth_png_encode_host.argtypes = [vp, vp, f32, f32, f32, vp, C.c_int64, C.POINTER(C.c_int64), vp, vp]
    L.sage_png_destroy.argtypes = [vp]
    return L


@pytest.fixture(scope="module")
def plan_emu(tmp_path_factory):
    L = _build("plan_emu", tmp_path_factory)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.sage_grid_inflate.argtypes = [i32, vp, i32, i32, dbl, dbl, vp, vp]
    L.sage_grid_components.argtypes = [i32, vp, i32, i32, vp, vp]
    L.sage_planner_create.argtypes = [C.POINTER(vp), i32, vp, i32, i32, i32, i32, i64]
    L.sage_planner_destroy.argtypes = [vp]
    L.sage_plan_astar.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    return L


def test_jpeg_kernels_emulated_match_pillow_and_oracle(jpeg_emu, oracle):
    from PIL import Image
    rng = np.random.default_rng(7)
    for (H, W), q in (((24, 40), 75), ((17, 33), 90)):
        yy, xx = np.mgrid[0:H, 0:W]
        a = np.stack([127 + 120 * np.sin(xx / 7 + yy / 11), xx * 255 / (W - 1), yy * 255 / (H - 1)], -1).astype(np.uint8)
        m = rng.uniform(size=(H, W)) < 0.2
        a[m] = rng.integers(0, 256, (int(m.sum()), 3), dtype=np.uint8)
        fr = np.ascontiguousarray(np.stack([a, a[::-1].copy()]))
        h = C.c_void_p()
        assert jpeg_emu.sage_jpeg_create(C.byref(h), 0, W, H, q, 2, W * H * 4 + 8192) == 0
        cap = W * H * 4 + 8192
        out = np.zeros((2, cap), np.uint8)
        sizes = np.zeros(2, np.int64)
        rc = jpeg_emu.sage_jpeg_encode_host(h, fr.ctypes.data, 2, out.ctypes.data, cap, sizes.ctypes.data, None)
        jpeg_emu.sage_jpeg_destroy(h)
        assert rc == 0, jpeg_emu.sage_last_error()
        for f in range(2):
            got = out[f, :sizes[f]].tobytes()
            b = io.BytesIO()
            Image.fromarray(fr[f]).save(b, format="JPEG", quality=q)
            assert got == b.getvalue() == oracle.jpeg_encode(fr[f], q), (H, W, q, f)


def test_png_kernels_emulated_match_oracle_stream(png_emu, oracle):
    """csrc/png.cu (quantise + Sub filter + in-group run tokens + fixed-Huffman bit packing + Adler-32 + chunk
    CRC folding) under the emulator: the file equals the oracle's restatement byte for byte and Pillow decodes it
    (verifying every checksum) to the reference client's uint16 pixels (vlm_client_modular.py:354-355)."""
    from PIL import Image
    rng = np.random.default_rng(11)
    cases = [rng.uniform(-1, 8, size=(9, 21)).astype(np.float32),
             np.full((6, 40), 6.5, np.float32),
             np.array([[0.5]], np.float32),
             np.clip(rng.uniform(0, 9, size=(20, 17000 // 20)).astype(np.float32), 0.1, 6.5)]     # > one 32 KB IDAT chunk
    z = np.zeros((5, 64), np.float32)
    z[:, 14:18] = 3.0
    cases.append(z)
    for depth in cases:
        H, W = depth.shape
        h = C.c_void_p()
        assert png_emu.sage_png_create(C.byref(h), 0, W, H, 1) == 0, png_emu.sage_last_error()
        cap = C.c_int64(0)
        png_emu.sage_png_max_bytes(h, C.byref(cap))
        out = np.zeros(cap.value, np.uint8)
        twin = np.zeros((H, W), np.uint16)
        n = C.c_int64(0)
        d = np.ascontiguousarray(depth)
        rc = png_emu.sage_depth_png_encode_host(h, d.ctypes.data, 0.0, 6.5, 10000.0, out.ctypes.data, cap.value, C.byref(n),
                                                twin.ctypes.data, None)
        png_emu.sage_png_destroy(h)
        assert rc == 0, png_emu.sage_last_error()
        got = out[:n.value].tobytes()
        want = (np.clip(depth, 0.0, 6.5) * 10000.0).astype(np.uint16)
        assert np.array_equal(twin, want)
        assert got == oracle.depth_png_encode(depth), depth.shape
        im = Image.open(io.BytesIO(got))
        im.load()
        assert np.array_equal(np.asarray(im).astype(np.uint16), want)


def test_plan_kernels_emulated_match_reference_goldens(plan_emu, oracle):
    z = np.load(ROOT / "tests" / "golden" / "plan_rooms_a.npz")
    raw, grid = np.ascontiguousarray(z["raw"]), np.ascontiguousarray(z["grid"])
    H, W = grid.shape
    out = np.zeros_like(raw)
    assert plan_emu.sage_grid_inflate(0, raw.ctypes.data, H, W, float(z["scale"]), float(z["radius"]), out.ctypes.data, None) == 0
    assert np.array_equal(out, grid)
    lab = np.zeros((H, W), np.int32)
    assert plan_emu.sage_grid_components(0, grid.ctypes.data, H, W, lab.ctypes.data, None) == 0
    assert np.array_equal(lab, oracle.grid_components(grid))
    starts, goals = np.ascontiguousarray(z["starts"]), np.ascontiguousarray(z["goals"])
    Q = len(starts)
    h = C.c_void_p()
    assert plan_emu.sage_planner_create(C.byref(h), 0, grid.ctypes.data, H, W, 8, 0, 0) == 0
    mp = 4 * (H + W)
    plen = np.zeros(Q, np.int32)
    pxy = np.zeros((Q, mp, 2), np.int32)
    assert plan_emu.sage_plan_astar(h, starts.ctypes.data, goals.ctypes.data, Q, plen.ctypes.data, pxy.ctypes.data, None, None) == 0
    plan_emu.sage_planner_destroy(h)
    for k in range(Q):
        want = z["path_xy"][z["path_off"][k]:z["path_off"][k + 1]]
        assert plen[k] == z["path_len"][k] and np.array_equal(pxy[k, :plen[k]], want), k


def test_render_pipeline_emulated_matches_oracle(oracle, sage, tmp_path_factory, monkeypatch):
    """The whole render path (project, 3-kernel radix passes, SH colour, load-balanced emission,
    by-bin sort with fused ranges, bin order, warp-per-bin composite, C-ABI layer with its direct-launch
    path) under the emulator, bit for bit against the oracle: RGB, RGBA, depth, tile-row shards, an
    instance-capacity overflow and the empty scene."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    sys_path = str(EMU)
    import sys
    if sys_path not in sys.path:
        sys.path.insert(0, sys_path)
    import run_render_emu as R
    out = tmp_path_factory.mktemp("emu_render") / "librender_emu.so"
    try:
        L = R.load(str(out))
    except subprocess.CalledProcessError as e:
        pytest.skip(f"emulated build failed: {e}")
    cloud, _ = sage.synthetic_room(0, 2500)
    cam = sage.Camera.look_heading((1.5, 1.0, 1.2), np.radians(30.0), 160, 120)
    assert R.check(L, sage, oracle, cloud, cam, "room", rows_split=[0, 3, 4, 8])
    cam = sage.Camera.look_heading((3.0, 2.0, 1.2), np.radians(200.0), 97, 61)        # ragged edges
    assert R.check(L, sage, oracle, cloud, cam, "ragged")
    wall, _ = sage.synthetic_room(8, 400)
    cam = sage.Camera.look_heading((0.2, 2.0, 1.2), np.radians(180.0), 192, 112)      # footprints of hundreds of bins
    assert R.check(L, sage, oracle, wall, cam, "wall close-up")
    assert R.check(L, sage, oracle, wall.subset(np.zeros(0, np.int64)), cam, "empty")
    # the composite's other instantiations (these small whole frames get E = 4 + the per-stage box test by default)
    cam = sage.Camera.look_heading((1.5, 1.0, 1.2), np.radians(30.0), 160, 120)
    monkeypatch.setenv("SAGE_COMPOSITE_WIDE_MAX_CTAS", "0")
    assert R.check(L, sage, oracle, cloud, cam, "room, 2 entries per trip, index walk")
    monkeypatch.setenv("SAGE_COMPOSITE_PREFILTER", "1")
    assert R.check(L, sage, oracle, cloud, cam, "room, 2 entries per trip, box test")
    monkeypatch.delenv("SAGE_COMPOSITE_WIDE_MAX_CTAS")
    monkeypatch.setenv("SAGE_COMPOSITE_PREFILTER", "0")
    assert R.check(L, sage, oracle, cloud, cam, "room, 4 entries per trip, index walk")
