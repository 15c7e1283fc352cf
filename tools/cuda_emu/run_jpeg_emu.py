"""Development check (no GPU): csrc/jpeg.cu's kernels under the lock-step emulator vs Pillow.
Usage: python tools/cuda_emu/run_jpeg_emu.py   (builds /tmp/libjpeg_emu.so first)"""
import ctypes as C, io, subprocess, sys
from pathlib import Path
import numpy as np
from PIL import Image

HERE = Path(__file__).resolve().parent
subprocess.run(["g++", "-std=c++20", "-O1", "-x", "c++", "-DSAGE_CUDA_EMU", "-shared", "-fPIC", "-pthread", "-w",
                "-o", "/tmp/libjpeg_emu.so", str(HERE / "jpeg_emu.cpp")], check=True)
L = C.CDLL("/tmp/libjpeg_emu.so")
L.sage_last_error.restype = C.c_char_p
vp = C.c_void_p
L.sage_jpeg_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
L.sage_jpeg_encode_host.argtypes = [vp, vp, C.c_int, vp, C.c_int64, vp, vp]
L.sage_jpeg_destroy.argtypes = [vp]


def pil(a, q):
    b = io.BytesIO(); Image.fromarray(a).save(b, format="JPEG", quality=q); return b.getvalue()


def emu(frames, q, max_bytes=0):
    B, H, W, _ = frames.shape
    h = vp()
    assert L.sage_jpeg_create(C.byref(h), 0, W, H, q, B, max_bytes) == 0, L.sage_last_error()
    cap = W * H * 3 + 8192
    out = np.zeros((B, cap), np.uint8); sizes = np.zeros(B, np.int64)
    rc = L.sage_jpeg_encode_host(h, frames.ctypes.data, B, out.ctypes.data, cap, sizes.ctypes.data, None)
    L.sage_jpeg_destroy(h)
    assert rc == 0, (L.sage_last_error(), sizes)
    return [out[f, :sizes[f]].tobytes() for f in range(B)]


rng = np.random.default_rng(0)
bad = 0
cases = [(16, 16), (24, 40), (17, 33), (1, 1), (9, 7), (40, 56), (31, 50), (8, 100)]
if len(sys.argv) > 1: cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (H, W) in cases:
    for kind in ("noise", "smooth", "flat", "sparse"):
        B = 2
        fr = []
        for f in range(B):
            if kind == "noise": a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            elif kind == "flat": a = np.full((H, W, 3), rng.integers(0, 256), np.uint8)
            elif kind == "sparse":
                a = np.full((H, W, 3), 255, np.uint8); a[rng.integers(0, H), rng.integers(0, W)] = 0
            else:
                yy, xx = np.mgrid[0:H, 0:W]
                a = np.stack([127 + 120 * np.sin(xx / 7 + yy / 11 + f), xx * 255 / max(W - 1, 1), yy * 255 / max(H - 1, 1)], -1).astype(np.uint8)
            fr.append(a)
        fr = np.ascontiguousarray(np.stack(fr))
        for q in (75, 100, 20):
            got = emu(fr, q, W * H * 4 + 8192)
            for f in range(B):
                want = pil(fr[f], q)
                if got[f] != want:
                    bad += 1
                    k = next((i for i in range(min(len(got[f]), len(want))) if got[f][i] != want[i]), None)
                    print("MISMATCH", (H, W), kind, q, f, len(got[f]), len(want), "first diff at", k)
        print((H, W), kind, "ok" if not bad else f"bad={bad}", flush=True)
print("bad", bad)
sys.exit(1 if bad else 0)
