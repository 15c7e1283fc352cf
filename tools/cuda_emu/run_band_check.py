#!/usr/bin/env python
"""Brute-force check (CPU, no GPU) that the culling of the render path is conservative against the blend's ROUNDED exponent:
tools/cuda_emu/band_check.cpp compiles K1a's pixel box and K2's band / quadrant-mask helpers (csrc/sage_internal.cuh) with g++
and tests random splats - isotropic blobs to needles thousands of pixels long, any angle, centres on and off the screen - against
every pixel of a 1920 x 1080 image.  Prints one JSON line; exit status 1 on a violation.

    python tools/cuda_emu/run_band_check.py [n_splats, default 400] [seed] [--no-deflate]
"""
import ctypes as C
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def build(out, extra=()):
    subprocess.run(["g++", *extra, "-std=c++20", "-O2", "-DSAGE_CUDA_EMU", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off",
                    "-fno-fast-math", "-mfma", "-include", str(HERE / "cuda_emu.h"), "-x", "c++", str(HERE / "band_check.cpp"),
                    "-o", str(out)], check=True)
    return C.CDLL(str(out))


def random_splats(n, seed, W, H):
    rng = np.random.default_rng(seed)
    smaj = np.exp(rng.uniform(np.log(0.6), np.log(4000.0), n))
    smin = np.exp(rng.uniform(np.log(0.55), np.log(smaj)))                    # >= sqrt(0.3): the spec's blur floor
    needle = rng.uniform(size=n) < 0.4
    smin[needle] = rng.uniform(0.55, 1.5, needle.sum())                      # thin whatever the length
    th = rng.uniform(0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    a = smaj ** 2 * c * c + smin ** 2 * s * s
    b = (smaj ** 2 - smin ** 2) * c * s
    d = smaj ** 2 * s * s + smin ** 2 * c * c
    det = a * d - b * b
    LOG2E = 1.4426950408889634
    rec = np.zeros((n, 6), np.float32)
    rec[:, 0] = rng.uniform(-0.5 * W, 1.5 * W, n)
    rec[:, 1] = rng.uniform(-0.5 * H, 1.5 * H, n)
    rec[:, 2] = -0.5 * LOG2E * d / det
    rec[:, 3] = LOG2E * b / det
    rec[:, 4] = -0.5 * LOG2E * a / det
    rec[:, 5] = rng.uniform(-7.9, 0.0, n)
    return rec


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if len(args) > 0 else 400
    seed = int(args[1]) if len(args) > 1 else 0
    W, H = 1920, 1080
    # --no-deflate: build the helpers WITHOUT the rounding allowance (SAGE_CULL_EPS = 0) to see the violations it prevents
    extra = ("-DSAGE_CULL_EPS=0.0f",) if "--no-deflate" in sys.argv else ()
    with tempfile.TemporaryDirectory() as tmp:
        L = build(Path(tmp) / "libband_check.so", extra)
        rec = random_splats(n, seed, W, H)
        out = (C.c_long * 6)()
        L.band_check(rec.ctypes.data_as(C.c_void_p), C.c_long(n), W, H, out)
    res = {"splats": n, "seed": seed, "image": [W, H], "tiles_of_the_boxes": out[0], "tiles_emitted": out[1],
           "quadrant_bits_emitted": out[2], "pixels_passing_the_rounded_test": out[3], "violations": out[4],
           "splats_culled_by_the_box": out[5], "rounding_allowance": "off" if extra else "on"}
    print(json.dumps(res))
    return 1 if out[4] else 0


if __name__ == "__main__":
    sys.exit(main())
