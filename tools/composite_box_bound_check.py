#!/usr/bin/env python
"""The composite's per-stage box test (composite.cu, PRE: an entry of the stage is dropped when even the maximum of its
exponent over the bounding box of the quadrant's live pixels, plus a rounding allowance, stays below log2(1/255)) must be
CONSERVATIVE: never drop an entry a live pixel needs.  Brute force on random conics - isotropic to needle-thin, any angle,
centres up to thousands of pixels away, boxes of 1 x 1 to 8 x 8 pixels: the binary32 bound as the kernel computes it against
the exponent of every pixel centre in the box (binary64 + the loop's own rounding allowance).  Prints the number of wrong
skips (must be 0), the smallest margin near the threshold, and how many keeps were needless.

    python tools/composite_box_bound_check.py [rounds of 200 000 trials, default 40]
"""
import sys
import numpy as np
rng = np.random.default_rng(0)
f32 = np.float32
def trial(n):
    # random conic from covariance: sigma major/minor in pixels, angle
    smaj = np.exp(rng.uniform(np.log(0.6), np.log(3000), n)); smin = np.exp(rng.uniform(np.log(0.55), np.log(smaj)))
    th = rng.uniform(0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    # covariance
    a = smaj**2*c*c + smin**2*s*s; b = (smaj**2 - smin**2)*c*s; d = smaj**2*s*s + smin**2*c*c
    det = a*d - b*b
    A = d/det; B = -b/det; C = a/det      # conic
    LOG2E = 1.4426950408889634
    qa = (-0.5*LOG2E*A).astype(f32); qb = (-LOG2E*B).astype(f32); qc = (-0.5*LOG2E*C).astype(f32)
    lo = rng.uniform(-7.9, 0.0, n).astype(f32)
    gx = rng.uniform(-2000, 4000, n).astype(f32); gy = rng.uniform(-2000, 3000, n).astype(f32)
    # live box inside an 8x8 quadrant
    qx = rng.integers(0, 240, n)*8; qy = rng.integers(0, 135, n)*8
    x0 = qx + rng.integers(0, 8, n); x1 = np.minimum(x0 + rng.integers(0, 8, n), qx+7)
    y0 = qy + rng.integers(0, 8, n); y1 = np.minimum(y0 + rng.integers(0, 8, n), qy+7)
    X0 = (x0+0.5).astype(f32); X1 = (x1+0.5).astype(f32); Y0 = (y0+0.5).astype(f32); Y1 = (y1+0.5).astype(f32)
    ax0 = X0-gx; ax1 = X1-gx; ay0 = Y0-gy; ay1 = Y1-gy
    def form(dx, dy): return (dx*(qa*dx + qb*dy) + (qc*dy)*dy).astype(f32)
    with np.errstate(all="ignore"):
        kc = ((f32(-0.5)*qb)/qc).astype(f32); ka = ((f32(-0.5)*qb)/qa).astype(f32)
        clamp = lambda v, lo_, hi_: np.minimum(np.maximum(v, lo_), hi_).astype(f32)
        fm = form(ax0, clamp(kc*ax0, ay0, ay1))
        fm = np.maximum(fm, form(ax1, clamp(kc*ax1, ay0, ay1)))
        fm = np.maximum(fm, form(clamp(ka*ay0, ax0, ax1), ay0))
        fm = np.maximum(fm, form(clamp(ka*ay1, ax0, ax1), ay1))
    inside = (ax0 <= 0) & (ax1 >= 0) & (ay0 <= 0) & (ay1 >= 0)
    fm = np.where(inside, f32(0), fm)
    mx = np.maximum(np.abs(ax0), np.abs(ax1)); my = np.maximum(np.abs(ay0), np.abs(ay1))
    mag = (np.abs(qa)*mx*mx + np.abs(qb)*mx*my + np.abs(qc)*my*my).astype(f32)
    slack = (f32(0.02) + f32(2e-6)*mag).astype(f32)
    bound = (lo + fm + slack).astype(f32)
    # brute force over pixel centres in the box, kernel formula in float32 (fma emulated in float64 then rounded per op is hard: use float64 exact + report magnitude)
    worst = np.full(n, -np.inf)
    for i in range(8):
        for j in range(8):
            px = (x0+i+0.5); py = (y0+j+0.5)
            ok = (x0+i <= x1) & (y0+j <= y1)
            dx = (px.astype(f32) - gx).astype(np.float64); dy = (py.astype(f32) - gy).astype(np.float64)
            L = dx*(qa.astype(np.float64)*dx + qb.astype(np.float64)*dy) + (qc.astype(np.float64)*dy*dy + lo.astype(np.float64))
            worst = np.where(ok, np.maximum(worst, L), worst)
    # kernel's own float32 error allowance: 4 ulp of mag
    kerr = 4*6e-8*mag.astype(np.float64)
    margin = bound.astype(np.float64) - (worst + kerr)
    relevant = worst + kerr >= -7.99435344 - 1.0       # near or above the threshold: where a wrong skip could matter
    return margin, relevant, mag, bound, worst
tot = 0; minm = np.inf; bad = 0; tight = []
ROUNDS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for _ in range(ROUNDS):
    m, rel, mag, bound, worst = trial(200000)
    tot += len(m)
    # a wrong skip: bound < L255 while some pixel has worst+kerr >= L255
    wrong = (bound.astype(np.float64) < -7.99435344) & (worst + 4*6e-8*mag >= -7.99435344)
    bad += int(wrong.sum())
    minm = min(minm, float(m[rel].min()) if rel.any() else np.inf)
    # how loose: fraction of cases kept though worst < L255 - 0.5 (false keeps)
    kept = bound.astype(np.float64) >= -7.99435344
    tight.append((int(kept.sum()), int((kept & (worst < -8.5)).sum())))
print("trials", tot, "wrong skips", bad, "min margin near threshold", minm)
k = np.array(tight).sum(0); print("kept", k[0], "of which needless (true max < -8.5)", k[1], k[1]/max(k[0],1))
sys.exit(1 if bad else 0)
