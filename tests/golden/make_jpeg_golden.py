"""Golden vectors for the frame hand-off row (JPEG), produced by the library the reference calls.

The reference encodes every frame with Pillow defaults (generate_images.py:431-432,
vlm_client_modular.py:351,402,453).  This script runs that exact call —
``Image.fromarray(a).save(buf, format='JPEG'[, quality=q])`` — on small seeded arrays and stores
input + output bytes, so the pin travels to machines without the authoring container.

    python tests/golden/make_jpeg_golden.py      # rewrites tests/golden/jpeg_pillow.npz
Authoring container: Pillow 12.2.0, libjpeg-turbo (features.version('jpg') = 6.2).
"""
import io
from pathlib import Path

import numpy as np
from PIL import Image, features


def cases():
    rng = np.random.default_rng(20260117)
    out = []
    for (H, W) in [(16, 16), (24, 40), (17, 33), (1, 1), (9, 7), (32, 48), (23, 57)]:
        yy, xx = np.mgrid[0:H, 0:W]
        smooth = np.stack([127 + 120 * np.sin(xx / 7 + yy / 11), xx * 255 / max(W - 1, 1),
                           yy * 255 / max(H - 1, 1)], -1).astype(np.uint8)
        noise = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        mixed = smooth.copy()
        m = rng.uniform(size=(H, W)) < 0.15
        mixed[m] = noise[m]
        white = np.full((H, W, 3), 255, np.uint8)
        white[rng.integers(0, H), rng.integers(0, W)] = 0
        for name, a in (("smooth", smooth), ("noise", noise), ("mixed", mixed), ("white", white)):
            for q in (None, 50, 90, 100, 12):
                out.append((f"{H}x{W}_{name}_q{q}", np.ascontiguousarray(a), q))
    return out


def pillow_bytes(a, q):
    b = io.BytesIO()
    if q is None:
        Image.fromarray(a).save(b, format="JPEG")          # the reference's literal call
    else:
        Image.fromarray(a).save(b, format="JPEG", quality=q)
    return b.getvalue()


if __name__ == "__main__":
    d = {}
    for name, a, q in cases():
        d[name + "_in"] = a
        d[name + "_out"] = np.frombuffer(pillow_bytes(a, q), np.uint8)
    d["_meta"] = np.array([f"Pillow {Image.__version__ if hasattr(Image, '__version__') else ''} "
                           f"jpg {features.version('jpg')} turbo {features.check_feature('libjpeg_turbo')}"])
    p = Path(__file__).resolve().parent / "jpeg_pillow.npz"
    np.savez_compressed(p, **d)
    print(p, p.stat().st_size, "bytes,", len(d) // 2, "cases")
