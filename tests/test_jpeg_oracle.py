"""CPU: the JPEG oracle (oracle/jpeg_oracle.c) is PINNED against the library the reference calls
for its frame hand-off — Pillow's Image.save(format='JPEG') (generate_images.py:431-432,
vlm_client_modular.py:351,402,453): byte-identical output on the committed golden vectors and on
fresh arrays encoded by the Pillow installed next to the tests."""
import io
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden" / "jpeg_pillow.npz"


def _pillow(a, q=None):
    from PIL import Image
    b = io.BytesIO()
    kw = {} if q is None else {"quality": q}
    Image.fromarray(a).save(b, format="JPEG", **kw)
    return b.getvalue()


def test_oracle_matches_committed_pillow_goldens(oracle):
    z = np.load(GOLDEN)
    names = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    assert len(names) >= 100
    for n in names:
        q = n.rsplit("_q", 1)[1]
        q = 75 if q == "None" else int(q)           # Pillow's default quality is 75
        got = oracle.jpeg_encode(z[n + "_in"], q)
        assert got == z[n + "_out"].tobytes(), n


@pytest.mark.parametrize("hw", [(480, 640), (1080, 1920), (768, 1024), (1077, 1918), (33, 15), (2, 3)])
def test_oracle_matches_live_pillow(oracle, hw):
    H, W = hw
    rng = np.random.default_rng(H * 10007 + W)
    yy, xx = np.mgrid[0:H, 0:W]
    a = np.stack([127 + 120 * np.sin(xx / 37 + yy / 11), xx * 255 / max(W - 1, 1), yy * 255 / max(H - 1, 1)], -1)
    a = a.astype(np.uint8)
    m = rng.uniform(size=(H, W)) < 0.05
    a[m] = rng.integers(0, 256, (int(m.sum()), 3), dtype=np.uint8)
    for q in (None, 95):
        assert oracle.jpeg_encode(a, 75 if q is None else q) == _pillow(a, q), (hw, q)


def test_oracle_every_quality(oracle):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    a[:, :28] = (np.linspace(0, 255, 28)[None, :, None]).astype(np.uint8)
    for q in range(1, 101):
        assert oracle.jpeg_encode(a, q) == _pillow(a, q), q


def test_oracle_stream_decodes_and_coefs_shape(oracle):
    from PIL import Image
    rng = np.random.default_rng(9)
    a = np.clip(rng.normal(128, 30, (50, 70, 3)), 0, 255).astype(np.uint8)
    data, coefs = oracle.jpeg_encode(a, 90, want_coefs=True)
    assert coefs.shape == (5 * 4 * 6, 64)
    dec = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert dec.shape == a.shape
    assert data[:2] == b"\xff\xd8" and data[-2:] == b"\xff\xd9"
