"""CPU: the oracle's restatement of the device depth-PNG stream is a valid PNG (Pillow and OpenCV decode it, zlib
accepts the Adler-32, Pillow checks every chunk CRC) of exactly the uint16 pixels the reference's client computes
(vlm_client_modular.py:354-364)."""
import io
import zlib

import numpy as np
import pytest
from PIL import Image


def _depths():
    rng = np.random.default_rng(0)
    yield "noise", rng.uniform(-1, 8, size=(37, 53)).astype(np.float32)
    d = np.full((64, 96), np.inf, np.float32)               # nothing hit: clipped to 6.5 -> long runs
    d[20:40, 10:50] = np.linspace(0.1, 6.5, 40, dtype=np.float32)[None, :]
    yield "flat_and_ramp", d
    yield "one_pixel", np.array([[1.23456]], np.float32)
    yield "narrow", rng.uniform(0, 6.5, size=(5, 1)).astype(np.float32)
    yield "clipped_facade", np.clip(rng.uniform(0, 9, size=(48, 64)).astype(np.float32), 0.1, 6.5)
    yield "wide", rng.uniform(0, 6.5, size=(3, 700)).astype(np.float32)
    z = np.zeros((40, 64), np.float32)
    z[:, 31:35] = 3.0                                          # runs crossing the 32-byte group boundary
    yield "group_boundary", z


@pytest.mark.parametrize("name,depth", list(_depths()), ids=[n for n, _ in _depths()])
def test_depth_png_stream_decodes_to_the_reference_pixels(oracle, name, depth):
    want = (np.clip(depth, 0.0, 6.5) * 10000.0).astype(np.uint16)          # the client's own expression
    assert np.array_equal(oracle.depth_quantise(depth), want)
    png = oracle.depth_png_encode(depth)
    im = Image.open(io.BytesIO(png))
    im.load()                                                                 # verifies chunk CRCs and the zlib stream
    assert im.mode in ("I;16", "I;16B", "I") and im.size == (depth.shape[1], depth.shape[0])
    assert np.array_equal(np.asarray(im).astype(np.uint16), want)
    try:
        import cv2
        dec = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
        assert dec.dtype == np.uint16 and np.array_equal(dec.reshape(want.shape), want)
    except ImportError:
        pass


def test_depth_png_flat_regions_compress(oracle):
    d = np.full((240, 320), 6.5, np.float32)
    png = oracle.depth_png_encode(d)
    assert len(png) < d.size * 2 // 8                                      # runs: far below the raw 2 bytes per pixel
    idat = b"".join(_chunks(png, b"IDAT"))
    raw = zlib.decompress(idat)
    assert len(raw) == 240 * (1 + 640) and raw[0] == 1


def _chunks(png, typ):
    o = 8
    while o < len(png):
        n = int.from_bytes(png[o:o + 4], "big")
        if png[o + 4:o + 8] == typ:
            yield png[o + 8:o + 8 + n]
        o += 12 + n
