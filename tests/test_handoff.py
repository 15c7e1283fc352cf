"""CPU: the Pillow hook of handoff.py — pre-encoded bytes are used iff the pixels being saved are
bit-identical to the encoded frame and the caller asked for a default JPEG; everything else falls
through to Pillow.  (The "device" bytes here are made by Pillow itself; byte-equality of the real
device encoder with Pillow is tests/test_jpeg_gpu.py.)"""
import io

import numpy as np
import pytest
from PIL import Image


def _pillow(a, **kw):
    b = io.BytesIO()
    Image.fromarray(np.asarray(a)).save(b, format="JPEG", **kw)
    return b.getvalue()


@pytest.fixture()
def hook(sage):
    from sage3d_official_b200 import handoff
    handoff.install_pillow_hook()
    h0 = handoff.hook_stats()
    yield handoff, h0
    handoff.uninstall_pillow_hook()


def _frame(handoff, H=48, W=64, seed=0, valid=lambda: True):
    rng = np.random.default_rng(seed)
    pristine = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    marker = b"\xff\xd8SAGE-PRECODED" + bytes([seed]) + b"\xff\xd9"     # recognisable stand-in for device bytes
    return handoff.DeviceFrame(pristine.copy(), marker, pristine, valid), pristine, marker


def test_untouched_frame_uses_precoded_bytes(hook):
    handoff, h0 = hook
    fr, pristine, marker = _frame(handoff)
    im = Image.fromarray(fr)
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == marker
    b = io.BytesIO(); im.convert("RGB").save(b, format="jpeg")        # vlm_client_modular.py:448-453
    assert b.getvalue() == marker
    assert handoff.hook_stats()["hits"] == h0["hits"] + 2
    assert np.array_equal(np.asarray(im), pristine)        # a real image, pixels intact


def test_path_save_by_extension(hook, tmp_path):
    handoff, _ = hook
    fr, pristine, marker = _frame(handoff, seed=1)
    rgba = handoff.DeviceFrame(np.concatenate([pristine, np.full((48, 64, 1), 255, np.uint8)], 2), marker, pristine, lambda: True)
    Image.fromarray(rgba[:, :, :3].copy()).save(tmp_path / "a_001.jpg")       # generate_images.py:431-432
    assert (tmp_path / "a_001.jpg").read_bytes() == marker
    Image.fromarray(fr).save(tmp_path / "a.png")                               # other formats: Pillow
    assert np.array_equal(np.asarray(Image.open(tmp_path / "a.png")), pristine)


def test_any_difference_falls_back_to_pillow(hook):
    handoff, _ = hook
    fr, pristine, marker = _frame(handoff, seed=2)
    # options
    b = io.BytesIO(); Image.fromarray(fr).save(b, format="JPEG", quality=90)
    assert b.getvalue() == _pillow(pristine, quality=90)
    # the array was modified after get_rgb()
    fr2 = fr.copy(); fr2[3, 4, 1] ^= 1
    b = io.BytesIO(); Image.fromarray(fr2).save(b, format="JPEG")
    assert b.getvalue() == _pillow(fr2)
    # a different channel slice of an RGBA frame has the right shape but the wrong pixels
    rgba = handoff.DeviceFrame(np.concatenate([pristine, np.full((48, 64, 1), 255, np.uint8)], 2), marker, pristine, lambda: True)
    gba = rgba[:, :, 1:4].copy()
    b = io.BytesIO(); Image.fromarray(gba).save(b, format="JPEG")
    assert b.getvalue() == _pillow(gba)
    # the image was drawn on after fromarray (copy-on-write of the read-only core image)
    im = Image.fromarray(fr)
    im.putpixel((0, 0), (1, 2, 3))
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() != marker and b.getvalue() == _pillow(np.asarray(im))
    from PIL import ImageDraw
    im = Image.fromarray(fr)
    ImageDraw.Draw(im).line((0, 0, 10, 10), fill=(255, 0, 0))
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == _pillow(np.asarray(im))
    # resized / converted images are new objects
    b = io.BytesIO(); Image.fromarray(fr).resize((32, 24)).save(b, format="JPEG")
    assert b.getvalue()[:2] == b"\xff\xd8" and b.getvalue() != marker
    b = io.BytesIO(); Image.fromarray(fr).convert("L").save(b, format="JPEG")
    assert b.getvalue() != marker
    # the engine has rendered another frame since (pinned copy no longer this frame)
    stale, _, _ = _frame(handoff, seed=3, valid=lambda: False)
    b = io.BytesIO(); Image.fromarray(stale).save(b, format="JPEG")
    assert b.getvalue() == _pillow(stale)
    # plain arrays are untouched by the hook
    b = io.BytesIO(); Image.fromarray(pristine).save(b, format="JPEG")
    assert b.getvalue() == _pillow(pristine)


def test_history_of_converted_images_shares_pixels(hook):
    handoff, _ = hook
    fr, pristine, marker = _frame(handoff, seed=4)
    im = Image.fromarray(fr)
    hist = [im.convert("RGB") for _ in range(8)]
    hist[3].putpixel((5, 5), (9, 9, 9))                 # mutating one copy must not leak into the others
    for k, h in enumerate(hist):
        b = io.BytesIO(); h.save(b, format="JPEG")
        if k == 3:
            assert b.getvalue() == _pillow(np.asarray(h)) and np.asarray(h)[5, 5, 0] == 9
        else:
            assert b.getvalue() == marker
            assert np.array_equal(np.asarray(h), pristine)


def test_device_frame_is_an_ordinary_array(hook):
    handoff, _ = hook
    fr, pristine, _ = _frame(handoff, seed=5)
    assert isinstance(fr, np.ndarray) and fr.flags.writeable and fr.dtype == np.uint8
    assert (fr.astype(np.float32).mean() - pristine.mean()) < 1e-3
    import pickle
    back = pickle.loads(pickle.dumps(fr))
    assert type(back) is np.ndarray and np.array_equal(back, pristine)
    fr[0, 0, 0] ^= 255                                   # caller-owned: writable, and then no longer "the frame"
    assert handoff.precoded_jpeg(fr) is None


def test_lazy_image_materialises_on_demand_and_stays_safe(hook):
    handoff, _ = hook
    fr, pristine, marker = _frame(handoff, seed=6)
    im = Image.fromarray(fr)
    fr[:] = 0                                            # the caller reuses its array: the image owns its pixels
    assert handoff._core(im) is None                      # nothing unpacked yet
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == marker and handoff._core(im) is None
    assert im.size == (64, 48) and im.mode == "RGB"
    assert np.array_equal(np.asarray(im), pristine)      # materialised from the private copy
    assert handoff._core(im) is not None
    # once real pixel storage exists the precoded bytes are gone for good: Pillow encodes (for the real
    # device JPEG that is the same byte stream — tests/test_jpeg_gpu.py — here the marker is a fake)
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == _pillow(pristine)
    thumb = Image.fromarray(handoff.DeviceFrame(pristine.copy(), marker, pristine, lambda: True))
    thumb.thumbnail((16, 16))                            # in-place resize replaces the core image
    b = io.BytesIO(); thumb.save(b, format="JPEG")
    assert b.getvalue() != marker and thumb.size[0] <= 16


def test_inplace_mutators_and_stock_image_behaviour(hook):
    """the in-place mutators that ignore `readonly`, pixel access, pickling and equality
    all behave as on the image stock `Image.fromarray` returns — and never leave stale JPEG bytes."""
    import pickle
    handoff, _ = hook
    fr, pristine, marker = _frame(handoff, seed=7)
    other = np.random.default_rng(99).integers(0, 256, pristine.shape, dtype=np.uint8)
    # frombytes() on a never-loaded image: works, and the new pixels are what gets saved
    im = Image.fromarray(fr)
    im.frombytes(other.tobytes())
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == _pillow(other)
    # load() then frombytes(): same
    im = Image.fromarray(fr)
    im.load()
    im.frombytes(other.tobytes())
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == _pillow(other)
    # putdata()
    im = Image.fromarray(fr)
    im.putdata([tuple(int(v) for v in px) for px in other.reshape(-1, 3)])
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == _pillow(other)
    # pixel access is writable, as on a stock fromarray image
    im = Image.fromarray(fr)
    px = im.load()
    px[2, 3] = (9, 8, 7)
    want = pristine.copy(); want[3, 2] = (9, 8, 7)
    assert np.array_equal(np.asarray(im), want)
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == _pillow(want)
    # equality with the stock image, both ways round
    im = Image.fromarray(fr)
    stock = handoff._plain_fromarray(pristine)
    assert im == stock and stock == im and not (im != stock)
    # pickling (multiprocessing hand-off): degrades to a plain Image with the same pixels
    im = Image.fromarray(fr)
    back = pickle.loads(pickle.dumps(im))
    assert type(back) is Image.Image and np.array_equal(np.asarray(back), pristine)
    b = io.BytesIO(); im.save(b, format="JPEG")          # pickling read the pixels: Pillow encodes from now on
    assert b.getvalue() == _pillow(pristine)
    # copy() of an untouched image keeps the precoded bytes; of a touched one it does not
    im = Image.fromarray(fr)
    b = io.BytesIO(); im.copy().save(b, format="JPEG")
    assert b.getvalue() == marker


def test_rgba_root_frame_fast_rgb_copy_and_private_twin(hook):
    """Camera.get_rgba() frames: frame[:, :, :3].copy() (generate_images.py:431) takes the packed-conversion
    path, is an ordinary writable array with the same bytes NumPy's strided copy gives, and the hook uses the
    engine's private twin as the image's pixels (no copy) only while the caller's pixels equal it."""
    handoff, h0 = hook
    rng = np.random.default_rng(7)
    pristine = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    twin = pristine.copy().view(handoff.PrivateTwin)
    marker = b"\xff\xd8SAGE-RGBA\xff\xd9"
    rgba = np.concatenate([pristine, np.full((48, 64, 1), 255, np.uint8)], 2)
    fr = handoff.DeviceFrame(rgba, marker, twin, lambda: True, rgba_root=True)
    rgb = fr[:, :, :3].copy()
    assert rgb.shape == (48, 64, 3) and rgb.flags.c_contiguous and rgb.flags.writeable
    assert np.array_equal(np.asarray(rgb), pristine)
    im = Image.fromarray(rgb)
    b = io.BytesIO(); im.save(b, format="JPEG")
    assert b.getvalue() == marker and handoff.hook_stats()["hits"] == h0["hits"] + 1
    assert im._sage_arr is twin                                    # zero-copy: the twin is the pixel store
    rgb[0, 0, 0] ^= 1                                              # the caller's array is its own: the image is unaffected
    assert np.array_equal(np.asarray(im), pristine)
    # other slices / orders / a caller that drew on the RGBA frame first: plain NumPy semantics, Pillow encodes
    assert np.array_equal(np.asarray(fr[:, :, 1:4].copy()), rgba[:, :, 1:4])
    assert np.array_equal(np.asarray(fr[::2, :, :3].copy()), rgba[::2, :, :3])
    assert fr[:, :, :3].copy(order="F").flags.f_contiguous
    fr[5, 5, 0] ^= 0x40
    drawn = fr[:, :, :3].copy()
    assert np.array_equal(np.asarray(drawn), np.asarray(fr)[:, :, :3])
    b = io.BytesIO(); Image.fromarray(drawn).save(b, format="JPEG")
    assert b.getvalue() == _pillow(np.asarray(drawn)) and b.getvalue() != marker
    # an RGBA frame without a JPEG (hand-off off) still gets the fast copy, and plain Pillow
    plain = handoff.DeviceFrame(np.concatenate([pristine, np.full((48, 64, 1), 255, np.uint8)], 2), None, None, lambda: True, rgba_root=True)
    c = plain[:, :, :3].copy()
    assert np.array_equal(np.asarray(c), pristine)
    b = io.BytesIO(); Image.fromarray(c).save(b, format="JPEG")
    assert b.getvalue() == _pillow(pristine)


def test_cv2_hook_returns_device_png_only_for_identical_uint16_pixels(sage):
    """vlm_client_modular.py:354-364 verbatim on a DeviceDepth: the metadata survives np.clip / * / astype, and
    cv2.imencode('.png', ...) hands back the precoded file only while the pixels equal the encoded image."""
    cv2 = pytest.importorskip("cv2")
    from sage3d_official_b200 import handoff
    rng = np.random.default_rng(3)
    depth = np.clip(rng.uniform(0, 9, size=(48, 64)).astype(np.float32), 0.1, 6.5)
    pixels = (np.clip(depth, 0.0, 6.5) * 10000.0).astype(np.uint16)
    marker = b"\x89PNG-SAGE-PRECODED"
    live = [True]
    handoff.install_cv2_hook()
    try:
        h0 = handoff.cv2_hook_stats()["hits"]
        depth_image = handoff.DeviceDepth(depth.copy(), marker, pixels, lambda: live[0])
        depth_array_clamped = np.clip(depth_image, 0.0, 6.5)
        depth_array_encoded = (depth_array_clamped * 10000.0).astype(np.uint16)
        ok, buf = cv2.imencode('.png', depth_array_encoded)
        assert ok and buf.tobytes() == marker and handoff.cv2_hook_stats()["hits"] == h0 + 1
        # a different scale, a touched pixel, an encoder option, a stale frame, another format: OpenCV encodes
        for arr, params in (((depth_array_clamped * 1000.0).astype(np.uint16), None),
                            (depth_array_encoded + np.uint16(1), None),
                            (depth_array_encoded, [cv2.IMWRITE_PNG_COMPRESSION, 9])):
            ok, buf = cv2.imencode('.png', arr) if params is None else cv2.imencode('.png', arr, params)
            assert ok and buf.tobytes() != marker
            assert np.array_equal(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED), np.asarray(arr))
        live[0] = False
        ok, buf = cv2.imencode('.png', depth_array_encoded)
        assert buf.tobytes() != marker
        live[0] = True
        ok, buf = cv2.imencode('.jpg', np.zeros((8, 8, 3), np.uint8))
        assert ok and buf.tobytes()[:2] == b"\xff\xd8"
        assert handoff.cv2_hook_stats()["hits"] == h0 + 1
    finally:
        handoff.uninstall_cv2_hook()
    assert not getattr(cv2.imencode, "_sage_hook", False)
