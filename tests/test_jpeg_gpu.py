"""GPU: the device JPEG encoder (csrc/jpeg.cu, through the C-ABI) against the oracle, the committed
Pillow goldens and the live Pillow — BYTE-IDENTICAL files (integer work: the bar is bit-exact).
Stage-level checks (quantised coefficients, bit offsets) localise a failure."""
import io
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden" / "jpeg_pillow.npz"


def _pillow(a, q=None):
    from PIL import Image
    b = io.BytesIO()
    kw = {} if q is None else {"quality": q}
    Image.fromarray(a).save(b, format="JPEG", **kw)
    return b.getvalue()


def _gpu_encode(sage, frames, q, max_bytes=0):
    import torch
    frames = np.ascontiguousarray(frames)
    if frames.ndim == 3:
        frames = frames[None]
    B, H, W, _ = frames.shape
    enc = sage.JpegEncoder(W, H, quality=q, max_batch=B, max_bytes_per_frame=max_bytes)
    out = enc.encode(torch.from_numpy(frames).cuda())
    return out, enc


def _first_diff(a, b):
    return next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))


def test_goldens_byte_exact(sage, oracle):
    z = np.load(GOLDEN)
    names = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    for n in names:
        q = n.rsplit("_q", 1)[1]
        q = 75 if q == "None" else int(q)
        a = z[n + "_in"]
        (got,), enc = _gpu_encode(sage, a, q, max_bytes=a.shape[0] * a.shape[1] * 4 + 8192)
        want = z[n + "_out"].tobytes()
        if got != want:
            _, ocoef = oracle.jpeg_encode(a, q, want_coefs=True)
            gcoef, _, _ = enc.debug(0)
            bad = np.argwhere(gcoef != ocoef)
            pytest.fail(f"{n}: stream differs at byte {_first_diff(got, want)} (len {len(got)} vs {len(want)}); "
                        f"{len(bad)} coefficient mismatches, first {bad[:3].tolist()}")


@pytest.mark.parametrize("hw", [(480, 640), (768, 1024), (1080, 1920), (1077, 1918), (35, 1000)])
def test_full_sizes_vs_oracle_and_pillow(sage, oracle, hw):
    H, W = hw
    rng = np.random.default_rng(H * 31 + W)
    yy, xx = np.mgrid[0:H, 0:W]
    a = np.stack([127 + 120 * np.sin(xx / 37 + yy / 11), xx * 255 / max(W - 1, 1), yy * 255 / max(H - 1, 1)], -1)
    a = a.astype(np.uint8)
    m = rng.uniform(size=(H, W)) < 0.05
    a[m] = rng.integers(0, 256, (int(m.sum()), 3), dtype=np.uint8)
    (got,), enc = _gpu_encode(sage, a, 75)
    want, ocoef = oracle.jpeg_encode(a, 75, want_coefs=True)
    gcoef, off, total = enc.debug(0)
    assert np.array_equal(gcoef, ocoef), f"{np.count_nonzero(gcoef != ocoef)} coefficients differ"
    assert np.all(np.diff(off.astype(np.int64)) > 0) and off[0] == 0 and total > off[-1]
    assert got == want, f"differs at byte {_first_diff(got, want)}; len {len(got)} vs {len(want)}"
    assert got == _pillow(a), "oracle and GPU agree but differ from the live Pillow"


def test_noise_all_ff_and_stuffing(sage, oracle):
    # white noise at quality 100 (long codes, many 0xFF bytes to stuff) and an all-white frame
    rng = np.random.default_rng(3)
    H, W = 200, 312
    a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    for q in (100, 75, 5):
        (got,), _ = _gpu_encode(sage, a, q, max_bytes=H * W * 4 + 8192)
        assert got == oracle.jpeg_encode(a, q), q
        assert got.count(b"\xff\x00") > 0 or q == 5
    w = np.full((H, W, 3), 255, np.uint8)
    (got,), _ = _gpu_encode(sage, w, 75)
    assert got == oracle.jpeg_encode(w, 75) == _pillow(w)


def test_batch_and_device_output(sage, oracle):
    import torch
    rng = np.random.default_rng(11)
    B, H, W = 5, 120, 176
    fr = np.clip(rng.normal(128, 40, (B, H, W, 3)) + np.linspace(-60, 60, W)[None, None, :, None], 0, 255).astype(np.uint8)
    got, enc = _gpu_encode(sage, fr, 75)
    for f in range(B):
        assert got[f] == oracle.jpeg_encode(fr[f], 75), f
    out, sizes = enc.encode_device(torch.from_numpy(fr).cuda())
    torch.cuda.synchronize()
    out, sizes = out.cpu().numpy(), sizes.cpu().numpy()
    for f in range(B):
        assert out[f, :sizes[f]].tobytes() == got[f]


def test_capacity_overflow_is_reported(sage):
    import torch
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    enc = sage.JpegEncoder(64, 64, quality=100, max_batch=1, max_bytes_per_frame=2048)
    with pytest.raises(sage.SageError, match="did not fit"):
        enc.encode(torch.from_numpy(a).cuda())
    # the encoder stays usable
    flat = np.full((64, 64, 3), 90, np.uint8)
    assert enc.encode(torch.from_numpy(flat).cuda())[0] == _pillow(flat, 100)


def test_render_jpeg_handoff_matches_pillow_of_the_frame(sage, oracle):
    """env.get_rgb() -> PIL JPEG on the host (reference) == render -> JPEG on the device."""
    cloud, _ = sage.synthetic_room(0, 50_000)
    scene = sage.GaussianScene(cloud)
    cams = [sage.Camera.look_heading((1.5 + 0.2 * i, 1.0 + 0.1 * i, 1.2), np.radians(30.0 + 25 * i), 640, 480)
            for i in range(5)]
    enc = sage.JpegEncoder(640, 480)
    got = dict(scene.render_batch_jpeg(cams, enc))
    assert sorted(got) == list(range(5))
    for i, cam in enumerate(cams):
        rgb, _ = scene.render_host(cam, want_depth=False)
        assert got[i] == _pillow(rgb), i
        assert got[i] == oracle.jpeg_encode(rgb, 75)
    # optional raw frame + depth alongside the bytes
    rgb_a, dep_a = scene.render_jpeg_async(cams[2], enc, 0, want_rgb=True, want_depth=True)
    data = scene.jpeg_wait(enc, 0)
    rgb_h, dep_h = scene.render_host(cams[2])
    assert data == got[2] and np.array_equal(rgb_a, rgb_h) and np.array_equal(dep_a, dep_h)


def test_decoded_psnr_property_1080p(sage):
    """Size-independent property at the bench resolution: the stream decodes and is close to the
    source (a wrong Huffman/bit-offset anywhere breaks the decode of everything after it)."""
    from PIL import Image
    cloud, layout = sage.synthetic_interior(1, 200_000)
    scene = sage.GaussianScene(cloud)
    pos, heading = sage.camera_path(layout, 4, seed=1)[2]
    cam = sage.Camera.look_heading(pos, heading, 1920, 1080)
    enc = sage.JpegEncoder(1920, 1080)
    rgb, _ = scene.render_jpeg_async(cam, enc, 1, want_rgb=True)
    data = scene.jpeg_wait(enc, 1)
    dec = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.float64)
    mse = np.mean((dec - rgb.astype(np.float64)) ** 2)
    assert dec.shape == (1080, 1920, 3) and 10 * np.log10(255.0 ** 2 / max(mse, 1e-9)) > 25.0
    assert data == _pillow(rgb)


def test_env_facade_handoff_serves_the_devices_bytes(sage, tmp_path):
    """SimpleVLNEnv(jpeg_handoff=True): the caller's own Image.fromarray(rgb).save(..., 'JPEG')
    (run_benchmark.py:1421, vlm_client_modular.py:445-455) returns the device's bytes — which equal
    what Pillow makes of the same pixels — and a modified frame falls back to Pillow."""
    from PIL import Image
    from sage3d_official_b200 import handoff
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    cloud, _ = sage.synthetic_room(0, 30_000)
    try:
        env = SimpleVLNEnv(str(tmp_path / "in_memory.npz"), resolution=(640, 480), gaussians=cloud, jpeg_handoff=True)
        env.set_start_pose([1.5, 1.0, 0.85], [0.0, 0.0, 0.2, 0.98])
        h0 = handoff.hook_stats()["hits"]
        for step in range(3):
            rgb = env.get_rgb()
            assert isinstance(rgb, np.ndarray) and rgb.shape == (480, 640, 3) and rgb.flags.writeable
            plain = np.array(rgb, copy=True).view(np.ndarray)
            b = io.BytesIO(); Image.fromarray(rgb).convert("RGB").save(b, format="JPEG")
            assert b.getvalue() == _pillow(plain), step
            env.apply_cmd_for(0.0, 0.0, 0.4, 1.0)           # turn: a new frame next time
        assert handoff.hook_stats()["hits"] == h0 + 3
        rgb = env.get_rgb()
        rgb[10:20, 10:20] = 0                               # caller draws on its array
        b = io.BytesIO(); Image.fromarray(rgb).save(b, format="JPEG")
        assert b.getvalue() == _pillow(np.asarray(rgb)) and handoff.hook_stats()["hits"] == h0 + 3
        old = env.get_rgb()
        env.apply_cmd_for(0.0, 0.0, 0.4, 1.0)
        env.get_rgb()                                       # the pinned copy now holds another frame
        b = io.BytesIO(); Image.fromarray(old).save(b, format="JPEG")
        assert b.getvalue() == _pillow(np.asarray(old)) and handoff.hook_stats()["hits"] == h0 + 3
        # depth and the frame itself are the same with and without the hand-off
        env2 = SimpleVLNEnv(str(tmp_path / "in_memory.npz"), resolution=(640, 480), gaussians=cloud, jpeg_handoff=False)
        env2.set_start_pose([1.5, 1.0, 0.85], [0.0, 0.0, 0.2, 0.98])
        env.set_start_pose([1.5, 1.0, 0.85], [0.0, 0.0, 0.2, 0.98])
        assert np.array_equal(env.get_rgb(), env2.get_rgb()) and np.array_equal(env.get_depth(), env2.get_depth())
        assert type(env2.get_rgb()) is np.ndarray
    finally:
        handoff.uninstall_pillow_hook()


def test_isaac_shim_handoff_generate_images_save_line(sage, tmp_path):
    """generate_images.py:426-432 verbatim against the stand-ins with the hand-off on: the .jpg on
    disk is the device's stream and equals Pillow's encode of the same frame."""
    from PIL import Image
    from sage3d_official_b200 import handoff
    from sage3d_official_b200.scene_io import save_ply
    import sage3d_official_b200.isaac_shim as shim
    cloud, _ = sage.synthetic_room(4, 20_000)
    save_ply(tmp_path / "0009.ply", cloud)
    (tmp_path / "0009.usda").write_text('#usda 1.0\ndef Xform "World" { def "gauss" ( prepend references = @./0009.usdz@</World/gauss> ) {} }\n')
    try:
        shim.install(jpeg_handoff=True)
        from omni.isaac.core import World
        from omni.isaac.core.utils.stage import open_stage
        from omni.isaac.sensor import Camera
        assert open_stage(usd_path=str(tmp_path / "0009.usda"))
        world = World()
        cam = Camera(prim_path="/World/NaVILACamera", frequency=30, resolution=(1024, 768))   # generate_images.py:43
        cam.initialize()
        h0 = handoff.hook_stats()["hits"]
        for idx, (x, qz) in enumerate([(1.5, 0.1), (2.0, 0.5), (2.5, -0.3)]):
            position = np.array([x, 1.0, 0.3], dtype=np.float32)
            position[2] = 1.2
            cam.set_world_pose(position=position, orientation=np.array([0.0, 0.0, qz, 0.9], dtype=np.float32))
            for _ in range(3):
                world.step(render=True)
            img = cam.get_rgba()
            img_array = img[:, :, :3].copy()
            Image.fromarray(img_array).save(tmp_path / f"s_t_{idx:03d}.jpg")
            got = (tmp_path / f"s_t_{idx:03d}.jpg").read_bytes()
            assert got == _pillow(np.asarray(img_array)), idx
        assert handoff.hook_stats()["hits"] == h0 + 3
    finally:
        shim.install(jpeg_handoff=False)
        handoff.uninstall_pillow_hook()
        import omni.usd
        omni.usd.get_context().close_stage()
