"""GPU: the device depth-PNG encoder (csrc/png.cu) against the oracle's restatement (byte for byte) and against
the reference client's own host path (vlm_client_modular.py:354-364) — identical uint16 pixels out of any reader."""
import io

import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(1)
    yield "noise_odd", rng.uniform(-1, 8, size=(37, 53)).astype(np.float32)
    d = np.full((96, 128), np.inf, np.float32)
    d[20:60, 10:90] = np.linspace(0.1, 6.5, 80, dtype=np.float32)[None, :]
    yield "flat_and_ramp", d
    yield "one_pixel", np.array([[1.23456]], np.float32)
    yield "column", rng.uniform(0, 6.5, size=(70, 1)).astype(np.float32)
    yield "wide", rng.uniform(0, 6.5, size=(3, 2000)).astype(np.float32)
    yield "vga_clipped", np.clip(rng.uniform(0, 9, size=(480, 640)).astype(np.float32), 0.1, 6.5)
    z = np.zeros((40, 64), np.float32)
    z[:, 31:35] = 3.0
    yield "group_boundary", z
    yield "big_values", np.full((33, 65), 6.5, np.float32)      # 65000 = 0xFDE8: 9-bit literals


@pytest.mark.parametrize("name,depth", list(_cases()), ids=[n for n, _ in _cases()])
def test_device_png_is_the_oracle_stream(sage, oracle, name, depth):
    import torch
    H, W = depth.shape
    enc = sage.DepthPngEncoder(W, H)
    png, pixels = enc.encode(torch.from_numpy(depth).cuda(), want_pixels=True)
    want = (np.clip(depth, 0.0, 6.5) * 10000.0).astype(np.uint16)          # the client's own expression
    assert np.array_equal(pixels, want)
    assert png == oracle.depth_png_encode(depth), name
    im = Image.open(io.BytesIO(png))
    im.load()                                                                 # chunk CRCs + Adler-32 verified by the reader
    assert np.array_equal(np.asarray(im).astype(np.uint16), want)
    enc.close()


def test_other_clip_and_scale_and_argument_checks(sage, oracle):
    import torch
    rng = np.random.default_rng(2)
    depth = rng.uniform(0, 60, size=(50, 70)).astype(np.float32)
    enc = sage.DepthPngEncoder(70, 50)
    png = enc.encode(torch.from_numpy(depth).cuda(), lo=0.5, hi=50.0, scale=1000.0)
    assert png == oracle.depth_png_encode(depth, 0.5, 50.0, 1000.0)
    with pytest.raises(sage.SageError):
        enc.encode(torch.from_numpy(depth).cuda(), lo=0.0, hi=10.0, scale=10000.0)      # 100000 does not fit uint16
    with pytest.raises(ValueError):
        enc.encode(torch.zeros((50, 71), device="cuda"))
    enc.close()


def test_rendered_frame_depth_png_fused_with_the_frame(sage, oracle):
    """sage_frame_depth_png_async on the slot a frame was rendered on: 1080p, through the C-ABI."""
    import torch
    cloud, layout = sage.synthetic_room(0, 60_000)
    cam = sage.Camera.look_heading((1.5, 1.0, 1.2), np.radians(30.0), 1920, 1080)
    scene = sage.GaussianScene(cloud)
    scene.set_depth_clip(0.1, 6.5)                              # as the facade does
    enc = sage.DepthPngEncoder(1920, 1080)
    rgb, depth = scene.render_host_async(cam, 1, want_depth=True)
    twin = torch.zeros((1080, 1920), dtype=torch.uint16).pin_memory()
    scene.depth_png_async(enc, 1, twin=twin)
    png = scene.depth_png_wait(enc, 1)
    scene.host_wait(1)
    want = (np.clip(depth, 0.0, 6.5) * 10000.0).astype(np.uint16)
    assert np.array_equal(twin.numpy(), want)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(png))).astype(np.uint16), want)
    assert png == oracle.depth_png_encode(depth)
    assert len(png) < want.nbytes * 1.13                        # fixed Huffman: at most 9 bits per byte
    # a slot whose frame has no depth is refused
    scene.render_host_async(cam, 2, want_depth=False)
    scene.host_wait(2)
    with pytest.raises(sage.SageError):
        scene.depth_png_async(enc, 2)
    enc.close(); scene.close()


def test_facade_depth_handoff_through_the_clients_own_lines(sage, tmp_path):
    cv2 = pytest.importorskip("cv2")
    from sage3d_official_b200 import handoff
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    cloud, layout = sage.synthetic_room(1, 40_000)
    try:
        env = SimpleVLNEnv(str(tmp_path / "mem.npz"), resolution=(640, 480), gaussians=cloud, jpeg_handoff=True)
        env.set_start_pose([1.5, 1.0, 0.85], [0.0, 0.0, 0.0, 1.0])
        h0 = handoff.cv2_hook_stats()["hits"]
        depth_image = env.get_depth()
        plain = np.asarray(depth_image).copy()
        # vlm_client_modular.py:354-358
        depth_array_clamped = np.clip(depth_image, 0.0, 6.5)
        depth_array_encoded = (depth_array_clamped * 10000.0).astype(np.uint16)
        _, depth_buffer = cv2.imencode('.png', depth_array_encoded)
        depth_bytes = depth_buffer.tobytes()
        assert handoff.cv2_hook_stats()["hits"] == h0 + 1
        handoff.uninstall_cv2_hook()
        _, host_buffer = cv2.imencode('.png', (np.clip(plain, 0.0, 6.5) * 10000.0).astype(np.uint16))
        dev = cv2.imdecode(np.frombuffer(depth_bytes, np.uint8), cv2.IMREAD_UNCHANGED)
        ref = cv2.imdecode(host_buffer, cv2.IMREAD_UNCHANGED)
        assert dev.dtype == np.uint16 and np.array_equal(dev, ref)
        env.close()
    finally:
        handoff.uninstall_cv2_hook()
        handoff.uninstall_pillow_hook()
