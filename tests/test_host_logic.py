"""CPU: host-side logic of the drop-in — camera conventions, scene generators, map generator."""
import json

import numpy as np
import pytest


def test_camera_axes_follow_isaac_world_convention(sage):
    # identity orientation (w=1): camera looks along +X, +Z up  
    cam = sage.Camera.from_isaac_pose((0, 0, 1.2), (1, 0, 0, 0), 640, 480)
    p = np.array([5.0, 0.0, 1.2])                       # straight ahead
    pc = cam.R.astype(np.float64) @ p + cam.t
    assert pc[2] == pytest.approx(5.0) and abs(pc[0]) < 1e-6 and abs(pc[1]) < 1e-6
    left = cam.R.astype(np.float64) @ np.array([5.0, 1.0, 1.2]) + cam.t
    assert left[0] < 0                                   # +Y world (left) -> negative image x
    up = cam.R.astype(np.float64) @ np.array([5.0, 0.0, 2.2]) + cam.t
    assert up[1] < 0                                     # +Z world (up) -> negative image y (y down)


def test_unnormalised_orientation_is_normalised(sage):
    a = sage.Camera.from_isaac_pose((1, 2, 1.2), (0.7, 0, 0, 0.2), 64, 48)
    b = sage.Camera.from_isaac_pose((1, 2, 1.2), (7.0, 0, 0, 2.0), 64, 48)
    assert np.allclose(a.R, b.R, atol=1e-6) and np.allclose(a.t, b.t, atol=1e-5)
    with pytest.raises(ValueError):
        sage.Camera.from_isaac_pose((0, 0, 0), (0, 0, 0, 0), 64, 48)


def test_intrinsics_match_reference_focal_and_aperture(sage):
    fx, fy, cx, cy = sage.Camera.intrinsics(640, 480)
    assert fx == pytest.approx(8.0 / 20.955 * 640, rel=1e-6) and fx == fy      # 244.3 px
    assert (cx, cy) == (320.0, 240.0)
    assert sage.Camera.intrinsics(1920, 1080)[0] == pytest.approx(733.0, abs=0.1)


def test_usd_model_transform_fold(sage):
    from sage3d_official_b200.camera import USD_GAUSS_MODEL_TO_WORLD as M
    p_model = np.array([0.3, -1.2, 2.0])
    p_world = M @ p_model
    assert np.allclose(p_world, [p_model[0], p_model[2], -p_model[1]])       # (x, z, -y)
    cw = sage.Camera.from_isaac_pose((1, 2, 1.2), (0.9, 0, 0, 0.43), 64, 48)
    cm = sage.Camera.from_isaac_pose((1, 2, 1.2), (0.9, 0, 0, 0.43), 64, 48, model_to_world=M)
    assert np.allclose(cw.R.astype(np.float64) @ p_world + cw.t, cm.R.astype(np.float64) @ p_model + cm.t, atol=1e-5)
    assert np.allclose(M @ cm.campos.astype(np.float64), [1, 2, 1.2], atol=1e-6)


def test_generators_are_seeded_and_shaped(sage):
    a, la = sage.synthetic_room(0, 2000)
    b, _ = sage.synthetic_room(0, 2000)
    c, _ = sage.synthetic_room(1, 2000)
    for x, y in zip(a.as_tuple(), b.as_tuple()):
        assert np.array_equal(x, y) and x.dtype == np.float32 and x.flags["C_CONTIGUOUS"]
    assert not np.array_equal(a.xyz, c.xyz)
    assert a.sh.shape == (2000, 16, 3) and a.sh_degree == 3
    big, layout = sage.synthetic_interior(1, 5000)
    assert big.n == 5000 and len(layout["walls"]) == 6 and len(layout["boxes"]) == 40
    assert big.xyz[:, 0].min() > -0.5 and big.xyz[:, 0].max() < 12.5
    poses = sage.camera_path(layout, 32, seed=0)
    assert len(poses) == 32 and all(p[0][2] == 1.2 for p in poses)


def test_layout_semantic_map_has_reference_format(sage, tmp_path):
    from sage3d_official_b200.scenes import layout_semantic_map, write_semantic_map
    _, layout = sage.synthetic_room(0, 100)
    inst = layout_semantic_map(layout)
    labels = {i["category_label"] for i in inst}
    assert {"floor", "wall", "Unable Area"} <= labels
    y, x = inst[0]["mask_coords_m"][0]
    assert isinstance(y, str) and isinstance(x, str) and len(y.split(".")[1]) == 2   # 2-decimal strings, [y, x]
    p = write_semantic_map(layout, tmp_path / "map.json")
    assert json.load(open(p))[1]["category_label"] == "wall"


def test_composite_box_bound_is_conservative():
    """tools/composite_box_bound_check.py: the binary32 box bound of k_composite's per-stage filter never falls below
    the exponent of a pixel in the box (400 000 random conics and boxes, needle-thin splats included)."""
    import subprocess
    import sys
    from pathlib import Path
    tool = Path(__file__).resolve().parents[1] / "tools" / "composite_box_bound_check.py"
    r = subprocess.run([sys.executable, str(tool), "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "wrong skips 0" in r.stdout


def test_culling_is_conservative_against_the_rounded_exponent():
    """tools/cuda_emu/run_band_check.py: K1a's pixel box and K2's band / quadrant-mask helpers (compiled from
    csrc/sage_internal.cuh with g++) against the blend's own binary32 exponent at every pixel of a 1080p image, for
    random splats from blobs to needles thousands of pixels long: no pixel that passes the alpha >= 1/255 test may lie
    outside the box, in a tile that is not emitted, or in a quadrant whose bit is clear."""
    import json
    import shutil
    import subprocess
    import sys
    from pathlib import Path
    import pytest
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    tool = Path(__file__).resolve().parents[1] / "tools" / "cuda_emu" / "run_band_check.py"
    r = subprocess.run([sys.executable, str(tool), "300", "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["violations"] == 0 and out["pixels_passing_the_rounded_test"] > 1_000_000
