"""Environment façade (SimpleVLNEnv drop-in) against golden trajectories recorded from the
UNMODIFIED reference motion/pose code (tests/golden/make_env_golden.py).  Positions, yaw,
consecutive-collision counters and the camera pose handed to the renderer must match exactly."""
import glob
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(str(Path(__file__).parent / "golden" / "env_episode_*.npz")))


class OracleDetector:
    """Stand-in detector for the CPU run: same interface as the product's detector, lookups by the
    oracle (test infrastructure).  The product's GPU detector is used in the gpu-marked test."""

    def __init__(self, oracle, map_json, np2):
        from sage3d_official_b200.collision_detector import build_obstacle_grid
        self.o = oracle
        self.grid, (self.min_x, self.max_x, self.min_y, self.max_y) = build_obstacle_grid(json.loads(map_json), 0.08, 0.05)
        self.np2 = np2
        self.calls = 0

    def check_many(self, positions, segments=None):
        self.calls += 1
        pos = np.asarray(positions)
        pos = pos if pos.dtype == np.float32 else pos.astype(np.float64)
        return self.o.collide_grid(self.grid, self.min_x, self.max_x, self.min_y, self.max_y, 0.05, pos, np2_promote=self.np2)

    def check_collision_3d(self, p):
        return bool(self.check_many(np.asarray(p)[None, :])[0])


def _replay(env, z):
    env.set_start_pose([float(v) for v in z["start_position"]], [float(v) for v in z["start_rotation"]])
    got = {"pos": [], "f32": [], "yaw": [], "cc": [], "cam_pos": [], "cam_ori": []}

    def snap():
        got["pos"].append(np.asarray(env._pos, np.float64).copy())
        got["f32"].append(env._pos.dtype == np.float32)
        got["yaw"].append(env.get_yaw())
        got["cc"].append(env.consecutive_collisions)
        got["cam_pos"].append(np.asarray(env._cam_position, np.float64).copy())
        got["cam_ori"].append(np.asarray(env._cam_orientation, np.float64).copy())

    snap()
    for c in z["cmds"]:
        env.apply_cmd_for(*[float(v) for v in c])
        snap()
    assert np.array_equal(np.array(got["pos"]), z["pos"]), "agent positions diverge from the reference"
    assert np.array_equal(np.array(got["f32"]), z["pos_is_f32"])
    assert np.array_equal(np.array(got["yaw"]), z["yaw"])
    assert np.array_equal(np.array(got["cc"]), z["cc"])
    assert np.array_equal(np.array(got["cam_pos"]), z["cam_pos"])
    assert np.array_equal(np.array(got["cam_ori"]), z["cam_ori"])


def _same_numpy_major(z):
    return str(z["numpy_version"]).split(".")[0] == np.__version__.split(".")[0]


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_motion_logic_matches_reference_cpu(sage, oracle, path, tmp_path, monkeypatch):
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    z = np.load(path)
    if not _same_numpy_major(z):
        pytest.skip("golden recorded under a different NumPy major (dtype propagation differs)")
    from oracle_backend import OracleScene
    cloud, _ = sage.synthetic_room(0, 300)
    monkeypatch.setattr(SimpleVLNEnv, "_scene_factory", OracleScene)      # no GPU here: frames by the oracle
    env = SimpleVLNEnv(str(tmp_path / "in_memory.npz"), resolution=(64, 48), gaussians=cloud)
    det = OracleDetector(oracle, str(z["map_json"]), np2=int(np.__version__.split(".")[0]) >= 2)
    env.collision_detector = det
    _replay(env, z)
    # batching: at most 2 dtype groups x 2 phases per command instead of <= 60 sequential lookups
    assert det.calls <= 4 * len(z["cmds"])
    rgb, dep = env.get_rgb(), env.get_depth()
    assert rgb.shape == (48, 64, 3) and dep.shape == (48, 64) and float(dep.min()) >= 0.1 and float(dep.max()) <= 6.5
    assert env.perf["prefetch_hits"] >= 1          # the frame was submitted right after the last pose change
    assert env.get_collision_count() == 0          # _check_collision is dead code in the reference too
    # a scene path without a Gaussian payload is an error, not a black frame
    with pytest.raises(FileNotFoundError):
        SimpleVLNEnv(str(tmp_path / "missing_scene.usda"), resolution=(64, 48))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_facade_on_gpu_matches_reference(sage, oracle, path, tmp_path):
    from conftest import oracle_camera
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    from sage3d_official_b200.scene_io import save_npz
    z = np.load(path)
    if not _same_numpy_major(z):
        pytest.skip("golden recorded under a different NumPy major")
    cloud, layout = (sage.synthetic_room(0, 20_000) if "room" in path else sage.synthetic_interior(1, 60_000))
    save_npz(tmp_path / "scene.npz", cloud)
    mp = tmp_path / "map.json"
    mp.write_text(str(z["map_json"]))
    env = SimpleVLNEnv(str(tmp_path / "scene.npz"), resolution=(160, 120), map_json_path=str(mp))
    logs = []
    env.set_log_function(logs.append)              # this is what creates the detector (simple_env.py:119-123)
    assert env.collision_detector is not None
    _replay(env, z)
    assert env.collide_launches <= 4 * len(z["cmds"])
    # the frame the façade returns == the oracle's frame for the camera the façade built
    rgb, depth = env.get_rgb(), env.get_depth()
    assert rgb.shape == (120, 160, 3) and rgb.dtype == np.uint8 and depth.dtype == np.float32
    o_rgb, o_dep, _ = oracle.render(oracle_camera(oracle, env._camera()), *cloud.as_tuple())
    assert np.array_equal(rgb, o_rgb)
    assert np.array_equal(depth, np.clip(o_dep, 0.1, 6.5))
    n = env.render_calls
    env.get_rgb(); env.get_depth()
    assert env.render_calls == n                   # 1 engine frame per pose, not 5 + 21
    env.close()


@pytest.mark.gpu
def test_ply_roundtrip_and_usda_stage_resolution(sage, tmp_path):
    from sage3d_official_b200.scene_io import save_ply, load_ply, resolve_stage_assets
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    cloud, _ = sage.synthetic_room(2, 3000)
    save_ply(tmp_path / "0001.ply", cloud)
    back = load_ply(tmp_path / "0001.ply")
    for a, b in zip(cloud.as_tuple(), back.as_tuple()):
        assert np.array_equal(a, b)
    usda = tmp_path / "0001.usda"
    usda.write_text('#usda 1.0\ndef Xform "World" {\n def "gauss" ( prepend references = @./0001.usdz@</World/gauss> ) { float3 xformOp:rotateXYZ = (-90, 0, 0) }\n}\n')
    g, c, is_stage = resolve_stage_assets(str(usda))
    assert g == str(tmp_path / "0001.ply") and c is None and is_stage
    env = SimpleVLNEnv(str(usda), resolution=(96, 64))
    env.set_start_pose([1.0, 1.0, 0.85], [0.0, 0.0, 0.0, 1.0])
    assert env._model_to_world is not None          # rotateXYZ(-90,0,0) folded into the view
    assert env.get_rgb().shape == (64, 96, 3)
    env.close()


@pytest.mark.gpu
def test_isaac_shim_serves_generate_images_call_sequence(sage, oracle, tmp_path):
    """The calls generate_images.py makes (lines 320-350, 417-428), against the stand-ins."""
    import sys
    from conftest import oracle_camera
    from sage3d_official_b200.scene_io import save_ply
    from sage3d_official_b200.camera import USD_GAUSS_MODEL_TO_WORLD as M
    import sage3d_official_b200.isaac_shim as shim
    shim.install()
    import omni.usd
    from omni.isaac.kit import SimulationApp
    from omni.isaac.core import World
    from omni.isaac.core.utils.stage import open_stage
    from omni.isaac.sensor import Camera
    from pxr import Gf, UsdGeom, UsdLux

    cloud, _ = sage.synthetic_room(2, 20_000)
    model = sage.GaussianCloud(*(a.copy() for a in cloud.as_tuple()))
    model.xyz[:] = (cloud.xyz.astype(np.float64) @ M).astype(np.float32)    # PLY (model) space
    save_ply(tmp_path / "0007.ply", model)
    usda = tmp_path / "0007.usda"
    usda.write_text('#usda 1.0\ndef Xform "World" { def "gauss" ( prepend references = @./0007.usdz@</World/gauss> ) {} }\n')

    app = SimulationApp({"headless": True})
    omni.usd.get_context().close_stage()
    assert not open_stage(usd_path=str(tmp_path / "missing.usda"))
    assert open_stage(usd_path=str(usda))
    stage = omni.usd.get_context().get_stage()
    if not stage.GetPrimAtPath("/World/EnvLight"):
        dome = UsdLux.DomeLight.Define(stage, "/World/EnvLight")
        dome.CreateIntensityAttr(30000.0)
        dome.CreateColorAttr(Gf.Vec3f(1.0, 1.0, 1.0))
    assert stage.GetPrimAtPath("/World/EnvLight")
    world = World()
    world.reset()
    for _ in range(5):
        world.step(render=True)
    cam = Camera(prim_path="/World/NaVILACamera", frequency=30, resolution=(256, 192))
    cam.initialize()
    UsdGeom.Camera(stage.GetPrimAtPath("/World/NaVILACamera")).GetFocalLengthAttr().Set(8.0)
    position = np.array([1.5, 1.0, 0.3], dtype=np.float32)
    position[2] = 1.2
    rotation = [-0.2588, 0.0, 0.0, 0.9659]                                   # trajectory [qx,qy,qz,qw] passed verbatim
    cam.set_world_pose(position=position, orientation=np.array(rotation, dtype=np.float32))
    for _ in range(3):
        world.step(render=True)
    img = cam.get_rgba()
    assert img.shape == (192, 256, 4) and img.dtype == np.uint8 and (img[..., 3] == 255).all()
    ref_cam = sage.Camera.from_isaac_pose(position, np.array(rotation, dtype=np.float32), 256, 192, model_to_world=M)
    o_rgb, o_dep, _ = oracle.render(oracle_camera(oracle, ref_cam), *model.as_tuple())
    assert np.array_equal(img[:, :, :3], o_rgb)
    assert np.array_equal(cam.get_current_frame()["distance_to_image_plane"], o_dep)
    world.clear()
    omni.usd.get_context().close_stage()
    app.close()
