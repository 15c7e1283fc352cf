#!/usr/bin/env python
"""Golden trajectories for the environment façade's pose + motion logic, produced by driving the
UNMODIFIED reference ``SimpleVLNEnv`` methods (set_start_pose, apply_cmd_for and everything they
call: simple_env.py:1149-1353, 1987-2250) with Isaac Sim stubbed out:

  * ``imageio`` / ``isaacsim`` are replaced by empty stub modules (import-time only);
  * the instance is created without ``__init__`` (which would boot Isaac) and given a fake
    ``world`` (step() is a no-op) and a fake ``cam`` that records ``set_world_pose`` arguments;
  * the collision detector is the reference's own ``SemanticMap2DCollisionDetector``.

Outputs tests/golden/env_episode_<case>.npz: the map JSON, the command script and, after every
command, agent position (+ its dtype), yaw, consecutive_collisions and the camera position /
orientation handed to Isaac.  /root/reference does not exist on the GPU box, hence the fixtures.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REF = "/root/reference/Code/benchmark/environment_evaluation"
HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
os.environ["SILENT_LOGGING_MODE"] = "1"

sys.modules.setdefault("imageio", types.ModuleType("imageio"))
isaac = types.ModuleType("isaacsim")
simapp = types.ModuleType("isaacsim.simulation_app")
simapp.SimulationApp = type("SimulationApp", (), {})
sys.modules["isaacsim"] = isaac
sys.modules["isaacsim.simulation_app"] = simapp
sys.path.insert(0, REF)
import simple_env as ref_env  # noqa: E402  (the reference, unmodified)
from collision_detector import SemanticMap2DCollisionDetector  # noqa: E402

sys.path.insert(0, str(ROOT))
import sage3d_official_b200 as sage  # noqa: E402  (only for the synthetic layouts / maps)
from sage3d_official_b200.scenes import layout_semantic_map  # noqa: E402


class FakeWorld:
    def step(self, render=False):
        pass


class FakeCam:
    def __init__(self):
        self.position = None
        self.orientation = None

    def set_world_pose(self, position=None, orientation=None):
        self.position = np.array(position)
        self.orientation = np.array(orientation)


def make_reference_env(map_path):
    env = object.__new__(ref_env.SimpleVLNEnv)
    env.is_stop_called = False
    env.consecutive_collisions = 0
    env._total_collision_count = 0
    env._debug_disable_collision = False
    env._collision_detected = False
    env._log_function = None
    env.world = FakeWorld()
    env.cam = FakeCam()
    env._pos = np.array([0.0, 0.0, 0.85], dtype=np.float32)
    env._yaw = 0.0
    env.collision_detector = SemanticMap2DCollisionDetector(map_path, robot_radius_m=0.08, scale=0.05)
    return env


def script(rng, n):
    """NaVILA/NaVid-like discrete actions plus a few continuous ones."""
    cmds = []
    for _ in range(n):
        r = rng.uniform()
        if r < 0.55:
            cmds.append((0.25, 0.0, 0.0, 1.0))                         # forward 25 cm
        elif r < 0.70:
            cmds.append((0.0, 0.0, math_rad(15.0), 1.0))               # turn left 15 deg
        elif r < 0.85:
            cmds.append((0.0, 0.0, -math_rad(15.0), 1.0))
        elif r < 0.93:
            cmds.append((rng.uniform(0.0, 0.5), rng.uniform(-0.2, 0.2), rng.uniform(-0.5, 0.5), rng.uniform(0.2, 1.0)))
        else:
            cmds.append((0.03, 0.0, 0.0, 1.0))                         # tiny move (< 5 cm rule)
    return np.array(cmds, dtype=np.float64)


def math_rad(d):
    return float(np.radians(d))


def run_case(name, layout, start_xy, heading2d, seed, n_cmds):
    inst = layout_semantic_map(layout)
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(inst, f)
        map_path = f.name
    env = make_reference_env(map_path)
    det = env.collision_detector
    # 2-D map frame -> trajectory ("3-D") frame is a translation (trajectory_2d_to_3d.py:56-63)
    sx, sy = det.min_x + det.max_x, det.min_y + det.max_y
    position = [start_xy[0] - sx, start_xy[1] - sy, 0.85]
    psi = heading2d + np.pi
    psi = (psi + np.pi) % (2 * np.pi) - np.pi
    rotation = [-float(np.sin(psi / 2)), 0.0, 0.0, float(np.cos(psi / 2))]       # trajectory_2d_to_3d.py:160-171
    with contextlib.redirect_stdout(io.StringIO()):      # the reference's _log prints unconditionally
        env.set_start_pose(position, rotation)
    rng = np.random.default_rng(seed)
    cmds = script(rng, n_cmds)
    rec = {"pos": [], "pos_is_f32": [], "yaw": [], "cc": [], "cam_pos": [], "cam_ori": []}

    def snap():
        rec["pos"].append(np.asarray(env._pos, np.float64).copy())
        rec["pos_is_f32"].append(env._pos.dtype == np.float32)
        rec["yaw"].append(env._yaw)
        rec["cc"].append(env.consecutive_collisions)
        rec["cam_pos"].append(np.asarray(env.cam.position, np.float64).copy())
        rec["cam_ori"].append(np.asarray(env.cam.orientation, np.float64).copy())

    snap()
    for c in cmds:
        with contextlib.redirect_stdout(io.StringIO()):
            env.apply_cmd_for(*[float(v) for v in c])
        snap()
    os.unlink(map_path)
    moved = np.linalg.norm(np.diff(np.array(rec["pos"])[:, :2], axis=0), axis=1)
    print(name, "cmds", len(cmds), "moved steps", int((moved > 1e-6).sum()), "blocked", int((moved < 1e-6).sum()),
          "max consecutive", max(rec["cc"]), "f64 positions", int((~np.array(rec["pos_is_f32"])).sum()))
    np.savez_compressed(HERE / f"env_episode_{name}.npz", map_json=json.dumps(inst), start_position=np.array(position),
                        start_rotation=np.array(rotation), cmds=cmds, pos=np.array(rec["pos"]),
                        pos_is_f32=np.array(rec["pos_is_f32"]), yaw=np.array(rec["yaw"]), cc=np.array(rec["cc"]),
                        cam_pos=np.array(rec["cam_pos"]), cam_ori=np.array(rec["cam_ori"]), numpy_version=np.__version__)


def main():
    _, room = sage.synthetic_room(0, 10)
    _, interior = sage.synthetic_interior(1, 10)
    run_case("room", room, (1.0, 1.0), 0.5, seed=1, n_cmds=400)
    p0, h0 = sage.camera_path(interior, 1, seed=3)[0]
    run_case("interior", interior, (float(p0[0]), float(p0[1])), float(h0), seed=2, n_cmds=600)


if __name__ == "__main__":
    main()
