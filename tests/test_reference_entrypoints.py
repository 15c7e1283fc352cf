"""CPU: the reference's own entry points load UNMODIFIED on top of the drop-ins (no GPU needed up to
argument parsing).  Skipped where /root/reference is absent (the GPU box)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/Code")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not present")


@pytest.mark.parametrize("script,needle", [
    ("benchmark/environment_evaluation/run_benchmark.py", "--max-steps"),
    ("data_pipeline/training_data_construction/generate_images.py", "--instance-id"),
])
def test_unmodified_reference_script_runs_against_dropins(script, needle, tmp_path):
    env = dict(os.environ, SILENT_LOGGING_MODE="1")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_sage_bench.py"), str(REF / script), "--help"],
                       capture_output=True, text=True, cwd=tmp_path, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert needle in r.stdout


def test_dropin_modules_expose_the_members_run_benchmark_uses():
    sys.path.insert(0, str(ROOT))
    from sage3d_official_b200.simple_env import SimpleVLNEnv
    from sage3d_official_b200.collision_detector import SemanticMap2DCollisionDetector
    for m in ("set_log_function", "set_collision_detection", "set_start_pose", "reset_episode_time",
              "update_time_and_reset_collision", "get_rgb", "get_depth", "get_rgbd", "apply_cmd_for", "get_agent_pos",
              "get_yaw", "get_collision_count", "load_scene", "update_map", "close", "write_video"):
        assert callable(getattr(SimpleVLNEnv, m)), m                      # SURVEY.md §8b seam 1
    for m in ("check_collision_3d", "check_path_collision_3d", "check_collision_at_position", "get_collision_info",
              "forward_position_mapping", "_world_to_pixel", "_pixel_to_world"):
        assert callable(getattr(SemanticMap2DCollisionDetector, m)), m
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_backend import OracleScene
    import sage3d_official_b200 as sage
    SimpleVLNEnv._scene_factory = OracleScene                 # CPU: frames by the oracle (test double)
    try:
        env = SimpleVLNEnv(str(ROOT / "in_memory.npz"), resolution=(32, 24), gaussians=sage.synthetic_room(0, 50)[0])
    finally:
        SimpleVLNEnv._scene_factory = None
    for a in ("is_stop_called", "_current_time", "_episode_start_time", "_collision_detected",
              "consecutive_collisions", "_total_collision_count", "semantic_map_path"):
        assert hasattr(env, a), a
