"""TEST launcher: the reference's entry points UNMODIFIED on top of the drop-in modules, with the
façade's engine hooks pointed at the CPU oracle (tests/oracle_backend.py) — what
tools/launch_sage_bench.py does with the GPU engine, for a machine without one.  Test infrastructure."""
import runpy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    script = sys.argv[1]
    import sage3d_official_b200.simple_env as fast_env
    import sage3d_official_b200.collision_detector as fast_det
    from oracle_backend import OracleScene, make_oracle_detector
    fast_env.SimpleVLNEnv._scene_factory = OracleScene
    fast_env.SimpleVLNEnv._detector_factory = make_oracle_detector()
    sys.modules["simple_env"] = fast_env
    sys.modules["collision_detector"] = fast_det
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
