#!/usr/bin/env python
"""Stage the reference's OWN Python for this path under oracle/_ref/ (git-ignored; it travels to the
GPU box with gpurun, where /root/reference does not exist).  TEST / BENCH INFRASTRUCTURE ONLY.

Copied verbatim, never edited, never committed:
  Code/benchmark/environment_evaluation/*.py   run_benchmark.py (the config-5 harness), its helpers
      (measures, task_types, episodes_adapter, vlm_client_modular, object_based_success) and
      collision_detector.py — the reference's check_collision_3d, timed beside the GPU collide
      in bench.py (cpu_baseline kind "reference") and used as the live oracle of the collide tests.
The reference's simple_env.py is NOT staged: it needs Isaac Sim and is what the drop-in replaces.

    python oracle/make_ref.py [/root/reference]
"""
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref" / "environment_evaluation"
SKIP = {"simple_env.py"}


def make(ref_root="/root/reference") -> Path:
    src = Path(ref_root) / "Code" / "benchmark" / "environment_evaluation"
    if not src.is_dir():
        raise FileNotFoundError(f"{src} not found (the reference tree is only present in the authoring container)")
    DEST.mkdir(parents=True, exist_ok=True)
    n = 0
    for f in sorted(src.glob("*.py")):
        if f.name in SKIP:
            continue
        shutil.copyfile(f, DEST / f.name)
        n += 1
    (DEST.parent / "README").write_text(
        "Verbatim copies of reference files made by oracle/make_ref.py; git-ignored, not product source.\n")
    return DEST


def ref_dir():
    """oracle/_ref/environment_evaluation if staged, else the live reference tree, else None."""
    if (DEST / "collision_detector.py").exists():
        return DEST
    live = Path("/root/reference/Code/benchmark/environment_evaluation")
    return live if live.is_dir() else None


if __name__ == "__main__":
    d = make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print(f"staged {len(list(d.glob('*.py')))} reference files under {d}")
