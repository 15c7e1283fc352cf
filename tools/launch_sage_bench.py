#!/usr/bin/env python
"""Run the reference's entry points UNMODIFIED on top of this engine.

    python tools/launch_sage_bench.py /path/to/run_benchmark.py  [its own arguments ...]
    python tools/launch_sage_bench.py /path/to/generate_images.py [its own arguments ...]

`run_benchmark.py` finds the drop-in `simple_env` / `collision_detector` modules in sys.modules
(it force-inserts its own directory at sys.path[0], so only pre-registration wins);
`generate_images.py` finds the Isaac stand-ins of `isaac_shim`.

Frame hand-off (handoff.py) is ON here unless SAGE_JPEG_HANDOFF=0: every frame is also JPEG-encoded
on the device, and the scripts' own `Image.fromarray(frame).save(..., 'JPEG')` calls pick those
bytes up (identical to what Pillow would have produced) instead of encoding on the host.
"""
import runpy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = sys.argv[1]
    import os
    os.environ.setdefault("SAGE_JPEG_HANDOFF", "1")
    import sage3d_official_b200.simple_env as fast_env
    import sage3d_official_b200.collision_detector as fast_det
    import sage3d_official_b200.isaac_shim as shim
    shim.install()
    sys.modules["simple_env"] = fast_env
    sys.modules["collision_detector"] = fast_det
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
