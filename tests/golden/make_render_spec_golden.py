"""Frozen-spec fixtures for the render path.

Nothing under the reference pins a pixel (its renderer is Isaac Sim's closed RTX path: DESIGN.md section 1), so
these are NOT reference outputs: they are digests of the frames the CPU oracle (oracle/sage_oracle.c, the frozen
spec of DESIGN.md section 2) produces for config 1 (50 k synthetic room, 640 x 480, four seeded poses) and two small
hostile cases.  They make the spec itself a committed artefact: a change of the oracle's arithmetic - deliberate or
not - shows up as a changed digest in review, and the GPU tests compare the CUDA frames with these committed digests
as well as with the oracle built on the box.

    python tests/golden/make_render_spec_golden.py      # rewrites tests/golden/render_spec_digests.json
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).with_name("render_spec_digests.json")


def cases(sage):
    room, _ = sage.synthetic_room(0, 50_000)
    poses = [((1.5, 1.0, 1.2), 30.0), ((3.0, 2.0, 1.2), 200.0), ((5.2, 3.4, 1.2), -135.0), ((0.6, 0.5, 1.2), 45.0)]
    for k, (p, h) in enumerate(poses):
        yield f"config1_room50k_640x480_pose{k}", room, sage.Camera.look_heading(p, np.radians(h), 640, 480)
    small, _ = sage.synthetic_room(3, 4000, sh_degree=2)
    yield "room4k_deg2_333x187", small, sage.Camera.look_heading((2.0, 1.5, 1.0), np.radians(75.0), 333, 187)
    yield "room50k_1920x1080_pose2", room, sage.Camera.look_heading((5.2, 3.4, 1.2), np.radians(-135.0), 1920, 1080)


def cloud_digest(cloud, cam):
    """Digest of the INPUTS (the generator uses NumPy transcendentals, which may differ in the last bit between CPU
    families: a test on a machine whose inputs differ skips the comparison instead of failing it)."""
    h = hashlib.sha256()
    for a in cloud.as_tuple():
        h.update(np.ascontiguousarray(a, np.float32).tobytes())
    for a in (cam.R, cam.t, cam.campos, [cam.fx, cam.fy, cam.cx, cam.cy]):
        h.update(np.ascontiguousarray(a, np.float32).tobytes())
    return h.hexdigest()


def digest(rgb, depth):
    return {"rgb_sha256": hashlib.sha256(np.ascontiguousarray(rgb).tobytes()).hexdigest(),
            "depth_sha256": hashlib.sha256(np.ascontiguousarray(depth, np.float32).tobytes()).hexdigest(),
            "rgb_mean": float(np.mean(rgb)), "covered_px": int(np.isfinite(depth).sum())}


def main():
    import sage3d_official_b200 as sage
    import oracle
    oracle.build()
    out = {}
    for name, cloud, cam in cases(sage):
        oc = oracle.make_camera(cam.R, cam.t, cam.campos, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
        rgb, depth, st = oracle.render(oc, *cloud.as_tuple())
        out[name] = dict(digest(rgb, depth), visible=st["visible"], inputs_sha256=cloud_digest(cloud, cam))
        print(name, out[name])
    OUT.write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
