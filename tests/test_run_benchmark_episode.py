"""CPU: BASELINE config 5 in miniature — the UNMODIFIED reference run_benchmark.py (its step loop
:1331-1698) drives the drop-in SimpleVLNEnv for a few steps against the scripted socket VLM
(tools/fake_vlm_server.py, the length-prefixed JSON protocol of vlm_client_modular.py:440-491).
Here the façade's engine hooks point at the CPU oracle (tests/oracle_backend.py); on the GPU box the
same harness runs with the real engine (tools/run_benchmark_episode.py, profiles/episode_r02_*.json).
Skipped where neither /root/reference nor oracle/_ref is present."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))


def _have_reference():
    return any((d / "run_benchmark.py").exists() for d in (ROOT / "oracle" / "_ref" / "environment_evaluation",
                                                           Path("/root/reference/Code/benchmark/environment_evaluation")))


@pytest.mark.skipif(not _have_reference(), reason="reference run_benchmark.py not present")
def test_unmodified_run_benchmark_steps_through_the_dropin(oracle):
    import run_benchmark_episode as H
    # 20 steps: more frames than the facade's pinned-frame pool lends out (run_benchmark.py:1422 keeps them all)
    out = H.run(steps=20, width=64, height=48, n=3000, timeout=600,
                launcher=[sys.executable, str(ROOT / "tests" / "launch_with_oracle.py")])
    assert out["returncode"] == 0, out.get("stderr_tail")
    assert out["steps"] == 20 and out["vlm_requests"] == 20, out
    assert out["vlm_bad_images"] == 0 and out["vlm_images"] >= 20         # real JPEGs of the frames reached the server
    assert out["engine_frames"] >= 20 and out["resolution"] == [64, 48]   # a frame per step; SAGE_RESOLUTION reached the ctor
    assert out["collide_launches"] >= 1 and out["measurements_keys"]


def test_fake_vlm_server_speaks_the_clients_protocol():
    """Round trip with the reference client's own framing (8-byte big-endian length + JSON)."""
    import base64
    import io
    import json
    import socket

    import numpy as np
    from PIL import Image
    from fake_vlm_server import FakeVLMServer
    srv = FakeVLMServer(script=("move forward 0.25 meters", "turn left 15 degrees")).start()
    try:
        answers = []
        for k in range(3):
            buf = io.BytesIO()
            Image.fromarray(np.full((8, 8, 3), 40 * k, np.uint8)).save(buf, format="JPEG")
            data = json.dumps({"images": [base64.b64encode(buf.getvalue()).decode()], "query": "go"}).encode()
            s = socket.create_connection((srv.host, srv.port), timeout=10)
            s.sendall(len(data).to_bytes(8, "big")); s.sendall(data)
            size = int.from_bytes(s.recv(8), "big")
            resp = b""
            while len(resp) < size:
                resp += s.recv(4096)
            s.close()
            answers.append(json.loads(resp.decode()))
    finally:
        st = srv.stop()
    assert answers == ["move forward 0.25 meters", "turn left 15 degrees", "move forward 0.25 meters"]
    assert st["requests"] == 3 and st["images"] == 3 and st["bad_images"] == 0
