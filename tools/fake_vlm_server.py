#!/usr/bin/env python
"""Scripted stand-in for the VLM servers of the reference (evaluation_model/*), speaking the
length-prefixed JSON socket protocol of its client — ``SocketProtocol._send_rgb_request``
(Code/benchmark/environment_evaluation/vlm_client_modular.py:440-491):

    request   8-byte big-endian length | JSON {"images": [base64 JPEG, ...], "query": str}
    response  8-byte big-endian length | JSON (here: a text action, parsed by TextOutputParser :221-300)

One connection per query, as the client opens them.  The action script is deterministic (a fixed
cycle of text actions, 0 ms think time) so that an episode through the UNMODIFIED run_benchmark.py is
reproducible.  The server also keeps what it saw — requests, bytes, the
JPEG sizes and SOI markers of the frames — so a harness can check the frames really arrived.

    python tools/fake_vlm_server.py --port 54321 [--script forward,forward,left,...] [--stats out.json]
"""
from __future__ import annotations

import argparse
import base64
import json
import socket
import threading
import time

DEFAULT_SCRIPT = ("move forward 0.25 meters", "move forward 0.25 meters", "turn left 15 degrees",
                  "move forward 0.25 meters", "turn right 15 degrees", "move forward 0.25 meters",
                  "turn right 15 degrees", "move forward 0.25 meters", "turn left 15 degrees")


class FakeVLMServer:
    def __init__(self, host: str = "127.0.0.1", port: int = 0, script=DEFAULT_SCRIPT, think_s: float = 0.0,
                 check_images: bool = True):
        self.script = tuple(script)
        self.think_s = float(think_s)
        self.check_images = check_images
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind((host, port))
        self.sock.listen(16)
        self.host, self.port = self.sock.getsockname()
        self.stats = {"requests": 0, "bytes_in": 0, "images": 0, "jpeg_bytes": 0, "bad_images": 0,
                      "serve_s": 0.0, "first_query": None}
        self._stop = threading.Event()
        self._thread = None

    @staticmethod
    def _recv_exact(conn, n: int) -> bytes:
        buf = bytearray()
        while len(buf) < n:
            chunk = conn.recv(min(1 << 20, n - len(buf)))
            if not chunk:
                raise ConnectionError("peer closed the connection")
            buf += chunk
        return bytes(buf)

    def _serve_one(self, conn) -> None:
        t0 = time.perf_counter()
        size = int.from_bytes(self._recv_exact(conn, 8), "big")
        req = json.loads(self._recv_exact(conn, size).decode("utf-8"))
        images = req.get("images", [])
        self.stats["requests"] += 1
        self.stats["bytes_in"] += size + 8
        self.stats["images"] += len(images)
        if self.stats["first_query"] is None:
            self.stats["first_query"] = str(req.get("query", ""))[:200]
        if self.check_images:
            for b64 in images:
                raw = base64.b64decode(b64)
                self.stats["jpeg_bytes"] += len(raw)
                if raw[:2] != b"\xff\xd8" or raw[-2:] != b"\xff\xd9":
                    self.stats["bad_images"] += 1
        if self.think_s > 0:
            time.sleep(self.think_s)
        action = self.script[(self.stats["requests"] - 1) % len(self.script)]
        data = json.dumps(action).encode("utf-8")
        conn.sendall(len(data).to_bytes(8, "big"))
        conn.sendall(data)
        self.stats["serve_s"] += time.perf_counter() - t0

    def serve_forever(self) -> None:
        self.sock.settimeout(0.2)
        while not self._stop.is_set():
            try:
                conn, _ = self.sock.accept()
            except socket.timeout:
                continue
            except OSError:
                break
            with conn:
                try:
                    conn.settimeout(60)
                    self._serve_one(conn)
                except Exception as e:          # a broken request must not take the server down
                    self.stats.setdefault("errors", []).append(repr(e))

    def start(self) -> "FakeVLMServer":
        self._thread = threading.Thread(target=self.serve_forever, daemon=True)
        self._thread.start()
        return self

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        try:
            self.sock.close()
        except OSError:
            pass
        return dict(self.stats)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=54321)
    ap.add_argument("--script", default=None, help="comma-separated text actions (default: a 9-step forward/turn cycle)")
    ap.add_argument("--think-ms", type=float, default=0.0)
    ap.add_argument("--stats", default=None, help="write the request statistics here on SIGINT/exit")
    args = ap.parse_args()
    script = tuple(s.strip() for s in args.script.split(",")) if args.script else DEFAULT_SCRIPT
    srv = FakeVLMServer(args.host, args.port, script, args.think_ms / 1000.0)
    print(f"[fake_vlm] listening on {srv.host}:{srv.port}, {len(script)}-action script", flush=True)
    try:
        srv.serve_forever()
    except KeyboardInterrupt:
        pass
    st = srv.stop()
    if args.stats:
        with open(args.stats, "w") as f:
            json.dump(st, f)
    print(json.dumps(st))


if __name__ == "__main__":
    main()
