"""Multi-GPU rendering: one process per GPU, the scene replicated, every frame split into
contiguous tile-row blocks.

Two exchange paths:

* ``PeerFrames`` (NVLink, the product path) — each rank's composite kernel stores its finished
  pixels straight into every rank's frame buffer through peer-mapped pointers
  (``sage_render_peers``): the strip exchange is fused into the kernel's epilogue, nothing is
  staged or copied afterwards.  A one-element NCCL all-reduce on the render stream closes the frame.
* ``gather_strips`` — a plain ``all_gather`` of padded strips over whatever backend
  ``torch.distributed`` was initialised with.  It is the fallback where peer mapping is not
  available and the path the CPU (gloo, world_size 2) tests exercise.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check

TILE = 16


def tile_row_cuts(height: int, world: int) -> List[int]:
    """Boundaries of the per-rank tile-row blocks: 68 rows over 8 ranks -> 9,9,9,9,8,8,8,8."""
    ty = (height + TILE - 1) // TILE
    base, rem = divmod(ty, world)
    cuts = [0]
    for r in range(world):
        cuts.append(cuts[-1] + base + (1 if r < rem else 0))
    return cuts


def balanced_cuts(row_cost: Sequence[float], world: int) -> List[int]:
    """Tile-row boundaries that equalise a per-row cost (e.g. instances per row of a calibration
    frame plus a per-pixel term) instead of the row count; every rank keeps at least one row."""
    cost = np.asarray(row_cost, dtype=np.float64)
    n = len(cost)
    if world >= n:
        return list(range(n)) + [n] * (world - n + 1)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target))
        k = min(max(k, cuts[-1] + 1), n - (world - r))
        if k - 1 > cuts[-1] and abs(cum[k - 1] - target) < abs(cum[k] - target):
            k -= 1
        cuts.append(k)
    cuts.append(n)
    return cuts


def rows_of(rank: int, height: int, world: int) -> Tuple[int, int]:
    cuts = tile_row_cuts(height, world)
    return cuts[rank], cuts[rank + 1]


def pixel_rows(rows: Tuple[int, int], height: int) -> Tuple[int, int]:
    return rows[0] * TILE, min(rows[1] * TILE, height)


def gather_strips(strip, height: int, width: int, rank: int, world: int, group=None):
    """All-gather per-rank strips (tensor [rows_r, W, ...]) into the full frame on every rank.

    Strips are padded to the tallest block so one ``all_gather`` moves them (uneven last blocks)."""
    import torch
    import torch.distributed as dist
    cuts = tile_row_cuts(height, world)
    heights = [pixel_rows((cuts[r], cuts[r + 1]), height) for r in range(world)]
    max_h = max(b - a for a, b in heights)
    pad_shape = (max_h,) + tuple(strip.shape[1:])
    padded = torch.zeros(pad_shape, dtype=strip.dtype, device=strip.device)
    padded[: strip.shape[0]] = strip
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    frame = torch.empty((height,) + tuple(strip.shape[1:]), dtype=strip.dtype, device=strip.device)
    for r, (a, b) in enumerate(heights):
        frame[a:b] = out[r][: b - a]
    return frame


class PeerFrames:
    """Frame buffers of all ranks, mapped into this process (CUDA IPC over NVLink)."""

    def __init__(self, rgb_buffers: Sequence, depth_buffers: Sequence, rank: int, world: int):
        import torch
        import torch.distributed as dist
        self.rank, self.world = rank, world
        self._opened = []
        L = _lib.lib()

        def export(t):
            h = (C.c_uint8 * 64)()
            off = C.c_uint64(0)
            check(L.sage_ipc_export(C.c_void_p(t.data_ptr()), C.cast(h, C.c_void_p), C.byref(off)))
            return bytes(h), int(off.value)

        # one row of frame counters per buffer slot: flags[b][r] = last frame rank r completed into slot b
        self._flags = torch.zeros((len(rgb_buffers), 8), dtype=torch.int32, device=rgb_buffers[0].device)
        self._seq = [0] * len(rgb_buffers)
        mine = {"rgb": [export(t) for t in rgb_buffers], "depth": [export(t) for t in depth_buffers],
                "flags": [export(self._flags[b]) for b in range(len(rgb_buffers))]}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        torch.cuda.synchronize()

        def open_all(kind, local):
            table = []
            for b, t in enumerate(local):
                ptrs = []
                for r in range(world):
                    if r == rank:
                        ptrs.append(int(t.data_ptr()))
                    else:
                        hb, off = everyone[r][kind][b]
                        buf = (C.c_uint8 * 64).from_buffer_copy(hb)
                        p = C.c_void_p()
                        check(L.sage_ipc_open(C.cast(buf, C.c_void_p), C.c_uint64(off), C.byref(p)))
                        self._opened.append((p.value, off))
                        ptrs.append(int(p.value))
                table.append(ptrs)
            return table

        self.rgb = open_all("rgb", rgb_buffers)
        self.depth = open_all("depth", depth_buffers)
        self.flags = open_all("flags", [self._flags[b] for b in range(len(rgb_buffers))])
        # ready-made pointer tables for sage_render_peers / sage_peer_barrier: nothing is rebuilt per frame
        tab = lambda ptrs: (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in ptrs])
        self.rgb = [tab(p) for p in self.rgb]
        self.depth = [tab(p) for p in self.depth]
        self._flag_rows = [tab(p) for p in self.flags]
        self._flag = torch.zeros(1, dtype=torch.int32, device=rgb_buffers[0].device)
        self.use_nccl = os.environ.get("SAGE_PEER_BARRIER", "flags") == "nccl"
        dist.barrier()

    def close_frame(self, b: int = 0, stream=None):
        """Device-side frame barrier on the current stream for buffer slot ``b``: once it has run on every
        rank, all ranks' strips have landed in all copies of the slot.  Counters in peer memory
        (``sage_peer_barrier``: one 32-thread kernel, no collective); ``SAGE_PEER_BARRIER=nccl`` switches back
        to the one-element NCCL all-reduce for comparison."""
        if self.use_nccl:
            import torch.distributed as dist
            dist.all_reduce(self._flag)
            return
        import torch
        self._seq[b] += 1
        st = stream if stream is not None else torch.cuda.current_stream()
        check(_lib.lib().sage_peer_barrier(self._flag_rows[b], self.world, self.rank, self._seq[b], C.c_void_p(st.cuda_stream)))

    def close(self):
        L = _lib.lib()
        for p, off in self._opened:
            L.sage_ipc_close(C.c_void_p(p), C.c_uint64(off))
        self._opened = []


class SharedHostFrames:
    """Host-facing multi-GPU delivery: ONE pinned host frame (per pipeline slot) shared by all ranks of
    the node; every rank copies the rows it rendered straight into it over its own PCIe link
    (``sage_render_host_async`` with its tile-row range) — N links instead of one, and no NVLink hop
    or collective on the host-facing path.  A frame is delivered when all N strips have landed:
    ranks publish a per-slot frame counter in the same shared segment, rank 0 (the caller's side)
    waits for all of them and acknowledges, which is also what lets a slot be reused.

    The segment is POSIX shared memory, page-locked in every process with ``cudaHostRegister``."""

    def __init__(self, height: int, width: int, depth: int, rank: int, world: int, timeout_s: float = 60.0,
                 pin: bool = True):
        from multiprocessing import shared_memory
        import torch.distributed as dist
        self.H, self.W, self.d, self.rank, self.world = int(height), int(width), int(depth), rank, world
        self.timeout_s = timeout_s
        al = lambda n: (n + 4095) // 4096 * 4096
        rgb_b, dep_b = al(self.H * self.W * 3), al(self.H * self.W * 4)
        ctrl_off = self.d * (rgb_b + dep_b)
        size = ctrl_off + 4096
        name = [None]
        if rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=size)
            name = [self._shm.name]
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
        self._buf = np.ndarray((size,), np.uint8, buffer=self._shm.buf)
        self._rt = None
        if pin:                     # (pin=False: the CPU protocol test, no device in the process)
            import torch
            self._rt = torch.cuda.cudart()
            err = self._rt.cudaHostRegister(self._buf.ctypes.data, size, 0)
            if int(err) != 0:
                raise RuntimeError(f"cudaHostRegister failed on the shared frame segment: {err}")
        self.rgb, self.dep = [], []
        for k in range(self.d):
            o = k * (rgb_b + dep_b)
            self.rgb.append(self._buf[o:o + self.H * self.W * 3].reshape(self.H, self.W, 3))
            self.dep.append(self._buf[o + rgb_b:o + rgb_b + self.H * self.W * 4].view(np.float32).reshape(self.H, self.W))
        ctrl = self._buf[ctrl_off:ctrl_off + 4096].view(np.int64)
        self.done = ctrl[:self.d * world].reshape(self.d, world)      # done[slot][rank] = frames of that slot landed
        self.consumed = ctrl[self.d * world:self.d * world + self.d]  # consumed[slot] = frames rank 0 has taken
        if rank == 0:
            ctrl[:] = 0
        dist.barrier()

    def _spin(self, cond, what: str):
        t0 = time.perf_counter()
        while not cond():
            if time.perf_counter() - t0 > self.timeout_s:
                raise TimeoutError(f"SharedHostFrames: rank {self.rank} timed out waiting for {what}")

    def submit(self, scene, cam, rows, i: int) -> None:
        """Start frame i (slot i % depth): this rank's rows, into the shared host frame."""
        slot, gen = i % self.d, i // self.d
        if gen > 0:       # the slot's previous frame must have been delivered before it is overwritten
            self._spin(lambda: self.consumed[slot] >= gen, f"slot {slot} to be consumed")
        scene.render_host_async(cam, slot, rows=rows, out=(self.rgb[slot], self.dep[slot]))

    def finish(self, scene, i: int, release: bool = True) -> None:
        """Frame i: this rank's strip has landed; rank 0 additionally waits for every rank's strip — the
        frame is then complete in :meth:`frame`.  ``release=False`` keeps the slot (rank 0 reads the frame
        first and calls :meth:`release` itself)."""
        slot, gen = i % self.d, i // self.d
        scene.host_wait(slot)
        self.done[slot, self.rank] = gen + 1
        if self.rank == 0:
            self._spin(lambda: bool((self.done[slot] >= gen + 1).all()), f"all strips of frame {i}")
            if release:
                self.release(i)

    def release(self, i: int) -> None:
        """Rank 0: frame i has been taken; its slot may be overwritten."""
        if self.rank == 0:
            self.consumed[i % self.d] = i // self.d + 1

    def frame(self, i: int):
        return self.rgb[i % self.d], self.dep[i % self.d]

    def drain(self, scene) -> None:
        for k in range(self.d):
            scene.host_wait(k)

    def close(self) -> None:
        try:
            if self._rt is not None:
                self._rt.cudaHostUnregister(self._buf.ctypes.data)
        except Exception:
            pass
        self.rgb, self.dep, self.done, self.consumed, self._buf = [], [], None, None, None
        try:
            self._shm.close()
            if self.rank == 0:
                self._shm.unlink()
        except Exception:
            pass
