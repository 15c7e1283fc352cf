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

        mine = {"rgb": [export(t) for t in rgb_buffers], "depth": [export(t) for t in depth_buffers]}
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
        self._flag = torch.zeros(1, dtype=torch.int32, device=rgb_buffers[0].device)
        dist.barrier()

    def close_frame(self):
        """Device-side frame barrier on the current stream: once it has run on every rank, all
        ranks' strips have landed in all frame buffers."""
        import torch.distributed as dist
        dist.all_reduce(self._flag)

    def close(self):
        L = _lib.lib()
        for p, off in self._opened:
            L.sage_ipc_close(C.c_void_p(p), C.c_uint64(off))
        self._opened = []
