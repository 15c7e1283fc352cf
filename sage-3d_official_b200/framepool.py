"""Pinned host frame buffers lent to callers as NumPy views (shared by the two seams)."""
from __future__ import annotations

import weakref


class FramePool:
    """Pinned host frame buffers handed out as NumPy views and taken back when the last reference to
    the view (or to anything derived from it) is gone."""

    def __init__(self, make_pair, limit: int = 16):
        self._make, self._limit = make_pair, limit
        self._free, self._n = [], 0

    def take(self):
        """(pair, token) — token must be passed to lend(); None when the pool is exhausted."""
        if self._free:
            return self._free.pop()
        if self._n >= self._limit:
            return None
        self._n += 1
        return self._make()

    def lend(self, pair, arrays):
        """The arrays (views of the pair's buffers) go to the caller; the pair returns to the pool
        when all of them have been garbage-collected."""
        state = {"left": len(arrays)}

        def back(_state=state, _pair=pair, _pool=weakref.ref(self)):
            _state["left"] -= 1
            pool = _pool()
            if _state["left"] == 0 and pool is not None:
                pool._free.append(_pair)

        for a in arrays:
            weakref.finalize(a, back)

    def give_back(self, pair):
        self._free.append(pair)
