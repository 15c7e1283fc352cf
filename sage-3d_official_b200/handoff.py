"""Frame hand-off without host-side JPEG work.

After the render is ~1 ms, the reference's step is dominated by what it does to the frame on the
host: ``run_benchmark.py:1421`` wraps every frame with ``Image.fromarray(rgb)``, and the VLM client
re-encodes its whole 8-frame history with Pillow on EVERY query
(``vlm_client_modular.py:445-455``: ``img.convert('RGB'); img.save(buffer, format='JPEG')``);
``generate_images.py:431-432`` does ``Image.fromarray(img_array).save(path)`` per waypoint.
A 1080p Pillow encode costs ~13 ms of one host core.

The engine already produced those exact bytes on the GPU (csrc/jpeg.cu is byte-identical to
Pillow's default encoder).  This module lets the unmodified callers pick them up:

* :class:`DeviceFrame` — the ``np.ndarray`` the façade returns from ``get_rgb()`` /
  ``Camera.get_rgba()``; an ordinary, writable, caller-owned array that also remembers the JPEG
  the device made of the frame and where the untouched frame lives.
* :func:`install_pillow_hook` — wraps ``PIL.Image.fromarray`` (nothing else).  If the array it is
  given IS, bit for bit, the frame the device encoded (checked with a full compare against the
  engine's pinned copy, ~0.3 ms), the returned image is a :class:`PrecodedImage` whose
  ``save(fp, format='JPEG')`` / ``save('x.jpg')`` with no encoder options writes the device's
  bytes.  Any mutation (Pillow mutators copy-on-write a read-only image; we detect the swap), any
  option (quality=, optimize=, ...), any other format or mode falls through to Pillow itself.
  The image owns a private copy of the pixels and unpacks it into Pillow storage only when something
  asks for pixels (``load()``): a frame that is only wrapped, kept in a history and saved as JPEG —
  the VLM client's whole use of it — never pays Pillow's per-frame unpack either.

The rule is airtight by construction: pre-encoded bytes are used only when the pixels being saved
are provably the pixels that were encoded, and for those pixels the two encoders agree byte for
byte (tests/test_jpeg_gpu.py).  The hook is opt-in: the launchers install it
(tools/launch_sage_bench.py, isaac_shim.install(jpeg_handoff=True)) or ``SAGE_JPEG_HANDOFF=1``.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np

_JPEG_EXT = (".jpg", ".jpeg", ".jpe", ".jfif")


class DeviceFrame(np.ndarray):
    """An RGB(A) frame that remembers its device-encoded JPEG.

    ``_sage`` = (jpeg bytes, pristine RGB array the bytes were made from, validity callable).
    Views and copies inherit it; whether it still applies is decided by content, not provenance
    (see :func:`precoded_jpeg`)."""

    def __new__(cls, arr, jpeg: Optional[bytes], pristine: np.ndarray, valid: Callable[[], bool]):
        obj = np.asarray(arr).view(cls)
        obj._sage = (jpeg, pristine, valid) if jpeg is not None else None
        return obj

    def __array_finalize__(self, obj):
        self._sage = getattr(obj, "_sage", None)

    def __reduce__(self):          # pickling / multiprocessing: degrade to a plain array
        return np.asarray(self).__reduce__()


def _same_bytes(a: np.ndarray, b: np.ndarray) -> bool:
    """Exact equality of two same-shape uint8 arrays; 8 bytes per comparison when both are contiguous."""
    if a.flags.c_contiguous and b.flags.c_contiguous and a.nbytes % 8 == 0:
        return bool(np.array_equal(a.reshape(-1).view(np.uint64), b.reshape(-1).view(np.uint64)))
    return bool(np.array_equal(a, b))


def precoded_jpeg(arr) -> Optional[bytes]:
    """The device's JPEG of ``arr`` if ``arr`` is bit-identical to the encoded frame, else None."""
    meta = getattr(arr, "_sage", None)
    if meta is None or not isinstance(arr, np.ndarray):
        return None
    jpeg, pristine, valid = meta
    try:
        if arr.dtype != np.uint8 or arr.shape != pristine.shape or not valid():
            return None
        if not _same_bytes(np.asarray(arr), pristine):
            return None
    except Exception:
        return None
    return jpeg


_hook_state = {"installed": False, "orig": None, "cls": None, "hits": 0, "misses": 0}


def hook_stats() -> dict:
    return {"hits": _hook_state["hits"], "misses": _hook_state["misses"]}


def _core(im):
    """The image's core storage or None, without tripping Pillow >= 10's asserting ``im`` property
    (older Pillow keeps it as a plain attribute)."""
    d = im.__dict__
    return d["_im"] if "_im" in d else d.get("im")


def _precoded_class():
    if _hook_state["cls"] is not None:
        return _hook_state["cls"]
    from PIL import Image

    class PrecodedImage(Image.Image):
        """An RGB image that can hand out the JPEG its pixels already have.  Pixel storage is
        created lazily (``load()``, which every Pillow operation calls before touching ``im``): a
        frame that is only ever saved as a default JPEG never pays Pillow's 6 MB unpack."""
        _sage_jpeg: Optional[bytes] = None
        _sage_im = None          # the core image the bytes describe, once materialised
        _sage_arr = None         # private copy of the pixels, until materialised

        def load(self):
            if self._sage_arr is not None and _core(self) is None:
                core = _hook_state["orig"](self._sage_arr)
                self.im = core.im
                self._sage_im = self.im
                self._sage_arr = None            # (readonly stays 1: Pillow's mutators copy on write)
            return super().load()

        def _sage_valid(self) -> bool:
            if self._sage_jpeg is None or self.mode != "RGB" or self.info:
                return False
            if self._sage_arr is not None:                  # never materialised: nothing could have changed
                return _core(self) is None
            return bool(self.readonly) and _core(self) is self._sage_im

        def save(self, fp, format=None, **params):
            if not params and self._sage_valid():
                fmt = format
                if fmt is None and isinstance(fp, (str, bytes, os.PathLike)):
                    name = os.fsdecode(fp).lower()
                    fmt = "JPEG" if name.endswith(_JPEG_EXT) else None
                if isinstance(fmt, str) and fmt.upper() == "JPEG":
                    _hook_state["hits"] += 1
                    if isinstance(fp, (str, bytes, os.PathLike)):
                        with open(fp, "wb") as f:
                            f.write(self._sage_jpeg)
                    else:
                        fp.write(self._sage_jpeg)
                        if hasattr(fp, "flush"):
                            fp.flush()
                    return None
            _hook_state["misses"] += 1
            return super().save(fp, format, **params)

        def convert(self, mode=None, *args, **kwargs):
            # img.convert('RGB') on an RGB image is an identity copy (vlm_client_modular.py:448-451):
            # share the (read-only) pixels instead of copying 6 MB; Pillow's mutators copy on write
            if mode in (None, "RGB") and not args and not kwargs and self._sage_valid():
                if self._sage_arr is not None:
                    return _lazy(self._sage_arr, self._sage_jpeg)
                new = self._new(self.im)
                return _promote(new, self._sage_jpeg)
            return super().convert(mode, *args, **kwargs)

    _hook_state["cls"] = PrecodedImage
    return PrecodedImage


def _promote(im, jpeg: bytes):
    im.__class__ = _precoded_class()
    im._sage_jpeg = jpeg
    im._sage_im = im.im
    im._sage_arr = None
    im.readonly = 1
    return im


def _lazy(arr: np.ndarray, jpeg: bytes):
    """A PrecodedImage over a PRIVATE pixel array (never written again); no Pillow storage yet."""
    im = _precoded_class()()
    im._size = (int(arr.shape[1]), int(arr.shape[0]))
    im._mode = "RGB"
    im.readonly = 1              # from birth: putpixel / paste / ImageDraw then _copy() before writing
    im._sage_jpeg = jpeg
    im._sage_arr = arr
    im._sage_im = None
    return im


def install_pillow_hook() -> bool:
    """Wrap ``PIL.Image.fromarray`` once (idempotent).  Returns True if the hook is in place."""
    if _hook_state["installed"]:
        return True
    try:
        from PIL import Image
    except Exception:
        return False
    orig = Image.fromarray
    _hook_state["orig"] = orig
    _precoded_class()

    def fromarray(obj, mode=None, *a, **kw):
        if mode is None and not a and not kw:
            data = precoded_jpeg(obj)
            if data is not None and obj.ndim == 3 and obj.shape[2] == 3:
                # own the pixels (the caller may write to its array later); ~0.3 ms per 1080p frame
                return _lazy(np.array(obj, dtype=np.uint8, order="C", copy=True).view(np.ndarray), data)
            return orig(obj)
        return orig(obj, mode, *a, **kw)

    fromarray.__doc__ = orig.__doc__
    fromarray._sage_hook = True
    Image.fromarray = fromarray
    _hook_state["installed"], _hook_state["orig"] = True, orig
    return True


def uninstall_pillow_hook() -> None:
    if _hook_state["installed"]:
        from PIL import Image
        Image.fromarray = _hook_state["orig"]
        _hook_state["installed"], _hook_state["orig"] = False, None


def handoff_requested() -> bool:
    return os.environ.get("SAGE_JPEG_HANDOFF", "0").strip().lower() in ("1", "true", "yes", "on")
