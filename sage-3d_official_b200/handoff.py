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
  bytes.  Any option (quality=, optimize=, ...), any other format or mode falls through to Pillow.
  The image owns a private copy of the pixels and has NO Pillow pixel storage until something asks
  for pixels: a frame that is only wrapped, kept in a history and saved as JPEG — the VLM client's
  whole use of it — never pays Pillow's per-frame unpack.

The rule holds by construction: the precoded bytes exist only while the image is "virtual".  Every
Pillow path to the pixels goes through ``im``; the first access materialises ordinary, writable
storage from the private copy and drops the bytes for good — including the in-place mutators that
ignore ``readonly`` (``frombytes``, ``putdata``).  From then on the object is an ordinary image:
writable through ``load()[x, y] = v``, equal to ``Image.fromarray(rgb)``, picklable (it reduces to a
plain ``Image``).  For the untouched pixels the two encoders agree byte for byte
(tests/test_jpeg_gpu.py).  The hook is opt-in: the launchers install it
(tools/launch_sage_bench.py, isaac_shim.install(jpeg_handoff=True)) or ``SAGE_JPEG_HANDOFF=1``.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np

_JPEG_EXT = (".jpg", ".jpeg", ".jpe", ".jfif")


class DeviceFrame(np.ndarray):
    """An RGB(A) frame that remembers its device-encoded JPEG.

    ``_sage`` = (jpeg bytes | None, pristine RGB array the bytes were made from, validity callable).
    Views and copies inherit it; whether it still applies is decided by content, not provenance
    (see :func:`precoded_jpeg`).

    An RGBA frame (``Camera.get_rgba()``) additionally knows it is one (``_sage_rgba`` = the root
    array): ``frame[:, :, :3].copy()`` — what generate_images.py:431 does to every frame, a 3-of-4-byte
    strided copy NumPy runs at ~0.5 GB/s — is served by OpenCV's packed conversion instead (~10x
    faster, same bytes)."""

    def __new__(cls, arr, jpeg: Optional[bytes], pristine: Optional[np.ndarray], valid: Callable[[], bool],
                rgba_root: bool = False):
        obj = np.asarray(arr).view(cls)
        obj._sage = (jpeg, pristine, valid) if (jpeg is not None and pristine is not None) else None
        obj._sage_rgba = np.asarray(arr) if rgba_root else None
        return obj

    def __array_finalize__(self, obj):
        self._sage = getattr(obj, "_sage", None)
        self._sage_rgba = getattr(obj, "_sage_rgba", None)

    def __reduce__(self):          # pickling / multiprocessing: degrade to a plain array
        return np.asarray(self).__reduce__()

    def _is_rgb_view_of_root(self) -> bool:
        root = self._sage_rgba
        return (root is not None and self.ndim == 3 and self.shape[2] == 3 and root.ndim == 3 and root.shape[2] == 4
                and self.dtype == np.uint8 and self.shape[:2] == root.shape[:2] and self.strides == root.strides
                and self.__array_interface__["data"][0] == root.__array_interface__["data"][0])

    def copy(self, order="C"):
        if order in ("C", "K", "A") and self._is_rgb_view_of_root():
            rgb = _rgba_to_rgb(self._sage_rgba)
            if rgb is not None:
                out = rgb.view(DeviceFrame)
                out._sage, out._sage_rgba = self._sage, None      # packed RGB: comparable with the pristine twin
                return out
        return super().copy(order)


def _rgba_to_rgb(rgba: np.ndarray) -> Optional[np.ndarray]:
    """Packed (H,W,3) copy of the colour channels of a C-contiguous (H,W,4) array, or None without OpenCV."""
    try:
        import cv2
        if rgba.flags.c_contiguous:
            return cv2.cvtColor(rgba, cv2.COLOR_RGBA2RGB)
    except Exception:
        pass
    return None


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


class PrivateTwin(np.ndarray):
    """The engine's own copy of an encoded frame (never handed to a caller, never written): the hook may
    use it as a PrecodedImage's pixels without copying."""
    _sage_private = True


_hook_state = {"installed": False, "orig": None, "hits": 0, "misses": 0}


def hook_stats() -> dict:
    return {"hits": _hook_state["hits"], "misses": _hook_state["misses"]}


def _core(im):
    """The image's pixel storage or None, without materialising a virtual PrecodedImage."""
    return im.__dict__.get("_im")


def _plain_fromarray(arr):
    """Pillow's own ``Image.fromarray`` (the un-hooked one)."""
    if _hook_state["orig"] is not None:
        return _hook_state["orig"](arr)
    from PIL import Image
    return Image.fromarray(arr)


def _rebuild_plain(mode, size, data):
    from PIL import Image
    return Image.frombytes(mode, size, data)


try:
    from PIL import Image as _PILImage
except Exception:                                   # Pillow is optional: without it there is no hook
    _PILImage = None


if _PILImage is not None:

    class PrecodedImage(_PILImage.Image):
        """An RGB image that can hand out the JPEG its pixels already have.

        The invariant is structural: the precoded bytes are valid ONLY while the image has never
        been given Pillow pixel storage.  The pixels live in a private NumPy copy; the moment anything
        reaches for ``im`` — ``load()``, ``frombytes()``, ``putdata()``, ``tobytes()``, a filter, a
        comparison — the storage is materialised from that copy and the JPEG is dropped for good, so no
        in-place mutator (not even the ones that ignore ``readonly``) can leave stale bytes behind.
        After that the object behaves exactly like the image ``Image.fromarray`` would have returned
        (writable, comparable, picklable)."""

        _sage_jpeg: Optional[bytes] = None
        _sage_arr = None         # private copy of the pixels while the image is still "virtual"

        # ``im`` is what every Pillow code path touches before reading or writing pixels (a plain
        # attribute up to Pillow 9, a property asserting on None from 10 on): materialise on demand.
        @property
        def im(self):
            core = self.__dict__.get("_im")
            if core is None and self._sage_arr is not None:
                arr, self._sage_arr, self._sage_jpeg = self._sage_arr, None, None
                core = _plain_fromarray(arr).im
                self.__dict__["_im"] = core
                self.readonly = 0
            if core is None:
                raise ValueError("image has no pixel storage")
            return core

        @im.setter
        def im(self, core):
            self._sage_arr, self._sage_jpeg = None, None
            self.__dict__["_im"] = core

        # Pillow >= 10 keeps the storage in ``_im`` and reads it directly in a few places
        @property
        def _im(self):
            if self.__dict__.get("_im") is None and self._sage_arr is not None:
                return self.im
            return self.__dict__.get("_im")

        @_im.setter
        def _im(self, core):
            if core is not None:
                self._sage_arr, self._sage_jpeg = None, None
            self.__dict__["_im"] = core

        def load(self):
            self.im                                  # materialise (drops the precoded state)
            return super().load()

        def _sage_valid(self) -> bool:
            return (self._sage_jpeg is not None and self._sage_arr is not None and self.__dict__.get("_im") is None
                    and self.mode == "RGB" and not self.info)

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
            # share the private (never written) pixel array instead of copying 6 MB
            if mode in (None, "RGB") and not args and not kwargs and self._sage_valid():
                return _lazy(self._sage_arr, self._sage_jpeg)
            return super().convert(mode, *args, **kwargs)

        def copy(self):
            if self._sage_valid():
                return _lazy(self._sage_arr, self._sage_jpeg)
            return super().copy()

        def __eq__(self, other):
            if not isinstance(other, _PILImage.Image):
                return NotImplemented
            return (self.mode == other.mode and self.size == other.size and self.info == other.info
                    and self.getpalette() == other.getpalette() and self.tobytes() == other.tobytes())

        __hash__ = None

        def __reduce__(self):                        # pickling / multiprocessing: a plain Image of the same pixels
            return (_rebuild_plain, (self.mode, self.size, self.tobytes()))

else:
    PrecodedImage = None


def _lazy(arr: np.ndarray, jpeg: bytes):
    """A PrecodedImage over a PRIVATE pixel array (never written again); no Pillow storage yet."""
    im = PrecodedImage()
    im._size = (int(arr.shape[1]), int(arr.shape[0]))
    im._mode = "RGB"
    im.readonly = 0
    im._sage_jpeg = jpeg
    im._sage_arr = arr
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

    def fromarray(obj, mode=None, *a, **kw):
        if mode is None and not a and not kw:
            data = precoded_jpeg(obj)
            if data is not None and obj.ndim == 3 and obj.shape[2] == 3:
                pristine = obj._sage[1]
                if getattr(pristine, "_sage_private", False):
                    # the twin is never handed to the caller and just compared equal: it IS the pixels
                    return _lazy(pristine, data)
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


# ---- depth: uint16 PNG (vlm_client_modular.py:354-364) -----------------------------------------------
class DeviceDepth(np.ndarray):
    """A depth map that remembers the device-made PNG of its quantised form.

    ``_sage_png`` = (png bytes, the uint16 pixels the file holds, validity callable).  NumPy carries the
    attribute through the client's own ``np.clip(depth, 0.0, 6.5) * 10000.0`` and ``.astype(np.uint16)``
    (vlm_client_modular.py:354-355); whether the file still applies to the array that finally reaches the
    encoder is decided by comparing pixels (:func:`precoded_png`), never by provenance."""

    def __new__(cls, arr, png: Optional[bytes], pixels_u16: Optional[np.ndarray], valid: Callable[[], bool]):
        obj = np.asarray(arr).view(cls)
        obj._sage_png = (png, pixels_u16, valid) if (png is not None and pixels_u16 is not None) else None
        return obj

    def __array_finalize__(self, obj):
        self._sage_png = getattr(obj, "_sage_png", None)

    def __reduce__(self):
        return np.asarray(self).__reduce__()


def precoded_png(arr) -> Optional[bytes]:
    """The device's PNG if ``arr`` is, bit for bit, the uint16 image that file holds; else None."""
    meta = getattr(arr, "_sage_png", None)
    if meta is None or not isinstance(arr, np.ndarray):
        return None
    png, pixels, valid = meta
    try:
        if arr.dtype != np.uint16 or arr.shape != pixels.shape or not valid():
            return None
        a, b = np.ascontiguousarray(np.asarray(arr)).reshape(-1), pixels.reshape(-1)
        if a.nbytes % 8 == 0 and b.flags.c_contiguous:
            same = bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))
        else:
            same = bool(np.array_equal(a, b))
        return png if same else None
    except Exception:
        return None


_cv2_state = {"installed": False, "orig": None, "hits": 0, "misses": 0}


def cv2_hook_stats() -> dict:
    return {"hits": _cv2_state["hits"], "misses": _cv2_state["misses"]}


def install_cv2_hook() -> bool:
    """Wrap ``cv2.imencode`` once: ``cv2.imencode('.png', depth_u16)`` with no parameters, on an array whose
    pixels equal the device-encoded image, returns the device's file (a valid PNG of the same pixels — not
    OpenCV's bytes, see csrc/png.cu).  Everything else goes to OpenCV."""
    if _cv2_state["installed"]:
        return True
    try:
        import cv2
    except Exception:
        return False
    orig = cv2.imencode

    def imencode(ext, img, params=None):
        if params is None and isinstance(ext, str) and ext.lower() == ".png":
            data = precoded_png(img)
            if data is not None:
                _cv2_state["hits"] += 1
                return True, np.frombuffer(data, np.uint8).copy()
        _cv2_state["misses"] += 1
        plain = np.asarray(img)
        return orig(ext, plain) if params is None else orig(ext, plain, params)

    imencode._sage_hook = True
    cv2.imencode = imencode
    _cv2_state["installed"], _cv2_state["orig"] = True, orig
    return True


def uninstall_cv2_hook() -> None:
    if _cv2_state["installed"]:
        import cv2
        cv2.imencode = _cv2_state["orig"]
        _cv2_state["installed"], _cv2_state["orig"] = False, None


def handoff_requested() -> bool:
    return os.environ.get("SAGE_JPEG_HANDOFF", "0").strip().lower() in ("1", "true", "yes", "on")
