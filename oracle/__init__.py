"""CPU oracle bindings — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package never does.

The C source (``sage_oracle.c``) documents which reference lines each function restates and
what pins it: grid-collide is pinned by golden vectors generated from the unmodified reference
class; render and capsule-collide are **parity unpinned** (the reference has no source, test or
golden vector for them) and this oracle is the frozen specification.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libsage_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (seconds).  Rebuilds when the source is newer."""
    newest = max(p.stat().st_mtime for p in _HERE.glob("*.c"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libsage_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class Camera(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("campos", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("W", C.c_int32), ("H", C.c_int32), ("near_z", C.c_float), ("far_z", C.c_float)]


SPLAT_DTYPE = np.dtype([(k, np.float32) for k in
                        ("gx", "gy", "qa", "qb", "qc", "lo", "z", "r", "g", "b", "cov_a", "cov_c")])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.so_exp2.restype = C.c_float
        _lib.so_exp2.argtypes = [C.c_float]
        _lib.so_exp.restype = C.c_float
        _lib.so_exp.argtypes = [C.c_float]
        _lib.so_render.restype = C.c_int
        _lib.so_abi_version.restype = C.c_int
    return _lib


def make_camera(R, t, campos, fx, fy, cx, cy, W, H, near=0.1, far=50.0) -> Camera:
    cam = Camera()
    cam.R[:] = [float(v) for v in np.asarray(R, np.float32).reshape(9)]
    cam.t[:] = [float(v) for v in np.asarray(t, np.float32).reshape(3)]
    cam.campos[:] = [float(v) for v in np.asarray(campos, np.float32).reshape(3)]
    cam.fx, cam.fy, cam.cx, cam.cy = float(fx), float(fy), float(cx), float(cy)
    cam.W, cam.H = int(W), int(H)
    cam.near_z, cam.far_z = float(near), float(far)
    return cam


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty=C.c_float):
    return a.ctypes.data_as(C.POINTER(ty))


def exp2(x: float) -> float:
    return float(lib().so_exp2(C.c_float(x)))


def set_threads(n: int) -> None:
    lib().so_set_threads(C.c_int(n))


def project(cam: Camera, xyz, log_scale, quat, opacity_raw, sh):
    """Per-Gaussian projection.  Returns (records[SPLAT_DTYPE], visible[u8])."""
    xyz, log_scale, quat, opacity_raw, sh = map(_f32, (xyz, log_scale, quat, opacity_raw, sh))
    n = xyz.shape[0]
    nb = sh.shape[1]
    deg = int(round(np.sqrt(nb))) - 1
    out = np.zeros(n, dtype=SPLAT_DTYPE)
    vis = np.zeros(n, dtype=np.uint8)
    lib().so_project(C.byref(cam), C.c_int64(n), C.c_int(deg), _p(xyz), _p(log_scale), _p(quat),
                     _p(opacity_raw), _p(sh), out.ctypes.data_as(C.c_void_p), _p(vis, C.c_uint8))
    return out, vis


def render(cam: Camera, xyz, log_scale, quat, opacity_raw, sh, brute: bool = False, want_T: bool = False):
    """Full frame.  Returns (rgb u8 [H,W,3], depth f32 [H,W], stats dict[, final T f32 [H,W]])."""
    xyz, log_scale, quat, opacity_raw, sh = map(_f32, (xyz, log_scale, quat, opacity_raw, sh))
    n = xyz.shape[0]
    nb = sh.shape[1]
    deg = int(round(np.sqrt(nb))) - 1
    H, W = cam.H, cam.W
    rgb = np.zeros((H, W, 3), np.uint8)
    depth = np.zeros((H, W), np.float32)
    fT = np.zeros((H, W), np.float32) if want_T else None
    stats = np.zeros(3, np.int64)
    rc = lib().so_render(C.byref(cam), C.c_int64(n), C.c_int(deg), _p(xyz), _p(log_scale), _p(quat),
                         _p(opacity_raw), _p(sh), C.c_int(1 if brute else 0), _p(rgb, C.c_uint8),
                         _p(depth), _p(fT) if want_T else None, _p(stats, C.c_int64))
    if rc != 0:
        raise RuntimeError(f"so_render failed rc={rc}")
    st = {"visible": int(stats[0]), "instances": int(stats[1]), "threads": int(stats[2])}
    return (rgb, depth, st, fT) if want_T else (rgb, depth, st)


def collide_grid(grid_map, min_x, max_x, min_y, max_y, scale, pos, np2_promote=None):
    """Restatement of SemanticMap2DCollisionDetector.check_collision_3d over a batch.

    pos float32 -> ``np2_promote`` picks NumPy>=2 (True) or NumPy 1.26 (False) scalar promotion;
    pos float64 -> all-double path (promotion-independent).
    """
    grid_map = None if grid_map is None else np.ascontiguousarray(grid_map, np.uint8)
    H, W = (0, 0) if grid_map is None else grid_map.shape
    pos = np.asarray(pos)
    mp = None if grid_map is None else _p(grid_map, C.c_uint8)
    dbl = [C.c_double(float(v)) for v in (min_x, max_x, min_y, max_y, scale)]
    if pos.dtype == np.float64:
        pos = np.ascontiguousarray(pos.reshape(-1, 3))
        hit = np.zeros(pos.shape[0], np.uint8)
        lib().so_collide_grid_f64(mp, C.c_int(H), C.c_int(W), *dbl, _p(pos, C.c_double),
                                  C.c_int(pos.shape[0]), _p(hit, C.c_uint8))
    else:
        if np2_promote is None:
            np2_promote = int(np.__version__.split(".")[0]) >= 2
        pos = _f32(pos).reshape(-1, 3)
        hit = np.zeros(pos.shape[0], np.uint8)
        lib().so_collide_grid_f32(mp, C.c_int(H), C.c_int(W), *dbl, _p(pos), C.c_int(pos.shape[0]),
                                  C.c_int(1 if np2_promote else 0), _p(hit, C.c_uint8))
    return hit.astype(bool)


def collide_capsule(mu, log_scale, quat, pos, radius=0.1, z0=0.6, z1=1.1):
    mu, log_scale, quat, pos = map(_f32, (mu, log_scale, quat, pos))
    pos = pos.reshape(-1, 3)
    m, P = mu.shape[0], pos.shape[0]
    hit = np.zeros(P, np.uint8)
    first = np.zeros(P, np.int32)
    lib().so_collide_capsule(C.c_int64(m), _p(mu), _p(log_scale), _p(quat), _p(pos), C.c_int(P),
                             C.c_float(radius), C.c_float(z0), C.c_float(z1), _p(hit, C.c_uint8),
                             _p(first, C.c_int32))
    return hit.astype(bool), first


def jpeg_encode(rgb, quality: int = 75, want_coefs: bool = False):
    """Restatement of Pillow's ``Image.fromarray(rgb).save(buf, format='JPEG', quality=quality)``
    (jpeg_oracle.c).  rgb u8 [H,W,3] -> bytes (and, optionally, the quantised coefficients
    int16 [blocks,64] in zig-zag / scan order)."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    H, W, ch = rgb.shape
    assert ch == 3
    cap = W * H * 4 + 8192
    out = np.zeros(cap, np.uint8)
    n = C.c_int64(0)
    nblk = ((W + 15) // 16) * ((H + 15) // 16) * 6
    coefs = np.zeros((nblk, 64), np.int16) if want_coefs else None
    L = lib()
    L.so_jpeg_encode_coefs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                       C.POINTER(C.c_int64), C.c_void_p]
    rc = L.so_jpeg_encode_coefs(rgb.ctypes.data, W, H, int(quality), out.ctypes.data, cap, C.byref(n),
                                coefs.ctypes.data if want_coefs else None)
    if rc != 0:
        raise RuntimeError(f"so_jpeg_encode failed rc={rc}")
    data = out[:n.value].tobytes()
    return (data, coefs) if want_coefs else data


def grid_inflate(raw, scale: float, radius: float):
    """Restatement of ``distance_transform_edt(raw == 0, sampling=scale) <= radius`` (plan_oracle.c)."""
    raw = np.ascontiguousarray(raw, np.uint8)
    H, W = raw.shape
    out = np.zeros((H, W), np.uint8)
    L = lib()
    L.so_grid_inflate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
    rc = L.so_grid_inflate(raw.ctypes.data, H, W, float(scale), float(radius), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"so_grid_inflate rc={rc}")
    return out


def grid_components(grid):
    """4-connected components of the free cells: int32 labels (smallest linear index; -1 = obstacle)."""
    grid = np.ascontiguousarray(grid, np.uint8)
    H, W = grid.shape
    lab = np.zeros((H, W), np.int32)
    L = lib()
    L.so_grid_components.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rc = L.so_grid_components(grid.ctypes.data, H, W, lab.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"so_grid_components rc={rc}")
    return lab


def astar(grid, start, goal, max_path: int = 0, want_expansions: bool = False):
    """Restatement of ``astar_pixel(grid, start, goal)`` (vln_trajectory_generator.py:253-286):
    list of (x, y) from start to goal, or None."""
    grid = np.ascontiguousarray(grid, np.uint8)
    H, W = grid.shape
    mp = int(max_path) if max_path else H * W
    path = np.zeros((mp, 2), np.int32)
    nexp = C.c_int64(0)
    L = lib()
    L.so_astar.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    n = L.so_astar(grid.ctypes.data, H, W, int(start[0]), int(start[1]), int(goal[0]), int(goal[1]),
                   path.ctypes.data, mp, C.byref(nexp))
    if n < 0:
        raise RuntimeError(f"so_astar rc={n}")
    res = None if n == 0 else [(int(x), int(y)) for x, y in path[:n]]
    return (res, int(nexp.value)) if want_expansions else res


# ---- depth PNG (csrc/png.cu) --------------------------------------------------------------------
def depth_quantise(depth, lo=0.0, hi=6.5, scale=10000.0):
    """The reference's own sequence (vlm_client_modular.py:354-355): binary32 clip, binary32 multiply,
    truncation to uint16."""
    d = np.asarray(depth, np.float32)
    return (np.clip(d, np.float32(lo), np.float32(hi)) * np.float32(scale)).astype(np.uint16)


def depth_png_encode(depth, lo=0.0, hi=6.5, scale=10000.0):
    """NumPy restatement of the device's depth PNG stream, byte for byte (csrc/png.cu): Sub filter
    (bpp 2), one fixed-Huffman deflate block, tokens per group of 32 filtered bytes (a run of >= 4 equal
    bytes inside a group = literal + (length, distance 1)), zlib wrapper, 32 KB IDAT chunks.
    The format itself is pinned by zlib/Pillow decoding it (tests/test_png_oracle.py)."""
    import struct
    import zlib
    q = depth_quantise(depth, lo, hi, scale)
    H, W = q.shape
    raw = np.empty((H, 2 * W), np.uint8)
    raw[:, 0::2] = (q >> 8).astype(np.uint8)
    raw[:, 1::2] = (q & 0xFF).astype(np.uint8)
    sub = raw.copy()
    sub[:, 2:] = raw[:, 2:] - raw[:, :-2]
    L = 1 + 2 * W
    F = np.empty((H, L), np.uint8)
    F[:, 0] = 1
    F[:, 1:] = sub
    G = (L + 31) // 32
    P = np.zeros((H, G * 32), np.int32)
    P[:, :L] = F
    have = np.zeros((H, G * 32), bool)
    have[:, :L] = True
    P, have = P.reshape(H, G, 32), have.reshape(H, G, 32)
    eq = np.zeros_like(have)
    eq[:, :, 1:] = (P[:, :, 1:] == P[:, :, :-1]) & have[:, :, 1:]
    lane = np.arange(32)
    start = np.maximum.accumulate(np.where(~eq, lane, 0), axis=2)
    nxt = np.where(~eq, lane, 32)                              # a lane that starts a run (or is past the data)
    end = np.full_like(start, 32)
    end[:, :, :-1] = np.minimum.accumulate(nxt[:, :, ::-1], axis=2)[:, :, ::-1][:, :, 1:]
    rl, k = end - start, lane - start
    lit = have & ((k == 0) | (rl < 4))
    mat = have & ~lit & (k == 1)
    rev = lambda v, n: sum(((v >> j) & 1) << (n - 1 - j) for j in range(n))
    code = np.zeros(P.shape, np.int64)
    nb = np.zeros(P.shape, np.int64)
    small = lit & (P < 144)
    big = lit & (P >= 144)
    code[small], nb[small] = rev(0x30 + P[small], 8), 8
    code[big], nb[big] = rev(0x190 + (P[big] - 144), 9), 9
    ln = rl - 1
    a, b, c = mat & (ln <= 10), mat & (ln > 10) & (ln <= 18), mat & (ln > 18)
    code[a], nb[a] = rev(ln[a] - 2, 7), 12
    code[b], nb[b] = rev(9 + ((ln[b] - 11) >> 1), 7) | (((ln[b] - 11) & 1) << 7), 13
    code[c], nb[c] = rev(13 + ((ln[c] - 19) >> 2), 7) | (((ln[c] - 19) & 3) << 7), 14
    tok = nb.ravel() > 0
    cv, nv = code.ravel()[tok], nb.ravel()[tok]
    bits = ((cv[:, None] >> np.arange(14)[None, :]) & 1).astype(np.uint8)
    valid = np.arange(14)[None, :] < nv[:, None]
    stream = np.concatenate([np.array([1, 1, 0], np.uint8), bits[valid], np.zeros(7, np.uint8)])
    deflate = np.packbits(stream, bitorder="little").tobytes()
    z = b"\x78\x01" + deflate + struct.pack(">I", zlib.adler32(F.tobytes()))
    chunk = lambda t, d: struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 16, 0, 0, 0, 0))
    for o in range(0, len(z), 32768):
        out += chunk(b"IDAT", z[o:o + 32768])
    return out + chunk(b"IEND", b"")
