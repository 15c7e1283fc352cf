"""Scene ingest for the render path: standard 3DGS PLY / NPZ -> GaussianCloud,
and resolution of the assets a SAGE-3D USDA stage refers to.

The reference reaches its Gaussians through a USDA stage that references ``{id}.usdz[gauss.usda]``
and a collision payload ``{id}_collision.usd`` (Code/benchmark/scene_data/sage3d_usda_builder.py:93-149,
Data/template.usda:115-124,156-165).  USDZ/NuRec payloads cannot be parsed here (3DGRUT is not
vendored), so the engine looks for the scene's PLY (the InteriorGS source format, README.md:224-266)
or an ``.npz`` next to the stage.

PLY layout (the de-facto 3DGS export): binary little-endian float32 vertex properties
``x y z [nx ny nz] f_dc_0..2 f_rest_0..(3*((deg+1)^2-1)-1) opacity scale_0..2 rot_0..3``;
f_rest is channel-major (all R coefficients, then G, then B); opacity is a logit, scales are logs,
rot is (w,x,y,z) un-normalised.
"""
from __future__ import annotations

import os
import re
from pathlib import Path
from typing import Optional

import numpy as np

from .scenes import GaussianCloud

_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def _read_ply(path):
    """Header + every fixed-size element of a binary little-endian PLY: {element: structured array}."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = (tok[1], int(tok[2]), [])
                elements.append(cur)
            elif tok[0] == "property" and cur is not None:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported (element {cur[0]})")
                cur[2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian":
            raise ValueError(f"{path}: need a binary_little_endian PLY")
        out = {}
        for name, count, props in elements:
            data = np.fromfile(f, dtype=np.dtype(props), count=count)
            if data.shape[0] != count:
                raise ValueError(f"{path}: expected {count} {name} records, file holds {data.shape[0]}")
            out[name] = data
    return out


def load_ply(path) -> GaussianCloud:
    """Standard 3DGS PLY, or — recognised by its ``chunk`` element — the PlayCanvas compressed PLY
    that InteriorGS ships (``3dgs_compressed.ply``, README.md:212-243)."""
    elems = _read_ply(path)
    if "chunk" in elems:
        return _decode_compressed(path, elems)
    if "vertex" not in elems:
        raise ValueError(f"{path}: need a binary_little_endian PLY with a vertex element")
    data = elems["vertex"]
    n = data.shape[0]
    names = set(data.dtype.names)
    need = {"x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
            "rot_0", "rot_1", "rot_2", "rot_3"}
    if not need <= names:
        raise ValueError(f"{path}: missing 3DGS properties {sorted(need - names)}")
    n_rest = len([k for k in names if k.startswith("f_rest_")])
    nb = n_rest // 3 + 1
    deg = int(round(np.sqrt(nb))) - 1
    if (deg + 1) ** 2 != nb or deg > 3:
        raise ValueError(f"{path}: {n_rest} f_rest properties do not form SH degree 0..3")
    f32 = lambda cols: np.ascontiguousarray(np.stack([data[c].astype(np.float32) for c in cols], 1))
    sh = np.zeros((n, nb, 3), np.float32)
    sh[:, 0, :] = f32(["f_dc_0", "f_dc_1", "f_dc_2"])
    if n_rest:
        rest = f32([f"f_rest_{i}" for i in range(n_rest)]).reshape(n, 3, nb - 1)   # channel-major
        sh[:, 1:, :] = rest.transpose(0, 2, 1)
    return GaussianCloud(f32(["x", "y", "z"]), f32(["scale_0", "scale_1", "scale_2"]),
                         f32(["rot_0", "rot_1", "rot_2", "rot_3"]),
                         np.ascontiguousarray(data["opacity"].astype(np.float32)), sh)


# ---- PlayCanvas compressed PLY (InteriorGS ``3dgs_compressed.ply``) -------------------------------
# The reference's README has the user run playcanvas/splat-transform (not vendored, a node tool) to
# expand this file, then 3DGRUT to wrap the result into a USDZ (README.md:224-266).  The format, as
# published with the PlayCanvas engine / splat-transform: splats in chunks of 256;
#   element chunk: min/max of position (6 floats), of log-scale (6), and - newer files - of colour (6);
#   element vertex: four uint32 per splat:
#     packed_position  11-10-11 bits (x: 31..21, y: 20..11, z: 10..0), unorm, lerp(chunk min, max)
#     packed_rotation  2-10-10-10: index of the largest of (rot_0..rot_3) = (w, x, y, z) in bits 31..30,
#                      the other three in index order, each (v * sqrt(1/2) + 0.5) as unorm10; the largest
#                      is rebuilt as sqrt(1 - sum of squares) (sign folded so that it is positive)
#     packed_scale     11-10-11 like the position, lerp(chunk min_scale, max_scale) (log scales)
#     packed_color     8-8-8-8 (r: 31..24, g, b, a): rgb = lerp(chunk min, max) (or the byte / 255 in
#                      old files), f_dc = (rgb - 0.5) / C0; a = sigmoid(opacity)  ->  logit on load
#   element sh (optional): uchar f_rest_0..(3k-1), k = 3 / 8 / 15, channel-major like the plain PLY;
#                      value = (c - 0.5) * 8 with c = 0 | 1 | (byte + 0.5) / 256 for byte 0 | 255 | other
SH_C0 = 0.28209479177387814
_CHUNK = 256


def _unorm(v, bits):
    return (v & np.uint32((1 << bits) - 1)).astype(np.float32) / np.float32((1 << bits) - 1)


def _lerp(lo, hi, t):
    return lo * (np.float32(1.0) - t) + hi * t


def _decode_compressed(path, elems) -> GaussianCloud:
    chunk, vert = elems["chunk"], elems.get("vertex")
    need_v = {"packed_position", "packed_rotation", "packed_scale", "packed_color"}
    if vert is None or not need_v <= set(vert.dtype.names):
        raise ValueError(f"{path}: compressed PLY without packed vertex properties")
    n = vert.shape[0]
    if chunk.shape[0] * _CHUNK < n:
        raise ValueError(f"{path}: {chunk.shape[0]} chunks cannot hold {n} splats")
    ci = np.arange(n) // _CHUNK
    cf = lambda k: chunk[k].astype(np.float32)[ci]

    def unpack_111011(word, lo_names, hi_names):
        t = np.stack([_unorm(word >> np.uint32(21), 11), _unorm(word >> np.uint32(11), 10), _unorm(word, 11)], 1)
        lo = np.stack([cf(k) for k in lo_names], 1)
        hi = np.stack([cf(k) for k in hi_names], 1)
        return np.ascontiguousarray(_lerp(lo, hi, t), dtype=np.float32)

    xyz = unpack_111011(vert["packed_position"], ("min_x", "min_y", "min_z"), ("max_x", "max_y", "max_z"))
    log_scale = unpack_111011(vert["packed_scale"], ("min_scale_x", "min_scale_y", "min_scale_z"),
                              ("max_scale_x", "max_scale_y", "max_scale_z"))
    # rotation: smallest-three
    w = vert["packed_rotation"]
    norm = np.float32(np.sqrt(2.0))                 # 1 / (sqrt(2) / 2)
    abc = np.stack([_unorm(w >> np.uint32(20), 10), _unorm(w >> np.uint32(10), 10), _unorm(w, 10)], 1)
    abc = (abc - np.float32(0.5)) * norm
    m = np.sqrt(np.maximum(np.float32(0.0), np.float32(1.0) - (abc * abc).sum(1, dtype=np.float32)))
    largest = (w >> np.uint32(30)).astype(np.int64)
    quat = np.empty((n, 4), np.float32)              # (rot_0..rot_3) = (w, x, y, z)
    rows = np.arange(n)
    others = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])[largest]
    quat[rows, largest] = m
    for k in range(3):
        quat[rows, others[:, k]] = abc[:, k]
    # colour + opacity
    c = vert["packed_color"]
    rgba = np.stack([_unorm(c >> np.uint32(24), 8), _unorm(c >> np.uint32(16), 8), _unorm(c >> np.uint32(8), 8),
                     _unorm(c, 8)], 1)
    rgb = rgba[:, :3]
    if "min_r" in chunk.dtype.names:
        lo = np.stack([cf(k) for k in ("min_r", "min_g", "min_b")], 1)
        hi = np.stack([cf(k) for k in ("max_r", "max_g", "max_b")], 1)
        rgb = _lerp(lo, hi, rgb)
    f_dc = (rgb - np.float32(0.5)) / np.float32(SH_C0)
    a = rgba[:, 3].astype(np.float64)
    with np.errstate(divide="ignore"):
        opacity = -np.log(1.0 / a - 1.0)             # a = 0 -> -inf, a = 1 -> +inf, as splat-transform
    # the engine drops non-finite attributes at load; keep them renderable instead:
    opacity = np.clip(opacity, -30.0, 30.0).astype(np.float32)
    sh_el = elems.get("sh")
    n_rest = 0 if sh_el is None else len(sh_el.dtype.names)
    nb = n_rest // 3 + 1
    deg = int(round(np.sqrt(nb))) - 1
    if (deg + 1) ** 2 != nb or deg > 3 or (sh_el is not None and sh_el.shape[0] != n):
        raise ValueError(f"{path}: {n_rest} compressed f_rest properties do not form SH degree 0..3")
    sh = np.zeros((n, nb, 3), np.float32)
    sh[:, 0, :] = f_dc
    if n_rest:
        raw = np.stack([sh_el[f"f_rest_{i}"] for i in range(n_rest)], 1)
        cc = (raw.astype(np.float32) + np.float32(0.5)) / np.float32(256.0)
        cc = np.where(raw == 0, np.float32(0.0), np.where(raw == 255, np.float32(1.0), cc))
        rest = ((cc - np.float32(0.5)) * np.float32(8.0)).reshape(n, 3, nb - 1)
        sh[:, 1:, :] = rest.transpose(0, 2, 1)
    return GaussianCloud(xyz, log_scale, quat, np.ascontiguousarray(opacity), sh)


def save_compressed_ply(path, cloud: GaussianCloud) -> None:
    """Writer of the same format (the packing rules of splat-transform's compressed-PLY writer): used
    to build test fixtures and to ship a scene at ~1/4 of the plain PLY's size (16 B + SH bytes per splat)."""
    n, nb = cloud.n, cloud.sh.shape[1]
    nchunk = (n + _CHUNK - 1) // _CHUNK
    pad = nchunk * _CHUNK - n
    ext = lambda a: np.concatenate([a, np.repeat(a[-1:], pad, 0)]) if pad else a
    pos = ext(cloud.xyz.astype(np.float32)).reshape(nchunk, _CHUNK, 3)
    scl = np.clip(ext(cloud.log_scale.astype(np.float32)), -20.0, 20.0).reshape(nchunk, _CHUNK, 3)
    rgb = (ext(cloud.sh[:, 0, :].astype(np.float32)) * np.float32(SH_C0) + np.float32(0.5)).reshape(nchunk, _CHUNK, 3)
    chunk = np.zeros(nchunk, dtype=[(k, "<f4") for k in (
        "min_x", "min_y", "min_z", "max_x", "max_y", "max_z",
        "min_scale_x", "min_scale_y", "min_scale_z", "max_scale_x", "max_scale_y", "max_scale_z",
        "min_r", "min_g", "min_b", "max_r", "max_g", "max_b")])

    def pack_unorm(v, bits):
        return np.clip(np.floor(v * ((1 << bits) - 1) + 0.5), 0, (1 << bits) - 1).astype(np.uint32)

    def quantise(block, lo_names, hi_names):
        lo, hi = block.min(1), block.max(1)
        for i, (a, b) in enumerate(zip(lo_names, hi_names)):
            chunk[a], chunk[b] = lo[:, i], hi[:, i]
        span = np.where(hi - lo > 0, hi - lo, 1.0).astype(np.float32)
        return ((block - lo[:, None, :]) / span[:, None, :]).reshape(-1, 3)[:n]

    tp = quantise(pos, ("min_x", "min_y", "min_z"), ("max_x", "max_y", "max_z"))
    ts = quantise(scl, ("min_scale_x", "min_scale_y", "min_scale_z"), ("max_scale_x", "max_scale_y", "max_scale_z"))
    tc = quantise(rgb, ("min_r", "min_g", "min_b"), ("max_r", "max_g", "max_b"))
    pack_111011 = lambda t: (pack_unorm(t[:, 0], 11) << np.uint32(21)) | (pack_unorm(t[:, 1], 10) << np.uint32(11)) | pack_unorm(t[:, 2], 11)
    q = cloud.quat.astype(np.float64)
    q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-30)
    largest = np.abs(q).argmax(1)
    rows = np.arange(n)
    q = np.where(q[rows, largest][:, None] < 0, -q, q)
    others = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])[largest]
    rot = largest.astype(np.uint32)
    for k in range(3):
        rot = (rot << np.uint32(10)) | pack_unorm(q[rows, others[:, k]] * (np.sqrt(2.0) * 0.5) + 0.5, 10)
    alpha = 1.0 / (1.0 + np.exp(-cloud.opacity.astype(np.float64)))
    col = (pack_unorm(tc[:, 0], 8) << np.uint32(24)) | (pack_unorm(tc[:, 1], 8) << np.uint32(16)) | \
          (pack_unorm(tc[:, 2], 8) << np.uint32(8)) | pack_unorm(alpha, 8)
    vert = np.zeros(n, dtype=[("packed_position", "<u4"), ("packed_rotation", "<u4"), ("packed_scale", "<u4"), ("packed_color", "<u4")])
    vert["packed_position"], vert["packed_rotation"], vert["packed_scale"], vert["packed_color"] = pack_111011(tp), rot, pack_111011(ts), col
    n_rest = 3 * (nb - 1)
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element chunk {nchunk}\n".encode())
        for k in chunk.dtype.names:
            f.write(f"property float {k}\n".encode())
        f.write(f"element vertex {n}\n".encode())
        for k in vert.dtype.names:
            f.write(f"property uint {k}\n".encode())
        if n_rest:
            f.write(f"element sh {n}\n".encode())
            for i in range(n_rest):
                f.write(f"property uchar f_rest_{i}\n".encode())
        f.write(b"end_header\n")
        chunk.tofile(f)
        vert.tofile(f)
        if n_rest:
            rest = cloud.sh[:, 1:, :].astype(np.float32).transpose(0, 2, 1).reshape(n, n_rest)     # channel-major
            np.clip(np.floor((rest / 8.0 + 0.5) * 256.0), 0, 255).astype(np.uint8).tofile(f)


def save_ply(path, cloud: GaussianCloud) -> None:
    n, nb = cloud.n, cloud.sh.shape[1]
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + \
            [f"f_rest_{i}" for i in range(3 * (nb - 1))] + ["opacity", "scale_0", "scale_1", "scale_2",
                                                            "rot_0", "rot_1", "rot_2", "rot_3"]
    arr = np.zeros(n, dtype=[(k, "<f4") for k in names])
    for i, k in enumerate("xyz"):
        arr[k] = cloud.xyz[:, i]
    for c in range(3):
        arr[f"f_dc_{c}"] = cloud.sh[:, 0, c]
    rest = cloud.sh[:, 1:, :].transpose(0, 2, 1).reshape(n, -1)
    for i in range(rest.shape[1]):
        arr[f"f_rest_{i}"] = rest[:, i]
    arr["opacity"] = cloud.opacity
    for i in range(3):
        arr[f"scale_{i}"] = cloud.log_scale[:, i]
    for i in range(4):
        arr[f"rot_{i}"] = cloud.quat[:, i]
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {n}\n".encode())
        for k in names:
            f.write(f"property float {k}\n".encode())
        f.write(b"end_header\n")
        arr.tofile(f)


def save_npz(path, cloud: GaussianCloud, **extra) -> None:
    np.savez(path, xyz=cloud.xyz, log_scale=cloud.log_scale, quat=cloud.quat, opacity=cloud.opacity, sh=cloud.sh, **extra)


def load_npz(path) -> GaussianCloud:
    z = np.load(path)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return GaussianCloud(f32(z["xyz"]), f32(z["log_scale"]), f32(z["quat"]), f32(z["opacity"]), f32(z["sh"]))


def load_cloud(path) -> GaussianCloud:
    p = str(path).lower()
    if p.endswith(".ply"):
        return load_ply(path)
    if p.endswith(".npz"):
        return load_npz(path)
    raise ValueError(f"unsupported Gaussian file {path}")


def morton_order(xyz: np.ndarray, bits: int = 10) -> np.ndarray:
    """Permutation sorting points along a 3-D Morton curve (scene-ingest locality: neighbouring
    indices become spatial neighbours, so culling and record gathers stay coherent)."""
    lo, hi = xyz.min(0), xyz.max(0)
    q = ((xyz - lo) / np.maximum(hi - lo, 1e-9) * ((1 << bits) - 1)).astype(np.uint64)
    code = np.zeros(xyz.shape[0], np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(code, kind="stable")


def resolve_stage_assets(scene_usd_path: str):
    """Return (gaussian_file | None, collision_file | None, is_usd_stage) for a scene path.

    A ``.ply``/``.npz`` path is taken as is.  For a USDA stage the payload references are read from
    the text (``@...usdz@[</...>]`` and ``@..._collision.usd@``, template.usda:116,157) and a
    same-stem ``.ply``/``.npz`` is searched next to the payload and next to the stage."""
    p = Path(scene_usd_path)
    low = p.name.lower()
    if low.endswith((".ply", ".npz")):
        col = p.with_name(p.stem + "_collision.npz")
        return (str(p) if p.exists() else None), (str(col) if col.exists() else None), False
    gauss = coll = None
    stems, dirs = [p.stem], [p.parent]
    if p.exists() and low.endswith((".usda", ".usd")):
        try:
            text = p.read_text(errors="replace")
            for m in re.finditer(r"@([^@]+?\.usdz)@", text):
                q = Path(m.group(1))
                stems.append(q.stem)
                dirs.append(q.parent if q.is_absolute() else (p.parent / q.parent))
            for m in re.finditer(r"@([^@]+?_collision\.usd[a-z]?)@", text):
                q = Path(m.group(1))
                dirs.append(q.parent if q.is_absolute() else (p.parent / q.parent))
        except OSError:
            pass
    for d in dirs:
        for s in stems:
            # plain / compressed PLY or NPZ named after the scene, or the InteriorGS layout {id}/3dgs_compressed.ply
            for c in (Path(d) / f"{s}.ply", Path(d) / f"{s}.npz", Path(d) / f"{s}.compressed.ply",
                      Path(d) / s / "3dgs_compressed.ply", Path(d) / f"{s}_3dgs_compressed.ply"):
                if gauss is None and c.exists():
                    gauss = str(c)
            # collision bodies: ready-made ellipsoids, else hull meshes the loader fits on the fly
            for c in (Path(d) / f"{s}_collision.npz", Path(d) / f"{s}_collision.usda", Path(d) / f"{s}_collision.obj"):
                if coll is None and c.exists():
                    coll = str(c)
    return gauss, coll, True


def load_collision_bodies(path):
    """(mu, log_scale, quat) of a collision payload: ``.npz`` of ellipsoids as is; hull meshes
    (``.usda`` / ``.obj``, template.usda:156-165) through :mod:`collision_mesh` (stage rotation folded)."""
    if str(path).lower().endswith(".npz"):
        z = np.load(path)
        return z["mu"], z["log_scale"], z["quat"]
    from .collision_mesh import convert
    return convert(path)
