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


def load_ply(path) -> GaussianCloud:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n, props, in_vertex = None, None, [], False
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
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian" or n is None:
            raise ValueError(f"{path}: need a binary_little_endian PLY with a vertex element")
        data = np.fromfile(f, dtype=np.dtype(props), count=n)
    if data.shape[0] != n:
        raise ValueError(f"{path}: expected {n} vertices, file holds {data.shape[0]}")
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
            for ext in (".ply", ".npz"):
                c = Path(d) / f"{s}{ext}"
                if gauss is None and c.exists():
                    gauss = str(c)
            c = Path(d) / f"{s}_collision.npz"
            if coll is None and c.exists():
                coll = str(c)
    return gauss, coll, True
