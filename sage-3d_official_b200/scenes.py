"""Seeded synthetic scenes for the render/collide path.

InteriorGS assets are not in the reference tree and there is no network, so the BASELINE configs
run on generated interiors: Gaussians hugging the faces of a room shell, interior partition walls
and furniture boxes, with splat sizes tied to the sampling density (as a trained scene's are).

Also generates the matching 2-D semantic map (the JSON format parsed by the reference's
collision detector, semantic_map_builder.py:176-197 / collision_detector.py:55-93) and a set of
collision ellipsoids for the capsule test.
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np


@dataclass
class GaussianCloud:
    """Raw 3DGS attributes, host side (float32, C-contiguous)."""
    xyz: np.ndarray          # [N,3]
    log_scale: np.ndarray    # [N,3]
    quat: np.ndarray         # [N,4] (w,x,y,z), not necessarily normalised
    opacity: np.ndarray      # [N]   raw (pre-sigmoid)
    sh: np.ndarray           # [N,(deg+1)^2,3]

    @property
    def n(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def sh_degree(self) -> int:
        return int(round(np.sqrt(self.sh.shape[1]))) - 1

    def subset(self, idx) -> "GaussianCloud":
        return GaussianCloud(*(np.ascontiguousarray(a[idx]) for a in
                               (self.xyz, self.log_scale, self.quat, self.opacity, self.sh)))

    def as_tuple(self):
        return self.xyz, self.log_scale, self.quat, self.opacity, self.sh


@dataclass
class _Patch:
    origin: np.ndarray
    u: np.ndarray
    v: np.ndarray
    color: np.ndarray

    @property
    def area(self) -> float:
        return float(np.linalg.norm(self.u) * np.linalg.norm(self.v))

    @property
    def normal(self) -> np.ndarray:
        n = np.cross(self.u, self.v)
        return n / np.linalg.norm(n)


def _box_patches(lo, hi, rng, inward: bool, skip_bottom: bool = False):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces = [
        (lo, ex, ey), (lo + ez, ex, ey),          # z-, z+
        (lo, ex, ez), (lo + ey, ex, ez),          # y-, y+
        (lo, ey, ez), (lo + ex, ey, ez),          # x-, x+
    ]
    if skip_bottom:
        faces = faces[1:]
    base = rng.uniform(-0.8, 0.8, size=3)
    out = []
    for o, u, v in faces:
        out.append(_Patch(np.array(o, float), u, v, base + rng.uniform(-0.25, 0.25, size=3)))
    return out


def _matrix_to_quat_wxyz(Rm: np.ndarray) -> np.ndarray:
    """Batch rotation matrices [N,3,3] -> quaternions [N,4] (w,x,y,z)."""
    m = Rm
    tr = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    q = np.zeros((m.shape[0], 4))
    c0 = tr > 0
    s = np.sqrt(np.maximum(tr[c0] + 1.0, 1e-12)) * 2
    q[c0] = np.stack([0.25 * s, (m[c0, 2, 1] - m[c0, 1, 2]) / s, (m[c0, 0, 2] - m[c0, 2, 0]) / s,
                      (m[c0, 1, 0] - m[c0, 0, 1]) / s], 1)
    rest = ~c0
    for i in np.nonzero(rest)[0]:
        a = m[i]
        k = int(np.argmax([a[0, 0], a[1, 1], a[2, 2]]))
        j, l = (k + 1) % 3, (k + 2) % 3
        s = np.sqrt(max(1.0 + a[k, k] - a[j, j] - a[l, l], 1e-12)) * 2
        v = np.zeros(4)
        v[0] = (a[l, j] - a[j, l]) / s
        v[1 + k] = 0.25 * s
        v[1 + j] = (a[j, k] + a[k, j]) / s
        v[1 + l] = (a[l, k] + a[k, l]) / s
        q[i] = v
    return q


def _sample_patches(patches, n: int, rng, sh_degree: int, size_scale: float) -> GaussianCloud:
    areas = np.array([p.area for p in patches])
    total = float(areas.sum())
    spacing = np.sqrt(total / max(n, 1))
    which = rng.choice(len(patches), size=n, p=areas / total)
    O = np.stack([p.origin for p in patches])[which]
    U = np.stack([p.u for p in patches])[which]
    Vv = np.stack([p.v for p in patches])[which]
    Nn = np.stack([p.normal for p in patches])[which]
    col = np.stack([p.color for p in patches])[which]
    a, b = rng.uniform(size=n), rng.uniform(size=n)
    thickness = 0.09 * spacing                     # 5 mm at the config-1 density (5.5 cm spacing)
    xyz = O + a[:, None] * U + b[:, None] * Vv + Nn * rng.normal(0.0, thickness, size=(n, 1))
    # tangential log-scale U(ln 0.36 s, ln 1.45 s): (0.02, 0.08) m at the config-1 density
    lo, hi = np.log(0.36 * spacing * size_scale), np.log(1.45 * spacing * size_scale)
    t1, t2 = rng.uniform(lo, hi, size=n), rng.uniform(lo, hi, size=n)
    log_scale = np.stack([t1, t2, np.minimum(t1, t2) - np.log(5.0)], 1)
    uh = U / np.linalg.norm(U, axis=1, keepdims=True)
    vh = np.cross(Nn, uh)
    th = rng.uniform(0.0, 2 * np.pi, size=n)
    c, s = np.cos(th)[:, None], np.sin(th)[:, None]
    Rm = np.stack([uh * c + vh * s, -uh * s + vh * c, Nn], axis=2)   # columns = local axes
    quat = _matrix_to_quat_wxyz(Rm) * rng.uniform(0.5, 2.0, size=(n, 1))  # un-normalised on purpose
    opacity = rng.uniform(0.0, 4.0, size=n)
    nb = (sh_degree + 1) ** 2
    sh = np.zeros((n, nb, 3))
    sh[:, 0, :] = np.clip(col + rng.uniform(-0.35, 0.35, size=(n, 3)), -1.0, 1.0)
    if nb > 1:
        sh[:, 1:, :] = rng.normal(0.0, 0.05, size=(n, nb - 1, 3))
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    return GaussianCloud(f32(xyz), f32(log_scale), f32(quat), f32(opacity), f32(sh))


def _furniture(rng, count: int, room_lo, room_hi, margin: float = 0.3):
    boxes = []
    lo, hi = np.asarray(room_lo, float), np.asarray(room_hi, float)
    for _ in range(count):
        size = rng.uniform(0.4, 1.5, size=3)
        size[2] = rng.uniform(0.4, 1.2)
        p = rng.uniform(lo[:2] + margin, np.maximum(hi[:2] - margin - size[:2], lo[:2] + margin + 1e-3))
        boxes.append((np.array([p[0], p[1], lo[2]]), np.array([p[0] + size[0], p[1] + size[1], lo[2] + size[2]])))
    return boxes


def synthetic_room(seed: int = 0, n: int = 50_000, sh_degree: int = 3, size=(6.0, 4.0, 3.0),
                   n_furniture: int = 8, size_scale: float = 1.0):
    """Config 1: one 6 x 4 x 3 m room, Z-up, 70 % shell / 30 % furniture.  Returns (cloud, layout)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.zeros(3), np.asarray(size, float)
    shell = _box_patches(lo, hi, rng, inward=True)
    boxes = _furniture(rng, n_furniture, lo, hi)
    furn = [p for b in boxes for p in _box_patches(b[0], b[1], rng, inward=False, skip_bottom=True)]
    n_shell = int(round(0.7 * n))
    a = _sample_patches(shell, n_shell, rng, sh_degree, size_scale)
    b = _sample_patches(furn, n - n_shell, rng, sh_degree, size_scale * 0.8)
    cloud = GaussianCloud(*(np.concatenate(p) for p in zip(a.as_tuple(), b.as_tuple())))
    perm = rng.permutation(cloud.n)           # no spatial order in the input arrays
    return cloud.subset(perm), {"lo": lo, "hi": hi, "walls": [], "boxes": boxes}


def synthetic_interior(seed: int = 1, n: int = 2_000_000, sh_degree: int = 3, size=(12.0, 10.0, 3.0),
                       n_furniture: int = 40, size_scale: float = 1.0):
    """Config 2: a 12 x 10 x 3 m multi-room floor (2 x 2 rooms joined by doorways)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.zeros(3), np.asarray(size, float)
    shell = _box_patches(lo, hi, rng, inward=True)
    th = 0.12
    mx, my = size[0] / 2.0, size[1] / 2.0
    door = 1.0
    # partition walls as thin boxes, each split around a doorway
    walls = [
        (np.array([mx - th / 2, 0.0, 0.0]), np.array([mx + th / 2, my / 2 - door / 2, size[2]])),
        (np.array([mx - th / 2, my / 2 + door / 2, 0.0]), np.array([mx + th / 2, my + my / 2 - door / 2, size[2]])),
        (np.array([mx - th / 2, my + my / 2 + door / 2, 0.0]), np.array([mx + th / 2, size[1], size[2]])),
        (np.array([0.0, my - th / 2, 0.0]), np.array([mx / 2 - door / 2, my + th / 2, size[2]])),
        (np.array([mx / 2 + door / 2, my - th / 2, 0.0]), np.array([mx + mx / 2 - door / 2, my + th / 2, size[2]])),
        (np.array([mx + mx / 2 + door / 2, my - th / 2, 0.0]), np.array([size[0], my + th / 2, size[2]])),
    ]
    wall_p = [p for w in walls for p in _box_patches(w[0], w[1], rng, inward=False, skip_bottom=True)]
    boxes = []
    per = max(1, n_furniture // 4)
    for qx in range(2):
        for qy in range(2):
            rlo = np.array([qx * mx + th, qy * my + th, 0.0])
            rhi = np.array([(qx + 1) * mx - th, (qy + 1) * my - th, size[2]])
            boxes += _furniture(rng, per, rlo, rhi)
    furn = [p for b in boxes for p in _box_patches(b[0], b[1], rng, inward=False, skip_bottom=True)]
    n_shell = int(round(0.55 * n))
    n_wall = int(round(0.20 * n))
    parts = [_sample_patches(shell, n_shell, rng, sh_degree, size_scale),
             _sample_patches(wall_p, n_wall, rng, sh_degree, size_scale),
             _sample_patches(furn, n - n_shell - n_wall, rng, sh_degree, size_scale * 0.8)]
    cloud = GaussianCloud(*(np.concatenate(p) for p in zip(*(c.as_tuple() for c in parts))))
    perm = rng.permutation(cloud.n)
    return cloud.subset(perm), {"lo": lo, "hi": hi, "walls": walls, "boxes": boxes}


def camera_path(layout, n_poses: int = 256, seed: int = 0, height: float = 1.2):
    """Seeded walk through the free space of a layout: list of (position[3], heading_rad)."""
    rng = np.random.default_rng(seed + 1000)
    lo, hi = layout["lo"], layout["hi"]
    blockers = list(layout["walls"]) + list(layout["boxes"])

    def free(p):
        if not (lo[0] + 0.4 < p[0] < hi[0] - 0.4 and lo[1] + 0.4 < p[1] < hi[1] - 0.4):
            return False
        for a, b in blockers:
            if a[0] - 0.3 < p[0] < b[0] + 0.3 and a[1] - 0.3 < p[1] < b[1] + 0.3:
                return False
        return True

    p = None
    while p is None:
        c = rng.uniform(lo[:2] + 0.5, hi[:2] - 0.5)
        if free(c):
            p = c
    heading = rng.uniform(-np.pi, np.pi)
    poses = []
    while len(poses) < n_poses:
        poses.append((np.array([p[0], p[1], height]), float(heading)))
        heading += rng.normal(0.0, 0.12)
        step = np.array([np.cos(heading), np.sin(heading)]) * 0.05
        if free(p + step):
            p = p + step
        else:
            heading += rng.uniform(0.6, 2.4) * rng.choice([-1.0, 1.0])
    return poses


def layout_semantic_map(layout, step: float = 0.05):
    """2-D semantic map instances (the JSON list the reference's detector parses).

    Walls (shell perimeter + partitions) are 'wall'; furniture footprints are 'Unable Area'
    (label casing as written by semantic_map_builder.py:266); the floor is a 'floor' instance so
    the bounds cover the room.  Coordinates are 2-decimal strings, [y, x] order
    (semantic_map_builder.py:176-181).
    """
    lo, hi = layout["lo"], layout["hi"]

    def cells(x0, x1, y0, y1):
        xs = np.arange(np.round(x0 / step), np.round(x1 / step) + 1) * step
        ys = np.arange(np.round(y0 / step), np.round(y1 / step) + 1) * step
        return [[f"{y:.2f}", f"{x:.2f}"] for y in ys for x in xs]

    inst = []

    def add(label, coords, iid):
        inst.append({"category_id": iid, "category_label": label, "instance_id": iid, "item_id": iid,
                     "mask_coords_m": coords})

    add("floor", cells(lo[0], hi[0], lo[1], hi[1]), 0)
    t = step
    perim = (cells(lo[0], hi[0], lo[1], lo[1] + t) + cells(lo[0], hi[0], hi[1] - t, hi[1]) +
             cells(lo[0], lo[0] + t, lo[1], hi[1]) + cells(hi[0] - t, hi[0], lo[1], hi[1]))
    add("wall", perim, 1)
    k = 2
    for a, b in layout["walls"]:
        add("wall", cells(a[0], b[0], a[1], b[1]), k)
        k += 1
    for a, b in layout["boxes"]:
        add("Unable Area", cells(a[0], b[0], a[1], b[1]), k)
        k += 1
    return inst


def write_semantic_map(layout, path, step: float = 0.05):
    with open(path, "w") as f:
        json.dump(layout_semantic_map(layout, step), f)
    return path


def collision_ellipsoids(layout, seed: int = 0, per_m2: float = 40.0):
    """Collision bodies as ellipsoids covering walls and furniture: (mu[M,3], log_scale[M,3], quat[M,4])."""
    rng = np.random.default_rng(seed + 77)
    lo, hi = layout["lo"], layout["hi"]
    th = 0.06
    slabs = [
        (np.array([lo[0] - th, lo[1], lo[2]]), np.array([lo[0], hi[1], hi[2]])),
        (np.array([hi[0], lo[1], lo[2]]), np.array([hi[0] + th, hi[1], hi[2]])),
        (np.array([lo[0], lo[1] - th, lo[2]]), np.array([hi[0], lo[1], hi[2]])),
        (np.array([lo[0], hi[1], lo[2]]), np.array([hi[0], hi[1] + th, hi[2]])),
    ] + list(layout["walls"]) + list(layout["boxes"])
    mus, lss, qs = [], [], []
    for a, b in slabs:
        d = b - a
        area = 2 * (d[0] * d[1] + d[1] * d[2] + d[0] * d[2])
        m = max(4, int(area * per_m2 / 6))
        mu = rng.uniform(a, b, size=(m, 3))
        r = rng.uniform(0.04, 0.12, size=(m, 3))
        ang = rng.uniform(0, 2 * np.pi, size=m)
        q = np.stack([np.cos(ang / 2), np.zeros(m), np.zeros(m), np.sin(ang / 2)], 1)
        mus.append(mu); lss.append(np.log(r)); qs.append(q)
    f32 = lambda x: np.ascontiguousarray(np.concatenate(x), dtype=np.float32)
    return f32(mus), f32(lss), f32(qs)
