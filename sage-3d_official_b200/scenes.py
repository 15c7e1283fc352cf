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


def _recycle_large_allocations():
    """Serve large numpy temporaries from the heap (and keep it) instead of fresh mmap regions:
    in the sandboxed containers this runs in, first-touch page faults cost ~30 ms/MB."""
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 1 << 30)                    # M_MMAP_THRESHOLD
        libc.mallopt(-1, ctypes.c_int(2 ** 31 - 1))  # M_TRIM_THRESHOLD
    except Exception:
        pass


_recycle_large_allocations()


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
    """Batch rotation matrices [N,3,3] -> quaternions [N,4] (w,x,y,z); vectorised, picks the
    numerically largest component as pivot."""
    m = Rm
    m00, m11, m22 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    cand = np.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], 1)
    k = np.argmax(cand, axis=1)
    s = 2.0 * np.sqrt(np.maximum(cand[np.arange(m.shape[0]), k], 1e-12))
    a = (m[:, 2, 1] - m[:, 1, 2]) / s
    b = (m[:, 0, 2] - m[:, 2, 0]) / s
    c = (m[:, 1, 0] - m[:, 0, 1]) / s
    d = (m[:, 0, 1] + m[:, 1, 0]) / s
    e = (m[:, 0, 2] + m[:, 2, 0]) / s
    f = (m[:, 1, 2] + m[:, 2, 1]) / s
    q4 = 0.25 * s
    q = np.where((k == 0)[:, None], np.stack([q4, a, b, c], 1),
        np.where((k == 1)[:, None], np.stack([a, q4, d, e], 1),
        np.where((k == 2)[:, None], np.stack([b, d, q4, f], 1), np.stack([c, e, f, q4], 1))))
    return q


def _sample_groups(groups, n: int, rng, sh_degree: int, chunk: int = 262_144) -> GaussianCloud:
    """Sample n Gaussians over several patch groups.

    groups: list of (patches, fraction_of_n, size_scale).  Patches are drawn from one combined
    distribution, so the output arrays have no spatial order.  Work proceeds in chunks written into
    preallocated float32 outputs (first-touch page faults are expensive in the sandboxes this
    runs in; temporaries stay chunk-sized and are recycled by the allocator).
    """
    P_o, P_u, P_v, P_n, P_col, P_w, P_sp, P_ss = [], [], [], [], [], [], [], []
    for patches, frac, size_scale in groups:
        areas = np.array([p.area for p in patches])
        g_n = max(frac * n, 1.0)
        spacing = np.sqrt(areas.sum() / g_n)
        for p, ar in zip(patches, areas):
            P_o.append(p.origin); P_u.append(p.u); P_v.append(p.v); P_n.append(p.normal)
            P_col.append(p.color); P_w.append(frac * ar / areas.sum()); P_sp.append(spacing); P_ss.append(size_scale)
    P_o, P_u, P_v, P_n, P_col = (np.stack(x) for x in (P_o, P_u, P_v, P_n, P_col))
    P_w = np.asarray(P_w); P_w = P_w / P_w.sum()
    P_sp, P_ss = np.asarray(P_sp), np.asarray(P_ss)
    P_uh = P_u / np.linalg.norm(P_u, axis=1, keepdims=True)
    P_vh = np.cross(P_n, P_uh)
    nb = (sh_degree + 1) ** 2
    xyz = np.empty((n, 3), np.float32)
    log_scale = np.empty((n, 3), np.float32)
    quat = np.empty((n, 4), np.float32)
    opacity = np.empty(n, np.float32)
    sh = np.empty((n, nb, 3), np.float32)
    for s0 in range(0, n, chunk):
        m = min(chunk, n - s0)
        which = rng.choice(len(P_w), size=m, p=P_w)
        spacing = P_sp[which]
        a, b = rng.uniform(size=m), rng.uniform(size=m)
        thickness = 0.09 * spacing                 # 5 mm at the config-1 density (5.5 cm spacing)
        Nn = P_n[which]
        xyz[s0:s0 + m] = (P_o[which] + a[:, None] * P_u[which] + b[:, None] * P_v[which]
                          + Nn * (rng.standard_normal(m) * thickness)[:, None])
        # tangential log-scale U(ln 0.36 s, ln 1.45 s): (0.02, 0.08) m at the config-1 density
        lo = np.log(0.36 * spacing * P_ss[which])
        t1 = lo + rng.uniform(size=m) * np.log(1.45 / 0.36)
        t2 = lo + rng.uniform(size=m) * np.log(1.45 / 0.36)
        log_scale[s0:s0 + m, 0] = t1
        log_scale[s0:s0 + m, 1] = t2
        log_scale[s0:s0 + m, 2] = np.minimum(t1, t2) - np.log(5.0)
        th = rng.uniform(0.0, 2 * np.pi, size=m)
        c, sn = np.cos(th)[:, None], np.sin(th)[:, None]
        uh, vh = P_uh[which], P_vh[which]
        Rm = np.stack([uh * c + vh * sn, -uh * sn + vh * c, Nn], axis=2)   # columns = local axes
        quat[s0:s0 + m] = _matrix_to_quat_wxyz(Rm) * rng.uniform(0.5, 2.0, size=(m, 1))  # un-normalised on purpose
        opacity[s0:s0 + m] = rng.uniform(0.0, 4.0, size=m)
        sh[s0:s0 + m, 0, :] = np.clip(P_col[which] + rng.uniform(-0.35, 0.35, size=(m, 3)), -1.0, 1.0)
        if nb > 1:
            sh[s0:s0 + m, 1:, :] = rng.standard_normal(size=(m, nb - 1, 3), dtype=np.float32) * np.float32(0.05)
    return GaussianCloud(xyz, log_scale, quat, opacity, sh)


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
    cloud = _sample_groups([(shell, 0.7, size_scale), (furn, 0.3, size_scale * 0.8)], n, rng, sh_degree)
    return cloud, {"lo": lo, "hi": hi, "walls": [], "boxes": boxes}


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
    cloud = _sample_groups([(shell, 0.55, size_scale), (wall_p, 0.20, size_scale),
                            (furn, 0.25, size_scale * 0.8)], n, rng, sh_degree)
    return cloud, {"lo": lo, "hi": hi, "walls": walls, "boxes": boxes}


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


def to_trajectory_frame(xyz: np.ndarray, map_bounds) -> np.ndarray:
    """Points of a layout given in 2-D-semantic-map coordinates, moved into the frame the agent and
    the USD stage live in.  The reference's 3-D trajectories are its 2-D ones mirrored about the map
    centre and negated (trajectory_2d_to_3d.py:33-63; undone by collision_detector.py:146-189), a pure
    translation: x3 = x2 - (min_x + max_x), y3 = y2 - (min_y + max_y).  ``map_bounds`` =
    (min_x, max_x, min_y, max_y) of the obstacle grid built from the same map."""
    min_x, max_x, min_y, max_y = map_bounds
    out = np.array(xyz, dtype=np.float32, copy=True)
    out[:, 0] -= np.float32(min_x + max_x)
    out[:, 1] -= np.float32(min_y + max_y)
    return out


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
