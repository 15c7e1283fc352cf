"""Collision meshes -> collision ellipsoids (the second half of scene ingest).

The reference's stage pulls ``{id}_collision.usd`` in as a payload under ``/World/scene_collision``
(Data/template.usda:156-165; sage3d_usda_builder.py:93-149): convex-decomposed hulls, one Mesh prim per
hull, rotated 180 degrees about Z by the prim's ``xformOp:orient = (6.1e-17, 0, 0, 1)`` (:161).  The
capsule test of this engine works on ellipsoids ``(mu, R, semi-axes)``; this module
turns each hull into one ENCLOSING ellipsoid, so a capsule that misses the ellipsoid misses the hull.

Inputs: text USD (``.usda``: Mesh prims with ``points`` / ``faceVertexCounts`` / ``faceVertexIndices``,
the xformOps of every ancestor applied) and Wavefront OBJ (one body per ``o``/``g`` group — what V-HACD /
CoACD write).  Binary ``.usd``/``.usdc`` crate files need ``pxr`` (used when importable); without it
the file must be converted once with ``usdcat in.usd -o out.usda``.

    python -m sage3d_official_b200.collision_mesh scene_collision.usda scene_collision.npz [--mvee]

Output ``.npz``: mu [M,3] f32, log_scale [M,3] f32 (log of the semi-axes), quat [M,4] f32 (w,x,y,z) —
what ``SimpleVLNEnv`` loads next to the Gaussian payload (scene_io.resolve_stage_assets).
"""
from __future__ import annotations

import re
import sys
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

# the collision prim's own transform in the stage (template.usda:161): 180 degrees about +Z
STAGE_COLLISION_ROTATION = np.diag([-1.0, -1.0, 1.0])
MIN_SEMI_AXIS = 1e-3          # metres: flat hulls (a wall panel) still get a body of 2 mm thickness


# ---- geometry -----------------------------------------------------------------------------------
def _quat_wxyz(Rm: np.ndarray) -> np.ndarray:
    """Rotation matrix (det +1) -> unit quaternion (w, x, y, z), w >= 0."""
    m = Rm
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return -q if q[0] < 0 else q


def fit_ellipsoid_pca(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Enclosing ellipsoid aligned with the principal axes of the vertices.

    Axes = eigenvectors of the vertex covariance; centre = midpoint of the extent along each axis;
    semi-axes = half extents e scaled by k = max_i sqrt(sum((p_i / e)^2)) <= sqrt(3), the smallest factor
    for which every vertex lies inside.  Returns (mu[3], semi_axes[3], R[3,3]) with R's columns the axes."""
    p = np.asarray(points, np.float64).reshape(-1, 3)
    c0 = p.mean(0)
    cov = (p - c0).T @ (p - c0) / max(len(p), 1)
    _, vec = np.linalg.eigh(cov)
    if np.linalg.det(vec) < 0:
        vec[:, 0] = -vec[:, 0]
    loc = (p - c0) @ vec
    lo, hi = loc.min(0), loc.max(0)
    ctr = 0.5 * (lo + hi)
    e = np.maximum(0.5 * (hi - lo), MIN_SEMI_AXIS)
    k = float(np.sqrt((((loc - ctr) / e) ** 2).sum(1)).max()) if len(p) else 1.0
    k = max(k, 1.0)
    return c0 + vec @ ctr, e * k, vec


def fit_ellipsoid_mvee(points: np.ndarray, tol: float = 1e-4, max_iter: int = 2000):
    """Minimum-volume enclosing ellipsoid (Khachiyan's barycentric coordinate ascent), then inflated by
    the residual so that every vertex is inside.  Tighter than the PCA fit for elongated, skewed hulls."""
    p = np.unique(np.asarray(points, np.float64).reshape(-1, 3), axis=0)
    n = len(p)
    if n < 4 or np.linalg.matrix_rank(p - p.mean(0), tol=1e-9) < 3:
        return fit_ellipsoid_pca(points)
    Q = np.vstack([p.T, np.ones(n)])
    u = np.full(n, 1.0 / n)
    for _ in range(max_iter):
        X = Q @ (u[:, None] * Q.T)
        M = np.einsum("ij,ji->i", Q.T, np.linalg.solve(X, Q))
        j = int(M.argmax())
        step = (M[j] - 4.0) / (4.0 * (M[j] - 1.0))
        nu = (1.0 - step) * u
        nu[j] += step
        done = np.linalg.norm(nu - u) < tol
        u = nu
        if done:
            break
    c = p.T @ u
    A = np.linalg.inv(p.T @ (u[:, None] * p) - np.outer(c, c)) / 3.0
    w, vec = np.linalg.eigh(A)
    if np.linalg.det(vec) < 0:
        vec[:, 0] = -vec[:, 0]
    semi = 1.0 / np.sqrt(np.maximum(w, 1e-18))
    semi = np.maximum(semi, MIN_SEMI_AXIS)
    loc = (p - c) @ vec
    k = float(np.sqrt(((loc / semi) ** 2).sum(1)).max())
    return c, semi * max(k, 1.0), vec


def split_components(n_points: int, counts: np.ndarray, indices: np.ndarray) -> List[np.ndarray]:
    """Vertex index sets of the connected components of a face list (union-find over shared vertices)."""
    parent = np.arange(n_points)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    o = 0
    for c in counts:
        f = indices[o:o + c]
        o += c
        r = find(int(f[0]))
        for v in f[1:]:
            rv = find(int(v))
            if rv != r:
                parent[rv] = r
    used = np.unique(indices)
    roots = np.array([find(int(v)) for v in used])
    return [used[roots == r] for r in np.unique(roots)]


# ---- USDA text --------------------------------------------------------------------------------------
_NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|nan)"


def _numbers(text: str) -> np.ndarray:
    return np.array(re.findall(_NUM, text), dtype=np.float64)


def _quat_to_matrix(w, x, y, z) -> np.ndarray:
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _euler_xyz(rx, ry, rz) -> np.ndarray:
    """USD rotateXYZ: R = Rz @ Ry @ Rx on column vectors (degrees)."""
    a, b, c = np.radians([rx, ry, rz])
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _local_matrix(attrs: Dict[str, str]) -> np.ndarray:
    """4x4 local-to-parent matrix (column-vector convention) from a prim's xformOps."""
    order = re.findall(r'"(xformOp:[^"]+)"', attrs.get("xformOpOrder", ""))
    M = np.eye(4)
    for op in order:                      # USD: first op in the list is the outermost
        val = attrs.get(op)
        if val is None:
            continue
        v = _numbers(val)
        L = np.eye(4)
        kind = op.split(":")[1]
        if kind == "translate":
            L[:3, 3] = v[:3]
        elif kind == "scale":
            L[:3, :3] = np.diag(v[:3])
        elif kind == "orient":
            L[:3, :3] = _quat_to_matrix(*v[:4])
        elif kind == "rotateXYZ":
            L[:3, :3] = _euler_xyz(*v[:3])
        elif kind in ("rotateX", "rotateY", "rotateZ"):
            e = [0.0, 0.0, 0.0]
            e["XYZ".index(kind[-1])] = v[0]
            L[:3, :3] = _euler_xyz(*e)
        elif kind == "transform":
            L = v[:16].reshape(4, 4).T    # USD stores row-vector matrices
        else:
            raise ValueError(f"unsupported xformOp {op}")
        M = M @ L
    return M


def _parse_prims(text: str):
    """Yield (type, name, attrs, world_matrix) for every prim of a .usda text, depth first."""
    # strip comments and the leading layer metadata
    text = re.sub(r"#[^\n]*", "", text)
    pos = 0
    head = re.compile(r'\b(def|over|class)\s+(?:(\w+)\s+)?"([^"]+)"\s*(\([^{}]*?\))?\s*\{', re.S)

    def body_end(start):
        depth, i, n = 1, start, len(text)
        in_str = False
        while i < n and depth:
            ch = text[i]
            if ch == '"':
                in_str = not in_str
            elif not in_str:
                if ch == "{":
                    depth += 1
                elif ch == "}":
                    depth -= 1
            i += 1
        return i

    def walk(lo, hi, parent_M):
        i = lo
        while True:
            m = head.search(text, i, hi)
            if not m:
                return
            end = body_end(m.end())
            body = text[m.end():end - 1]
            # attributes of THIS prim: the body with child prims cut out
            own, j = [], 0
            for c in head.finditer(body):
                if c.start() < j:
                    continue
                own.append(body[j:c.start()])
                j = body_end_local(body, c.end())
            own.append(body[j:])
            attrs = _attributes("".join(own))
            M = parent_M @ _local_matrix(attrs)
            yield m.group(2) or "", m.group(3), attrs, M
            yield from walk(m.end(), end - 1, M)
            i = end

    def body_end_local(s, start):
        depth, i = 1, start
        in_str = False
        while i < len(s) and depth:
            ch = s[i]
            if ch == '"':
                in_str = not in_str
            elif not in_str:
                if ch == "{":
                    depth += 1
                elif ch == "}":
                    depth -= 1
            i += 1
        return i

    yield from walk(pos, len(text), np.eye(4))


def _attributes(body: str) -> Dict[str, str]:
    """'type name = value' statements of a prim body -> {name: value text}; values may span lines."""
    out = {}
    for m in re.finditer(r"(?:custom\s+|uniform\s+|varying\s+)*[\w:\[\]]+\s+([\w:]+)\s*=\s*", body):
        name = m.group(1)
        i = m.end()
        if i < len(body) and body[i] in "[(":
            close = {"[": "]", "(": ")"}[body[i]]
            depth, j = 0, i
            while j < len(body):
                if body[j] == body[i]:
                    depth += 1
                elif body[j] == close:
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            out[name] = body[i:j + 1]
        else:
            out[name] = body[i:body.find("\n", i) if body.find("\n", i) >= 0 else len(body)]
    return out


def load_usda_meshes(path) -> List[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """[(points_world[n,3], faceVertexCounts, faceVertexIndices)] for every Mesh prim of a text USD file."""
    with open(path, "rb") as f:
        head = f.read(8)
    if head.startswith(b"PXR-USDC"):
        return _load_usd_pxr(path)
    text = open(path, "r", errors="replace").read()
    meshes = []
    for typ, _name, attrs, M in _parse_prims(text):
        if typ != "Mesh" or "points" not in attrs:
            continue
        pts = _numbers(attrs["points"]).reshape(-1, 3)
        pts = pts @ M[:3, :3].T + M[:3, 3]
        idx = _numbers(attrs.get("faceVertexIndices", "")).astype(np.int64)
        cnt = _numbers(attrs.get("faceVertexCounts", "")).astype(np.int64)
        meshes.append((pts, cnt, idx))
    return meshes


def _load_usd_pxr(path):
    try:
        from pxr import Usd, UsdGeom       # noqa: WPS433 (optional dependency)
    except ImportError as e:
        raise RuntimeError(f"{path} is a binary USD crate file and 'pxr' is not installed: convert it once with "
                           f"'usdcat {path} -o {str(path).rsplit('.', 1)[0]}.usda'") from e
    stage = Usd.Stage.Open(str(path))
    cache = UsdGeom.XformCache()
    meshes = []
    for prim in stage.Traverse():
        if prim.IsA(UsdGeom.Mesh):
            g = UsdGeom.Mesh(prim)
            pts = np.array(g.GetPointsAttr().Get(), np.float64).reshape(-1, 3)
            M = np.array(cache.GetLocalToWorldTransform(prim), np.float64).T
            pts = pts @ M[:3, :3].T + M[:3, 3]
            meshes.append((pts, np.array(g.GetFaceVertexCountsAttr().Get(), np.int64),
                           np.array(g.GetFaceVertexIndicesAttr().Get(), np.int64)))
    return meshes


def load_obj_meshes(path):
    """One (points, counts, indices) per ``o``/``g`` group of a Wavefront OBJ (global 1-based vertex indices)."""
    verts, groups, cur = [], [], None
    for line in open(path, "r", errors="replace"):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            verts.append([float(t[1]), float(t[2]), float(t[3])])
        elif t[0] in ("o", "g"):
            cur = []
            groups.append(cur)
        elif t[0] == "f":
            if cur is None:
                cur = []
                groups.append(cur)
            cur.append([int(x.split("/")[0]) for x in t[1:]])
    V = np.array(verts, np.float64).reshape(-1, 3)
    out = []
    for faces in groups:
        if not faces:
            continue
        flat = np.array([(i - 1) if i > 0 else (len(V) + i) for f in faces for i in f], np.int64)
        used, inv = np.unique(flat, return_inverse=True)
        out.append((V[used], np.array([len(f) for f in faces], np.int64), inv.astype(np.int64)))
    return out


# ---- the converter ----------------------------------------------------------------------------------
def meshes_to_ellipsoids(meshes: Iterable, method: str = "pca", stage_rotation: bool = True,
                         split: bool = True):
    """(mu[M,3], log_scale[M,3], quat[M,4]) float32: one enclosing ellipsoid per convex body."""
    fit = fit_ellipsoid_mvee if method == "mvee" else fit_ellipsoid_pca
    mus, lss, qs = [], [], []
    for pts, cnt, idx in meshes:
        parts = split_components(len(pts), cnt, idx) if (split and len(idx)) else [np.arange(len(pts))]
        for sel in parts:
            p = pts[sel]
            if len(p) == 0:
                continue
            if stage_rotation:
                p = p @ STAGE_COLLISION_ROTATION.T
            mu, semi, Rm = fit(p)
            mus.append(mu)
            lss.append(np.log(semi))
            qs.append(_quat_wxyz(Rm))
    if not mus:
        z = np.zeros((0, 3), np.float32)
        return z, z.copy(), np.zeros((0, 4), np.float32)
    f32 = lambda a: np.ascontiguousarray(np.array(a), dtype=np.float32)
    return f32(mus), f32(lss), f32(qs)


def convert(src, dst: Optional[str] = None, method: str = "pca", stage_rotation: bool = True):
    """Collision file (.usda / .usd / .obj) -> ellipsoids; written to ``dst`` (.npz) when given."""
    low = str(src).lower()
    meshes = load_obj_meshes(src) if low.endswith(".obj") else load_usda_meshes(src)
    mu, ls, q = meshes_to_ellipsoids(meshes, method=method, stage_rotation=stage_rotation)
    if dst is not None:
        np.savez(dst, mu=mu, log_scale=ls, quat=q)
    return mu, ls, q


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--mvee", action="store_true", help="minimum-volume enclosing ellipsoids (slower, tighter)")
    ap.add_argument("--no-stage-rotation", action="store_true",
                    help="keep mesh coordinates (default: apply the stage's 180-degree Z rotation, template.usda:161)")
    a = ap.parse_args(argv)
    mu, _, _ = convert(a.src, a.dst, "mvee" if a.mvee else "pca", not a.no_stage_rotation)
    print(f"{a.src}: {mu.shape[0]} collision bodies -> {a.dst}")


if __name__ == "__main__":
    main(sys.argv[1:])
