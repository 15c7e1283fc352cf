"""Scene ingest: the InteriorGS compressed PLY decoder (README.md:212-243 — today two external
tools) and the collision-hull -> ellipsoid converter (template.usda:156-165).  CPU only.

The compressed-PLY format is third-party (PlayCanvas splat-transform, not vendored): the decoder is pinned by
a hand-packed file whose every field is computed below from the published packing rules, and by a
round trip through the writer within the quantisation steps."""
import struct

import numpy as np
import pytest

import sage3d_official_b200 as sage
from sage3d_official_b200 import collision_mesh as cm
from sage3d_official_b200.scene_io import (SH_C0, load_cloud, load_collision_bodies, load_ply, resolve_stage_assets,
                                           save_compressed_ply, save_ply)


def _write_compressed(path, chunk_floats, verts, sh_bytes=None, colour_range=True):
    names = ["min_x", "min_y", "min_z", "max_x", "max_y", "max_z", "min_scale_x", "min_scale_y", "min_scale_z",
             "max_scale_x", "max_scale_y", "max_scale_z"]
    if colour_range:
        names += ["min_r", "min_g", "min_b", "max_r", "max_g", "max_b"]
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element chunk {len(chunk_floats)}\n".encode())
        for k in names:
            f.write(f"property float {k}\n".encode())
        f.write(f"element vertex {len(verts)}\n".encode())
        for k in ("packed_position", "packed_rotation", "packed_scale", "packed_color"):
            f.write(f"property uint {k}\n".encode())
        if sh_bytes is not None:
            f.write(f"element sh {len(verts)}\n".encode())
            for i in range(len(sh_bytes[0])):
                f.write(f"property uchar f_rest_{i}\n".encode())
        f.write(b"end_header\n")
        for c in chunk_floats:
            assert len(c) == len(names)
            f.write(struct.pack(f"<{len(c)}f", *c))
        for v in verts:
            f.write(struct.pack("<4I", *v))
        if sh_bytes is not None:
            for row in sh_bytes:
                f.write(bytes(row))


def test_compressed_ply_known_answer(tmp_path):
    # one chunk; position box [0,2]x[-1,1]x[10,12], log-scale box [-5,-3]^3, colour range [0.25, 0.75]
    chunk = [0, -1, 10, 2, 1, 12, -5, -5, -5, -3, -3, -3, 0.25, 0.25, 0.25, 0.75, 0.75, 0.75]
    # splat 0: position t = (2047, 0, 1023) -> (2, -1, 10 + 2*1023/2047)
    pos0 = (2047 << 21) | (0 << 11) | 1023
    # rotation: largest = index 0 (w); others x, y, z = (0.5, -0.5, 0) * -> unorm10 of v*sqrt(.5)+.5
    enc = lambda v: int(np.floor((v * np.sqrt(0.5) + 0.5) * 1023 + 0.5))
    a, b, c = enc(0.5), enc(-0.5), enc(0.0)
    rot0 = (0 << 30) | (a << 20) | (b << 10) | c
    scl0 = (0 << 21) | (1023 << 11) | 2047          # -> (-5, -3, -3)
    col0 = (255 << 24) | (0 << 16) | (51 << 8) | 128  # r -> .75, g -> .25, b -> .25 + .5 * 51/255 = .35, alpha 128/255
    # splat 1: largest = index 2 (y): stored order (w, x, z)
    rot1 = (2 << 30) | (enc(0.1) << 20) | (enc(0.2) << 10) | enc(-0.3)
    v1 = (0, rot1, 0, (0 << 24) | (0 << 16) | (0 << 8) | 255)
    sh = [[0] * 9, [255] * 9]
    sh[0][0], sh[0][3], sh[0][6] = 128, 64, 1           # first coefficient of R, G, B
    p = tmp_path / "3dgs_compressed.ply"
    _write_compressed(p, [chunk], [(pos0, rot0, scl0, col0), v1], sh)
    cloud = load_ply(p)
    assert cloud.n == 2 and cloud.sh.shape == (2, 4, 3)
    np.testing.assert_allclose(cloud.xyz[0], [2.0, -1.0, 10 + 2 * 1023 / 2047], rtol=0, atol=1e-6)
    np.testing.assert_allclose(cloud.log_scale[0], [-5.0, -3.0, -3.0], rtol=0, atol=1e-6)
    dec = lambda u: (u / 1023.0 - 0.5) * np.sqrt(2.0)
    xyz = np.array([dec(a), dec(b), dec(c)])
    np.testing.assert_allclose(cloud.quat[0], [np.sqrt(1 - (xyz ** 2).sum()), *xyz], atol=1e-6)
    assert abs(cloud.quat[0, 1] - 0.5) < 1.5e-3 and abs(cloud.quat[0, 2] + 0.5) < 1.5e-3
    w1, x1, z1 = dec(enc(0.1)), dec(enc(0.2)), dec(enc(-0.3))
    np.testing.assert_allclose(cloud.quat[1], [w1, x1, np.sqrt(1 - w1 * w1 - x1 * x1 - z1 * z1), z1], atol=1e-6)
    np.testing.assert_allclose(cloud.sh[0, 0], (np.array([0.75, 0.25, 0.35]) - 0.5) / SH_C0, atol=1e-6)
    assert abs(cloud.opacity[0] - (-np.log(255 / 128 - 1))) < 1e-6
    assert cloud.opacity[1] == 30.0                    # alpha 1 -> +inf in splat-transform; kept finite here
    np.testing.assert_allclose(cloud.sh[0, 1], [((128.5 / 256) - 0.5) * 8, ((64.5 / 256) - 0.5) * 8, ((1.5 / 256) - 0.5) * 8], atol=1e-6)
    assert np.all(cloud.sh[0, 2:] == -4.0) and np.all(cloud.sh[1, 1:] == 4.0)      # byte 0 -> c = 0, byte 255 -> c = 1


def test_compressed_ply_without_colour_range_and_sh(tmp_path):
    chunk = [0, 0, 0, 1, 1, 1, -1, -1, -1, 1, 1, 1]
    col = (255 << 24) | (128 << 16) | (0 << 8) | 64
    p = tmp_path / "old.ply"
    _write_compressed(p, [chunk], [(0, 0, 0, col)], None, colour_range=False)
    cloud = load_ply(p)
    assert cloud.sh.shape == (1, 1, 3)
    np.testing.assert_allclose(cloud.sh[0, 0], (np.array([1.0, 128 / 255, 0.0]) - 0.5) / SH_C0, atol=1e-6)


@pytest.mark.parametrize("n,deg", [(1000, 3), (257, 1), (256, 0)])
def test_compressed_ply_round_trip(tmp_path, n, deg):
    cloud, _ = sage.synthetic_room(3, n, sh_degree=deg)
    p = tmp_path / "c.ply"
    save_compressed_ply(p, cloud)
    back = load_cloud(p)
    assert back.n == n and back.sh.shape == cloud.sh.shape
    ext = cloud.xyz.max(0) - cloud.xyz.min(0)
    assert np.all(np.abs(back.xyz - cloud.xyz) <= ext / 1023 * 0.51 + 1e-6)             # half a step of the coarsest axis
    assert np.abs(back.log_scale - cloud.log_scale).max() <= (cloud.log_scale.max() - cloud.log_scale.min()) / 1023 * 0.51 + 1e-6
    q = cloud.quat / np.linalg.norm(cloud.quat, axis=1, keepdims=True)
    assert np.abs((q * back.quat).sum(1)).min() > 1 - 1e-5
    assert np.abs(back.sh[:, 0] - cloud.sh[:, 0]).max() <= 0.51 / 255 / SH_C0 + 1e-6
    if deg:
        assert np.abs(back.sh[:, 1:] - cloud.sh[:, 1:]).max() <= 8 / 256 * 0.51 + 1e-6
    a = 1 / (1 + np.exp(-cloud.opacity.astype(np.float64)))
    b = 1 / (1 + np.exp(-back.opacity.astype(np.float64)))
    assert np.abs(a - b).max() <= 0.51 / 255
    # the size is what the format promises: 16 B + SH bytes per splat (+ 72 B per chunk)
    assert p.stat().st_size < n * (16 + 3 * ((deg + 1) ** 2 - 1)) + (n // 256 + 1) * 72 + 4096
    # a plain PLY still takes the plain path
    save_ply(tmp_path / "p.ply", cloud)
    assert np.array_equal(load_cloud(tmp_path / "p.ply").xyz, cloud.xyz)


def test_stage_resolution_finds_interiorgs_layout(tmp_path):
    cloud, _ = sage.synthetic_room(0, 300)
    (tmp_path / "0001_839920").mkdir()
    save_compressed_ply(tmp_path / "0001_839920" / "3dgs_compressed.ply", cloud)
    (tmp_path / "0001_839920.usda").write_text('#usda 1.0\nover "gauss" ( prepend references = @0001_839920.usdz@ ) {}\n')
    gauss, coll, is_stage = resolve_stage_assets(str(tmp_path / "0001_839920.usda"))
    assert gauss is not None and gauss.endswith("3dgs_compressed.ply") and coll is None and is_stage
    assert load_cloud(gauss).n == 300


# ---- collision hulls -> ellipsoids -------------------------------------------------------------------
def _box_mesh(lo, hi):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    pts = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    quads = [[0, 1, 3, 2], [4, 6, 7, 5], [0, 4, 5, 1], [2, 3, 7, 6], [0, 2, 6, 4], [1, 5, 7, 3]]
    return pts, np.full(6, 4), np.array(quads).ravel()


def _inside(points, mu, semi, Rm, slack=1e-6):
    loc = (points - mu) @ Rm
    return np.all(((loc / semi) ** 2).sum(1) <= 1 + slack)


@pytest.mark.parametrize("fit", [cm.fit_ellipsoid_pca, cm.fit_ellipsoid_mvee])
def test_fitted_ellipsoid_encloses_every_vertex(fit):
    rng = np.random.default_rng(5)
    for _ in range(20):
        A = rng.normal(size=(3, 3)) * rng.uniform(0.05, 1.0, size=3)
        pts = rng.normal(size=(rng.integers(4, 40), 3)) @ A.T + rng.uniform(-5, 5, size=3)
        mu, semi, Rm = fit(pts)
        assert abs(np.linalg.det(Rm) - 1) < 1e-9 and np.all(semi > 0)
        assert _inside(pts, mu, semi, Rm)
    # a flat hull (wall panel) still gives a body, at least MIN_SEMI_AXIS thick
    flat = np.array([[0, 0, 0], [2, 0, 0], [2, 0, 3], [0, 0, 3.0]])
    mu, semi, Rm = fit(flat)
    assert semi.min() >= cm.MIN_SEMI_AXIS and _inside(flat, mu, semi, Rm)


def test_mvee_is_not_larger_than_pca_on_a_box():
    pts, _, _ = _box_mesh([0, 0, 0], [2, 1, 0.5])
    v = lambda semi: float(np.prod(semi))
    assert v(cm.fit_ellipsoid_mvee(pts)[1]) <= v(cm.fit_ellipsoid_pca(pts)[1]) * 1.001
    # the MVEE of a box is the sqrt(3)-scaled inscribed ellipsoid
    np.testing.assert_allclose(np.sort(cm.fit_ellipsoid_mvee(pts)[1]), np.sqrt(3) * np.array([0.25, 0.5, 1.0]), rtol=2e-3)


USDA = '''#usda 1.0
(
    defaultPrim = "World"
    upAxis = "Z"
)

def Xform "World"
{
    double3 xformOp:translate = (10, 0, 0)
    uniform token[] xformOpOrder = ["xformOp:translate"]

    def Xform "hulls"
    {
        quatf xformOp:orient = (0.70710678, 0, 0, 0.70710678)
        uniform token[] xformOpOrder = ["xformOp:orient"]

        def Mesh "hull_0"
        {
            int[] faceVertexCounts = [4, 4, 4, 4, 4, 4]
            int[] faceVertexIndices = [0, 1, 3, 2, 4, 6, 7, 5, 0, 4, 5, 1, 2, 3, 7, 6, 0, 2, 6, 4, 1, 5, 7, 3]
            point3f[] points = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (2, 0, 0), (2, 0, 1), (2, 1, 0), (2, 1, 1)]
            uniform token subdivisionScheme = "none"
        }

        def Mesh "two_hulls_in_one_prim"
        {
            int[] faceVertexCounts = [3, 3]
            int[] faceVertexIndices = [0, 1, 2, 3, 4, 5]
            point3f[] points = [(5, 5, 0), (6, 5, 0), (5, 6, 0.5), (-5, -5, 0), (-6, -5, 0), (-5, -6, 1e-1)]
        }
    }
}
'''


def test_usda_meshes_with_ancestor_transforms_and_component_split(tmp_path):
    p = tmp_path / "0001_collision.usda"
    p.write_text(USDA)
    meshes = cm.load_usda_meshes(p)
    assert len(meshes) == 2
    # hull_0: rotated 90 degrees about Z (x, y) -> (-y, x), then translated by (10, 0, 0)
    pts = meshes[0][0]
    np.testing.assert_allclose(pts.min(0), [9, 0, 0], atol=1e-6)
    np.testing.assert_allclose(pts.max(0), [10, 2, 1], atol=1e-6)
    mu, ls, q = cm.meshes_to_ellipsoids(meshes, stage_rotation=False)
    assert mu.shape == (3, 3) and ls.shape == (3, 3) and q.shape == (3, 4)      # the second prim splits into two bodies
    np.testing.assert_allclose(mu[0], [9.5, 1.0, 0.5], atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)
    # stage rotation (template.usda:161): 180 degrees about Z
    mu_r, _, _ = cm.meshes_to_ellipsoids(meshes, stage_rotation=True)
    np.testing.assert_allclose(mu_r[0], [-9.5, -1.0, 0.5], atol=1e-5)
    # the facade's loader takes the mesh file directly
    mu2, ls2, q2 = load_collision_bodies(p)
    assert np.array_equal(mu2, mu_r) and mu2.dtype == np.float32
    # and the converter CLI writes what the facade reads
    cm.main([str(p), str(tmp_path / "0001_collision.npz")])
    mu3, _, _ = load_collision_bodies(tmp_path / "0001_collision.npz")
    assert np.array_equal(mu3, mu_r)


def test_obj_groups_become_bodies_and_capsule_agrees_with_hull(tmp_path):
    # two boxes as OBJ groups; the capsule oracle on the fitted ellipsoids must report a hit wherever the
    # agent's vertical segment actually passes through a box (enclosing fit => no false negatives)
    import oracle
    boxes = [([1, 1, 0], [2, 3, 1.5]), ([4, 0.5, 0], [4.4, 0.9, 2.0])]
    lines, off = [], 0
    for k, (lo, hi) in enumerate(boxes):
        pts, cnt, idx = _box_mesh(lo, hi)
        lines.append(f"o hull_{k}")
        lines += [f"v {x} {y} {z}" for x, y, z in pts]
        for f in idx.reshape(-1, 4):
            lines.append("f " + " ".join(str(int(i) + 1 + off) for i in f))
        off += len(pts)
    p = tmp_path / "scene_collision.obj"
    p.write_text("\n".join(lines) + "\n")
    mu, ls, q = cm.convert(p, stage_rotation=False)
    assert mu.shape[0] == 2
    rng = np.random.default_rng(0)
    pos = np.stack([rng.uniform(0, 5, 400), rng.uniform(0, 4, 400), np.full(400, 0.85)], 1).astype(np.float32)
    hit, _ = oracle.collide_capsule(mu, ls, q, pos, radius=0.0)
    truly_inside = np.zeros(400, bool)
    for lo, hi in boxes:
        truly_inside |= (pos[:, 0] >= lo[0]) & (pos[:, 0] <= hi[0]) & (pos[:, 1] >= lo[1]) & (pos[:, 1] <= hi[1])
    assert truly_inside.sum() > 10
    assert np.all(hit[truly_inside])
