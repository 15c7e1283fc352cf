"""CPU: the oracle (binary32, frozen op order, polynomial exp2) against an independent float64
NumPy statement of the textbook 3DGS forward pass (Kerbl et al. 2023 equations).  Render parity is
UNPINNED by the reference (no source / golden image), so this guards against common-mode formula
errors in the spec itself.  Tolerances: colour 2e-3 (pre-quantisation), depth 2e-3 m, excluding
pixels where a skip/stop decision is within 3e-4 (alpha; the blend's degree-3 exp2 is good to
7.5e-5) / 1e-3 (T), relative, of its threshold in the float64 evaluation.  Pixels nothing covers
carry depth +inf on both sides."""
import numpy as np
import pytest

from conftest import oracle_camera

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def sh_basis(d, deg):
    x, y, z = d
    b = [SH_C0]
    if deg > 0:
        b += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg > 2:
        b += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
              SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    return np.array(b)


def textbook_render(cam, cloud):
    f8 = lambda a: np.asarray(a, np.float64)
    R, t, cp = f8(cam.R), f8(cam.t), f8(cam.campos)
    W, H = cam.width, cam.height
    splats = []
    for i in range(cloud.n):
        mu = f8(cloud.xyz[i])
        tv = R @ mu + t
        if not (cam.near < tv[2] < cam.far):
            continue
        o = 1.0 / (1.0 + np.exp(-float(cloud.opacity[i])))
        if o < 1.0 / 255.0:
            continue
        q = f8(cloud.quat[i]); q = q / np.linalg.norm(q)
        w, x, y, z = q
        Rg = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                       [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                       [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        S = np.diag(np.exp(f8(cloud.log_scale[i])))
        Sigma = Rg @ S @ S @ Rg.T
        limx, limy = 1.3 * W / (2 * cam.fx), 1.3 * H / (2 * cam.fy)
        txc = np.clip(tv[0] / tv[2], -limx, limx) * tv[2]
        tyc = np.clip(tv[1] / tv[2], -limy, limy) * tv[2]
        J = np.array([[cam.fx / tv[2], 0, -cam.fx * txc / tv[2] ** 2], [0, cam.fy / tv[2], -cam.fy * tyc / tv[2] ** 2]])
        cov = J @ R @ Sigma @ R.T @ J.T + 0.3 * np.eye(2)
        det = np.linalg.det(cov)
        if det <= 0:
            continue
        conic = np.linalg.inv(cov)
        g = np.array([cam.fx * tv[0] / tv[2] + cam.cx, cam.fy * tv[1] / tv[2] + cam.cy])
        d = mu - cp; d = d / np.linalg.norm(d)
        col = np.maximum(sh_basis(d, cloud.sh_degree) @ f8(cloud.sh[i]) + 0.5, 0.0)
        splats.append((np.float32(tv[2]).view(np.uint32), i, tv[2], g, conic, o, col))
    splats.sort(key=lambda s: (int(s[0]), s[1]))
    C = np.zeros((H, W, 3)); D = np.zeros((H, W)); T = np.ones((H, W))
    done = np.zeros((H, W), bool); fragile = np.zeros((H, W), bool)
    ys, xs = np.mgrid[0:H, 0:W]
    px, py = xs + 0.5, ys + 0.5
    for _, _, z, g, conic, o, col in splats:
        dx, dy = px - g[0], py - g[1]
        power = -0.5 * (conic[0, 0] * dx * dx + conic[1, 1] * dy * dy) - conic[0, 1] * dx * dy
        a = np.minimum(0.99, o * np.exp(power))
        fragile |= (~done) & (np.abs(a - 1 / 255) < 3e-4 * (1 / 255))
        use = (~done) & (power <= 0) & (a >= 1 / 255)
        Tn = T * (1 - a)
        fragile |= use & (np.abs(Tn - 1e-4) < 1e-3 * 1e-4)
        stop = use & (Tn < 1e-4)
        done |= stop
        use &= ~stop
        wgt = np.where(use, a * T, 0.0)
        C += wgt[..., None] * col
        D += wgt * z
        T = np.where(use, Tn, T)
    cover = 1 - T
    depth = np.where(cover > 1e-6, D / np.maximum(cover, 1e-30), np.inf)
    fragile |= np.abs(cover - 1e-6) < 1e-7
    return C, depth, T, fragile


@pytest.mark.parametrize("seed,deg", [(11, 3), (12, 1), (13, 0)])
def test_oracle_follows_textbook_equations(sage, oracle, seed, deg):
    cloud, _ = sage.synthetic_room(seed, 400, sh_degree=deg, size_scale=2.5)
    cam = sage.Camera.look_heading((2.5, 1.8, 1.2), np.radians(200.0 + seed), 72, 54)
    C, depth, T, fragile = textbook_render(cam, cloud)
    rgb, dep, st, fT = oracle.render(oracle_camera(oracle, cam), *cloud.as_tuple(), brute=True, want_T=True)
    ok = ~fragile
    assert ok.mean() > 0.97
    assert np.max(np.abs(fT - T)[ok]) < 2e-3
    hit = ok & np.isfinite(depth)
    assert np.array_equal(np.isfinite(dep)[ok], np.isfinite(depth)[ok])
    assert np.max(np.abs(dep - depth)[hit]) < 2e-3
    ref8 = (np.clip(C, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    assert np.max(np.abs(rgb.astype(int) - ref8.astype(int))[ok]) <= 1
    assert (T < 0.9).mean() > 0.3, "scene too empty to mean anything"


def test_oracle_binned_equals_brute(sage, oracle):
    cloud, _ = sage.synthetic_room(4, 3000)
    cam = sage.Camera.look_heading((3.0, 2.0, 1.2), 2.0, 120, 90)
    oc = oracle_camera(oracle, cam)
    a = oracle.render(oc, *cloud.as_tuple(), brute=False)
    b = oracle.render(oc, *cloud.as_tuple(), brute=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_blend_exp2_and_log2_accuracy(oracle):
    """The two helper functions the exponent-domain blend adds to the spec."""
    import ctypes as C
    L = oracle.lib()
    L.so_exp2_fast.restype = C.c_float; L.so_exp2_fast.argtypes = [C.c_float]
    L.so_log2.restype = C.c_float; L.so_log2.argtypes = [C.c_float]
    xs = np.linspace(-7.99435344, -0.0144995697, 4001).astype(np.float32)
    got = np.array([L.so_exp2_fast(float(x)) for x in xs], np.float64)
    assert np.max(np.abs(got / 2.0 ** xs.astype(np.float64) - 1)) < 8e-5
    assert np.all(np.diff(got) >= 0), "alpha must not decrease with the exponent"
    os_ = np.concatenate([np.random.default_rng(0).uniform(1 / 300, 1, 4000), [1.0, 0.5, 1 / 255]]).astype(np.float32)
    lg = np.array([L.so_log2(float(o)) for o in os_], np.float64)
    assert np.max(np.abs(lg - np.log2(os_.astype(np.float64)))) < 1e-6
    assert L.so_log2(1.0) == 0.0 and L.so_log2(0.5) == -1.0 and L.so_log2(0.0) == -127.0


def test_exp2_polynomial_accuracy(oracle):
    xs = np.linspace(-30, 0, 20001).astype(np.float32).astype(np.float64)
    got = np.array([oracle.exp2(float(x)) for x in xs])
    assert np.max(np.abs(got / 2.0 ** xs - 1)) < 4e-7
    assert oracle.exp2(-1000.0) < 1e-37 and oracle.exp2(0.0) == pytest.approx(1.0, abs=2e-7)


def test_oracle_tile_mode_equals_brute_force_on_needle_splats(sage, oracle):
    """The oracle's tile-list mode is an acceleration of its brute-force mode (every splat against every pixel: the
    definition).  Needle-like splats are where a bounding box from the exact ellipse is NOT enough - the binary32
    exponent's rounding passes the 1/255 test beyond it (seed 2 at 1080p, pixel (462, 974), until round 2)."""
    from conftest import oracle_camera
    from test_render_parity_gpu import needle_scene
    for seed, res in ((2, (1920, 1080)), (0, (640, 480)), (1, (640, 480)), (3, (333, 187))):
        cloud, cam = needle_scene(sage, seed, *res)
        oc = oracle_camera(oracle, cam)
        a_rgb, a_dep, _ = oracle.render(oc, *cloud.as_tuple())
        b_rgb, b_dep, _ = oracle.render(oc, *cloud.as_tuple(), brute=True)
        assert np.array_equal(a_rgb, b_rgb) and np.array_equal(a_dep.view(np.uint32), b_dep.view(np.uint32)), (seed, res)
