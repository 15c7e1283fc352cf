"""GPU: the CUDA render path against a plain PyTorch fp32 statement of the textbook 3DGS forward pass (Kerbl et
al. 2023: EWA projection, 0.3 px blur, alpha = min(0.99, o exp(-1/2 d^T conic d)), skip below 1/255, stop when T
would fall below 1e-4, SH degree <= 3) — an implementation that shares nothing with the oracle or the kernels
(torch.exp / torch.sigmoid, no polynomial exp2, no binning, no frozen op order).

Tolerance (stated here because the two sides round differently and the discrete skip/stop decisions can flip on a
borderline splat): PSNR >= 50 dB and every channel within 2/255; depth within 1e-3 m on >= 99.5 % of the pixels both
sides cover (a flipped decision moves a pixel by one splat's weight) and never more than 0.05 m; coverage masks
equal except where coverage is within 1e-5 of the 1e-6 threshold."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def torch_render(cam, cloud, dev):
    import torch
    f = lambda a: torch.as_tensor(np.asarray(a, np.float32), device=dev)
    R, t, cp = f(cam.R).reshape(3, 3), f(cam.t), f(cam.campos)
    W, H = cam.width, cam.height
    mu, ls, q, op, sh = f(cloud.xyz), f(cloud.log_scale), f(cloud.quat), f(cloud.opacity), f(cloud.sh)
    tv = mu @ R.T + t
    z = tv[:, 2]
    o = torch.sigmoid(op)
    keep = (z > cam.near) & (z < cam.far) & (o >= 1.0 / 255.0)
    q = q / q.norm(dim=1, keepdim=True)
    w_, x, y, zz = q.unbind(1)
    Rg = torch.stack([1 - 2 * (y * y + zz * zz), 2 * (x * y - w_ * zz), 2 * (x * zz + w_ * y),
                      2 * (x * y + w_ * zz), 1 - 2 * (x * x + zz * zz), 2 * (y * zz - w_ * x),
                      2 * (x * zz - w_ * y), 2 * (y * zz + w_ * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    M = Rg * torch.exp(ls)[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    limx, limy = 1.3 * W / (2 * cam.fx), 1.3 * H / (2 * cam.fy)
    zc = z.clamp(min=1e-6)
    txc = (tv[:, 0] / zc).clamp(-limx, limx) * zc
    tyc = (tv[:, 1] / zc).clamp(-limy, limy) * zc
    J = torch.zeros((mu.shape[0], 2, 3), device=dev)
    J[:, 0, 0] = cam.fx / zc; J[:, 0, 2] = -cam.fx * txc / (zc * zc)
    J[:, 1, 1] = cam.fy / zc; J[:, 1, 2] = -cam.fy * tyc / (zc * zc)
    T_ = J @ R
    cov = T_ @ Sigma @ T_.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    keep &= (det > 0) & torch.isfinite(det)
    ca, cb, cc = c / det, -b / det, a / det
    gx = cam.fx * tv[:, 0] / zc + cam.cx
    gy = cam.fy * tv[:, 1] / zc + cam.cy
    d = mu - cp
    d = d / d.norm(dim=1, keepdim=True)
    dx, dy, dz = d.unbind(1)
    deg = cloud.sh_degree
    basis = [torch.full_like(dx, SH_C0)]
    if deg > 0:
        basis += [-SH_C1 * dy, SH_C1 * dz, -SH_C1 * dx]
    if deg > 1:
        xx, yy, z2, xy, yz, xz = dx * dx, dy * dy, dz * dz, dx * dy, dy * dz, dx * dz
        basis += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * z2 - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg > 2:
        basis += [SH_C3[0] * dy * (3 * xx - yy), SH_C3[1] * xy * dz, SH_C3[2] * dy * (4 * z2 - xx - yy),
                  SH_C3[3] * dz * (2 * z2 - 3 * xx - 3 * yy), SH_C3[4] * dx * (4 * z2 - xx - yy),
                  SH_C3[5] * dz * (xx - yy), SH_C3[6] * dx * (xx - 3 * yy)]
    col = (torch.stack(basis, 1)[:, :, None] * sh).sum(1).add(0.5).clamp(min=0.0)
    idx = torch.nonzero(keep).squeeze(1)
    keys = z[idx].view(torch.int32).to(torch.int64) * (1 << 32) + idx          # (depth bits, index)
    order = idx[torch.argsort(keys)]
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32) + 0.5,
                            torch.arange(W, device=dev, dtype=torch.float32) + 0.5, indexing="ij")
    C = torch.zeros((H, W, 3), device=dev); D = torch.zeros((H, W), device=dev); T = torch.ones((H, W), device=dev)
    done = torch.zeros((H, W), dtype=torch.bool, device=dev)
    for i in order.tolist():
        ddx, ddy = xs - gx[i], ys - gy[i]
        power = -0.5 * (ca[i] * ddx * ddx + cc[i] * ddy * ddy) - cb[i] * ddx * ddy
        al = (o[i] * torch.exp(power)).clamp(max=0.99)
        use = (~done) & (al >= 1.0 / 255.0)
        Tn = T * (1 - al)
        stop = use & (Tn < 1e-4)
        done |= stop
        use &= ~stop
        wgt = torch.where(use, al * T, torch.zeros_like(T))
        C += wgt[..., None] * col[i]
        D += wgt * z[i]
        T = torch.where(use, Tn, T)
    cover = 1 - T
    depth = torch.where(cover > 1e-6, D / cover.clamp(min=1e-30), torch.full_like(D, float("inf")))
    return C, depth, cover


@pytest.mark.parametrize("seed,n,deg,res", [(21, 2500, 3, (160, 120)), (22, 1500, 1, (97, 61)), (23, 800, 0, (64, 48))])
def test_cuda_frame_matches_plain_torch_fp32(sage, seed, n, deg, res):
    import torch
    cloud, _ = sage.synthetic_room(seed, n, sh_degree=deg, size_scale=2.0)
    cam = sage.Camera.look_heading((2.5, 1.8, 1.2), np.radians(190.0 + seed), res[0], res[1])
    scene = sage.GaussianScene(cloud)
    rgb, depth = scene.render(cam)
    torch.cuda.synchronize()
    C, ref_depth, cover = torch_render(cam, cloud, rgb.device)
    ref8 = (C.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8)
    diff = (rgb.to(torch.int32) - ref8.to(torch.int32)).abs()
    assert int(diff.max()) <= 2
    mse = float((diff.to(torch.float32) ** 2).mean())
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    assert psnr >= 50.0, psnr
    fin_a, fin_b = torch.isfinite(depth), torch.isfinite(ref_depth)
    sure = (cover - 1e-6).abs() > 1e-5
    assert bool((fin_a == fin_b)[sure].all())
    both = fin_a & fin_b
    err = (depth - ref_depth).abs()[both]
    assert float((err <= 1e-3).float().mean()) >= 0.995 and float(err.max()) <= 0.05, (float(err.max()), float((err <= 1e-3).float().mean()))
    assert float((cover > 0.1).float().mean()) > 0.3, "scene too empty to mean anything"
    scene.close()
