// K0 precompute (scene upload), K1a project + cull (per frame), K1b SH colour (per frame).
// The per-frame kernels read the camera from the frame context's device-resident parameter block
// (SageFrameParams), so their launch parameters are frame-invariant and a frame replays as a graph.
//
// Stands in for the per-Gaussian work hidden inside Isaac's world.step(render=True)
// (reference call sites simple_env.py:1369, generate_images.py:426).  The arithmetic follows
// DESIGN.md §2.2-2.3 operation by operation; the CPU oracle restates the same sequence.
#include "sage_internal.cuh"

// ---------------------------------------------------------------------------------------------
// K0: view-independent terms, computed once per scene: sigmoid(opacity), the 3-D covariance, and
// the SH coefficients repacked as 16-byte chunks, contiguous per Gaussian (a visible Gaussian's
// 192 B are six whole 32-byte sectors, so a scattered visible set wastes no DRAM traffic).
// Gaussians with a non-finite attribute or an SH coefficient >= 1e30 in magnitude are dropped at
// load (spec rule): their log2-opacity is stored as -127, which the per-frame cull rejects.
// Opacity is kept as lo = log2(sigmoid(raw)) (deterministic sage_log2): the blend adds it to the
// exponent instead of multiplying alpha by it.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_precompute(SageSceneDev sc, const float* __restrict__ xyz, const float* __restrict__ log_scale,
             const float* __restrict__ quat, const float* __restrict__ opacity,
             const float* __restrict__ sh)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sc.n) return;
    const int nf = 3 * (sc.deg + 1) * (sc.deg + 1);
    const float* src = sh + (size_t)nf * i;
    bool ok = true;
    for (int c = 0; c < sc.n_chunks; ++c) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = (4 * c + k < nf) ? src[4 * c + k] : 0.0f;
            ok = ok && (fabsf(v[k]) < 1e30f);          // false for NaN / inf as well
        }
        sc.shc[(size_t)i * sc.n_chunks + c] = make_float4(v[0], v[1], v[2], v[3]);
    }
    float chk = xyz[3 * i + 0] + xyz[3 * i + 1] + xyz[3 * i + 2] + log_scale[3 * i + 0] +
                log_scale[3 * i + 1] + log_scale[3 * i + 2] + quat[4 * i + 0] + quat[4 * i + 1] +
                quat[4 * i + 2] + quat[4 * i + 3] + opacity[i];
    ok = ok && (chk - chk == 0.0f);
    float o = ok ? 1.0f / (1.0f + sage_exp(-opacity[i])) : 0.0f;
    sc.geo0[i] = make_float4(xyz[3 * i + 0], xyz[3 * i + 1], xyz[3 * i + 2], sage_log2(o));

    float sx = sage_exp(log_scale[3 * i + 0]);
    float sy = sage_exp(log_scale[3 * i + 1]);
    float sz = sage_exp(log_scale[3 * i + 2]);
    float qw = quat[4 * i + 0], qx = quat[4 * i + 1], qy = quat[4 * i + 2], qz = quat[4 * i + 3];
    float qn = sqrtf(fmaf(qz, qz, fmaf(qy, qy, fmaf(qx, qx, qw * qw))));
    float w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
    float xx = x * x, yy = y * y, zz = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    float r00 = 1.0f - 2.0f * (yy + zz), r01 = 2.0f * (xy - wz), r02 = 2.0f * (xz + wy);
    float r10 = 2.0f * (xy + wz), r11 = 1.0f - 2.0f * (xx + zz), r12 = 2.0f * (yz - wx);
    float r20 = 2.0f * (xz - wy), r21 = 2.0f * (yz + wx), r22 = 1.0f - 2.0f * (xx + yy);
    float m00 = r00 * sx, m01 = r01 * sy, m02 = r02 * sz;
    float m10 = r10 * sx, m11 = r11 * sy, m12 = r12 * sz;
    float m20 = r20 * sx, m21 = r21 * sy, m22 = r22 * sz;
    float S00 = fmaf(m02, m02, fmaf(m01, m01, m00 * m00));
    float S01 = fmaf(m02, m12, fmaf(m01, m11, m00 * m10));
    float S02 = fmaf(m02, m22, fmaf(m01, m21, m00 * m20));
    float S11 = fmaf(m12, m12, fmaf(m11, m11, m10 * m10));
    float S12 = fmaf(m12, m22, fmaf(m11, m21, m10 * m20));
    float S22 = fmaf(m22, m22, fmaf(m21, m21, m20 * m20));
    sc.geo1[i] = make_float4(S00, S01, S02, S11);
    sc.geo2[i] = make_float2(S12, S22);
}

void sage_launch_precompute(const SageSceneDev& sc, const float* xyz, const float* log_scale,
                            const float* quat, const float* opacity, const float* sh, cudaStream_t st)
{
    if (sc.n <= 0) return;
    int blocks = (int)((sc.n + 255) / 256);
    SAGE_KLAUNCH(k_precompute, blocks, 256, st, sc, xyz, log_scale, quat, opacity, sh);
}

// ---------------------------------------------------------------------------------------------
// K1a: project, cull, conic, conservative pixel bounding box.  One thread per Gaussian; every
// per-Gaussian stream is a 16-byte-per-lane coalesced access.  Emits the depth key
// (bits(z) - bits(near), or CULLED) that the first radix pass compacts and sorts in one go.
// ---------------------------------------------------------------------------------------------
#ifdef SAGE_CUDA_EMU
constexpr int SAGE_STREAM_GRID = 3;
#else
constexpr int SAGE_STREAM_GRID = 148 * 8;      // fixed grid of the two streaming kernels (grid-stride loops)
#endif

__global__ void __launch_bounds__(256)
k_project(SageSceneDev sc, SageFrame fr)
{
    const SageCamDev& cam = fr.fp->cam;
    SAGE_DEP_LAUNCH_TOP();
    // grid-stride over the Gaussians (a fixed grid of 148 x 8 CTAs): at 2 M Gaussians one CTA per 256 of them was
    // 7813 CTA launches per frame for this kernel alone - at multi-GPU frame rates the CTA launch rate of the
    // chip, not the work, was the limit
    uint32_t n_vis = 0;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < sc.n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    bool vis = false;
    uint32_t key = SAGE_KEY_CULLED;
    if (i < sc.n) {
        const float4 g0 = __ldg(&sc.geo0[i]);
        const float mx = g0.x, my = g0.y, mz = g0.z, lo = g0.w;
        const float* R = cam.R;
        float tx = fmaf(R[2], mz, fmaf(R[1], my, fmaf(R[0], mx, cam.t[0])));
        float ty = fmaf(R[5], mz, fmaf(R[4], my, fmaf(R[3], mx, cam.t[1])));
        float tz = fmaf(R[8], mz, fmaf(R[7], my, fmaf(R[6], mx, cam.t[2])));
        if (tz > cam.near_z && tz < cam.far_z && lo >= SAGE_L255) {
            const float4 g1 = __ldg(&sc.geo1[i]);
            const float2 g2 = __ldg(&sc.geo2[i]);
            const float S00 = g1.x, S01 = g1.y, S02 = g1.z, S11 = g1.w, S12 = g2.x, S22 = g2.y;
            float txz = tx / tz, tyz = ty / tz;
            float ctx = fminf(cam.limx, fmaxf(-cam.limx, txz)) * tz;
            float cty = fminf(cam.limy, fmaxf(-cam.limy, tyz)) * tz;
            float tz2 = tz * tz;
            float j00 = cam.fx / tz;
            float j02 = -(cam.fx * ctx) / tz2;
            float j11 = cam.fy / tz;
            float j12 = -(cam.fy * cty) / tz2;
            float T00 = fmaf(j02, R[6], j00 * R[0]);
            float T01 = fmaf(j02, R[7], j00 * R[1]);
            float T02 = fmaf(j02, R[8], j00 * R[2]);
            float T10 = fmaf(j12, R[6], j11 * R[3]);
            float T11 = fmaf(j12, R[7], j11 * R[4]);
            float T12 = fmaf(j12, R[8], j11 * R[5]);
            float v00 = fmaf(S02, T02, fmaf(S01, T01, S00 * T00));
            float v01 = fmaf(S12, T02, fmaf(S11, T01, S01 * T00));
            float v02 = fmaf(S22, T02, fmaf(S12, T01, S02 * T00));
            float v10 = fmaf(S02, T12, fmaf(S01, T11, S00 * T10));
            float v11 = fmaf(S12, T12, fmaf(S11, T11, S01 * T10));
            float v12 = fmaf(S22, T12, fmaf(S12, T11, S02 * T10));
            float a = fmaf(T02, v02, fmaf(T01, v01, T00 * v00)) + SAGE_BLUR;
            float b = fmaf(T12, v02, fmaf(T11, v01, T10 * v00));
            float c = fmaf(T12, v12, fmaf(T11, v11, T10 * v10)) + SAGE_BLUR;
            float det = fmaf(a, c, -(b * b));
            if (det > 0.0f && det < INFINITY) {
                float conic_a = c / det, conic_b = -b / det, conic_c = a / det;
                float gx = fmaf(cam.fx, txz, cam.cx);
                float gy = fmaf(cam.fy, tyz, cam.cy);
                float qa = conic_a * (-0.5f * SAGE_LOG2E);
                float qb = conic_b * (-SAGE_LOG2E);
                float qc = conic_c * (-0.5f * SAGE_LOG2E);
                float chk = gx + gy + qa + qb + qc;

                // Conservative bounding rectangle of {alpha >= 1/255}, from the *rounded* conic the blend kernel
                // will use.  Not part of the spec: it only has to be conservative.  binary32 (round 1 used binary64:
                // two DSQRT and two DDIV sequences per visible splat, half of this kernel's instructions) with the
                // determinant by Kahan's fused-error form, so that a thin ellipse (A C ~ B^2) does not lose it to
                // cancellation, 0.3 % + 0.02 of slack on the log2 threshold and 0.1 % + 0.01 px on the half extents —
                // three orders above the rounding of the few operations involved.
                const float l2 = lo - SAGE_L255;                 // log2(255 o)
                const float t2 = l2 * 1.003f + 0.02f;
                float A = -qa, Cq = -qc;
                const float Bq = -0.5f * qb;
                sage_conic_deflate(&A, Bq, &Cq);                 // (the rounded exponent reaches further than the exact one: sage_internal.cuh)
                const float det2 = sage_det_kahan(A, Cq, Bq);
                const bool unbounded = !(det2 > 0.0f) || !(A > 0.0f) || !(Cq > 0.0f);
                float hx = unbounded ? 1e30f : sqrtf(t2 * Cq / det2);
                float hy = unbounded ? 1e30f : sqrtf(t2 * A / det2);
                if (!(hx < 1e9f)) hx = 1e30f;                    // overflow / NaN: cover everything
                if (!(hy < 1e9f)) hy = 1e30f;
                const float slack = 1e-3f * fmaxf(hx, hy) + 0.01f;
                const float xlo = floorf((gx - hx) - slack - 0.5f), xhi = ceilf((gx + hx) + slack - 0.5f);
                const float ylo = floorf((gy - hy) - slack - 0.5f), yhi = ceilf((gy + hy) + slack - 0.5f);
                const float Wm = (float)(cam.W - 1), Hm = (float)(cam.H - 1);
                if ((chk - chk == 0.0f) && xhi >= 0.0f && xlo <= Wm && yhi >= 0.0f && ylo <= Hm) {
                    int ix0 = (int)fmaxf(xlo, 0.0f), ix1 = (int)fminf(xhi, Wm);
                    int iy0 = (int)fmaxf(ylo, 0.0f), iy1 = (int)fminf(yhi, Hm);
                    // clip to the rows this call renders (multi-GPU row sharding)
                    iy0 = max(iy0, cam.bin_row0 * SAGE_BIN_H);
                    iy1 = min(iy1, min(cam.bin_row1 * SAGE_BIN_H, cam.H) - 1);
                    if (iy0 <= iy1 && ix0 <= ix1) {
                        vis = true;
                        key = __float_as_uint(tz) - __float_as_uint(cam.near_z);
                        fr.rec[3 * i + 0] = make_float4(gx, gy, qa, lo);
                        fr.rec[3 * i + 1] = make_float4(qb, qc, 0.0f, 0.0f);
                        // inclusive pixel bounding box (u16 each); the exact bin ranges are taken by
                        // k_emit, which walks only the visible set
                        fr.rect[i] = make_uint2((uint32_t)ix0 | ((uint32_t)ix1 << 16),
                                                (uint32_t)iy0 | ((uint32_t)iy1 << 16));
                    }
                }
            }
        }
        fr.dkey_src[i] = key;
    }
    n_vis += (uint32_t)__popc(__ballot_sync(0xffffffffu, vis));
    }
    // one atomic per warp for the visible counter
    if ((threadIdx.x & 31) == 0 && n_vis) atomicAdd(&fr.counters[CNT_VISIBLE], n_vis);
    SAGE_DEP_LAUNCH_END();
}

// ---------------------------------------------------------------------------------------------
// K1b: view-dependent SH colour of the visible Gaussians, in INDEX order (a thread per Gaussian
// checks its depth key): independent of the depth sort, so it runs on the frame's side stream /
// a parallel graph branch between the projection and the composite.  With a spatially ordered
// scene (Morton order at ingest) the visible set is contiguous runs: the 192-byte SH gathers of a
// warp land in neighbouring sectors.  Also writes the view depth z next to the colour (same fmaf
// chain as K1a: the same bits).
// ---------------------------------------------------------------------------------------------
__device__ __constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f,
                                            0.31539156525252005f, -1.0925484305920792f,
                                            0.5462742152960396f};
__device__ __constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,
                                            -0.4570457994644658f, 0.3731763325901154f,
                                            -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

template <int DEG>
__global__ void __launch_bounds__(256, 4)
k_sh_color(SageSceneDev sc, SageFrame fr)
{
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int NF = 3 * NB;
    constexpr int NC = (NF + 3) / 4;
    const SageCamDev& cam = fr.fp->cam;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sc.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (__ldg(&fr.dkey_src[i]) == SAGE_KEY_CULLED) continue;
    const float4 g0 = __ldg(&sc.geo0[i]);
    float shv[NC * 4];
    const float4* sp = sc.shc + (size_t)i * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float4 v = __ldg(&sp[c]);
        shv[4 * c + 0] = v.x; shv[4 * c + 1] = v.y; shv[4 * c + 2] = v.z; shv[4 * c + 3] = v.w;
    }
    const float dx = g0.x - cam.campos[0], dy = g0.y - cam.campos[1], dz = g0.z - cam.campos[2];
    float len = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    float x = dx / len, y = dy / len, z = dz / len;
    float basis[NB];
    basis[0] = 0.28209479177387814f;
    if constexpr (DEG > 0) {
        const float C1 = 0.4886025119029199f;
        basis[1] = -(C1 * y);
        basis[2] = C1 * z;
        basis[3] = -(C1 * x);
    }
    if constexpr (DEG > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        basis[4] = c_SH_C2[0] * xy;
        basis[5] = c_SH_C2[1] * yz;
        basis[6] = c_SH_C2[2] * ((2.0f * zz - xx) - yy);
        basis[7] = c_SH_C2[3] * xz;
        basis[8] = c_SH_C2[4] * (xx - yy);
        if constexpr (DEG > 2) {
            basis[9] = (c_SH_C3[0] * y) * (3.0f * xx - yy);
            basis[10] = (c_SH_C3[1] * xy) * z;
            basis[11] = (c_SH_C3[2] * y) * ((4.0f * zz - xx) - yy);
            basis[12] = (c_SH_C3[3] * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
            basis[13] = (c_SH_C3[4] * x) * ((4.0f * zz - xx) - yy);
            basis[14] = (c_SH_C3[5] * z) * (xx - yy);
            basis[15] = (c_SH_C3[6] * x) * (xx - 3.0f * yy);
        }
    }
    float acc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a = basis[0] * shv[c];
#pragma unroll
        for (int k = 1; k < NB; ++k) a = fmaf(basis[k], shv[3 * k + c], a);
        acc[c] = fmaxf(a + 0.5f, 0.0f);
    }
    const float* R = cam.R;
    const float tz = fmaf(R[8], g0.z, fmaf(R[7], g0.y, fmaf(R[6], g0.x, cam.t[2])));
    fr.rec[3 * (size_t)i + 2] = make_float4(acc[0], acc[1], acc[2], tz);
    }
}

void sage_launch_project(const SageSceneDev& sc, const SageFrame& fr, cudaStream_t st)
{
    if (sc.n <= 0) return;
    int blocks = (int)((sc.n + 255) / 256);
    if (blocks > SAGE_STREAM_GRID) blocks = SAGE_STREAM_GRID;
    SAGE_KLAUNCH(k_project, blocks, 256, st, sc, fr);
}

void sage_launch_sh_color(const SageSceneDev& sc, const SageFrame& fr, cudaStream_t st)
{
    if (sc.n <= 0) return;
    int blocks = (int)((sc.n + 255) / 256);
    if (blocks > SAGE_STREAM_GRID) blocks = SAGE_STREAM_GRID;
    switch (sc.deg) {
        case 0: SAGE_KLAUNCH(k_sh_color<0>, blocks, 256, st, sc, fr); break;
        case 1: SAGE_KLAUNCH(k_sh_color<1>, blocks, 256, st, sc, fr); break;
        case 2: SAGE_KLAUNCH(k_sh_color<2>, blocks, 256, st, sc, fr); break;
        default: SAGE_KLAUNCH(k_sh_color<3>, blocks, 256, st, sc, fr); break;
    }
}

void sage_project_configure(void)
{
#ifndef SAGE_CUDA_EMU
    sage_prefer_smem(k_project);
    sage_prefer_smem(k_sh_color<0>); sage_prefer_smem(k_sh_color<1>);
    sage_prefer_smem(k_sh_color<2>); sage_prefer_smem(k_sh_color<3>);
#endif
}
