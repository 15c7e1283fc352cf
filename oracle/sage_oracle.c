/*
 * sage_oracle.c — CPU ORACLE for the render-and-collide hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (sage-3d_official_b200/) never imports, links or calls it.
 *
 * What it restates, and what pins it:
 *
 *  (1) so_collide_grid*  — a C restatement of the reference's
 *      SemanticMap2DCollisionDetector.check_collision_3d
 *      (Code/benchmark/environment_evaluation/collision_detector.py:191-237, with
 *      forward_position_mapping :146-189 and _world_to_pixel :134-138).
 *      PINNED: tests/golden/collide_grid_*.npz were produced by importing the unmodified
 *      reference class in the authoring container (tests/golden/make_collide_golden.py).
 *
 *  (2) so_render / so_project — the frozen-spec 3DGS forward pass (SURVEY.md
 *      §8 a-spec).  The reference delegates this arithmetic to Isaac Sim's closed RTX
 *      renderer (simple_env.py:1369,1380,1425-1427; generate_images.py:426-428); no source,
 *      golden image or known-answer vector for it exists in the reference tree.
 *      ==> PARITY UNPINNED for render: this file *is* the specification the CUDA kernels
 *      are held bit-exact to; tests/test_oracle_textbook.py additionally checks it against an
 *      independent float64 NumPy statement of the textbook 3DGS equations.
 *
 *  (3) so_collide_capsule — ellipsoid-vs-agent-capsule overlap.  The
 *      reference has no such test (its PhysX fallback, simple_env.py:1878-1931, cannot fire
 *      for a kinematic body) ==> PARITY UNPINNED; this restatement is the spec.
 *
 * Numerics contract: every float operation below is an IEEE-754 binary32 operation in the
 * order written; fused multiply-adds appear only where fmaf() is written.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -mfma -fopenmp
 * so the compiler neither fuses nor re-associates.  The CUDA side is built with -fmad=false and
 * uses explicit fmaf/__ffma2_rn, which is what makes bit-exact parity possible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SO_TILE 16
#define SO_L255 (-7.99435344f)  /* log2(1/255) rounded to binary32: a splat is skipped where log2(alpha) < this */
#define SO_L099 (-0.0144995697f) /* log2(0.99) rounded to binary32: alpha is capped at 0.99 in the exponent domain */
#define SO_T_MIN 1e-4f
#define SO_BLUR 0.3f
#define SO_LOG2E 1.44269504f

typedef struct {
    float R[9];      /* model->camera rotation, row-major; camera axes: +x right, +y down, +z forward */
    float t[3];      /* model->camera translation */
    float campos[3]; /* camera centre in model space (SH view direction) */
    float fx, fy, cx, cy;
    int32_t W, H;
    float near_z, far_z;
} so_camera;

/* per-Gaussian projected record ("splat") */
typedef struct {
    float gx, gy;       /* mean in continuous pixel coordinates */
    float qa, qb, qc;   /* log2e-folded conic: power2 = qa dx^2 + qb dx dy + qc dy^2 */
    float lo;           /* log2(sigmoid(raw)) by so_log2 */
    float z;            /* planar view depth */
    float r, g, b;      /* SH colour, clamped >= 0 */
    float cov_a, cov_c; /* blurred 2-D covariance diagonal (oracle-side bounding only) */
} so_splat;

/* ------------------------------------------------------------------ exp2 */

static inline float so_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t so_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* Deterministic binary32 exp2: round-to-nearest split + degree-5 polynomial. */
float so_exp2(float x)
{
    x = fminf(fmaxf(x, -125.0f), 125.0f);
    const float magic = 12582912.0f; /* 1.5 * 2^23 */
    float t = x + magic;
    float n = t - magic;
    float f = x - n;
    float p = 0x1.5c08e4p-10f;
    p = fmaf(p, f, 0x1.3d0c52p-7f);
    p = fmaf(p, f, 0x1.c6b6e4p-5f);
    p = fmaf(p, f, 0x1.ebf918p-3f);
    p = fmaf(p, f, 0x1.62e428p-1f);
    p = fmaf(p, f, 0x1.000002p+0f);
    return so_u2f(so_f2u(p) + (so_f2u(t) << 23));
}

float so_exp(float x) { return so_exp2(x * SO_LOG2E); }

/* The blend's exp2: same split, degree-3 polynomial (max rel. error 7.5e-5, far below
 * the 1/255 colour quantum); callers guarantee x in [SO_L255, SO_L099], so no clamp. */
float so_exp2_fast(float x)
{
    const float magic = 12582912.0f;
    float t = x + magic;
    float n = t - magic;
    float f = x - n;
    float p = 0x1.c3f760p-5f;
    p = fmaf(p, f, 0x1.f0de1ap-3f);
    p = fmaf(p, f, 0x1.62f31ap-1f);
    p = fmaf(p, f, 0x1.fff692p-1f);
    return so_u2f(so_f2u(p) + (so_f2u(t) << 23));
}

/* Deterministic binary32 log2 for x in (0, 1]: exponent/mantissa split,
 * m in (sqrt(1/2), sqrt(2)], log(m) = 2 atanh((m-1)/(m+1)) as an odd series to s^9.
 * x == 0 (a Gaussian dropped at load) returns -127: below every threshold. */
float so_log2(float x)
{
    uint32_t u = so_f2u(x);
    float e = (float)((int)((u >> 23) & 255u) - 127);
    float m = so_u2f((u & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) { m = m * 0.5f; e = e + 1.0f; }
    float s = (m - 1.0f) / (m + 1.0f);
    float s2 = s * s;
    float p = 0x1.c71c72p-4f;             /* 1/9 */
    p = fmaf(p, s2, 0x1.24924ap-3f);      /* 1/7 */
    p = fmaf(p, s2, 0x1.99999ap-3f);      /* 1/5 */
    p = fmaf(p, s2, 0x1.555556p-2f);      /* 1/3 */
    p = fmaf(p, s2, 1.0f);
    float ln_m = (2.0f * s) * p;
    return fmaf(ln_m, SO_LOG2E, e);
}

/* ------------------------------------------------------------------ per-Gaussian projection */

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* View-independent part: 3-D covariance (6 unique) and opacity.  Returns 0 on success. */
static void so_cov3d(const float* ls, const float* q, float* S)
{
    float sx = so_exp(ls[0]), sy = so_exp(ls[1]), sz = so_exp(ls[2]);
    float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
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
    S[0] = fmaf(m02, m02, fmaf(m01, m01, m00 * m00)); /* 00 */
    S[1] = fmaf(m02, m12, fmaf(m01, m11, m00 * m10)); /* 01 */
    S[2] = fmaf(m02, m22, fmaf(m01, m21, m00 * m20)); /* 02 */
    S[3] = fmaf(m12, m12, fmaf(m11, m11, m10 * m10)); /* 11 */
    S[4] = fmaf(m12, m22, fmaf(m11, m21, m10 * m20)); /* 12 */
    S[5] = fmaf(m22, m22, fmaf(m21, m21, m20 * m20)); /* 22 */
}

static void so_sh_color(int deg, const float* sh /* [(deg+1)^2][3] */, float dx, float dy, float dz,
                        float* rgb)
{
    float len = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    float x = dx / len, y = dy / len, z = dz / len;
    float basis[16];
    int nb = (deg + 1) * (deg + 1);
    basis[0] = SH_C0;
    if (deg > 0) {
        basis[1] = -(SH_C1 * y);
        basis[2] = SH_C1 * z;
        basis[3] = -(SH_C1 * x);
    }
    if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        basis[4] = SH_C2[0] * xy;
        basis[5] = SH_C2[1] * yz;
        basis[6] = SH_C2[2] * ((2.0f * zz - xx) - yy);
        basis[7] = SH_C2[3] * xz;
        basis[8] = SH_C2[4] * (xx - yy);
        if (deg > 2) {
            basis[9] = (SH_C3[0] * y) * (3.0f * xx - yy);
            basis[10] = (SH_C3[1] * xy) * z;
            basis[11] = (SH_C3[2] * y) * ((4.0f * zz - xx) - yy);
            basis[12] = (SH_C3[3] * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
            basis[13] = (SH_C3[4] * x) * ((4.0f * zz - xx) - yy);
            basis[14] = (SH_C3[5] * z) * (xx - yy);
            basis[15] = (SH_C3[6] * x) * (xx - 3.0f * yy);
        }
    }
    for (int c = 0; c < 3; ++c) {
        float acc = basis[0] * sh[c];
        for (int k = 1; k < nb; ++k) acc = fmaf(basis[k], sh[3 * k + c], acc);
        rgb[c] = fmaxf(acc + 0.5f, 0.0f);
    }
}

/* Returns 1 if the Gaussian survives the spec's cull rules and fills *out. */
static int so_project_one(const so_camera* cam, int deg, const float* xyz, const float* ls,
                          const float* q, float op_raw, const float* sh, so_splat* out)
{
    const float* R = cam->R;
    /* load-time validity rule: every attribute finite, |SH coefficient| < 1e30 */
    {
        int nfl = 3 * (deg + 1) * (deg + 1);
        for (int k = 0; k < nfl; ++k)
            if (!(fabsf(sh[k]) < 1e30f)) return 0;
        float chk0 = xyz[0] + xyz[1] + xyz[2] + ls[0] + ls[1] + ls[2] + q[0] + q[1] + q[2] + q[3] + op_raw;
        if (!(chk0 - chk0 == 0.0f)) return 0;
    }
    float mx = xyz[0], my = xyz[1], mz = xyz[2];
    float tx = fmaf(R[2], mz, fmaf(R[1], my, fmaf(R[0], mx, cam->t[0])));
    float ty = fmaf(R[5], mz, fmaf(R[4], my, fmaf(R[3], mx, cam->t[1])));
    float tz = fmaf(R[8], mz, fmaf(R[7], my, fmaf(R[6], mx, cam->t[2])));
    if (!(tz > cam->near_z) || !(tz < cam->far_z)) return 0;

    float o = 1.0f / (1.0f + so_exp(-op_raw));
    float lo = so_log2(o);
    if (!(lo >= SO_L255)) return 0;       /* peak alpha below 1/255 */

    float S[6];
    so_cov3d(ls, q, S);

    float limx = 1.3f * ((0.5f * (float)cam->W) / cam->fx);
    float limy = 1.3f * ((0.5f * (float)cam->H) / cam->fy);
    float txz = tx / tz, tyz = ty / tz;
    float ctx = fminf(limx, fmaxf(-limx, txz)) * tz;
    float cty = fminf(limy, fmaxf(-limy, tyz)) * tz;
    float tz2 = tz * tz;
    float j00 = cam->fx / tz;
    float j02 = -(cam->fx * ctx) / tz2;
    float j11 = cam->fy / tz;
    float j12 = -(cam->fy * cty) / tz2;
    /* T = J * R  (2x3) */
    float T00 = fmaf(j02, R[6], j00 * R[0]);
    float T01 = fmaf(j02, R[7], j00 * R[1]);
    float T02 = fmaf(j02, R[8], j00 * R[2]);
    float T10 = fmaf(j12, R[6], j11 * R[3]);
    float T11 = fmaf(j12, R[7], j11 * R[4]);
    float T12 = fmaf(j12, R[8], j11 * R[5]);
    /* v0 = Sigma * T0^T, v1 = Sigma * T1^T */
    float v00 = fmaf(S[2], T02, fmaf(S[1], T01, S[0] * T00));
    float v01 = fmaf(S[4], T02, fmaf(S[3], T01, S[1] * T00));
    float v02 = fmaf(S[5], T02, fmaf(S[4], T01, S[2] * T00));
    float v10 = fmaf(S[2], T12, fmaf(S[1], T11, S[0] * T10));
    float v11 = fmaf(S[4], T12, fmaf(S[3], T11, S[1] * T10));
    float v12 = fmaf(S[5], T12, fmaf(S[4], T11, S[2] * T10));
    float a = fmaf(T02, v02, fmaf(T01, v01, T00 * v00)) + SO_BLUR;
    float b = fmaf(T12, v02, fmaf(T11, v01, T10 * v00));
    float c = fmaf(T12, v12, fmaf(T11, v11, T10 * v10)) + SO_BLUR;
    float det = fmaf(a, c, -(b * b));
    if (!(det > 0.0f) || !(det < INFINITY)) return 0;
    float conic_a = c / det, conic_b = -b / det, conic_c = a / det;

    out->gx = fmaf(cam->fx, txz, cam->cx);
    out->gy = fmaf(cam->fy, tyz, cam->cy);
    out->qa = conic_a * (-0.5f * SO_LOG2E);
    out->qb = conic_b * (-SO_LOG2E);
    out->qc = conic_c * (-0.5f * SO_LOG2E);
    out->lo = lo;
    out->z = tz;
    out->cov_a = a;
    out->cov_c = c;
    float rgb[3];
    so_sh_color(deg, sh, mx - cam->campos[0], my - cam->campos[1], mz - cam->campos[2], rgb);
    out->r = rgb[0];
    out->g = rgb[1];
    out->b = rgb[2];
    /* non-finite anywhere => the Gaussian is dropped (NaN quaternion, overflow ...) */
    float chk = out->gx + out->gy + out->qa + out->qb + out->qc + out->r + out->g + out->b;
    if (!(chk - chk == 0.0f)) return 0;
    return 1;
}

/* Exported per-Gaussian projection (element-wise parity tests of the projection kernel). */
void so_project(const so_camera* cam, int64_t n, int deg, const float* xyz, const float* log_scale,
                const float* quat, const float* opacity_raw, const float* sh, so_splat* out,
                uint8_t* visible)
{
    int nb = (deg + 1) * (deg + 1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        so_splat s;
        memset(&s, 0, sizeof s);
        int v = so_project_one(cam, deg, xyz + 3 * i, log_scale + 3 * i, quat + 4 * i,
                               opacity_raw[i], sh + (size_t)3 * nb * i, &s);
        if (!v) memset(&s, 0, sizeof s);
        out[i] = s;
        visible[i] = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------ full frame */

typedef struct { uint32_t key; uint32_t idx; } so_kv;

static int so_kv_cmp(const void* a, const void* b)
{
    const so_kv* x = (const so_kv*)a;
    const so_kv* y = (const so_kv*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

/* One pixel, one splat: the blend step of the spec.  Returns 1 when the pixel saturates. */
static inline int so_blend(const so_splat* s, float pxc, float pyc, float* T, float* C, float* Dz)
{
    float dx = pxc - s->gx;
    float dy = pyc - s->gy;
    float bdy = s->qb * dy;
    float cdy = s->qc * dy;
    float cdy2 = fmaf(cdy, dy, s->lo);           /* opacity folded into the exponent */
    float w = fmaf(s->qa, dx, bdy);
    float power2 = fmaf(dx, w, cdy2);            /* log2(alpha) before the cap */
    if (!(power2 >= SO_L255)) return 0;          /* alpha < 1/255 */
    float alpha = so_exp2_fast(fminf(power2, SO_L099));
    float Tn = *T * (1.0f - alpha);
    if (Tn < SO_T_MIN) return 1;
    float wgt = alpha * *T;
    C[0] = fmaf(s->r, wgt, C[0]);
    C[1] = fmaf(s->g, wgt, C[1]);
    C[2] = fmaf(s->b, wgt, C[2]);
    *Dz = fmaf(s->z, wgt, *Dz);
    *T = Tn;
    return 0;
}

static inline void so_finish(float T, const float* C, float Dz, uint8_t* rgb, float* depth, float* finalT)
{
    for (int c = 0; c < 3; ++c) {
        float v = fminf(fmaxf(C[c], 0.0f), 1.0f);
        rgb[c] = (uint8_t)(v * 255.0f + 0.5f);
    }
    float cover = 1.0f - T;
    if (depth) *depth = cover > 1e-6f ? Dz / cover : INFINITY;   /* nothing hit: Isaac's distance_to_image_plane is inf there */
    if (finalT) *finalT = T;
}

/*
 * Render one frame.  brute != 0: every pixel walks every visible Gaussian in depth order
 * (no binning at all) — the definition of the spec.  brute == 0: per-tile lists from a
 * generously conservative bounding box, mathematically identical, fast enough for 1080p/2M.
 * stats (optional, int64[3]): visible count V, tile-instances D (oracle's own binning), threads.
 * Returns 0 on success.
 */
int so_render(const so_camera* cam, int64_t n, int deg, const float* xyz, const float* log_scale,
              const float* quat, const float* opacity_raw, const float* sh, int brute,
              uint8_t* rgb_out, float* depth_out, float* finalT_out, int64_t* stats)
{
    const int W = cam->W, H = cam->H;
    if (W <= 0 || H <= 0 || n < 0) return 1;
    int nb = (deg + 1) * (deg + 1);
    so_splat* sp = (so_splat*)malloc((size_t)(n > 0 ? n : 1) * sizeof(so_splat));
    uint8_t* vis = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    if (!sp || !vis) { free(sp); free(vis); return 2; }

#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        vis[i] = (uint8_t)so_project_one(cam, deg, xyz + 3 * i, log_scale + 3 * i, quat + 4 * i,
                                         opacity_raw[i], sh + (size_t)3 * nb * i, &sp[i]);

    int64_t V = 0;
    for (int64_t i = 0; i < n; ++i) V += vis[i];
    so_kv* order = (so_kv*)malloc((size_t)(V > 0 ? V : 1) * sizeof(so_kv));
    if (!order) { free(sp); free(vis); return 2; }
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i)
        if (vis[i]) { order[k].key = so_f2u(sp[i].z); order[k].idx = (uint32_t)i; ++k; }
    qsort(order, (size_t)V, sizeof(so_kv), so_kv_cmp);

    int64_t D = 0;
    if (brute) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float T = 1.0f, C[3] = {0, 0, 0}, Dz = 0.0f;
                float pxc = (float)x + 0.5f, pyc = (float)y + 0.5f;
                for (int64_t r = 0; r < V; ++r)
                    if (so_blend(&sp[order[r].idx], pxc, pyc, &T, C, &Dz)) break;
                size_t p = (size_t)y * W + x;
                so_finish(T, C, Dz, rgb_out + 3 * p, depth_out ? depth_out + p : 0,
                          finalT_out ? finalT_out + p : 0);
            }
    } else {
        const int TX = (W + SO_TILE - 1) / SO_TILE, TY = (H + SO_TILE - 1) / SO_TILE;
        const int64_t NT = (int64_t)TX * TY;
        int64_t* cnt = (int64_t*)calloc((size_t)NT + 1, sizeof(int64_t));
        int32_t* rect = (int32_t*)malloc((size_t)(V > 0 ? V : 1) * 4 * sizeof(int32_t));
        if (!cnt || !rect) { free(cnt); free(rect); free(order); free(sp); free(vis); return 2; }
        /* conservative pixel bbox in double: alpha >= 1/255 needs d^T conic d <= 2 ln(255 o);
           1 % + 1 px of slack swallows the binary32 rounding of the blend path for ordinary splats.  For needle-like
           splats it does not: the exponent's terms (|qa| dx^2, |qb dx dy|, |qc| dy^2) reach 1e5 - 1e6 where their
           sum is ~ -8, so the ROUNDED exponent of the spec passes the threshold up to ~10 % beyond the exact ellipse
           along the needle (found by tests/test_render_parity_gpu.py::test_needle_splats..., where this mode dropped a
           contribution the brute-force mode - the definition - and the GPU both had).  Hence a second box from the
           rounded conic deflated by that rounding allowance, eps (|term| sum) with |dx dy| <= (dx^2 + dy^2) / 2:
           A' = A - eps (A + |B|), C' = C - eps (C + |B|); unbounded when the deflated form is no longer definite. */
        for (int64_t r = 0; r < V; ++r) {
            const so_splat* s = &sp[order[r].idx];
            double tau = 2.0 * (((double)s->lo - (double)SO_L255) / 1.4426950408889634) * 1.01 + 0.01;
            if (tau < 0.01) tau = 0.01;
            double hx = sqrt(tau * (double)s->cov_a) + 1.0, hy = sqrt(tau * (double)s->cov_c) + 1.0;
            {
                const double eps = 1e-6;
                double A = -(double)s->qa, B = -0.5 * (double)s->qb, C = -(double)s->qc;
                const double t2 = ((double)s->lo - (double)SO_L255) * 1.01 + 0.02;
                A -= eps * (A + fabs(B)); C -= eps * (C + fabs(B));
                const double det = A * C - B * B;
                if (!(det > 0.0) || !(A > 0.0) || !(C > 0.0)) { hx = 1e30; hy = 1e30; }
                else if (t2 > 0.0) {
                    const double hx2 = sqrt(t2 * C / det) * 1.001 + 1.0, hy2 = sqrt(t2 * A / det) * 1.001 + 1.0;
                    if (!(hx2 <= hx)) hx = hx2;
                    if (!(hy2 <= hy)) hy = hy2;
                }
            }
            double x0 = floor(((double)s->gx - hx) / SO_TILE), x1 = floor(((double)s->gx + hx) / SO_TILE);
            double y0 = floor(((double)s->gy - hy) / SO_TILE), y1 = floor(((double)s->gy + hy) / SO_TILE);
            int ix0 = x0 < 0 ? 0 : (x0 > TX ? TX : (int)x0), ix1 = x1 < -1 ? -1 : (x1 >= TX ? TX - 1 : (int)x1);
            int iy0 = y0 < 0 ? 0 : (y0 > TY ? TY : (int)y0), iy1 = y1 < -1 ? -1 : (y1 >= TY ? TY - 1 : (int)y1);
            rect[4 * r + 0] = ix0; rect[4 * r + 1] = iy0; rect[4 * r + 2] = ix1; rect[4 * r + 3] = iy1;
            for (int ty = iy0; ty <= iy1; ++ty)
                for (int tx = ix0; tx <= ix1; ++tx) cnt[(int64_t)ty * TX + tx + 1]++;
        }
        for (int64_t t = 0; t < NT; ++t) cnt[t + 1] += cnt[t];
        D = cnt[NT];
        uint32_t* list = (uint32_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(uint32_t));
        int64_t* fill = (int64_t*)malloc((size_t)(NT > 0 ? NT : 1) * sizeof(int64_t));
        if (!list || !fill) { free(list); free(fill); free(cnt); free(rect); free(order); free(sp); free(vis); return 2; }
        memcpy(fill, cnt, (size_t)NT * sizeof(int64_t));
        for (int64_t r = 0; r < V; ++r) /* depth order preserved inside every tile list */
            for (int ty = rect[4 * r + 1]; ty <= rect[4 * r + 3]; ++ty)
                for (int tx = rect[4 * r + 0]; tx <= rect[4 * r + 2]; ++tx)
                    list[fill[(int64_t)ty * TX + tx]++] = order[r].idx;
#pragma omp parallel for schedule(dynamic, 1)
        for (int64_t t = 0; t < NT; ++t) {
            int tx = (int)(t % TX), ty = (int)(t / TX);
            for (int yy = 0; yy < SO_TILE; ++yy)
                for (int xx = 0; xx < SO_TILE; ++xx) {
                    int x = tx * SO_TILE + xx, y = ty * SO_TILE + yy;
                    if (x >= W || y >= H) continue;
                    float T = 1.0f, C[3] = {0, 0, 0}, Dz = 0.0f;
                    float pxc = (float)x + 0.5f, pyc = (float)y + 0.5f;
                    for (int64_t j = cnt[t]; j < cnt[t + 1]; ++j)
                        if (so_blend(&sp[list[j]], pxc, pyc, &T, C, &Dz)) break;
                    size_t p = (size_t)y * W + x;
                    so_finish(T, C, Dz, rgb_out + 3 * p, depth_out ? depth_out + p : 0,
                              finalT_out ? finalT_out + p : 0);
                }
        }
        free(list); free(fill); free(cnt); free(rect);
    }
    if (stats) {
        stats[0] = V;
        stats[1] = D;
#ifdef _OPENMP
        stats[2] = omp_get_max_threads();
#else
        stats[2] = 1;
#endif
    }
    free(order); free(sp); free(vis);
    return 0;
}

void so_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ 2-D grid collide */

/*
 * check_collision_3d for one position whose mapped 2-D coordinate (px2d, py2d) is already a
 * double (collision_detector.py:204-234).  Python's round() is round-half-to-even == nearbyint()
 * under the default rounding mode; round(nan)/round(inf) raise -> the reference's blanket
 * `except` returns False (:235-237).
 */
static uint8_t so_grid_lookup(const uint8_t* map, int H, int W, double min_x, double min_y,
                              double scale, double px2d, double py2d)
{
    double fx = (px2d - min_x) / scale;
    double fy = (py2d - min_y) / scale;
    if (!(fx - fx == 0.0) || !(fy - fy == 0.0)) return 0; /* nan / inf -> exception -> False */
    double rx = nearbyint(fx), ry = nearbyint(fy);
    /* clamp far-out values before the integer conversion; anything beyond the 2-px margin is a hit */
    if (rx < -4.0 || rx > (double)W + 4.0 || ry < -4.0 || ry > (double)H + 4.0) return 1;
    long ix = (long)rx, iy = (long)ry;
    if (!(0 <= iy && iy < H && 0 <= ix && ix < W)) {
        const long margin = 2;
        if (-margin <= iy && iy < H + margin && -margin <= ix && ix < W + margin) {
            iy = iy < 0 ? 0 : (iy > H - 1 ? H - 1 : iy);
            ix = ix < 0 ? 0 : (ix > W - 1 ? W - 1 : ix);
        } else {
            return 1;
        }
    }
    return map[(size_t)iy * W + ix] == 1 ? 1 : 0;
}

/*
 * pos: float32 [P,3].  np2_promote != 0 reproduces NumPy >= 2 (NEP 50) scalar promotion:
 * `(min+max) - np.float32` is evaluated in binary32 (collision_detector.py:184-187 with the
 * env's float32 position, simple_env.py:1156).  np2_promote == 0 reproduces NumPy 1.26 (the
 * reference's pinned Isaac environment): the same expression is evaluated in binary64.
 */
void so_collide_grid_f32(const uint8_t* map, int H, int W, double min_x, double max_x, double min_y,
                         double max_y, double scale, const float* pos, int P, int np2_promote,
                         uint8_t* hit)
{
    double sx = min_x + max_x, sy = min_y + max_y;
    for (int i = 0; i < P; ++i) {
        float x = pos[3 * i + 0], y = pos[3 * i + 1];
        double px2d, py2d;
        if (np2_promote) {
            float nx = -x, ny = -y;
            px2d = (double)((float)sx - nx);
            py2d = (double)((float)sy - ny);
        } else {
            px2d = sx - (double)(-x);
            py2d = sy - (double)(-y);
        }
        hit[i] = map ? so_grid_lookup(map, H, W, min_x, min_y, scale, px2d, py2d) : 0;
    }
}

/* pos: float64 [P,3] (e.g. check_collision_at_position, collision_detector.py:263-272). */
void so_collide_grid_f64(const uint8_t* map, int H, int W, double min_x, double max_x, double min_y,
                         double max_y, double scale, const double* pos, int P, uint8_t* hit)
{
    double sx = min_x + max_x, sy = min_y + max_y;
    for (int i = 0; i < P; ++i) {
        double px2d = sx - (-pos[3 * i + 0]);
        double py2d = sy - (-pos[3 * i + 1]);
        hit[i] = map ? so_grid_lookup(map, H, W, min_x, min_y, scale, px2d, py2d) : 0;
    }
}

/* ------------------------------------------------------------------ capsule collide */

/*
 * Agent capsule = vertical segment (x, y, z0)-(x, y, z1) swept by a sphere of `radius`
 * (agent body: simple_env.py:238,754-755).  Body j = ellipsoid with centre mu, rotation from the
 * normalised (w,x,y,z) quaternion and semi-axes exp(log_scale).  Spec: inflate
 * every semi-axis by `radius`, map the segment into the inflated ellipsoid's unit-sphere frame,
 * minimise the squared norm over the segment in closed form; hit <=> min <= 1.
 * hit[p] = any body hit; first_body[p] = smallest hit body index or -1.
 */
void so_collide_capsule(int64_t m, const float* mu, const float* log_scale, const float* quat,
                        const float* pos, int P, float radius, float z0, float z1, uint8_t* hit,
                        int32_t* first_body)
{
    for (int p = 0; p < P; ++p) { hit[p] = 0; if (first_body) first_body[p] = -1; }
    float dzs = z1 - z0;
    for (int64_t j = 0; j < m; ++j) {
        const float* q = quat + 4 * j;
        float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
        float qn = sqrtf(fmaf(qz, qz, fmaf(qy, qy, fmaf(qx, qx, qw * qw))));
        float w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
        float xx = x * x, yy = y * y, zz = z * z;
        float xy = x * y, xz = x * z, yz = y * z;
        float wx = w * x, wy = w * y, wz = w * z;
        float r00 = 1.0f - 2.0f * (yy + zz), r01 = 2.0f * (xy - wz), r02 = 2.0f * (xz + wy);
        float r10 = 2.0f * (xy + wz), r11 = 1.0f - 2.0f * (xx + zz), r12 = 2.0f * (yz - wx);
        float r20 = 2.0f * (xz - wy), r21 = 2.0f * (yz + wx), r22 = 1.0f - 2.0f * (xx + yy);
        float e0 = so_exp(log_scale[3 * j + 0]) + radius;
        float e1 = so_exp(log_scale[3 * j + 1]) + radius;
        float e2 = so_exp(log_scale[3 * j + 2]) + radius;
        /* segment direction in the unit-sphere frame: (R^T * (0,0,dz)) / e */
        float b0 = (r20 * dzs) / e0, b1 = (r21 * dzs) / e1, b2 = (r22 * dzs) / e2;
        float bb = fmaf(b2, b2, fmaf(b1, b1, b0 * b0));
        for (int p = 0; p < P; ++p) {
            float d0 = pos[3 * p + 0] - mu[3 * j + 0];
            float d1 = pos[3 * p + 1] - mu[3 * j + 1];
            float d2 = z0 - mu[3 * j + 2];
            float a0 = fmaf(r20, d2, fmaf(r10, d1, r00 * d0)) / e0;
            float a1 = fmaf(r21, d2, fmaf(r11, d1, r01 * d0)) / e1;
            float a2 = fmaf(r22, d2, fmaf(r12, d1, r02 * d0)) / e2;
            float ab = fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
            float t = bb > 0.0f ? fminf(1.0f, fmaxf(0.0f, -ab / bb)) : 0.0f;
            float m0 = fmaf(t, b0, a0), m1 = fmaf(t, b1, a1), m2 = fmaf(t, b2, a2);
            float mm = fmaf(m2, m2, fmaf(m1, m1, m0 * m0));
            if (mm <= 1.0f) {
                hit[p] = 1;
                if (first_body && first_body[p] < 0) first_body[p] = (int32_t)j;
            }
        }
    }
}

int so_abi_version(void) { return 2; }
