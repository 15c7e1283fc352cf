// Development check of the culling helpers of csrc/sage_internal.cuh and of K1a's pixel box (no GPU): for every splat,
// brute-force EVERY pixel of the image with the blend's own binary32 exponent (the fmaf sequence of composite.cu /
// sage_oracle.c, not an exact formula) and verify that each pixel which passes the alpha >= 1/255 test lies inside the
// pixel box K1a would write, in a tile K2 would emit, with its 8 x 8-quadrant bit set.  "Conservative" has to hold against
// the ROUNDED exponent: for needle-like splats it passes beyond the exact ellipse (sage_conic_deflate).
// Build: g++ -std=c++20 -O2 -DSAGE_CUDA_EMU -fPIC -shared -pthread -w -ffp-contract=off -mfma -include tools/cuda_emu/cuda_emu.h
#define SAGE_CUDA_EMU 1
#include "../../sage-3d_official_b200/csrc/sage_internal.cuh"

// K1a's conservative pixel box (mirrors project.cu, "Conservative bounding rectangle of {alpha >= 1/255}"); false = culled
static bool pixel_box(float gx, float gy, float qa, float qb, float qc, float lo, int W, int H, int* ix0, int* ix1, int* iy0, int* iy1)
{
    const float l2 = lo - SAGE_L255;
    const float t2 = l2 * 1.003f + 0.02f;
    float A = -qa, Cq = -qc;
    const float Bq = -0.5f * qb;
    sage_conic_deflate(&A, Bq, &Cq);
    const float det2 = sage_det_kahan(A, Cq, Bq);
    const bool unbounded = !(det2 > 0.0f) || !(A > 0.0f) || !(Cq > 0.0f);
    float hx = unbounded ? 1e30f : sqrtf(t2 * Cq / det2);
    float hy = unbounded ? 1e30f : sqrtf(t2 * A / det2);
    if (!(hx < 1e9f)) hx = 1e30f;
    if (!(hy < 1e9f)) hy = 1e30f;
    const float slack = 1e-3f * fmaxf(hx, hy) + 0.01f;
    const float xlo = floorf((gx - hx) - slack - 0.5f), xhi = ceilf((gx + hx) + slack - 0.5f);
    const float ylo = floorf((gy - hy) - slack - 0.5f), yhi = ceilf((gy + hy) + slack - 0.5f);
    const float Wm = (float)(W - 1), Hm = (float)(H - 1);
    if (!(xhi >= 0.0f && xlo <= Wm && yhi >= 0.0f && ylo <= Hm)) return false;
    *ix0 = (int)fmaxf(xlo, 0.0f); *ix1 = (int)fminf(xhi, Wm);
    *iy0 = (int)fmaxf(ylo, 0.0f); *iy1 = (int)fminf(yhi, Hm);
    return *iy0 <= *iy1 && *ix0 <= *ix1;
}

// rec: n x 6 (gx gy qa qb qc lo).  out[0] tiles of the boxes, [1] tiles emitted, [2] quadrant bits emitted, [3] pixels that pass,
// [4] VIOLATIONS (a passing pixel outside the box / in a tile not emitted / with its quadrant bit clear), [5] splats culled by the box
extern "C" int band_check(const float* rec, long n, int W, int H, long* out)
{
    long tb = 0, te = 0, qe = 0, pass = 0, bad = 0, culled = 0;
    const int TX = (W + SAGE_BIN_W - 1) / SAGE_BIN_W, TY = (H + SAGE_BIN_H - 1) / SAGE_BIN_H;
    std::vector<uint8_t> mask((size_t)TX * TY);
    for (long i = 0; i < n; ++i) {
        const float gx = rec[6 * i], gy = rec[6 * i + 1], qa = rec[6 * i + 2], qb = rec[6 * i + 3], qc = rec[6 * i + 4], lo = rec[6 * i + 5];
        std::fill(mask.begin(), mask.end(), 0);
        int px0 = 0, px1 = -1, py0 = 0, py1 = -1;
        const bool vis = pixel_box(gx, gy, qa, qb, qc, lo, W, H, &px0, &px1, &py0, &py1);
        if (!vis) ++culled;
        if (vis) {                                                            // K2: the tiles and quadrant masks it emits
            tb += (long)(px1 / SAGE_BIN_W - px0 / SAGE_BIN_W + 1) * (py1 / SAGE_BIN_H - py0 / SAGE_BIN_H + 1);
            if (sage_box_cullable(px0, px1, py0, py1)) {
                const SageBand band = sage_band_make(gx, gy, qa, qb, qc, lo - SAGE_L255);
                for (int ty = py0 / SAGE_BIN_H; ty <= py1 / SAGE_BIN_H; ++ty) {
                    const SageRowItem it = sage_band_row(band, px0, px1, py0, py1, ty);
                    for (int tx = it.tx0; tx <= it.tx1; ++tx) mask[(size_t)ty * TX + tx] = (uint8_t)sage_row_quad_mask(it.top, it.bot, tx);
                }
            } else {
                for (int ty = py0 / SAGE_BIN_H; ty <= py1 / SAGE_BIN_H; ++ty)
                    for (int tx = px0 / SAGE_BIN_W; tx <= px1 / SAGE_BIN_W; ++tx)
                        mask[(size_t)ty * TX + tx] = (uint8_t)sage_box_quad_mask(px0, px1, py0, py1, tx, ty);
            }
            for (size_t t = 0; t < mask.size(); ++t) if (mask[t]) { ++te; qe += __builtin_popcount(mask[t]); }
        }
        for (int y = 0; y < H; ++y) {                                         // the blend's exponent at every pixel centre
            const float pyc = (float)y + 0.5f;
            const float dy = pyc - gy;
            const float t2x = qb * dy, t2y = qc * dy;
            const float c2 = fmaf(t2y, dy, lo);
            for (int x = 0; x < W; ++x) {
                const float dx = ((float)x + 0.5f) - gx;
                const float wv = fmaf(qa, dx, t2x);
                const float pw = fmaf(dx, wv, c2);
                if (!(pw >= SAGE_L255)) continue;
                ++pass;
                const int tx = x / SAGE_BIN_W, ty = y / SAGE_BIN_H;
                const unsigned bit = 1u << (((y % SAGE_BIN_H) / 8) * SAGE_QUADS_X + (x % SAGE_BIN_W) / 8);
                if (!vis || x < px0 || x > px1 || y < py0 || y > py1 || !(mask[(size_t)ty * TX + tx] & bit)) ++bad;
            }
        }
    }
    out[0] = tb; out[1] = te; out[2] = qe; out[3] = pass; out[4] = bad; out[5] = culled;
    return 0;
}
