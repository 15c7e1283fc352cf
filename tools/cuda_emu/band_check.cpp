// Development check of the exact tile-culling helpers in csrc/sage_internal.cuh (no GPU):
// for every splat, brute-force the pixel centres with alpha >= 1/255 (binary64 from the binary32
// conic, NO slack) and verify that each lies in an emitted tile with its 8x8-quadrant bit set.
#define SAGE_CUDA_EMU 1
#include "../../sage-3d_official_b200/csrc/sage_internal.cuh"

extern "C" int band_check(const float* rec /* n x 6: gx gy qa qb qc o */, long n, int W, int H, long* out /* 4: tiles_bbox, tiles_exact, quads_exact, violations */)
{
    long tb = 0, te = 0, qe = 0, bad = 0;
    for (long i = 0; i < n; ++i) {
        const float* r = rec + 6 * i;
        double bhx = 0, bhy = 0;
        const SageBand band = sage_band_make(r[0], r[1], r[2], r[3], r[4], r[5]);
        { const double t2d = (double)log2f(255.0f * r[5]) * 1.002 + 0.02, Ad = -(double)r[2], Bd = -0.5 * (double)r[3], Cd = -(double)r[4], dd = Ad * Cd - Bd * Bd; if (!(dd > 0 && Ad > 0 && Cd > 0)) continue; bhx = sqrt(t2d * Cd / dd); bhy = sqrt(t2d * Ad / dd); }
        double xlo = floor((double)r[0] - bhx - 0.5), xhi = ceil((double)r[0] + bhx - 0.5);
        double ylo = floor((double)r[1] - bhy - 0.5), yhi = ceil((double)r[1] + bhy - 0.5);
        if (!(xhi >= 0 && xlo <= W - 1 && yhi >= 0 && ylo <= H - 1)) continue;
        int ix0 = (int)fmax(xlo, 0.0), ix1 = (int)fmin(xhi, W - 1.0), iy0 = (int)fmax(ylo, 0.0), iy1 = (int)fmin(yhi, H - 1.0);
        if (ix0 > ix1 || iy0 > iy1) continue;
        tb += (long)(ix1 / 16 - ix0 / 16 + 1) * (iy1 / 16 - iy0 / 16 + 1);
        const double A = -(double)r[2], B = -0.5 * (double)r[3], C = -(double)r[4], t = log2(255.0 * (double)r[5]);
        for (int ty = iy0 / 16; ty <= iy1 / 16; ++ty) {
            const SageRowItem it = sage_band_row(band, ix0, ix1, iy0, iy1, ty);
            te += it.tx1 >= it.tx0 ? it.tx1 - it.tx0 + 1 : 0;
            for (int tx = it.tx0; tx <= it.tx1; ++tx) qe += __builtin_popcount(sage_row_quad_mask(it, tx));
            for (int y = (ty * 16 > iy0 ? ty * 16 : iy0); y <= (ty * 16 + 15 < iy1 ? ty * 16 + 15 : iy1); ++y)
                for (int x = 0; x < W; ++x) {
                    const double dx = x + 0.5 - (double)r[0], dy = y + 0.5 - (double)r[1];
                    if (A * dx * dx + 2 * B * dx * dy + C * dy * dy <= t) {
                        const int tx = x / 16;
                        const bool in_row = tx >= it.tx0 && tx <= it.tx1;
                        const unsigned bit = ((y & 8) ? 4u : 1u) << ((x & 8) ? 1 : 0);
                        if (!in_row || !(sage_row_quad_mask(it, tx) & bit)) ++bad;
                    }
                }
        }
    }
    out[0] = tb; out[1] = te; out[2] = qe; out[3] = bad;
    return 0;
}
