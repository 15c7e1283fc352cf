// K6 (2-D occupancy-grid collide) and K7 (ellipsoid-vs-capsule collide).
//
// K6 restates SemanticMap2DCollisionDetector.check_collision_3d (collision_detector.py:191-237,
// with forward_position_mapping :146-189 and _world_to_pixel :134-138) for a batch of positions.
// K7 is the 3-D agent-capsule test the reference only gestures at through PhysX
// (simple_env.py:1878-1931; agent body :238,754-755).
#include "sage_internal.cuh"

namespace {

// One lookup; (px2d, py2d) already mapped to the 2-D map frame in binary64.
__device__ __forceinline__ uint8_t grid_lookup(const uint8_t* __restrict__ map, int H, int W,
                                               double min_x, double min_y, double scale,
                                               double px2d, double py2d)
{
    double fx = (px2d - min_x) / scale;
    double fy = (py2d - min_y) / scale;
    if (!(fx - fx == 0.0) || !(fy - fy == 0.0)) return 0;   // round(nan|inf) raises -> except -> False
    double rx = rint(fx), ry = rint(fy);                    // Python round(): half-to-even
    if (rx < -4.0 || rx > (double)W + 4.0 || ry < -4.0 || ry > (double)H + 4.0) return 1;
    long long ix = (long long)rx, iy = (long long)ry;
    if (!(0 <= iy && iy < H && 0 <= ix && ix < W)) {
        const long long margin = 2;                          // collision_detector.py:213-221
        if (-margin <= iy && iy < H + margin && -margin <= ix && ix < W + margin) {
            iy = iy < 0 ? 0 : (iy > H - 1 ? H - 1 : iy);
            ix = ix < 0 ? 0 : (ix > W - 1 ? W - 1 : ix);
        } else {
            return 1;                                        // :223-225
        }
    }
    return map[(size_t)iy * W + ix] == 1 ? 1 : 0;
}

__global__ void k_collide_grid(const uint8_t* __restrict__ map, int H, int W, double min_x,
                               double max_x, double min_y, double max_y, double scale,
                               const float* __restrict__ pos32, const double* __restrict__ pos64,
                               int P, int np2_promote, uint8_t* __restrict__ hit)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const double sx = min_x + max_x, sy = min_y + max_y;
    double px2d, py2d;
    if (pos64) {
        px2d = sx - (-pos64[3 * i + 0]);
        py2d = sy - (-pos64[3 * i + 1]);
    } else {
        float x = pos32[3 * i + 0], y = pos32[3 * i + 1];
        if (np2_promote) {   // NumPy >= 2: python-float - np.float32 stays binary32
            float nx = -x, ny = -y;
            px2d = (double)((float)sx - nx);
            py2d = (double)((float)sy - ny);
        } else {             // NumPy 1.26: promoted to binary64
            px2d = sx - (double)(-x);
            py2d = sy - (double)(-y);
        }
    }
    hit[i] = map ? grid_lookup(map, H, W, min_x, min_y, scale, px2d, py2d) : 0;
}

// first colliding sample of each consecutive segment (the march of _try_direct_movement /
// _try_short_movement, simple_env.py:2116-2234, evaluated in one launch)
__global__ void k_first_hit(const uint8_t* __restrict__ hit, const int32_t* __restrict__ seg_len,
                            int n_segments, int32_t* __restrict__ first_hit)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segments) return;
    int start = 0;
    for (int k = 0; k < s; ++k) start += seg_len[k];
    int32_t f = -1;
    for (int k = 0; k < seg_len[s]; ++k)
        if (hit[start + k]) { f = k; break; }
    first_hit[s] = f;
}

// ---- capsule ---------------------------------------------------------------------------------

__global__ void k_bodies_pack(int64_t m, const float* __restrict__ mu, const float* __restrict__ ls,
                              const float* __restrict__ q, float4* __restrict__ c0,
                              float4* __restrict__ c1, float2* __restrict__ c2)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    c0[j] = make_float4(mu[3 * j], mu[3 * j + 1], mu[3 * j + 2], ls[3 * j]);
    c1[j] = make_float4(ls[3 * j + 1], ls[3 * j + 2], q[4 * j], q[4 * j + 1]);
    c2[j] = make_float2(q[4 * j + 2], q[4 * j + 3]);
}

constexpr int CAP_THREADS = 256;
constexpr int CAP_MAX_P = 256;   // poses per launch chunk held in shared memory

__global__ void __launch_bounds__(CAP_THREADS)
k_collide_capsule(int64_t m, const float4* __restrict__ c0, const float4* __restrict__ c1,
                  const float2* __restrict__ c2, const float* __restrict__ pos, int P, float radius,
                  float z0, float z1, uint32_t* __restrict__ hit_bits, int32_t* __restrict__ first_body)
{
    __shared__ float spx[CAP_MAX_P], spy[CAP_MAX_P];
    for (int p = threadIdx.x; p < P; p += CAP_THREADS) {
        spx[p] = pos[3 * p + 0];
        spy[p] = pos[3 * p + 1];
    }
    __syncthreads();
    int64_t j = (int64_t)blockIdx.x * CAP_THREADS + threadIdx.x;
    if (j >= m) return;
    const float4 a0 = __ldg(&c0[j]);
    const float4 a1 = __ldg(&c1[j]);
    const float2 a2 = __ldg(&c2[j]);
    const float mux = a0.x, muy = a0.y, muz = a0.z;
    const float qw = a1.z, qx = a1.w, qy = a2.x, qz = a2.y;
    float qn = sqrtf(fmaf(qz, qz, fmaf(qy, qy, fmaf(qx, qx, qw * qw))));
    float w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
    float xx = x * x, yy = y * y, zz = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    float r00 = 1.0f - 2.0f * (yy + zz), r01 = 2.0f * (xy - wz), r02 = 2.0f * (xz + wy);
    float r10 = 2.0f * (xy + wz), r11 = 1.0f - 2.0f * (xx + zz), r12 = 2.0f * (yz - wx);
    float r20 = 2.0f * (xz - wy), r21 = 2.0f * (yz + wx), r22 = 1.0f - 2.0f * (xx + yy);
    float e0 = sage_exp(a0.w) + radius;
    float e1 = sage_exp(a1.x) + radius;
    float e2 = sage_exp(a1.y) + radius;
    float dzs = z1 - z0;
    float b0 = (r20 * dzs) / e0, b1 = (r21 * dzs) / e1, b2 = (r22 * dzs) / e2;
    float bb = fmaf(b2, b2, fmaf(b1, b1, b0 * b0));
    float d2 = z0 - muz;
    for (int p = 0; p < P; ++p) {
        float d0 = spx[p] - mux;
        float d1 = spy[p] - muy;
        float q0 = fmaf(r20, d2, fmaf(r10, d1, r00 * d0)) / e0;
        float q1 = fmaf(r21, d2, fmaf(r11, d1, r01 * d0)) / e1;
        float q2 = fmaf(r22, d2, fmaf(r12, d1, r02 * d0)) / e2;
        float ab = fmaf(q2, b2, fmaf(q1, b1, q0 * b0));
        float t = bb > 0.0f ? fminf(1.0f, fmaxf(0.0f, -ab / bb)) : 0.0f;
        float m0 = fmaf(t, b0, q0), m1 = fmaf(t, b1, q1), m2 = fmaf(t, b2, q2);
        float mm = fmaf(m2, m2, fmaf(m1, m1, m0 * m0));
        if (mm <= 1.0f) {
            atomicOr(&hit_bits[p >> 5], 1u << (p & 31));
            if (first_body) atomicMin(&first_body[p], (int32_t)j);
        }
    }
}

}  // namespace

void sage_launch_collide_grid(const uint8_t* map, int H, int W, double min_x, double max_x,
                              double min_y, double max_y, double scale, const float* pos_f32,
                              const double* pos_f64, int P, int np2_promote, uint8_t* hit,
                              cudaStream_t st)
{
    if (P <= 0) return;
    k_collide_grid<<<(P + 63) / 64, 64, 0, st>>>(map, H, W, min_x, max_x, min_y, max_y, scale,
                                                pos_f32, pos_f64, P, np2_promote, hit);
}

void sage_launch_first_hit(const uint8_t* hit, const int32_t* seg_len, int n_segments,
                           int32_t* first_hit, cudaStream_t st)
{
    if (n_segments <= 0) return;
    k_first_hit<<<(n_segments + 31) / 32, 32, 0, st>>>(hit, seg_len, n_segments, first_hit);
}

void sage_launch_bodies_pack(int64_t m, const float* mu, const float* log_scale, const float* quat,
                             float4* c0, float4* c1, float2* c2, cudaStream_t st)
{
    if (m <= 0) return;
    k_bodies_pack<<<(int)((m + 255) / 256), 256, 0, st>>>(m, mu, log_scale, quat, c0, c1, c2);
}

int sage_capsule_max_poses(void) { return CAP_MAX_P; }

void sage_launch_collide_capsule(int64_t m, const float4* c0, const float4* c1, const float2* c2,
                                 const float* pos, int P, float radius, float z0, float z1,
                                 uint32_t* hit_bits, int32_t* first_body, cudaStream_t st)
{
    if (m <= 0 || P <= 0) return;
    k_collide_capsule<<<(int)((m + CAP_THREADS - 1) / CAP_THREADS), CAP_THREADS, 0, st>>>(
        m, c0, c1, c2, pos, P, radius, z0, z1, hit_bits, first_body);
}
