// Frame hand-off: baseline JPEG encoding of rendered frames on the GPU, so
// that what leaves the device is the ~0.3 MB byte stream the reference sends to its VLM / writes
// to disk, not the 6.2 MB RGB frame.
//
// The reference encodes on the host with Pillow defaults (quality 75, 4:2:0, Annex-K Huffman
// tables, no restart markers):
//     Image.fromarray(img_array).save(traj_dir / filename)   generate_images.py:431-432
//     pil_image.save(buffer, format='JPEG')                   vlm_client_modular.py:351,402,453
// The byte stream produced here is IDENTICAL to Pillow/libjpeg-turbo's for the same pixels
// (integer arithmetic end to end: 16-bit fixed-point colour conversion, 2x2 box chroma with the
// 1,2 bias, the "islow" 13-bit integer DCT, round-half-away quantisation, canonical Huffman
// codes), which is what lets the drop-in hand the caller pre-encoded bytes.
//
// One encode of B frames = 1 memset + 6 launches, no host synchronisation:
//   J1 k_jpeg_dct     warp / MCU (16x16 px): stage RGB, colour-convert, subsample, 96 1-D DCTs
//                     in shared memory, quantise, zig-zag -> coef[block][64] (int16) and the
//                     AC code length of every block
//   J2 k_jpeg_scan    CTA / frame: DC differences, exclusive scan of block bit lengths ->
//                     bit offset of every MCU, total bits, 1-fill of the last byte
//   J3 k_jpeg_huff    warp / MCU: canonical codes of 64 coefficients in parallel (ballot ->
//                     zero runs, warp scan -> bit positions), assembled in shared memory,
//                     whole words stored, boundary words OR-ed
//   J4 k_jpeg_ffcount / k_jpeg_ffscan / k_jpeg_stuff   0xFF byte stuffing as a two-level
//                     scan, header and EOI, final size
// HBM traffic per 1080p frame: 6.2 MB RGB read (L2-resident after the composite), 6.3 MB of
// coefficients written and read twice, < 1 MB of stream: latency-bound, not bandwidth-bound.
#include <cstring>
#include <vector>

#include "api_common.cuh"

namespace {

// ---- constant tables --------------------------------------------------------------------------
__constant__ uint8_t c_zigzag[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t h_zigzag[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ITU-T T.81 Annex K.1 quantisation tables (natural order) and K.3 Huffman specifications
const uint8_t h_qbase[2][64] = {
    {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};
// order: DC luma, AC luma, DC chroma, AC chroma (the order of the DHT segments in the file)
const uint8_t h_hbits[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
const uint8_t h_hclass[4] = {0x00, 0x10, 0x01, 0x11};
const uint8_t h_hv_dc[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t h_hv_ac_l[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t h_hv_ac_c[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

// Everything the kernels need that depends on (W, H, quality); lives in device memory.
struct JpegDev {
    uint16_t qdiv[2][64];      // quantiser * 8, natural order ([0] luma, [1] chroma)
    uint32_t dc_tab[2][12];    // len << 16 | code, indexed by the size category
    uint32_t ac_tab[2][256];   // len << 16 | code, indexed by run << 4 | size
    uint8_t header[640];       // SOI .. SOS
    int header_len;
    int W, H;
    int mcux, mcuy, nmcu, nblk;
    int y_wblk, y_hblk;        // real (non-dummy) luma blocks per row / column
    int ch_real;               // real chroma rows: ceil(H / 2)
    uint32_t raw_words;        // capacity of one frame's unstuffed scan, 32-bit words
    int n_chunks;              // raw capacity in 4096-byte chunks
};

struct JpegFrameInfo {         // per frame, written by J2 / J4b
    uint32_t total_bits;
    uint32_t raw_bytes;        // ceil(total_bits / 8)
    uint32_t overflow;         // 1: scan does not fit raw_words
    uint32_t ff_total;
};

#define JPEG_CHUNK 4096

// ---- shared arithmetic --------------------------------------------------------------------------
#define JFIX_0_298631336 2446
#define JFIX_0_390180644 3196
#define JFIX_0_541196100 4433
#define JFIX_0_765366865 6270
#define JFIX_0_899976223 7373
#define JFIX_1_175875602 9633
#define JFIX_1_501321110 12299
#define JFIX_1_847759065 15137
#define JFIX_1_961570560 16069
#define JFIX_2_053119869 16819
#define JFIX_2_562915447 20995
#define JFIX_3_072711026 25172

// One 8-point pass of libjpeg's jpeg_fdct_islow on p[0], p[S], ..., p[7S] (in place).
// PASS 0 = rows (results scaled up by 2^2), PASS 1 = columns (final descale).  All 32-bit:
// |input| <= 128*8*4 after pass 0, products stay below 2^31.
template <int S, int PASS>
__device__ __forceinline__ void fdct8(int32_t* p)
{
    const int32_t d0 = p[0], d1 = p[S], d2 = p[2 * S], d3 = p[3 * S], d4 = p[4 * S], d5 = p[5 * S], d6 = p[6 * S],
                  d7 = p[7 * S];
    int32_t t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6;
    int32_t t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    const int32_t t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int SH = PASS == 0 ? 13 - 2 : 13 + 2;
    constexpr int32_t RND = 1 << (SH - 1);
    if (PASS == 0) {
        p[0] = (t10 + t11) << 2;
        p[4 * S] = (t10 - t11) << 2;
    } else {
        p[0] = (t10 + t11 + 2) >> 2;
        p[4 * S] = (t10 - t11 + 2) >> 2;
    }
    int32_t z1 = (t12 + t13) * JFIX_0_541196100;
    p[2 * S] = (z1 + t13 * JFIX_0_765366865 + RND) >> SH;
    p[6 * S] = (z1 + t12 * (-JFIX_1_847759065) + RND) >> SH;
    z1 = t4 + t7;
    int32_t z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int32_t z5 = (z3 + z4) * JFIX_1_175875602;
    t4 *= JFIX_0_298631336;
    t5 *= JFIX_2_053119869;
    t6 *= JFIX_3_072711026;
    t7 *= JFIX_1_501321110;
    z1 *= -JFIX_0_899976223;
    z2 *= -JFIX_2_562915447;
    z3 *= -JFIX_1_961570560;
    z4 *= -JFIX_0_390180644;
    z3 += z5;
    z4 += z5;
    p[7 * S] = (t4 + z1 + z3 + RND) >> SH;
    p[5 * S] = (t5 + z2 + z4 + RND) >> SH;
    p[3 * S] = (t6 + z2 + z3 + RND) >> SH;
    p[S] = (t7 + z1 + z4 + RND) >> SH;
}

__device__ __forceinline__ int size_category(int a) { return 32 - __clz(a); }   // a >= 0

// The Huffman code of one zig-zag position of a block (jchuff.c encode_one_block).
// pos: 0..63; v: its quantised value (for pos 0: the DC DIFFERENCE); nz64: bit k set <=> AC
// position k is non-zero (bit 0 clear).  Result: up to three ZRL codes (zrl_n of them) followed
// by `bits` (len_b of them).  Total length = zrl_n * zrl_len + len_b.
struct HuffElem {
    uint32_t bits;
    int len_b;
    int zrl_n;
};

__device__ __forceinline__ HuffElem huff_elem(int pos, int v, unsigned long long nz64, const uint32_t* dc_tab,
                                              const uint32_t* ac_tab)
{
    HuffElem e;
    e.bits = 0; e.len_b = 0; e.zrl_n = 0;
    if (pos == 0) {
        const int a = v < 0 ? -v : v;
        const int nb = size_category(a);
        const uint32_t t = dc_tab[nb];
        const uint32_t vb = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
        e.bits = ((t & 0xFFFFu) << nb) | vb;
        e.len_b = (int)(t >> 16) + nb;
    } else if (v != 0) {
        const unsigned long long below = nz64 & ((1ull << pos) - 1ull);
        const int prev = below ? 63 - __clzll((long long)below) : 0;
        const int run = pos - prev - 1;
        const int a = v < 0 ? -v : v;
        const int nb = size_category(a);
        const uint32_t t = ac_tab[((run & 15) << 4) | nb];
        const uint32_t vb = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
        e.bits = ((t & 0xFFFFu) << nb) | vb;
        e.len_b = (int)(t >> 16) + nb;
        e.zrl_n = run >> 4;
    } else if (pos == 63) {
        const uint32_t t = ac_tab[0];        // EOB: the block ends in a run of zeros
        e.bits = t & 0xFFFFu;
        e.len_b = (int)(t >> 16);
    }
    return e;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// index of the block whose DC predicts block b = m*6+k (jchuff.c last_dc_val per component); -1: none
__device__ __forceinline__ int dc_pred_block(int m, int k)
{
    if (k >= 1 && k <= 3) return m * 6 + k - 1;
    if (m == 0) return -1;
    return (m - 1) * 6 + (k == 0 ? 3 : k);
}

// ---- J1: colour conversion, subsampling, DCT, quantisation -------------------------------------
constexpr int J1_WARPS = 4;

struct __align__(16) J1Warp {
    int16_t zz[6 * 64];        // quantised coefficients, zig-zag order (first: 16-byte aligned)
    int32_t blk[6 * 72];       // six 8x8 blocks, row stride 9 words (bank-conflict-free both ways)
    uint8_t rgb[16 * 48];      // the MCU's pixels
    uint8_t cb[256], cr[256];  // full-resolution chroma
};

__global__ void __launch_bounds__(J1_WARPS * 32) k_jpeg_dct(const JpegDev* __restrict__ jd,
                                                            const uint8_t* __restrict__ rgb_all, int bpp,
                                                            int16_t* __restrict__ coef_all,
                                                            uint16_t* __restrict__ acbits_all,
                                                            int16_t* __restrict__ dcs_all)
{
    __shared__ J1Warp s_warp[J1_WARPS];
    __shared__ uint16_t s_qdiv[2][64];
    __shared__ uint32_t s_ac[2][256];
    __shared__ uint8_t s_zz[64];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sage_dep_launch();
    for (int i = tid; i < 128; i += blockDim.x) s_qdiv[i >> 6][i & 63] = jd->qdiv[i >> 6][i & 63];
    for (int i = tid; i < 512; i += blockDim.x) s_ac[i >> 8][i & 255] = jd->ac_tab[i >> 8][i & 255];
    if (tid < 64) s_zz[tid] = c_zigzag[tid];
    __syncthreads();
    const int W = jd->W, H = jd->H;
    const int m = blockIdx.x * J1_WARPS + wid;
    if (m >= jd->nmcu) return;                       // warp-uniform; only __syncwarp below
    const int f = blockIdx.y;
    const int mx = m % jd->mcux, my = m / jd->mcux;
    const int px0 = mx * 16, py0 = my * 16;
    J1Warp& s = s_warp[wid];
    const uint8_t* src = rgb_all + (size_t)f * (size_t)W * (size_t)H * (size_t)bpp;      // bpp: 3 = RGB, 4 = RGBA (alpha ignored)

    // (1) stage the 16x16 pixels; rows/columns past the image replicate the last one
    //     (jcsample.c expand_right_edge, jcprepct.c expand_bottom_edge)
    const bool inside = (px0 + 16 <= W) && (py0 + 16 <= H);
    const bool aligned = (((size_t)W * 3) & 3) == 0 && (((uintptr_t)src) & 3) == 0;
    if (bpp == 4) {
        // RGBA frames (the composite's get_rgba() output mode): one 32-bit load per pixel, alpha dropped
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = lane + 32 * i;             // pixel: row = p / 16, col = p % 16
            const int sy = min(py0 + (p >> 4), H - 1), sx = min(px0 + (p & 15), W - 1);
            const uint32_t v = *reinterpret_cast<const uint32_t*>(src + ((size_t)sy * W + sx) * 4);
            s.rgb[p * 3] = (uint8_t)v; s.rgb[p * 3 + 1] = (uint8_t)(v >> 8); s.rgb[p * 3 + 2] = (uint8_t)(v >> 16);
        }
    } else if (inside && aligned) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(s.rgb);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int w = lane + 32 * i;             // 0..191: row = w / 12, word = w % 12
            const int row = w / 12, c = w - row * 12;
            dst[w] = *reinterpret_cast<const uint32_t*>(src + ((size_t)(py0 + row) * W + px0) * 3 + c * 4);
        }
    } else {
        for (int i = 0; i < 24; ++i) {
            const int b = lane + 32 * i;             // 0..767
            const int row = b / 48, cbyte = b - row * 48;
            const int px = cbyte / 3, ch = cbyte - px * 3;
            const int sy = min(py0 + row, H - 1), sx = min(px0 + px, W - 1);
            s.rgb[b] = src[((size_t)sy * W + sx) * 3 + ch];
        }
    }
    __syncwarp();

    // (2) RGB -> YCbCr (jccolor.c, 16-bit fixed point); luma goes straight into its DCT block
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = lane + 32 * i;                 // pixel: row = p / 16, col = p % 16
        const int r = s.rgb[p * 3], g = s.rgb[p * 3 + 1], b = s.rgb[p * 3 + 2];
        const int y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
        const int cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
        const int cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
        const int row = p >> 4, col = p & 15;
        const int k = ((row >> 3) << 1) | (col >> 3);
        s.blk[k * 72 + (row & 7) * 9 + (col & 7)] = y - 128;
        s.cb[p] = (uint8_t)cb;
        s.cr[p] = (uint8_t)cr;
    }
    __syncwarp();

    // (3) h2v2 chroma: 2x2 box, bias 1,2,1,2 along the row (jcsample.c h2v2_downsample); chroma
    //     rows below the last real one copy it (jcprepct.c pads the downsampled output)
    {
        const int last_real = jd->ch_real - 1 - my * 8;      // >= 0
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = lane + 32 * i;
            const int ly = c >> 3, lx = c & 7;
            const int sly = min(ly, last_real);
            const int bias = (lx & 1) ? 2 : 1;
            const uint8_t* a = s.cb + (2 * sly) * 16 + 2 * lx;
            s.blk[4 * 72 + ly * 9 + lx] = ((a[0] + a[1] + a[16] + a[17] + bias) >> 2) - 128;
            a = s.cr + (2 * sly) * 16 + 2 * lx;
            s.blk[5 * 72 + ly * 9 + lx] = ((a[0] + a[1] + a[16] + a[17] + bias) >> 2) - 128;
        }
    }
    __syncwarp();

    // (4) 48 row transforms, then 48 column transforms (jfdctint.c)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int t = lane + 32 * i;
        if (t < 48) fdct8<1, 0>(s.blk + (t >> 3) * 72 + (t & 7) * 9);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int t = lane + 32 * i;
        if (t < 48) fdct8<9, 1>(s.blk + (t >> 3) * 72 + (t & 7));
    }
    __syncwarp();

    // (5) quantise (jcdctmgr.c: divisor q*8, round half away from zero), zig-zag.  Luma blocks
    //     outside the real block grid are dummies: AC 0, DC copied (jccoefct.c compress_data)
    const int y_wblk = jd->y_wblk, y_hblk = jd->y_hblk;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int e = lane + 32 * i;                 // 0..383
        const int k = e >> 6, z = e & 63;
        const int nat = s_zz[z];
        int v = s.blk[k * 72 + (nat >> 3) * 9 + (nat & 7)];
        const int qv = s_qdiv[k >> 2][nat];
        const int a = v < 0 ? -v : v;
        int qd = (a + (qv >> 1)) / qv;
        v = v < 0 ? -qd : qd;
        if (k < 4) {
            const int bx = mx * 2 + (k & 1), by = my * 2 + (k >> 1);
            if (bx >= y_wblk || by >= y_hblk) v = 0;
        }
        s.zz[e] = (int16_t)v;
    }
    __syncwarp();
    if (lane == 0) {
        for (int k = 1; k < 4; ++k) {
            const int bx = mx * 2 + (k & 1), by = my * 2 + (k >> 1);
            if (by >= y_hblk) s.zz[k * 64] = s.zz[1 * 64];              // bottom dummy row: last block of the row above
            else if (bx >= y_wblk) s.zz[k * 64] = s.zz[(k - 1) * 64];   // right dummy: left neighbour
        }
    }
    __syncwarp();

    // (6) coefficients out (768 B per MCU, 16-byte stores) + AC code length per block
    {
        uint4* dst = reinterpret_cast<uint4*>(coef_all + ((size_t)f * jd->nblk + (size_t)m * 6) * 64);
        const uint4* sv = reinterpret_cast<const uint4*>(s.zz);
        dst[lane] = sv[lane];
        if (lane < 16) dst[32 + lane] = sv[32 + lane];
        // the six DC values again, densely packed: J2 reads nothing else of the coefficients
        if (lane < 6) dcs_all[(size_t)f * jd->nblk + (size_t)m * 6 + lane] = s.zz[lane * 64];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int lo = s.zz[k * 64 + lane], hi = s.zz[k * 64 + 32 + lane];
        const unsigned mlo = __ballot_sync(0xFFFFFFFFu, lo != 0) & ~1u;
        const unsigned mhi = __ballot_sync(0xFFFFFFFFu, hi != 0);
        const unsigned long long nz64 = (unsigned long long)mlo | ((unsigned long long)mhi << 32);
        const uint32_t* ac = s_ac[k >> 2];
        const int zl = (int)(ac[0xF0] >> 16);
        int len = 0;
        if (lane > 0) {
            const HuffElem e = huff_elem(lane, lo, nz64, nullptr, ac);
            len += e.zrl_n * zl + e.len_b;
        }
        const HuffElem e2 = huff_elem(lane + 32, hi, nz64, nullptr, ac);
        len += e2.zrl_n * zl + e2.len_b;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) len += __shfl_xor_sync(0xFFFFFFFFu, len, d);
        if (lane == 0) acbits_all[(size_t)f * jd->nblk + (size_t)m * 6 + k] = (uint16_t)len;
    }
}

// ---- J2: bit offsets ---------------------------------------------------------------------------
// One CTA per frame.  Phase A (thread per MCU, coalesced): bits of the MCU = its six AC lengths +
// the six DC-difference code lengths (predictors: the previous block of the same component,
// jchuff.c last_dc_val).  Phase B: exclusive scan, every thread owning a contiguous run.
__global__ void __launch_bounds__(1024) k_jpeg_scan(const JpegDev* __restrict__ jd,
                                                    const int16_t* __restrict__ dcs_all,
                                                    const uint16_t* __restrict__ acbits_all,
                                                    uint32_t* __restrict__ mcu_bitoff_all,
                                                    uint32_t* __restrict__ raw_all, JpegFrameInfo* __restrict__ info_all)
{
    __shared__ uint32_t s_dc[2][12];
    __shared__ uint32_t s_wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sage_dep_launch();                       // (PDL: the next kernel of the chain may be scheduled; it waits below)
    if (tid < 24) s_dc[tid / 12][tid % 12] = jd->dc_tab[tid / 12][tid % 12];
    __syncthreads();
    sage_dep_wait();                         // the DCT kernel's coefficients are complete and visible
    const int f = blockIdx.x;
    const int nmcu = jd->nmcu;
    const int16_t* dcs = dcs_all + (size_t)f * jd->nblk;
    const uint16_t* acbits = acbits_all + (size_t)f * jd->nblk;
    uint32_t* mcu_bitoff = mcu_bitoff_all + (size_t)f * nmcu;
    for (int m = tid; m < nmcu; m += 1024) {
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int pb = dc_pred_block(m, k);
            const int pred = pb >= 0 ? (int)dcs[pb] : 0;
            const int d = (int)dcs[m * 6 + k] - pred;
            const int nb = size_category(d < 0 ? -d : d);
            bits += (uint32_t)acbits[m * 6 + k] + (s_dc[k >> 2][nb] >> 16) + (uint32_t)nb;
        }
        mcu_bitoff[m] = bits;
    }
    __syncthreads();
    const int per = (nmcu + 1023) / 1024;
    const int m0 = min(tid * per, nmcu), m1 = min(m0 + per, nmcu);
    uint32_t sum = 0;
    for (int m = m0; m < m1; ++m) sum += mcu_bitoff[m];
    uint32_t incl = (uint32_t)warp_incl_scan((int)sum, lane);
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = s_wsum[lane];
        const uint32_t wi = (uint32_t)warp_incl_scan((int)w, lane);
        s_wsum[lane] = wi - w;                       // exclusive warp offsets
        if (lane == 31) {
            JpegFrameInfo& inf = info_all[f];
            const uint32_t total = wi;
            const uint32_t raw_bytes = (total + 7u) >> 3;
            inf.total_bits = total;
            inf.raw_bytes = raw_bytes;
            const bool ovf = (size_t)raw_bytes > (size_t)jd->raw_words * 4;
            inf.overflow = ovf ? 1u : 0u;
            inf.ff_total = 0;
            // jchuff.c flush_bits: the last byte is filled with 1-bits
            const uint32_t pad = (8u - (total & 7u)) & 7u;
            if (pad && !ovf) {
                const uint32_t off = total & 31u;     // pad never crosses a byte, hence never a word
                atomicOr(raw_all + (size_t)f * jd->raw_words + (total >> 5), ((1u << pad) - 1u) << (32u - off - pad));
            }
        }
    }
    __syncthreads();
    uint32_t run = s_wsum[wid] + incl - sum;
    for (int m = m0; m < m1; ++m) {
        const uint32_t b = mcu_bitoff[m];
        mcu_bitoff[m] = run;
        run += b;
    }
}

// ---- J3: Huffman coding --------------------------------------------------------------------------
constexpr int J3_WARPS = 4;
constexpr int J3_WORDS = 320;   // >= (6 * 1660 + 31) / 32 + 2: the longest possible MCU plus the bit shift

// OR `len` bits (right-aligned in v, len <= 33) into the msb-first bit string at bit position pos
__device__ __forceinline__ void smem_put(uint32_t* buf, int pos, unsigned long long v, int len)
{
    if (len == 0) return;
    const int w = pos >> 5, off = pos & 31;
    const unsigned long long x = v << (64 - off - len);      // off + len <= 31 + 33 = 64
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    if (hi) atomicOr(&buf[w], hi);
    if (lo) atomicOr(&buf[w + 1], lo);
}

__global__ void __launch_bounds__(J3_WARPS * 32) k_jpeg_huff(const JpegDev* __restrict__ jd,
                                                             const int16_t* __restrict__ coef_all,
                                                             const uint32_t* __restrict__ mcu_bitoff_all,
                                                             const JpegFrameInfo* __restrict__ info_all,
                                                             uint32_t* __restrict__ raw_all)
{
    __shared__ uint32_t s_buf[J3_WARPS][J3_WORDS];
    __shared__ uint32_t s_ac[2][256];
    __shared__ uint32_t s_dc[2][12];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sage_dep_launch();
    for (int i = tid; i < 512; i += blockDim.x) s_ac[i >> 8][i & 255] = jd->ac_tab[i >> 8][i & 255];
    if (tid < 24) s_dc[tid / 12][tid % 12] = jd->dc_tab[tid / 12][tid % 12];
    __syncthreads();
    sage_dep_wait();                         // bit offsets of every MCU (k_jpeg_scan)
    const int f = blockIdx.y;
    const int m = blockIdx.x * J3_WARPS + wid;
    if (m >= jd->nmcu || info_all[f].overflow) return;
    const int16_t* coef = coef_all + (size_t)f * jd->nblk * 64;
    uint32_t* buf = s_buf[wid];
    for (int i = lane; i < J3_WORDS; i += 32) buf[i] = 0;
    __syncwarp();
    const uint32_t base_bit = mcu_bitoff_all[(size_t)f * jd->nmcu + m];
    int rel = (int)(base_bit & 31u);
    int prev_dc = 0;
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
        const int b = m * 6 + k;
        const int lo = coef[(size_t)b * 64 + lane], hi = coef[(size_t)b * 64 + 32 + lane];
        const int dc = __shfl_sync(0xFFFFFFFFu, lo, 0);
        int pred;
        if (k >= 1 && k <= 3) pred = prev_dc;
        else {
            const int pb = dc_pred_block(m, k);
            pred = pb >= 0 ? (int)coef[(size_t)pb * 64] : 0;
        }
        prev_dc = dc;
        const unsigned mlo = __ballot_sync(0xFFFFFFFFu, lo != 0) & ~1u;
        const unsigned mhi = __ballot_sync(0xFFFFFFFFu, hi != 0);
        const unsigned long long nz64 = (unsigned long long)mlo | ((unsigned long long)mhi << 32);
        const uint32_t* ac = s_ac[k >> 2];
        const uint32_t zrl = ac[0xF0];
        const int zl = (int)(zrl >> 16);
        const unsigned long long zc = zrl & 0xFFFFu;
        const HuffElem e0 = huff_elem(lane, lane == 0 ? dc - pred : lo, nz64, s_dc[k >> 2], ac);
        const HuffElem e1 = huff_elem(lane + 32, hi, nz64, s_dc[k >> 2], ac);
        const int l0 = e0.zrl_n * zl + e0.len_b, l1 = e1.zrl_n * zl + e1.len_b;
        const int i0 = warp_incl_scan(l0, lane);
        const int tot0 = __shfl_sync(0xFFFFFFFFu, i0, 31);
        const int i1 = warp_incl_scan(l1, lane);
        const int tot1 = __shfl_sync(0xFFFFFFFFu, i1, 31);
        int p0 = rel + i0 - l0, p1 = rel + tot0 + i1 - l1;
        if (e0.zrl_n) {
            const unsigned long long z = e0.zrl_n == 1 ? zc : (e0.zrl_n == 2 ? (zc << zl) | zc : (zc << (2 * zl)) | (zc << zl) | zc);
            smem_put(buf, p0, z, e0.zrl_n * zl);
            p0 += e0.zrl_n * zl;
        }
        smem_put(buf, p0, e0.bits, e0.len_b);
        if (e1.zrl_n) {
            const unsigned long long z = e1.zrl_n == 1 ? zc : (e1.zrl_n == 2 ? (zc << zl) | zc : (zc << (2 * zl)) | (zc << zl) | zc);
            smem_put(buf, p1, z, e1.zrl_n * zl);
            p1 += e1.zrl_n * zl;
        }
        smem_put(buf, p1, e1.bits, e1.len_b);
        rel += tot0 + tot1;
    }
    __syncwarp();
    // whole words belong to this MCU alone: plain stores; the first and last word are shared
    // with the neighbouring MCUs (and the final 1-fill): OR into the zeroed stream
    const int nwords = (rel + 31) >> 5;
    uint32_t* raw = raw_all + (size_t)f * jd->raw_words + (base_bit >> 5);
    for (int w = lane; w < nwords; w += 32) {
        const uint32_t v = buf[w];
        if (w == 0 || w == nwords - 1) {
            if (v) atomicOr(raw + w, v);
        } else
            raw[w] = v;
    }
}

// ---- J4: 0xFF stuffing ------------------------------------------------------------------------
__device__ __forceinline__ int count_ff(uint32_t w)
{
    // bytes equal to 0xFF: a byte is 0xFF iff ~byte == 0
    const uint32_t x = ~w;
    int c = 0;
    c += (x & 0xFF000000u) == 0;
    c += (x & 0x00FF0000u) == 0;
    c += (x & 0x0000FF00u) == 0;
    c += (x & 0x000000FFu) == 0;
    return c;
}

// thread t of chunk c owns raw bytes [c*4096 + t*16, +16); bytes at or past raw_bytes do not exist
__device__ __forceinline__ uint4 load_raw16(const uint32_t* raw, uint32_t byte0, uint32_t raw_bytes, int* n_valid)
{
    uint4 v = make_uint4(0, 0, 0, 0);
    *n_valid = 0;
    if (byte0 < raw_bytes) {
        v = *reinterpret_cast<const uint4*>(raw + (byte0 >> 2));
        const uint32_t left = raw_bytes - byte0;
        *n_valid = left < 16u ? (int)left : 16;
    }
    return v;
}

__device__ __forceinline__ uint32_t raw_byte(const uint4& v, int i)
{
    const uint32_t w = i < 8 ? (i < 4 ? v.x : v.y) : (i < 12 ? v.z : v.w);
    return (w >> (24 - 8 * (i & 3))) & 0xFFu;
}

__device__ __forceinline__ int count_ff16(const uint4& v, int n_valid)
{
    if (n_valid == 16) return count_ff(v.x) + count_ff(v.y) + count_ff(v.z) + count_ff(v.w);
    int c = 0;
    for (int i = 0; i < n_valid; ++i) c += raw_byte(v, i) == 0xFFu;
    return c;
}

__global__ void __launch_bounds__(256) k_jpeg_ffcount(const JpegDev* __restrict__ jd,
                                                      const JpegFrameInfo* __restrict__ info_all,
                                                      const uint32_t* __restrict__ raw_all,
                                                      uint32_t* __restrict__ ffcnt_all)
{
    __shared__ int s_w[8];
    const int f = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sage_dep_launch();
    sage_dep_wait();
    const JpegFrameInfo inf = info_all[f];
    if (inf.overflow || (uint32_t)chunk * JPEG_CHUNK >= inf.raw_bytes) return;      // CTA-uniform
    int nv;
    const uint4 v = load_raw16(raw_all + (size_t)f * jd->raw_words, (uint32_t)chunk * JPEG_CHUNK + tid * 16, inf.raw_bytes, &nv);
    int c = count_ff16(v, nv);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if (lane == 0) s_w[wid] = c;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        ffcnt_all[(size_t)f * jd->n_chunks + chunk] = (uint32_t)t;
    }
}

// exclusive scan of the chunk counts (in place), header + EOI, final size
__global__ void __launch_bounds__(256) k_jpeg_ffscan(const JpegDev* __restrict__ jd, JpegFrameInfo* __restrict__ info_all,
                                                     uint32_t* __restrict__ ffcnt_all, uint8_t* __restrict__ out_all,
                                                     long long out_stride, long long* __restrict__ sizes)
{
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_carry;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sage_dep_launch();
    sage_dep_wait();
    JpegFrameInfo& inf = info_all[f];
    uint8_t* out = out_all + (size_t)f * (size_t)out_stride;
    if (inf.overflow) {
        if (tid == 0) sizes[f] = -1;
        return;
    }
    const int n = (int)((inf.raw_bytes + JPEG_CHUNK - 1) / JPEG_CHUNK);
    uint32_t* cnt = ffcnt_all + (size_t)f * jd->n_chunks;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + tid;
        const uint32_t c = i < n ? cnt[i] : 0u;
        const uint32_t incl = (uint32_t)warp_incl_scan((int)c, lane);
        if (lane == 31) s_w[wid] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < wid; ++w) woff += s_w[w];
        const uint32_t carry = s_carry;
        if (i < n) cnt[i] = carry + woff + incl - c;
        __syncthreads();
        if (tid == 255) s_carry = carry + woff + incl;
        __syncthreads();
    }
    const uint32_t ff_total = s_carry;
    const long long size = (long long)jd->header_len + (long long)inf.raw_bytes + (long long)ff_total + 2;
    if (tid == 0) {
        inf.ff_total = ff_total;
        sizes[f] = size <= out_stride ? size : -size;     // negative: the caller's buffer is too small
    }
    if (size <= out_stride) {
        for (int i = tid; i < jd->header_len; i += 256) out[i] = jd->header[i];
        if (tid == 0) {
            out[size - 2] = 0xFF;
            out[size - 1] = 0xD9;
        }
    }
}

__global__ void __launch_bounds__(256) k_jpeg_stuff(const JpegDev* __restrict__ jd,
                                                    const JpegFrameInfo* __restrict__ info_all,
                                                    const uint32_t* __restrict__ raw_all,
                                                    const uint32_t* __restrict__ ffoff_all, uint8_t* __restrict__ out_all,
                                                    long long out_stride)
{
    __shared__ int s_w[8];
    const int f = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sage_dep_wait();
    const JpegFrameInfo inf = info_all[f];
    if (inf.overflow || (uint32_t)chunk * JPEG_CHUNK >= inf.raw_bytes) return;      // CTA-uniform
    const long long size = (long long)jd->header_len + (long long)inf.raw_bytes + (long long)inf.ff_total + 2;
    if (size > out_stride) return;
    const uint32_t byte0 = (uint32_t)chunk * JPEG_CHUNK + tid * 16;
    int nv;
    const uint4 v = load_raw16(raw_all + (size_t)f * jd->raw_words, byte0, inf.raw_bytes, &nv);
    const int c = count_ff16(v, nv);
    const int incl = warp_incl_scan(c, lane);
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += s_w[w];
    uint8_t* o = out_all + (size_t)f * (size_t)out_stride + jd->header_len + byte0 +
                 ffoff_all[(size_t)f * jd->n_chunks + chunk] + (uint32_t)(woff + incl - c);
    for (int i = 0; i < nv; ++i) {
        const uint32_t b = raw_byte(v, i);
        *o++ = (uint8_t)b;
        if (b == 0xFFu) *o++ = 0;
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct HostHuff {
    uint16_t code[256];
    uint8_t len[256];
};

void derive_codes(const uint8_t* bits, const uint8_t* vals, HostHuff* t)
{
    memset(t, 0, sizeof *t);
    unsigned code = 0;
    int k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l - 1]; ++i) {
            t->code[vals[k]] = (uint16_t)code;
            t->len[vals[k]] = (uint8_t)l;
            ++code;
            ++k;
        }
        code <<= 1;
    }
}

}  // namespace

struct sage_jpeg {
    int device = 0;
    int W = 0, H = 0, quality = 75, max_batch = 1;
    int64_t max_bytes = 0;
    JpegDev host{};
    JpegDev* dev = nullptr;
    int16_t* coef = nullptr;
    uint16_t* acbits = nullptr;
    int16_t* dcs = nullptr;
    uint32_t* mcu_bitoff = nullptr;
    uint32_t* raw = nullptr;
    uint32_t* ffcnt = nullptr;
    JpegFrameInfo* info = nullptr;
    // staging of the host entry points
    uint8_t* out_dev = nullptr;
    long long* sizes_dev = nullptr;
    long long* sizes_pinned = nullptr;
    uint8_t* out_pinned = nullptr;      // [max_batch][max_bytes] host staging of the fused render -> JPEG path
    long long* guess = nullptr;         // [max_batch] bytes copied speculatively with the frame (last size + slack)
};

// Enqueue the encode of `batch` frames on `st` (no synchronisation).  The encoder's scratch is an
// array of max_batch frame slots; this call uses slots [slot0, slot0 + batch), so encodes running
// concurrently on different streams must use disjoint slot ranges.
int sage_jpeg_enqueue(sage_jpeg* j, const uint8_t* rgb_dev, int bpp, int batch, int slot0, uint8_t* out_dev, int64_t out_stride,
                      long long* sizes_dev, cudaStream_t st)
{
    REQUIRE(bpp == 3 || bpp == 4, "sage_jpeg_encode: frames are RGB (3 bytes per pixel) or RGBA (4)");
    REQUIRE(bpp == 3 || (((uintptr_t)rgb_dev) & 3) == 0, "sage_jpeg_encode: RGBA frames must be 4-byte aligned");
    REQUIRE(j != nullptr && rgb_dev != nullptr && out_dev != nullptr && sizes_dev != nullptr, "sage_jpeg_encode: NULL argument");
    REQUIRE(batch >= 1 && slot0 >= 0 && slot0 + batch <= j->max_batch, "sage_jpeg_encode: batch exceeds max_batch");
    REQUIRE(out_stride >= j->host.header_len + 2, "sage_jpeg_encode: out_stride too small");
    const JpegDev& h = j->host;
    const size_t z = (size_t)slot0;
    int16_t* coef = j->coef + z * h.nblk * 64;
    uint16_t* acbits = j->acbits + z * h.nblk;
    int16_t* dcs = j->dcs + z * h.nblk;
    uint32_t* bitoff = j->mcu_bitoff + z * h.nmcu;
    uint32_t* raw = j->raw + z * h.raw_words;
    uint32_t* ffcnt = j->ffcnt + z * h.n_chunks;
    JpegFrameInfo* info = j->info + z;
    CK(cudaMemsetAsync(raw, 0, (size_t)batch * h.raw_words * 4, st));
    SAGE_KLAUNCH(k_jpeg_dct, dim3((h.nmcu + J1_WARPS - 1) / J1_WARPS, batch), J1_WARPS * 32, st, j->dev, rgb_dev, bpp, coef, acbits, dcs);
    // the five kernels after the DCT are launched with programmatic stream serialisation: each may be scheduled
    // while its predecessor drains, loads its tables, then waits (griddepcontrol.wait) for the predecessor's data
    SAGE_KLAUNCH_PDL(k_jpeg_scan, batch, 1024, 0, st, j->dev, dcs, acbits, bitoff, raw, info);
    SAGE_KLAUNCH_PDL(k_jpeg_huff, dim3((h.nmcu + J3_WARPS - 1) / J3_WARPS, batch), J3_WARPS * 32, 0, st, j->dev, coef, bitoff, info, raw);
    SAGE_KLAUNCH_PDL(k_jpeg_ffcount, dim3(h.n_chunks, batch), 256, 0, st, j->dev, info, raw, ffcnt);
    SAGE_KLAUNCH_PDL(k_jpeg_ffscan, batch, 256, 0, st, j->dev, info, ffcnt, out_dev, (long long)out_stride, sizes_dev);
    SAGE_KLAUNCH_PDL(k_jpeg_stuff, dim3(h.n_chunks, batch), 256, 0, st, j->dev, info, raw, ffcnt, out_dev, (long long)out_stride);
    CK(cudaGetLastError());
    return 0;
}

// staging buffers of the host entry points (also used by scene_api.cu's fused render -> JPEG path)
uint8_t* sage_jpeg_stage_out(sage_jpeg* j) { return j->out_dev; }
long long* sage_jpeg_stage_sizes_dev(sage_jpeg* j) { return j->sizes_dev; }
long long* sage_jpeg_stage_sizes_host(sage_jpeg* j) { return j->sizes_pinned; }
int64_t sage_jpeg_stage_stride(const sage_jpeg* j) { return j->max_bytes; }
uint8_t* sage_jpeg_stage_out_host(sage_jpeg* j) { return j->out_pinned; }
long long* sage_jpeg_stage_guess(sage_jpeg* j) { return j->guess; }
int sage_jpeg_width(const sage_jpeg* j) { return j->W; }
int sage_jpeg_height(const sage_jpeg* j) { return j->H; }
int sage_jpeg_device(const sage_jpeg* j) { return j->device; }
int sage_jpeg_max_batch(const sage_jpeg* j) { return j->max_batch; }

extern "C" {

int sage_jpeg_destroy(sage_jpeg* j)
{
    if (!j) return 0;
    cudaSetDevice(j->device);
    cudaDeviceSynchronize();
    cudaFree(j->dev); cudaFree(j->coef); cudaFree(j->acbits); cudaFree(j->dcs); cudaFree(j->mcu_bitoff); cudaFree(j->raw);
    cudaFree(j->ffcnt); cudaFree(j->info); cudaFree(j->out_dev); cudaFree(j->sizes_dev);
    if (j->sizes_pinned) cudaFreeHost(j->sizes_pinned);
    if (j->out_pinned) cudaFreeHost(j->out_pinned);
    free(j->guess);
    cudaGetLastError();
    delete j;
    return 0;
}

int sage_jpeg_create(sage_jpeg** out, int device, int width, int height, int quality, int max_batch,
                     int64_t max_bytes_per_frame)
{
    REQUIRE(out != nullptr, "sage_jpeg_create: out is NULL");
    *out = nullptr;
    REQUIRE(sage_device_count() > 0, "sage_jpeg_create: no CUDA device (this library has no CPU fallback)");
    REQUIRE(width > 0 && height > 0 && width <= 16384 && height <= 16384 && (int64_t)width * height <= ((int64_t)1 << 26),
            "sage_jpeg_create: bad resolution (max 16384 per side, 2^26 pixels)");
    REQUIRE(quality >= 1 && quality <= 100, "sage_jpeg_create: quality must be 1..100");
    REQUIRE(max_batch >= 1 && max_batch <= 4096, "sage_jpeg_create: max_batch must be 1..4096");
    CK(cudaSetDevice(device));
    sage_jpeg* j = new sage_jpeg();
    j->device = device; j->W = width; j->H = height; j->quality = quality; j->max_batch = max_batch;
    JpegDev& h = j->host;
    memset(&h, 0, sizeof h);
    h.W = width; h.H = height;
    h.mcux = (width + 15) / 16; h.mcuy = (height + 15) / 16;
    h.nmcu = h.mcux * h.mcuy; h.nblk = h.nmcu * 6;
    h.y_wblk = (width + 7) / 8; h.y_hblk = (height + 7) / 8;
    h.ch_real = (height + 1) / 2;
    // jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    uint8_t q8[2][64];
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            long v = ((long)h_qbase[t][i] * scale + 50L) / 100L;
            if (v <= 0) v = 1;
            if (v > 255) v = 255;
            q8[t][i] = (uint8_t)v;
            h.qdiv[t][i] = (uint16_t)(v << 3);
        }
    HostHuff ht[4];
    const uint8_t* vals[4] = {h_hv_dc, h_hv_ac_l, h_hv_dc, h_hv_ac_c};
    for (int t = 0; t < 4; ++t) derive_codes(h_hbits[t], vals[t], &ht[t]);
    for (int c = 0; c < 2; ++c) {
        for (int i = 0; i < 12; ++i) h.dc_tab[c][i] = ((uint32_t)ht[2 * c].len[i] << 16) | ht[2 * c].code[i];
        for (int i = 0; i < 256; ++i) h.ac_tab[c][i] = ((uint32_t)ht[2 * c + 1].len[i] << 16) | ht[2 * c + 1].code[i];
    }
    // jcmarker.c: SOI, APP0 (JFIF 1.01, density 1:1), DQT x2, SOF0, DHT x4, SOS
    {
        uint8_t* p = h.header;
        auto u8 = [&](int v) { *p++ = (uint8_t)v; };
        auto u16 = [&](int v) { *p++ = (uint8_t)(v >> 8); *p++ = (uint8_t)v; };
        u16(0xFFD8);
        u16(0xFFE0); u16(16);
        const uint8_t jfif[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
        for (int i = 0; i < 14; ++i) u8(jfif[i]);
        for (int t = 0; t < 2; ++t) {
            u16(0xFFDB); u16(67); u8(t);
            for (int i = 0; i < 64; ++i) u8(q8[t][h_zigzag[i]]);
        }
        u16(0xFFC0); u16(17); u8(8); u16(height); u16(width); u8(3);
        const uint8_t comp[9] = {1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
        for (int i = 0; i < 9; ++i) u8(comp[i]);
        for (int t = 0; t < 4; ++t) {
            int n = 0;
            for (int i = 0; i < 16; ++i) n += h_hbits[t][i];
            u16(0xFFC4); u16(19 + n); u8(h_hclass[t]);
            for (int i = 0; i < 16; ++i) u8(h_hbits[t][i]);
            for (int i = 0; i < n; ++i) u8(vals[t][i]);
        }
        u16(0xFFDA); u16(12);
        const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
        for (int i = 0; i < 10; ++i) u8(sos[i]);
        h.header_len = (int)(p - h.header);
    }
    // capacity of one frame: default = the raw frame size (a JPEG of natural content is 10-20x smaller)
    int64_t mb = max_bytes_per_frame > 0 ? max_bytes_per_frame : (int64_t)width * height * 3 / 2 + 4096;
    if (mb < h.header_len + 2 + 64) mb = h.header_len + 2 + 64;
    mb = (mb + 15) & ~(int64_t)15;
    j->max_bytes = mb;
    const int64_t raw_bytes = (mb + JPEG_CHUNK - 1) / JPEG_CHUNK * JPEG_CHUNK;
    REQUIRE(raw_bytes / 4 < ((int64_t)1 << 31), "sage_jpeg_create: max_bytes_per_frame too large");
    h.raw_words = (uint32_t)(raw_bytes / 4);
    h.n_chunks = (int)(raw_bytes / JPEG_CHUNK);
    auto bail = [&](int rc) { sage_jpeg_destroy(j); return rc; };
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(sage_fail(#call, __FILE__, __LINE__, e__)); } while (0)
    const size_t B = (size_t)max_batch;
    CKB(dalloc(&j->dev, 1));
    CKB(cudaMemcpy(j->dev, &h, sizeof h, cudaMemcpyHostToDevice));
    CKB(dalloc(&j->coef, B * h.nblk * 64));
    CKB(dalloc(&j->acbits, B * h.nblk));
    CKB(dalloc(&j->dcs, B * h.nblk));
    CKB(dalloc(&j->mcu_bitoff, B * h.nmcu));
    CKB(dalloc(&j->raw, B * h.raw_words + 4));
    CKB(dalloc(&j->ffcnt, B * h.n_chunks));
    CKB(dalloc(&j->info, B));
    CKB(dalloc(&j->out_dev, B * (size_t)mb));
    CKB(dalloc(&j->sizes_dev, B));
    CKB(cudaHostAlloc((void**)&j->sizes_pinned, sizeof(long long) * B, cudaHostAllocDefault));
    CKB(cudaHostAlloc((void**)&j->out_pinned, B * (size_t)mb, cudaHostAllocDefault));
    j->guess = (long long*)calloc(B, sizeof(long long));
    if (!j->guess) { sage_jpeg_destroy(j); return 1; }
    for (size_t k = 0; k < B; ++k) j->guess[k] = mb < (1 << 18) ? mb : (1 << 18);
#undef CKB
    *out = j;
    return 0;
}

int sage_jpeg_max_bytes(sage_jpeg* j, int64_t* out)
{
    REQUIRE(j != nullptr && out != nullptr, "sage_jpeg_max_bytes: NULL argument");
    *out = j->max_bytes;
    return 0;
}

int sage_jpeg_encode(sage_jpeg* j, const uint8_t* rgb_dev, int batch, uint8_t* out_dev, int64_t out_stride,
                     int64_t* sizes_dev, sage_stream_t stream)
{
    REQUIRE(j != nullptr, "sage_jpeg_encode: NULL encoder");
    CK(cudaSetDevice(j->device));
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    return sage_jpeg_enqueue(j, rgb_dev, 3, batch, 0, out_dev, out_stride, reinterpret_cast<long long*>(sizes_dev), (cudaStream_t)stream);
}

int sage_jpeg_encode_host(sage_jpeg* j, const uint8_t* rgb_dev, int batch, uint8_t* out_host, int64_t out_stride,
                          int64_t* sizes_host, sage_stream_t stream)
{
    REQUIRE(j != nullptr && out_host != nullptr && sizes_host != nullptr, "sage_jpeg_encode_host: NULL argument");
    CK(cudaSetDevice(j->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (sage_jpeg_enqueue(j, rgb_dev, 3, batch, 0, j->out_dev, j->max_bytes, j->sizes_dev, st)) return 1;
    CK(cudaMemcpyAsync(j->sizes_pinned, j->sizes_dev, sizeof(long long) * batch, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int rc = 0;
    for (int f = 0; f < batch; ++f) {
        const long long sz = j->sizes_pinned[f];
        sizes_host[f] = sz;
        if (sz <= 0 || sz > out_stride) { rc = 1; continue; }
        CK(cudaMemcpyAsync(out_host + (size_t)f * out_stride, j->out_dev + (size_t)f * j->max_bytes, (size_t)sz,
                           cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    if (rc) return sage_fail("sage_jpeg_encode_host: a frame did not fit (sizes_host < 0: -needed, or > out_stride)", __FILE__, __LINE__);
    return 0;
}

// Debug / parity: quantised coefficients (zig-zag, scan order) and MCU bit offsets of the last encode.
int sage_jpeg_debug(sage_jpeg* j, int frame, int16_t* coef_host, uint32_t* mcu_bitoff_host, uint32_t* total_bits_host)
{
    REQUIRE(j != nullptr && frame >= 0 && frame < j->max_batch, "sage_jpeg_debug: bad argument");
    CK(cudaSetDevice(j->device));
    CK(cudaDeviceSynchronize());
    const JpegDev& h = j->host;
    if (coef_host) CK(cudaMemcpy(coef_host, j->coef + (size_t)frame * h.nblk * 64, sizeof(int16_t) * h.nblk * 64, cudaMemcpyDeviceToHost));
    if (mcu_bitoff_host) CK(cudaMemcpy(mcu_bitoff_host, j->mcu_bitoff + (size_t)frame * h.nmcu, sizeof(uint32_t) * h.nmcu, cudaMemcpyDeviceToHost));
    if (total_bits_host) {
        JpegFrameInfo inf;
        CK(cudaMemcpy(&inf, j->info + frame, sizeof inf, cudaMemcpyDeviceToHost));
        *total_bits_host = inf.total_bits;
    }
    return 0;
}

}  // extern "C"
