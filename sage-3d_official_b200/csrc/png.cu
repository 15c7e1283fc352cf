// Depth hand-off: clip(depth, lo, hi) * scale -> uint16 -> PNG, on the device.
//
// Replaces, behind the VLM clients of the benchmark,
//     depth_array_clamped = np.clip(depth_image, 0.0, 6.5)
//     depth_array_encoded = (depth_array_clamped * 10000.0).astype(np.uint16)
//     cv2.imencode('.png', depth_array_encoded)        (vlm_client_modular.py:354-364)
// which costs the host ~10 ms per 1080p frame after an 8.3 MB read-back of the float depth map.
//
// What is identical to the reference path: the uint16 PIXELS (binary32 clip, binary32 multiply,
// truncation — NumPy's own sequence), and therefore whatever any PNG reader decodes.  What is not:
// the file BYTES.  OpenCV hands the filtered rows to zlib (Z_RLE strategy, dynamic Huffman trees
// rebuilt every 16 K symbols, libpng's 8 KB IDAT chunking — all build- and version-dependent:
// zlib vs zlib-ng, libpng vs libspng); a PNG is defined by what it decodes to, so this encoder
// writes its own valid stream instead of chasing one build's bytes:
//   * filter type 1 (Sub, bpp = 2) on every row — the filter OpenCV selects;
//   * ONE deflate block with the FIXED Huffman code; tokens are made per group of 32 filtered
//     bytes: a run of >= 4 equal bytes inside a group becomes literal + (length, distance 1) —
//     zlib's RLE idea, without any dependence between groups, so every warp works alone;
//   * zlib wrapper with Adler-32 (exact 64-bit sums reduced mod 65521 at the end), IDAT chunks
//     of 32 KB, CRC-32 per chunk by polynomial folding (zlib's crc32_combine arithmetic).
//
//   P1 k_png_rows<false>  warp / row: quantise, filter, token bit counts, Adler sums, u16 twin
//   P2 k_png_scan         one CTA: exclusive scan of the rows' bit counts, stream length
//   P1 k_png_rows<true>   warp / row again: the same tokens, written at their bit offsets
//   P3 k_png_chunks       CTA / IDAT chunk: copy + CRC-32; header, IEND, final size
#include <string.h>

#include "api_common.cuh"

namespace {

constexpr int PNG_CHUNK = 32768;          // IDAT payload bytes per chunk
constexpr int PNG_PIECE = 128;            // bytes of a chunk each of the 256 CRC threads folds
constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr uint32_t ADLER_MOD = 65521u;

struct PngDev {
    int W, H, L;                 // L = 1 + 2 W filtered bytes per row
    uint32_t stream_words;       // capacity of the deflate bit stream (32-bit words)
    int max_chunks;
    uint8_t head[33];            // signature + IHDR chunk
    uint32_t crc_table[256];
    uint32_t fold[8];            // x^(8 * PNG_PIECE * 2^k) mod P: CRC tree combine
    uint32_t x2n[32];            // x^(2^k) mod P (zlib's x2n_table): shifts of the last, short chunk
    uint32_t k_full;             // constant part of a full chunk's CRC (type bytes, init and final xor)
    uint32_t raw_type;           // raw CRC of the four bytes "IDAT"
};

struct PngInfo {
    unsigned long long s1, s2;   // Adler sums: sum d_i, sum (n - i) d_i
    uint32_t total_bits;         // deflate bits (block header + tokens + end-of-block)
    uint32_t zlen;               // zlib stream bytes: 2 + deflate bytes + 4
    uint32_t overflow;           // the stream did not fit
    uint32_t pad;
};

__host__ __device__ inline uint32_t multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}

// x^(n * 2^k) mod P
__host__ __device__ inline uint32_t x2nmodp(const uint32_t* x2n, uint64_t n, unsigned k)
{
    uint32_t p = 1u << 31;
    while (n) {
        if (n & 1) p = multmodp(x2n[k & 31], p);
        n >>= 1;
        ++k;
    }
    return p;
}

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// filtered byte i (0 .. L-1) of a row whose pixels are q[]: i = 0 is the filter type (1 = Sub);
// raw byte j = i - 1 is the high (j even) or low (j odd) byte of pixel j / 2; Sub subtracts the raw
// byte two positions (one pixel) to the left
__device__ __forceinline__ uint32_t filt_byte(const float* __restrict__ drow, int i, int W, float lo, float hi, float scale)
{
    if (i == 0) return 1u;
    const int j = i - 1, px = j >> 1;
    auto quant = [&](int p) -> uint32_t {
        const float d = __ldg(&drow[p]);
        const float c = fminf(fmaxf(d, lo), hi);           // (a NaN depth becomes lo; the facade's frames are already clipped)
        const float v = __fmul_rn(c, scale);
        return (uint32_t)(uint16_t)(int)v;                   // truncation toward zero
    };
    const uint32_t cur = quant(px);
    const uint32_t a = (j & 1) ? (cur & 0xFFu) : (cur >> 8);
    uint32_t b = 0;
    if (px > 0) {
        const uint32_t prv = quant(px - 1);
        b = (j & 1) ? (prv & 0xFFu) : (prv >> 8);
    }
    return (a - b) & 0xFFu;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) k_png_rows(const PngDev* __restrict__ pd, const float* __restrict__ depth, float lo,
                                                  float hi, float scale, uint32_t* __restrict__ row_bits,
                                                  uint32_t* __restrict__ stream, PngInfo* __restrict__ info,
                                                  uint16_t* __restrict__ twin)
{
    __shared__ uint32_t s_win[4][12];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = blockIdx.x * 4 + w;
    const int W = pd->W, H = pd->H, L = pd->L;
    sage_dep_launch();
    if (WRITE) sage_dep_wait();               // (second pass: the rows' bit offsets come from k_png_scan)
    if (row >= H) return;
    const float* drow = depth + (size_t)row * W;
    if (!WRITE && twin != nullptr) {
        for (int p = lane; p < W; p += 32) {
            const float c = fminf(fmaxf(__ldg(&drow[p]), lo), hi);
            twin[(size_t)row * W + p] = (uint16_t)(int)__fmul_rn(c, scale);
        }
    }
    const unsigned long long n_total = (unsigned long long)H * (unsigned long long)L;
    unsigned long long s1 = 0, s2 = 0;
    uint32_t bits_run = WRITE ? row_bits[row] : 0u;       // WRITE: absolute bit offset of the row's first token
    if (WRITE && info->overflow) return;
    for (int g0 = 0; g0 < L; g0 += 32) {
        const int i = g0 + lane;
        const bool have = i < L;
        const uint32_t v = have ? filt_byte(drow, i, W, lo, hi, scale) : 0xFFFFFFFFu;
        if (!WRITE && have) {
            s1 += v;
            s2 += (n_total - ((unsigned long long)row * L + (unsigned long long)i)) * v;
        }
        // runs of equal bytes inside the group: eq = same as the byte of the lane below
        const uint32_t below = __shfl_up_sync(0xffffffffu, v, 1);
        const bool eq = have && lane > 0 && v == below;
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        // k = offset inside the run, rl = length of the run
        const unsigned upto = m & (0xFFFFFFFFu >> (31 - lane));               // eq bits of lanes <= this one
        const unsigned zeros_below = ~upto & (0xFFFFFFFFu >> (31 - lane));
        const int start = 31 - __clz((int)zeros_below);                           // lane 0 never has eq: always found
        const int k = lane - start;
        // first lane above this one that starts a new run (lanes past the data never have eq set)
        const unsigned sm = (lane == 31) ? 0u : ((~m) >> (lane + 1)) << (lane + 1);
        const int end = sm ? (__ffs((int)sm) - 1) : 32;
        const int rl = end - start;
        // tokens: literal at k = 0; with rl >= 4 one match of rl - 1 at k = 1; shorter runs are literals
        uint32_t code = 0;
        int nb = 0;
        if (have) {
            if (k == 0 || rl < 4) {
                if (v < 144u) { code = rev_bits(0x30u + v, 8); nb = 8; }
                else { code = rev_bits(0x190u + (v - 144u), 9); nb = 9; }
            } else if (k == 1) {
                const int len = rl - 1;                                              // 3 .. 31
                int sym, eb, ev;
                if (len <= 10) { sym = len - 3 + 1; eb = 0; ev = 0; }                // symbols 257..264 -> 7-bit codes 1..8
                else if (len <= 18) { sym = 9 + ((len - 11) >> 1); eb = 1; ev = (len - 11) & 1; }
                else { sym = 13 + ((len - 19) >> 2); eb = 2; ev = (len - 19) & 3; }  // 269..272 for 19..34
                code = rev_bits((uint32_t)sym, 7) | ((uint32_t)ev << 7);             // then the 5-bit distance code 0 (distance 1)
                nb = 7 + eb + 5;
            }
        }
        // exclusive scan of the token lengths
        int incl = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int gbits = __shfl_sync(0xffffffffu, incl, 31);
        if (WRITE) {
            const uint32_t g_off = bits_run;                                         // absolute bit position of the group
            const uint32_t w0 = g_off >> 5;
            if (lane < 12) s_win[w][lane] = 0;
            __syncwarp();
            if (nb) {
                const uint32_t pos = (g_off & 31u) + (uint32_t)(incl - nb);
                const unsigned long long bitsv = (unsigned long long)code << (pos & 31u);
                atomicOr(&s_win[w][pos >> 5], (uint32_t)bitsv);
                if (bitsv >> 32) atomicOr(&s_win[w][(pos >> 5) + 1], (uint32_t)(bitsv >> 32));
            }
            __syncwarp();
            const uint32_t last = ((g_off & 31u) + (uint32_t)gbits + 31u) >> 5;       // words touched
            if ((uint32_t)lane < last && w0 + lane < pd->stream_words) {
                const uint32_t val = s_win[w][lane];
                if (lane == 0 || (uint32_t)lane == last - 1) { if (val) atomicOr(&stream[w0 + lane], val); }
                else stream[w0 + lane] = val;
            }
            __syncwarp();
        }
        bits_run += (uint32_t)gbits;
    }
    if (!WRITE) {
        if (lane == 0) row_bits[row] = bits_run;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        // (reduced per row: the exact sums of a 16384 x 16384 image would not fit 64 bits)
        if (lane == 0) { atomicAdd(&info->s1, s1 % ADLER_MOD); atomicAdd(&info->s2, s2 % ADLER_MOD); }
    }
}

// exclusive scan of the rows' bit counts (in place -> absolute bit offsets), stream length, block header
__global__ void __launch_bounds__(1024) k_png_scan(const PngDev* __restrict__ pd, uint32_t* __restrict__ row_bits,
                                                   uint32_t* __restrict__ stream, PngInfo* __restrict__ info)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const int H = pd->H, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    sage_dep_launch();
    sage_dep_wait();
    if (threadIdx.x == 0) s_carry = 3u;                       // three block-header bits come first
    __syncthreads();
    for (int base = 0; base < H; base += 1024) {
        const int r = base + threadIdx.x;
        const uint32_t v = r < H ? row_bits[r] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[w] = incl;
        __syncthreads();
        uint32_t off = s_carry;
        for (int k = 0; k < w; ++k) off += s_w[k];
        if (r < H) row_bits[r] = off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long total = (unsigned long long)s_carry + 7ull;       // + end-of-block (seven 0 bits)
        const unsigned long long bytes = (total + 7ull) / 8ull;
        const bool over = bytes + 8ull > (unsigned long long)pd->stream_words * 4ull;
        info->overflow = over ? 1u : 0u;
        info->total_bits = (uint32_t)total;
        info->zlen = (uint32_t)(2ull + bytes + 4ull);
        if (!over) atomicOr(&stream[0], 3u);                   // BFINAL = 1, BTYPE = 01 (fixed Huffman), LSB first
    }
}

// zlib stream byte z (0 .. zlen-1): header, deflate bytes, Adler-32
__device__ __forceinline__ uint32_t zbyte(const uint32_t* __restrict__ stream, const PngInfo* __restrict__ info, uint32_t z, uint32_t adler)
{
    if (z == 0) return 0x78u;
    if (z == 1) return 0x01u;
    const uint32_t zlen = info->zlen;
    if (z >= zlen - 4) return (adler >> (8 * (zlen - 1 - z))) & 0xFFu;
    const uint32_t b = z - 2;
    return (stream[b >> 2] >> (8 * (b & 3))) & 0xFFu;
}

__global__ void __launch_bounds__(256) k_png_chunks(const PngDev* __restrict__ pd, const uint32_t* __restrict__ stream,
                                                    const PngInfo* __restrict__ info, uint8_t* __restrict__ out,
                                                    long long out_cap, long long* __restrict__ size_out)
{
    __shared__ uint32_t s_crc[256];
    sage_dep_wait();
    const uint32_t zlen = info->zlen;
    const uint32_t nchunks = (zlen + PNG_CHUNK - 1) / PNG_CHUNK;
    const long long total = 33ll + (long long)nchunks * 12ll + (long long)zlen + 12ll;
    const bool bad = info->overflow != 0u || total > out_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) *size_out = info->overflow ? -1ll : (total > out_cap ? -total : total);
    if (bad || blockIdx.x >= nchunks) return;
    const uint32_t a = (uint32_t)((1ull + info->s1) % ADLER_MOD);
    const uint32_t b = (uint32_t)(((unsigned long long)pd->H * (unsigned long long)pd->L + info->s2) % ADLER_MOD);
    const uint32_t adler = (b << 16) | a;
    const uint32_t c = blockIdx.x;
    const uint32_t z0 = c * PNG_CHUNK;
    const uint32_t len = min((uint32_t)PNG_CHUNK, zlen - z0);
    uint8_t* dst = out + 33 + (size_t)c * (PNG_CHUNK + 12);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < 33; i += 256) out[i] = pd->head[i];
        if (threadIdx.x < 12) {                                  // IEND
            const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
            out[total - 12 + threadIdx.x] = iend[threadIdx.x];
        }
    }
    // pieces are aligned to the END of the chunk (a zero prefix does not change a raw CRC), so the tree
    // below folds with the same eight constants whatever the chunk's length
    const long long p1 = (long long)len - (long long)(255 - threadIdx.x) * PNG_PIECE;       // end of this thread's piece
    const long long p0 = p1 - PNG_PIECE;
    uint32_t crc = 0;
    for (long long i = p0 < 0 ? 0 : p0; i < p1; ++i) {
        const uint32_t v = zbyte(stream, info, z0 + (uint32_t)i, adler);
        dst[8 + i] = (uint8_t)v;
        crc = pd->crc_table[(crc ^ v) & 0xFFu] ^ (crc >> 8);
    }
    s_crc[threadIdx.x] = crc;
    __syncthreads();
#pragma unroll 1
    for (int lvl = 0; lvl < 8; ++lvl) {
        const int stride = 1 << lvl;
        if ((threadIdx.x & (2 * stride - 1)) == 0) {
            const uint32_t left = s_crc[threadIdx.x], right = s_crc[threadIdx.x + stride];
            s_crc[threadIdx.x] = (left ? multmodp(pd->fold[lvl], left) : 0u) ^ right;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t k = pd->k_full;
        if (len != PNG_CHUNK) {
            const uint32_t sh_type = x2nmodp(pd->x2n, len, 3);            // x^(8 len)
            const uint32_t sh_init = x2nmodp(pd->x2n, (uint64_t)len + 4, 3);
            k = multmodp(sh_type, pd->raw_type) ^ multmodp(sh_init, 0xFFFFFFFFu) ^ 0xFFFFFFFFu;
        }
        const uint32_t crc32 = s_crc[0] ^ k;
        dst[0] = (uint8_t)(len >> 24); dst[1] = (uint8_t)(len >> 16); dst[2] = (uint8_t)(len >> 8); dst[3] = (uint8_t)len;
        dst[4] = 'I'; dst[5] = 'D'; dst[6] = 'A'; dst[7] = 'T';
        uint8_t* t = dst + 8 + len;
        t[0] = (uint8_t)(crc32 >> 24); t[1] = (uint8_t)(crc32 >> 16); t[2] = (uint8_t)(crc32 >> 8); t[3] = (uint8_t)crc32;
    }
}

uint32_t host_crc(const uint32_t* table, const uint8_t* p, size_t n, uint32_t crc)
{
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return crc;
}

}  // namespace

struct sage_png {
    int device = 0, W = 0, H = 0, max_batch = 1;
    int64_t max_bytes = 0;
    PngDev host{};
    PngDev* dev = nullptr;
    uint32_t* row_bits = nullptr;     // [B][H]
    uint32_t* stream = nullptr;       // [B][stream_words]
    PngInfo* info = nullptr;          // [B]
    uint8_t* out_dev = nullptr;       // [B][max_bytes] (host-facing paths)
    long long* sizes_dev = nullptr;
    long long* sizes_pinned = nullptr;
    uint8_t* out_pinned = nullptr;
};

int sage_png_destroy(sage_png* p)
{
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaFree(p->dev); cudaFree(p->row_bits); cudaFree(p->stream); cudaFree(p->info); cudaFree(p->out_dev); cudaFree(p->sizes_dev);
    if (p->sizes_pinned) cudaFreeHost(p->sizes_pinned);
    if (p->out_pinned) cudaFreeHost(p->out_pinned);
    delete p;
    return 0;
}

int sage_png_create(sage_png** out, int device, int width, int height, int max_batch)
{
    REQUIRE(out != nullptr, "sage_png_create: NULL out");
    REQUIRE(width >= 1 && height >= 1 && width <= 16384 && height <= 16384, "sage_png_create: width/height must be in [1, 16384]");
    REQUIRE(max_batch >= 1 && max_batch <= 64, "sage_png_create: max_batch must be in [1, 64]");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, "sage_png_create: no such CUDA device");
    CK(cudaSetDevice(device));
    sage_png* p = new sage_png();
    p->device = device; p->W = width; p->H = height; p->max_batch = max_batch;
    PngDev& h = p->host;
    h.W = width; h.H = height; h.L = 1 + 2 * width;
    // worst case: every byte a 9-bit literal
    const uint64_t worst_bits = 3ull + (uint64_t)height * (uint64_t)h.L * 9ull + 7ull;
    const uint64_t words = (worst_bits + 31) / 32 + 16;
    h.stream_words = (uint32_t)words;
    const uint64_t zmax = 2 + words * 4 + 4;
    h.max_chunks = (int)((zmax + PNG_CHUNK - 1) / PNG_CHUNK);
    p->max_bytes = (int64_t)(33 + (uint64_t)h.max_chunks * 12 + zmax + 12 + 15) & ~(int64_t)15;
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? CRC_POLY ^ (c >> 1) : c >> 1;
        h.crc_table[n] = c;
    }
    h.x2n[0] = 1u << 30;
    for (int k = 1; k < 32; ++k) h.x2n[k] = multmodp(h.x2n[k - 1], h.x2n[k - 1]);
    for (int k = 0; k < 8; ++k) h.fold[k] = x2nmodp(h.x2n, (uint64_t)PNG_PIECE << k, 3);
    const uint8_t idat[4] = {'I', 'D', 'A', 'T'};
    h.raw_type = host_crc(h.crc_table, idat, 4, 0);
    h.k_full = multmodp(x2nmodp(h.x2n, PNG_CHUNK, 3), h.raw_type) ^ multmodp(x2nmodp(h.x2n, PNG_CHUNK + 4, 3), 0xFFFFFFFFu) ^ 0xFFFFFFFFu;
    {
        uint8_t* q = h.head;
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        memcpy(q, sig, 8); q += 8;
        const uint8_t ihdr[17] = {'I', 'H', 'D', 'R', (uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                                  (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                                  16, 0, 0, 0, 0};                         // 16-bit greyscale, deflate, adaptive filtering, no interlace
        q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 13; q += 4;
        memcpy(q, ihdr, 17); q += 17;
        const uint32_t crc = host_crc(h.crc_table, ihdr, 17, 0xFFFFFFFFu) ^ 0xFFFFFFFFu;
        q[0] = (uint8_t)(crc >> 24); q[1] = (uint8_t)(crc >> 16); q[2] = (uint8_t)(crc >> 8); q[3] = (uint8_t)crc;
    }
    auto bail = [&](int rc) { sage_png_destroy(p); return rc; };
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(sage_fail(#call, __FILE__, __LINE__, e__)); } while (0)
    const size_t B = (size_t)max_batch;
    CKB(dalloc(&p->dev, 1));
    CKB(cudaMemcpy(p->dev, &h, sizeof h, cudaMemcpyHostToDevice));
    CKB(dalloc(&p->row_bits, B * (size_t)height));
    CKB(dalloc(&p->stream, B * (size_t)h.stream_words));
    CKB(dalloc(&p->info, B));
    CKB(dalloc(&p->out_dev, B * (size_t)p->max_bytes));
    CKB(dalloc(&p->sizes_dev, B));
    CKB(cudaHostAlloc((void**)&p->sizes_pinned, sizeof(long long) * B, cudaHostAllocDefault));
    CKB(cudaHostAlloc((void**)&p->out_pinned, B * (size_t)p->max_bytes, cudaHostAllocDefault));
#undef CKB
    *out = p;
    return 0;
}

int sage_png_max_bytes(sage_png* p, int64_t* out)
{
    REQUIRE(p != nullptr && out != nullptr, "sage_png_max_bytes: NULL argument");
    *out = p->max_bytes;
    return 0;
}

// enqueue the encode of one depth map on scratch slot `slot` (no synchronisation)
int sage_png_enqueue(sage_png* p, const float* depth_dev, float lo, float hi, float scale, int slot, uint8_t* out_dev,
                     int64_t out_cap, long long* size_dev, uint16_t* twin_dev, cudaStream_t st)
{
    REQUIRE(p != nullptr && depth_dev != nullptr && out_dev != nullptr && size_dev != nullptr, "sage_depth_png_encode: NULL argument");
    REQUIRE(slot >= 0 && slot < p->max_batch, "sage_depth_png_encode: slot exceeds max_batch");
    REQUIRE(lo <= hi && hi * scale < 65536.0f && lo * scale > -1.0f, "sage_depth_png_encode: clip(lo, hi) * scale must fit uint16");
    const PngDev& h = p->host;
    uint32_t* row_bits = p->row_bits + (size_t)slot * h.H;
    uint32_t* stream = p->stream + (size_t)slot * h.stream_words;
    PngInfo* info = p->info + slot;
    CK(cudaMemsetAsync(stream, 0, (size_t)h.stream_words * 4, st));
    CK(cudaMemsetAsync(info, 0, sizeof(PngInfo), st));
    const int blocks = (h.H + 3) / 4;
    SAGE_KLAUNCH(k_png_rows<false>, blocks, 128, st, p->dev, depth_dev, lo, hi, scale, row_bits, stream, info, twin_dev);
    SAGE_KLAUNCH_PDL(k_png_scan, 1, 1024, 0, st, p->dev, row_bits, stream, info);
    SAGE_KLAUNCH_PDL(k_png_rows<true>, blocks, 128, 0, st, p->dev, depth_dev, lo, hi, scale, row_bits, stream, info, (uint16_t*)nullptr);
    SAGE_KLAUNCH_PDL(k_png_chunks, h.max_chunks, 256, 0, st, p->dev, stream, info, out_dev, (long long)out_cap, size_dev);
    CK(cudaGetLastError());
    return 0;
}

int sage_depth_png_encode(sage_png* p, const float* depth_dev, float lo, float hi, float scale, uint8_t* out_dev,
                          int64_t out_cap, int64_t* size_dev, uint16_t* twin_dev, sage_stream_t stream)
{
    REQUIRE(p != nullptr, "sage_depth_png_encode: NULL encoder");
    CK(cudaSetDevice(p->device));
    return sage_png_enqueue(p, depth_dev, lo, hi, scale, 0, out_dev, out_cap, reinterpret_cast<long long*>(size_dev), twin_dev,
                            (cudaStream_t)stream);
}

int sage_depth_png_encode_host(sage_png* p, const float* depth_dev, float lo, float hi, float scale, uint8_t* out_host,
                               int64_t out_cap, int64_t* size_host, uint16_t* twin_host, sage_stream_t stream)
{
    REQUIRE(p != nullptr && out_host != nullptr && size_host != nullptr, "sage_depth_png_encode_host: NULL argument");
    CK(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    uint16_t* twin_dev = nullptr;
    if (twin_host) CK(dalloc(&twin_dev, (size_t)p->W * p->H));
    int rc = sage_png_enqueue(p, depth_dev, lo, hi, scale, 0, p->out_dev, p->max_bytes, p->sizes_dev, twin_dev, st);
    if (!rc) {
        CK(cudaMemcpyAsync(p->sizes_pinned, p->sizes_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const long long sz = p->sizes_pinned[0];
        *size_host = sz;
        if (sz <= 0 || sz > out_cap) {
            sage_fail("sage_depth_png_encode_host: out_host is too small (*size_host holds the size needed, negated)", __FILE__, __LINE__);
            rc = 1;
        } else {
            CK(cudaMemcpyAsync(out_host, p->out_dev, (size_t)sz, cudaMemcpyDeviceToHost, st));
            if (twin_host) CK(cudaMemcpyAsync(twin_host, twin_dev, (size_t)p->W * p->H * 2, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
    }
    if (twin_dev) cudaFree(twin_dev);
    return rc;
}

// accessors for the fused render -> PNG path (scene_api.cu)
uint8_t* sage_png_stage_out(sage_png* p) { return p->out_dev; }
uint8_t* sage_png_stage_out_host(sage_png* p) { return p->out_pinned; }
long long* sage_png_stage_sizes_dev(sage_png* p) { return p->sizes_dev; }
long long* sage_png_stage_sizes_host(sage_png* p) { return p->sizes_pinned; }
int64_t sage_png_stage_stride(const sage_png* p) { return p->max_bytes; }
int sage_png_width(const sage_png* p) { return p->W; }
int sage_png_height(const sage_png* p) { return p->H; }
int sage_png_device(const sage_png* p) { return p->device; }
int sage_png_max_batch(const sage_png* p) { return p->max_batch; }
