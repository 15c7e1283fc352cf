/*
 * jpeg_oracle.c — CPU ORACLE for the frame hand-off row.  TEST
 * INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it; the product never does.
 *
 * What it restates.  The reference turns every rendered frame into a JPEG on the host with
 * Pillow's defaults:
 *     Image.fromarray(img_array).save(traj_dir / filename)
 *         Code/data_pipeline/training_data_construction/generate_images.py:431-432
 *     pil_image.save(buffer, format='JPEG')            vlm_client_modular.py:351,402,453
 * i.e. baseline sequential JPEG, quality 75, 4:2:0 chroma, the Annex-K Huffman tables, no
 * restart markers, JFIF 1.01 header.  The arithmetic lives in a third-party dependency that is
 * not under /root/reference: Pillow (Data/isaac_sim_requirements.txt pins pillow==11.2.1 via the
 * Isaac env) bundling libjpeg-turbo (API level 62).  This file restates libjpeg's published
 * algorithm for that configuration:
 *   jccolor.c  rgb_ycc_convert      16-bit fixed-point RGB->YCbCr
 *   jcsample.c h2v2_downsample      2x2 box with the 1,2,1,2 rounding bias, right-edge replication
 *   jcprepct.c pre_process_data     bottom-edge replication to a full iMCU row
 *   jccoefct.c compress_data        dummy blocks right of / below the image (AC 0, DC copied)
 *   jfdctint.c jpeg_fdct_islow      the 13-bit "slow integer" forward DCT
 *   jcdctmgr.c quantize             divisor q*8, round half away from zero
 *   jchuff.c   encode_one_block     DC difference + run-length AC codes, 0xFF stuffing, 1-fill
 *   jcmarker.c                      SOI APP0 DQT DQT SOF0 DHTx4 SOS ... EOI
 *
 * PINNED: tests/test_jpeg_oracle.py encodes the same arrays with the container's Pillow
 * (12.2.0 / libjpeg-turbo, the same library family the reference calls) and requires the
 * byte streams to be IDENTICAL, over sizes that exercise every edge rule (multiples of 16,
 * odd sizes, 1x1, 1920x1080 whose last Y block row is a dummy row) and qualities 1..100.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const uint8_t ZIGZAG[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static const uint8_t QBASE[2][64] = {
    {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

/* Annex K.3 tables: [0]=DC luma, [1]=AC luma, [2]=DC chroma, [3]=AC chroma (file order) */
static const uint8_t HBITS[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
static const uint8_t HCLASS[4] = {0x00, 0x10, 0x01, 0x11};
static const uint8_t HV_DC[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t HV_AC_L[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t HV_AC_C[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

typedef struct {
    uint16_t code[256];
    uint8_t len[256];
} hufftab;

/* jchuff.c jpeg_make_c_derived_tbl: canonical codes in order of increasing length */
static void derive(const uint8_t* bits, const uint8_t* vals, hufftab* t) {
    memset(t, 0, sizeof(*t));
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

/* jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE) */
static void quant_tables(int quality, uint16_t q[2][64]) {
    if (quality <= 0) quality = 1;
    if (quality > 100) quality = 100;
    int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            long v = ((long)QBASE[t][i] * scale + 50L) / 100L;
            if (v <= 0) v = 1;
            if (v > 255) v = 255;
            q[t][i] = (uint16_t)v;
        }
}

typedef struct {
    uint8_t* p;
    int64_t n, cap;
    uint64_t acc; /* bit accumulator, msb first */
    int nacc;
    int overflow;
} bitw;

static void put_byte(bitw* w, int b) {
    if (w->n < w->cap) w->p[w->n] = (uint8_t)b;
    else w->overflow = 1;
    w->n++;
}
static void put_u16(bitw* w, int v) {
    put_byte(w, (v >> 8) & 0xFF);
    put_byte(w, v & 0xFF);
}
/* jchuff.c emit_bits: msb first, 0xFF followed by a stuffed 0x00 */
static void put_bits(bitw* w, unsigned code, int len) {
    w->acc = (w->acc << len) | (code & ((1u << len) - 1u));
    w->nacc += len;
    while (w->nacc >= 8) {
        int b = (int)((w->acc >> (w->nacc - 8)) & 0xFF);
        put_byte(w, b);
        if (b == 0xFF) put_byte(w, 0);
        w->nacc -= 8;
    }
}

#define CONST_BITS 13
#define PASS1_BITS 2
#define DESCALE(x, n) (((x) + (1L << ((n) - 1))) >> (n))
/* jfdctint.c jpeg_fdct_islow; d = 64 samples already level-shifted by -128, row-major */
static void fdct_islow(int32_t* d) {
    const long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633,
               F1501 = 12299, F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 8; ++i) {
            int32_t* p = pass == 0 ? d + 8 * i : d + i;
            int s = pass == 0 ? 1 : 8;
            long t0 = p[0] + p[7 * s], t7 = p[0] - p[7 * s];
            long t1 = p[s] + p[6 * s], t6 = p[s] - p[6 * s];
            long t2 = p[2 * s] + p[5 * s], t5 = p[2 * s] - p[5 * s];
            long t3 = p[3 * s] + p[4 * s], t4 = p[3 * s] - p[4 * s];
            long t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            int sh = pass == 0 ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS;
            if (pass == 0) {
                p[0] = (int32_t)((t10 + t11) << PASS1_BITS);
                p[4 * s] = (int32_t)((t10 - t11) << PASS1_BITS);
            } else {
                p[0] = (int32_t)DESCALE(t10 + t11, PASS1_BITS);
                p[4 * s] = (int32_t)DESCALE(t10 - t11, PASS1_BITS);
            }
            long z1 = (t12 + t13) * F0541;
            p[2 * s] = (int32_t)DESCALE(z1 + t13 * F0765, sh);
            p[6 * s] = (int32_t)DESCALE(z1 + t12 * (-F1847), sh);
            z1 = t4 + t7;
            long z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
            long z5 = (z3 + z4) * F1175;
            t4 *= F0298;
            t5 *= F2053;
            t6 *= F3072;
            t7 *= F1501;
            z1 *= -F0899;
            z2 *= -F2562;
            z3 *= -F1961;
            z4 *= -F0390;
            z3 += z5;
            z4 += z5;
            p[7 * s] = (int32_t)DESCALE(t4 + z1 + z3, sh);
            p[5 * s] = (int32_t)DESCALE(t5 + z2 + z4, sh);
            p[3 * s] = (int32_t)DESCALE(t6 + z2 + z3, sh);
            p[s] = (int32_t)DESCALE(t7 + z1 + z4, sh);
        }
    }
}

static int nbits_of(int v) {
    int n = 0;
    while (v) {
        ++n;
        v >>= 1;
    }
    return n;
}

/* jchuff.c encode_one_block; blk = 64 quantised coefficients in NATURAL order */
static void encode_block(bitw* w, const int16_t* blk, int last_dc, const hufftab* dc, const hufftab* ac) {
    int temp = blk[0] - last_dc, temp2 = temp;
    if (temp < 0) {
        temp = -temp;
        temp2--;
    }
    int nb = nbits_of(temp);
    put_bits(w, dc->code[nb], dc->len[nb]);
    if (nb) put_bits(w, (unsigned)temp2, nb);
    int r = 0;
    for (int k = 1; k < 64; ++k) {
        temp = blk[ZIGZAG[k]];
        if (temp == 0) {
            ++r;
            continue;
        }
        while (r > 15) {
            put_bits(w, ac->code[0xF0], ac->len[0xF0]);
            r -= 16;
        }
        temp2 = temp;
        if (temp < 0) {
            temp = -temp;
            temp2--;
        }
        nb = nbits_of(temp);
        int sym = (r << 4) + nb;
        put_bits(w, ac->code[sym], ac->len[sym]);
        put_bits(w, (unsigned)temp2, nb);
        r = 0;
    }
    if (r > 0) put_bits(w, ac->code[0], ac->len[0]);
}

static void header(bitw* w, int W, int H, uint16_t q[2][64]) {
    put_u16(w, 0xFFD8);
    put_u16(w, 0xFFE0);
    put_u16(w, 16);
    const uint8_t jfif[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    for (int i = 0; i < 14; ++i) put_byte(w, jfif[i]);
    for (int t = 0; t < 2; ++t) {
        put_u16(w, 0xFFDB);
        put_u16(w, 67);
        put_byte(w, t);
        for (int i = 0; i < 64; ++i) put_byte(w, q[t][ZIGZAG[i]]);
    }
    put_u16(w, 0xFFC0);
    put_u16(w, 17);
    put_byte(w, 8);
    put_u16(w, H);
    put_u16(w, W);
    put_byte(w, 3);
    const uint8_t comp[9] = {1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    for (int i = 0; i < 9; ++i) put_byte(w, comp[i]);
    const uint8_t* vals[4] = {HV_DC, HV_AC_L, HV_DC, HV_AC_C};
    for (int t = 0; t < 4; ++t) {
        int n = 0;
        for (int i = 0; i < 16; ++i) n += HBITS[t][i];
        put_u16(w, 0xFFC4);
        put_u16(w, 19 + n);
        put_byte(w, HCLASS[t]);
        for (int i = 0; i < 16; ++i) put_byte(w, HBITS[t][i]);
        for (int i = 0; i < n; ++i) put_byte(w, vals[t][i]);
    }
    put_u16(w, 0xFFDA);
    put_u16(w, 12);
    const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    for (int i = 0; i < 10; ++i) put_byte(w, sos[i]);
}

/* forward DCT + quantisation of one 8x8 block of `plane` (stride ps) at block (bx, by) */
static void dct_quant(const uint8_t* plane, int ps, int bx, int by, const uint16_t* q, int16_t* out) {
    int32_t d[64];
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) d[8 * y + x] = (int32_t)plane[(size_t)(by * 8 + y) * ps + bx * 8 + x] - 128;
    fdct_islow(d);
    for (int i = 0; i < 64; ++i) {
        int qv = q[i] << 3, t = d[i];
        if (t < 0) {
            t = -t;
            t += qv >> 1;
            t = t >= qv ? t / qv : 0;
            t = -t;
        } else {
            t += qv >> 1;
            t = t >= qv ? t / qv : 0;
        }
        out[i] = (int16_t)t;
    }
}

/* Encode an interleaved RGB8 image [H][W][3] exactly as Pillow's Image.save(format='JPEG',
 * quality=quality) does.  Returns 0 and the length in *out_len; 1 if cap was too small
 * (*out_len then holds the required size); 2 on bad arguments. */
static int jpeg_encode_impl(const uint8_t* rgb, int W, int H, int quality, uint8_t* out, int64_t cap, int64_t* out_len,
                            int16_t* coef_out) {
    if (W <= 0 || H <= 0 || W > 65535 || H > 65535) return 2;
    uint16_t q[2][64];
    quant_tables(quality, q);
    hufftab ht[4];
    derive(HBITS[0], HV_DC, &ht[0]);
    derive(HBITS[1], HV_AC_L, &ht[1]);
    derive(HBITS[2], HV_DC, &ht[2]);
    derive(HBITS[3], HV_AC_C, &ht[3]);

    const int mcux = (W + 15) / 16, mcuy = (H + 15) / 16;
    const int yw = mcux * 16, yh = mcuy * 16, cw = mcux * 8, ch = mcuy * 8;
    const int y_wblk = (W + 7) / 8, y_hblk = (H + 7) / 8; /* real (non-dummy) Y blocks */
    uint8_t* Y = (uint8_t*)malloc((size_t)yw * yh);
    uint8_t* Cbf = (uint8_t*)malloc((size_t)yw * yh);
    uint8_t* Crf = (uint8_t*)malloc((size_t)yw * yh);
    uint8_t* Cb = (uint8_t*)malloc((size_t)cw * ch);
    uint8_t* Cr = (uint8_t*)malloc((size_t)cw * ch);
    /* colour conversion with right-edge (jcsample.c expand_right_edge) and bottom-edge
     * (jcprepct.c expand_bottom_edge) replication */
    for (int y = 0; y < yh; ++y) {
        int sy = y < H ? y : H - 1;
        for (int x = 0; x < yw; ++x) {
            int sx = x < W ? x : W - 1;
            const uint8_t* p = rgb + ((size_t)sy * W + sx) * 3;
            long r = p[0], g = p[1], b = p[2];
            Y[(size_t)y * yw + x] = (uint8_t)((19595 * r + 38470 * g + 7471 * b + 32768) >> 16);
            Cbf[(size_t)y * yw + x] = (uint8_t)((-11059 * r - 21709 * g + 32768 * b + (128L << 16) + 32767) >> 16);
            Crf[(size_t)y * yw + x] = (uint8_t)((32768 * r - 27439 * g - 5329 * b + (128L << 16) + 32767) >> 16);
        }
    }
    /* h2v2 box filter, bias 1,2,1,2,... along a row (the right edge was replicated at full
     * resolution BEFORE filtering: jcsample.c).  Chroma rows below the last real one,
     * ceil(H/2)-1, are copies of it: jcprepct.c pads the DOWNSAMPLED output, it does not filter
     * replicated image rows. */
    const int ch_real = (H + 1) / 2;
    for (int y = 0; y < ch; ++y) {
        int sy = y < ch_real ? y : ch_real - 1;
        for (int x = 0; x < cw; ++x) {
            int bias = (x & 1) ? 2 : 1;
            const uint8_t* a = Cbf + (size_t)(2 * sy) * yw + 2 * x;
            Cb[(size_t)y * cw + x] = (uint8_t)((a[0] + a[1] + a[yw] + a[yw + 1] + bias) >> 2);
            a = Crf + (size_t)(2 * sy) * yw + 2 * x;
            Cr[(size_t)y * cw + x] = (uint8_t)((a[0] + a[1] + a[yw] + a[yw + 1] + bias) >> 2);
        }
    }
    bitw w = {out, 0, cap, 0, 0, 0};
    header(&w, W, H, q);
    int last_dc[3] = {0, 0, 0};
    int16_t blk[6][64];
    for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
            /* jccoefct.c compress_data: Y blocks beyond the real block grid are dummies */
            for (int yi = 0; yi < 2; ++yi)
                for (int xi = 0; xi < 2; ++xi) {
                    int bx = mx * 2 + xi, by = my * 2 + yi, k = yi * 2 + xi;
                    if (by < y_hblk && bx < y_wblk) dct_quant(Y, yw, bx, by, q[0], blk[k]);
                    else {
                        memset(blk[k], 0, sizeof(blk[k]));
                        /* right-edge dummy copies its left neighbour; bottom-row dummies copy the
                         * last block of the row above (index 1) */
                        blk[k][0] = by < y_hblk ? blk[k - 1][0] : blk[1][0];
                    }
                }
            dct_quant(Cb, cw, mx, my, q[1], blk[4]);
            dct_quant(Cr, cw, mx, my, q[1], blk[5]);
            if (coef_out) /* stage-level parity: zig-zag order, scan order, 6 blocks per MCU */
                for (int k = 0; k < 6; ++k)
                    for (int z = 0; z < 64; ++z) coef_out[(((size_t)my * mcux + mx) * 6 + k) * 64 + z] = blk[k][ZIGZAG[z]];
            for (int k = 0; k < 4; ++k) {
                encode_block(&w, blk[k], last_dc[0], &ht[0], &ht[1]);
                last_dc[0] = blk[k][0];
            }
            encode_block(&w, blk[4], last_dc[1], &ht[2], &ht[3]);
            last_dc[1] = blk[4][0];
            encode_block(&w, blk[5], last_dc[2], &ht[2], &ht[3]);
            last_dc[2] = blk[5][0];
        }
    put_bits(&w, 0x7F, 7); /* jchuff.c flush_bits: fill with ones */
    w.nacc = 0;
    put_u16(&w, 0xFFD9);
    free(Y);
    free(Cbf);
    free(Crf);
    free(Cb);
    free(Cr);
    *out_len = w.n;
    return w.overflow ? 1 : 0;
}

int so_jpeg_encode(const uint8_t* rgb, int W, int H, int quality, uint8_t* out, int64_t cap, int64_t* out_len) {
    return jpeg_encode_impl(rgb, W, H, quality, out, cap, out_len, NULL);
}

/* Same, and also the quantised coefficients: int16 [ceil(W/16)*ceil(H/16)*6][64], zig-zag order,
 * blocks in scan order (Y00 Y01 Y10 Y11 Cb Cr per MCU) — what the CUDA path exposes through
 * sage_jpeg_debug for stage-level parity. */
int so_jpeg_encode_coefs(const uint8_t* rgb, int W, int H, int quality, uint8_t* out, int64_t cap, int64_t* out_len,
                         int16_t* coef_out) {
    return jpeg_encode_impl(rgb, W, H, quality, out, cap, out_len, coef_out);
}
