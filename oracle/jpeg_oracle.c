/*
 * oracle/jpeg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPU restatement of the reference JPEG-encode path (TinyJPEG as vendored in
 * /root/reference/jpeg_enc.h) written from scratch so that it can
 *   (a) dump intermediate stages (quantised coefficients, per-block bit lengths,
 *       the unstuffed entropy stream) for stage-by-stage parity of the CUDA path,
 *   (b) define the three *extended* modes BASELINE.json asks for and the reference
 *       cannot produce (IJG quality 1..100, 4:2:0, grayscale).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product (imagecodecs_b200/) never does.
 *
 * PINNING: native modes (tje quality 1..3, 3/4 channels, 4:4:4) are byte-compared
 * with the compiled reference (oracle/_ref/libtje_ref.so) in tests/test_oracle.py
 * and against the SHA-256 known answers of SURVEY.md 8(c) committed under
 * tests/golden/.  Extended modes are "parity unpinned" (the reference returns 0
 * for them, jpeg_enc.h:1223-1226 / :954-956); they collapse to the native modes
 * at IJG Q=50 (tje 1) and Q=100 (tje 3), which IS tested, and are otherwise
 * validated by independent decoders (NanoJPEG, PIL, OpenCV).
 *
 * Must be compiled with -ffp-contract=off (no FMA): every float expression below
 * mirrors the reference's evaluation order in IEEE binary32.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_QMODE_TJE 0   /* quality 1..3, jpeg_enc.h:1231-1256 */
#define ORC_QMODE_IJG 1   /* quality 1..100, extended */

#define ORC_SUB_444 0
#define ORC_SUB_420 1     /* extended */

/* ------------------------------------------------------------------------- */
/* Tables (data of the reference, jpeg_enc.h:266-386).                        */
/* ------------------------------------------------------------------------- */

/* jpeg_enc.h:266-276 (Annex K.1 luminance, stored in natural order) */
static const uint8_t k_base_luma[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };

/* jpeg_enc.h:294-305 (the "from paper" chroma table; NOT Annex K.2) */
static const uint8_t k_base_chroma[64] = {
    16, 12, 14, 14, 18, 24, 49, 72, 11, 10, 16, 24, 40, 51, 61, 12,
    13, 17, 22, 35, 64, 92, 14, 16, 22, 37, 55, 78, 95, 19, 24, 29,
    56, 64, 87, 98, 26, 40, 51, 68, 81, 103, 112, 58, 57, 87, 109, 104,
    121, 100, 60, 69, 80, 103, 113, 120, 103, 55, 56, 62, 77, 92, 101, 99 };

/* jpeg_enc.h:376-386: zz[natural index] = position in the zigzag scan */
static const uint8_t k_zz[64] = {
    0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42,
    3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };

/* Huffman specs, Annex K.3.3 as carried by jpeg_enc.h:310-368.
 * Order of the four tables follows the enum at jpeg_enc.h:891-896. */
enum { HT_LUMA_DC = 0, HT_LUMA_AC = 1, HT_CHROMA_DC = 2, HT_CHROMA_AC = 3 };

static const uint8_t k_bits_luma_dc[16]   = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 };
static const uint8_t k_bits_chroma_dc[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
static const uint8_t k_vals_dc[12]        = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t k_bits_luma_ac[16]   = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d };
static const uint8_t k_bits_chroma_ac[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 };

static const uint8_t k_vals_luma_ac[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7,
    0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5,
    0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8,
    0xF9, 0xFA };

static const uint8_t k_vals_chroma_ac[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0,
    0x15, 0x62, 0x72, 0xD1, 0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26,
    0x27, 0x28, 0x29, 0x2A, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5,
    0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
    0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8,
    0xF9, 0xFA };

static const uint8_t* const k_ht_bits[4] = { k_bits_luma_dc, k_bits_luma_ac, k_bits_chroma_dc, k_bits_chroma_ac };
static const uint8_t* const k_ht_vals[4] = { k_vals_dc, k_vals_luma_ac, k_vals_dc, k_vals_chroma_ac };

/* ------------------------------------------------------------------------- */
/* Table construction                                                         */
/* ------------------------------------------------------------------------- */

/* Quantiser tables in the reference's *stored* order.
 * TJE mode: jpeg_enc.h:1230-1256.  IJG mode: SURVEY 8c-(i).  Returns 1/0. */
int orc_build_qt(int qmode, int quality, uint8_t qt_luma[64], uint8_t qt_chroma[64])
{
    if (qmode == ORC_QMODE_TJE) {
        if (quality < 1 || quality > 3) return 0;           /* jpeg_enc.h:1223 */
        if (quality == 3) {
            memset(qt_luma, 1, 64);
            memset(qt_chroma, 1, 64);
            return 1;
        }
        int div = (quality == 2) ? 10 : 1;                  /* jpeg_enc.h:1238-1241 */
        for (int i = 0; i < 64; ++i) {
            int l = k_base_luma[i] / div, c = k_base_chroma[i] / div;
            qt_luma[i] = (uint8_t)(l ? l : 1);
            qt_chroma[i] = (uint8_t)(c ? c : 1);
        }
        return 1;
    }
    if (qmode == ORC_QMODE_IJG) {
        if (quality < 1 || quality > 100) return 0;
        int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;
        for (int i = 0; i < 64; ++i) {
            int l = (k_base_luma[i] * s + 50) / 100, c = (k_base_chroma[i] * s + 50) / 100;
            qt_luma[i] = (uint8_t)(l < 1 ? 1 : l > 255 ? 255 : l);
            qt_chroma[i] = (uint8_t)(c < 1 ? 1 : c > 255 ? 255 : c);
        }
        return 1;
    }
    return 0;
}

/* Reciprocal AAN-scaled quantiser, natural order (jpeg_enc.h:974-986).
 * The expression is evaluated exactly as the reference writes it:
 * ((8 * aan[x]) * aan[y]) * qt[zz[i]]  with the int 8 / uint8 promoted to float. */
void orc_build_pqt(const uint8_t qt[64], float pqt[64])
{
    static const float aan[8] = { 1.0f, 1.387039845f, 1.306562965f, 1.175875602f,
                                  1.0f, 0.785694958f, 0.541196100f, 0.275899379f };
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
            int i = y * 8 + x;
            pqt[i] = 1.0f / (8 * aan[x] * aan[y] * qt[k_zz[i]]);
        }
}

/* Canonical Huffman code assignment, Annex C.2 (jpeg_enc.h:546-592, :907-946).
 * len[t][sym] = 0 for symbols the table does not define. */
void orc_build_huff(uint8_t len[4][256], uint16_t code[4][256])
{
    memset(len, 0, 4 * 256);
    memset(code, 0, 4 * 256 * sizeof(uint16_t));
    for (int t = 0; t < 4; ++t) {
        unsigned next = 0;
        int k = 0;
        for (int l = 1; l <= 16; ++l) {
            for (int j = 0; j < k_ht_bits[t][l - 1]; ++j, ++k) {
                uint8_t sym = k_ht_vals[t][k];
                len[t][sym] = (uint8_t)l;
                code[t][sym] = (uint16_t)next++;
            }
            next <<= 1;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Byte sink + header emission                                                */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint8_t* out; size_t cap; size_t n;             /* final JPEG bytes           */
    uint8_t* raw; size_t raw_cap; uint64_t raw_bits; /* unstuffed scan (optional) */
    uint64_t acc; int fill;                          /* MSB-first accumulator      */
} orc_sink;

static void put8(orc_sink* s, unsigned v)
{
    if (s->n < s->cap) s->out[s->n] = (uint8_t)v;
    s->n++;
}
static void put16(orc_sink* s, unsigned v) { put8(s, v >> 8); put8(s, v & 0xff); }   /* jpeg_enc.h:391-399 */
static void putn(orc_sink* s, const void* p, size_t n) { for (size_t i = 0; i < n; ++i) put8(s, ((const uint8_t*)p)[i]); }

static void put_dqt(orc_sink* s, const uint8_t* qt, int id)       /* jpeg_enc.h:498-509 */
{
    put16(s, 0xffdb); put16(s, 0x0043); put8(s, id); putn(s, qt, 64);
}
static void put_dht(orc_sink* s, int table, int cls, int id)      /* jpeg_enc.h:517-540 */
{
    int nv = 0;
    for (int i = 0; i < 16; ++i) nv += k_ht_bits[table][i];
    put16(s, 0xffc4); put16(s, 2 + 1 + 16 + nv); put8(s, (cls << 4) | id);
    putn(s, k_ht_bits[table], 16); putn(s, k_ht_vals[table], nv);
}

/* Everything before the entropy-coded segment (jpeg_enc.h:989-1077).
 * ncomp_out: 3 (YCbCr) or 1 (grayscale, extended). */
static void put_headers(orc_sink* s, int w, int h, int ncomp_out, int sub,
                        const uint8_t qt_luma[64], const uint8_t qt_chroma[64], int restart)
{
    static const char com[] = "Created by Tiny JPEG Encoder";
    put16(s, 0xffd8);
    put16(s, 0xffe0); put16(s, 16); putn(s, "JFIF", 5); put16(s, 0x0102);
    put8(s, 1); put16(s, 0x0060); put16(s, 0x0060); put8(s, 0); put8(s, 0);
    put16(s, 0xfffe); put16(s, 2 + (sizeof(com) - 1)); putn(s, com, sizeof(com) - 1);
    put_dqt(s, qt_luma, 0);
    if (ncomp_out == 3) put_dqt(s, qt_chroma, 1);
    put16(s, 0xffc0); put16(s, 8 + 3 * ncomp_out); put8(s, 8);
    put16(s, h); put16(s, w); put8(s, ncomp_out);
    for (int c = 0; c < ncomp_out; ++c) {
        put8(s, c + 1);
        put8(s, (c == 0 && sub == ORC_SUB_420) ? 0x22 : 0x11);
        put8(s, c == 0 ? 0 : 1);
    }
    put_dht(s, HT_LUMA_DC, 0, 0);
    put_dht(s, HT_LUMA_AC, 1, 0);
    if (ncomp_out == 3) {
        put_dht(s, HT_CHROMA_DC, 0, 1);
        put_dht(s, HT_CHROMA_AC, 1, 1);
    }
    if (restart > 0) { put16(s, 0xffdd); put16(s, 4); put16(s, restart); }   /* extended: DRI, MCUs per restart interval */
    put16(s, 0xffda); put16(s, 6 + 2 * ncomp_out); put8(s, ncomp_out);
    for (int c = 0; c < ncomp_out; ++c) { put8(s, c + 1); put8(s, c == 0 ? 0x00 : 0x11); }
    put8(s, 0); put8(s, 63); put8(s, 0);
}

/* Header only (for header-parity tests of the host emitter). */
size_t orc_headers_ex(int w, int h, int ncomp_out, int sub, int qmode, int quality, int restart,
                      uint8_t* out, size_t cap)
{
    uint8_t ql[64], qc[64];
    if (!orc_build_qt(qmode, quality, ql, qc)) return 0;
    orc_sink s; memset(&s, 0, sizeof s);
    s.out = out; s.cap = cap;
    put_headers(&s, w, h, ncomp_out, sub, ql, qc, restart);
    return s.n;
}
size_t orc_headers(int w, int h, int ncomp_out, int sub, int qmode, int quality,
                   uint8_t* out, size_t cap)
{
    return orc_headers_ex(w, h, ncomp_out, sub, qmode, quality, 0, out, cap);
}

/* ------------------------------------------------------------------------- */
/* Entropy-coded segment writer (jpeg_enc.h:613-643)                          */
/* ------------------------------------------------------------------------- */

static void put_bits(orc_sink* s, unsigned nbits, unsigned bits)
{
    if (nbits == 0) return;
    s->acc = (s->acc << nbits) | (bits & ((1u << nbits) - 1u));
    s->fill += (int)nbits;
    if (s->raw) {                                  /* unstuffed copy, bit-addressed */
        for (int b = (int)nbits - 1; b >= 0; --b) {
            uint64_t pos = s->raw_bits++;
            if ((pos >> 3) < s->raw_cap && ((bits >> b) & 1u))
                s->raw[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
        }
    } else {
        s->raw_bits += nbits;
    }
    while (s->fill >= 8) {
        unsigned c = (unsigned)(s->acc >> (s->fill - 8)) & 0xffu;
        put8(s, c);
        if (c == 0xff) put8(s, 0);                 /* jpeg_enc.h:634-638 */
        s->fill -= 8;
    }
}

/* Category and amplitude bits of a coefficient (jpeg_enc.h:598-610). */
static void vli(int v, unsigned* nbits, unsigned* bits)
{
    int a = v < 0 ? -v : v;
    if (v < 0) --v;
    unsigned n = 1;
    while (a >>= 1) ++n;
    *nbits = n;
    *bits = (unsigned)v & ((1u << n) - 1u);
}

/* ------------------------------------------------------------------------- */
/* One 8x8 data unit                                                          */
/* ------------------------------------------------------------------------- */

/* AAN forward DCT on 8 values with stride `st` (jpeg_enc.h:656-763).
 * Every intermediate is a separately rounded binary32 operation. */
static void aan8(float* d, int st)
{
    const float c4 = (float)0.707106781, c6 = (float)0.382683433;
    const float c2m6 = (float)0.541196100, c2p6 = (float)1.306562965;
    float s07 = d[0 * st] + d[7 * st], d07 = d[0 * st] - d[7 * st];
    float s16 = d[1 * st] + d[6 * st], d16 = d[1 * st] - d[6 * st];
    float s25 = d[2 * st] + d[5 * st], d25 = d[2 * st] - d[5 * st];
    float s34 = d[3 * st] + d[4 * st], d34 = d[3 * st] - d[4 * st];

    float e0 = s07 + s34, e3 = s07 - s34;
    float e1 = s16 + s25, e2 = s16 - s25;
    d[0 * st] = e0 + e1;
    d[4 * st] = e0 - e1;
    float r = (e2 + e3) * c4;
    d[2 * st] = e3 + r;
    d[6 * st] = e3 - r;

    float o0 = d34 + d25, o1 = d25 + d16, o2 = d16 + d07;
    float z5 = (o0 - o2) * c6;
    float z2 = c2m6 * o0 + z5;
    float z4 = c2p6 * o2 + z5;
    float z3 = o1 * c4;
    float z11 = d07 + z3, z13 = d07 - z3;
    d[5 * st] = z13 + z2;
    d[3 * st] = z13 - z2;
    d[1 * st] = z11 + z4;
    d[7 * st] = z11 - z4;
}

/* FDCT + quantise + zigzag (jpeg_enc.h:799-817): samples -> du[64] in zigzag order */
static void transform_block(const float* samples, const float* pqt, int du[64])
{
    float t[64];
    memcpy(t, samples, sizeof t);
    for (int r = 0; r < 8; ++r) aan8(t + 8 * r, 1);
    for (int c = 0; c < 8; ++c) aan8(t + c, 8);
    for (int i = 0; i < 64; ++i) {
        float v = t[i];
        v *= pqt[i];
        v = floorf(v + 1024 + 0.5f);      /* (v + 1024.0f) + 0.5f : two roundings */
        v -= 1024;
        du[k_zz[i]] = (int)v;
    }
}

/* Huffman symbols of one data unit (jpeg_enc.h:831-888). Returns bits emitted. */
static unsigned code_block(orc_sink* s, const int du[64], int* pred,
                           const uint8_t* dcl, const uint16_t* dcc,
                           const uint8_t* acl, const uint16_t* acc)
{
    uint64_t before = s->raw_bits;
    unsigned n, b;
    int diff = du[0] - *pred;
    *pred = du[0];
    if (diff) {
        vli(diff, &n, &b);
        put_bits(s, dcl[n], dcc[n]);
        put_bits(s, n, b);
    } else {
        put_bits(s, dcl[0], dcc[0]);
    }
    int last = 0;
    for (int i = 63; i > 0; --i) if (du[i]) { last = i; break; }
    int run = 0;
    for (int i = 1; i <= last; ++i) {
        if (du[i] == 0) {
            if (++run == 16) { put_bits(s, acl[0xf0], acc[0xf0]); run = 0; }
            continue;
        }
        vli(du[i], &n, &b);
        unsigned sym = ((unsigned)run << 4) | n;
        put_bits(s, acl[sym], acc[sym]);
        put_bits(s, n, b);
        run = 0;
    }
    if (last != 63) put_bits(s, acl[0], acc[0]);
    return (unsigned)(s->raw_bits - before);
}

/* ------------------------------------------------------------------------- */
/* Sample fetch                                                               */
/* ------------------------------------------------------------------------- */

typedef struct { const uint8_t* px; int w, h, ncomp; ptrdiff_t stride; } orc_img;

/* pixel with edge replication (jpeg_enc.h:1101-1116) */
static const uint8_t* pix(const orc_img* im, int x, int y)
{
    if (x >= im->w) x = im->w - 1;
    if (y >= im->h) y = im->h - 1;
    return im->px + (ptrdiff_t)y * im->stride + (ptrdiff_t)x * im->ncomp;
}

/* jpeg_enc.h:1118-1120; C's left-to-right evaluation made explicit */
static float rgb_y(const uint8_t* p)  { float r = p[0], g = p[1], b = p[2]; return ((0.299f * r + 0.587f * g) + 0.114f * b) - 128; }
static float rgb_cb(const uint8_t* p) { float r = p[0], g = p[1], b = p[2]; return (-0.1687f * r - 0.3313f * g) + 0.5f * b; }
static float rgb_cr(const uint8_t* p) { float r = p[0], g = p[1], b = p[2]; return (0.5f * r - 0.4187f * g) - 0.0813f * b; }

/* ------------------------------------------------------------------------- */
/* Whole-image encode                                                         */
/* ------------------------------------------------------------------------- */

/*
 * Returns 1 on success, 0 on rejected arguments (reference convention,
 * jpeg_enc.h:111-112).  *out_size always receives the number of bytes the stream
 * needs, even if it exceeded `cap`.
 *
 * Optional dumps (NULL to skip), all indexed by block in STREAM order
 * (jpeg_enc.h:1128-1154: per MCU Y,Cb,Cr; 4:2:0: Y00 Y01 Y10 Y11 Cb Cr):
 *   coef_dump   int16[nblocks*64]  quantised coefficients in zigzag order
 *   bits_dump   uint32[nblocks]    entropy-coded bits of each block
 *   raw_dump    the scan before 0xFF00 stuffing and before padding; raw_bits = its length
 */
/*
 * restart > 0 (EXTENDED, not in the reference encoder; opt-in): a DRI segment announces `restart`
 * MCUs per restart interval; after every interval but the last the entropy-coded data is padded to
 * a byte boundary with 1-bits (ITU-T T.81 F.1.2.3), RSTm (m = interval index mod 8) follows and the
 * DC predictors return to 0; the final interval is padded with 1-bits as well.  The decoded pixels
 * are those of the restart-free stream (checked with the reference's own decoder, jpeg_dec.h).
 */
int orc_encode_ex(const uint8_t* px, int w, int h, int ncomp, ptrdiff_t stride,
                  int qmode, int quality, int sub, int restart,
                  uint8_t* out, size_t cap, size_t* out_size,
                  int16_t* coef_dump, uint32_t* bits_dump,
                  uint8_t* raw_dump, size_t raw_cap, uint64_t* raw_bits);

int orc_encode(const uint8_t* px, int w, int h, int ncomp, ptrdiff_t stride,
               int qmode, int quality, int sub,
               uint8_t* out, size_t cap, size_t* out_size,
               int16_t* coef_dump, uint32_t* bits_dump,
               uint8_t* raw_dump, size_t raw_cap, uint64_t* raw_bits)
{
    return orc_encode_ex(px, w, h, ncomp, stride, qmode, quality, sub, 0, out, cap, out_size,
                         coef_dump, bits_dump, raw_dump, raw_cap, raw_bits);
}

int orc_encode_ex(const uint8_t* px, int w, int h, int ncomp, ptrdiff_t stride,
                  int qmode, int quality, int sub, int restart,
                  uint8_t* out, size_t cap, size_t* out_size,
                  int16_t* coef_dump, uint32_t* bits_dump,
                  uint8_t* raw_dump, size_t raw_cap, uint64_t* raw_bits)
{
    uint8_t qtl[64], qtc[64];
    float pql[64], pqc[64];
    uint8_t hl[4][256]; uint16_t hc[4][256];

    if (out_size) *out_size = 0;
    if (raw_bits) *raw_bits = 0;
    if (!orc_build_qt(qmode, quality, qtl, qtc)) return 0;
    if (ncomp != 1 && ncomp != 3 && ncomp != 4) return 0;           /* 1 = extended gray */
    if (w <= 0 || h <= 0 || w > 0xffff || h > 0xffff) return 0;    /* jpeg_enc.h:958-960 */
    if (sub != ORC_SUB_444 && sub != ORC_SUB_420) return 0;
    if (ncomp == 1 && sub != ORC_SUB_444) return 0;
    if (stride == 0) stride = (ptrdiff_t)w * ncomp;
    if (restart < 0 || restart > 0xffff) return 0;

    orc_build_pqt(qtl, pql);
    orc_build_pqt(qtc, pqc);
    orc_build_huff(hl, hc);

    orc_sink s; memset(&s, 0, sizeof s);
    s.out = out; s.cap = cap; s.raw = raw_dump; s.raw_cap = raw_cap;
    if (raw_dump) memset(raw_dump, 0, raw_cap);
    orc_img im = { px, w, h, ncomp, stride };

    put_headers(&s, w, h, ncomp == 1 ? 1 : 3, sub, qtl, qtc, restart);

    int pred[3] = { 0, 0, 0 };
    float blk[64];
    int du[64];
    size_t nb = 0;
    const int mcu_px = (ncomp != 1 && sub == ORC_SUB_420) ? 16 : 8;
    const long total_mcus = (long)((w + mcu_px - 1) / mcu_px) * ((h + mcu_px - 1) / mcu_px);
    long mcus_done = 0;

    /* end of an MCU: close the restart interval if one ends here and more MCUs follow */
#define MCU_DONE()                                                                     \
    do {                                                                               \
        ++mcus_done;                                                                   \
        if (restart > 0 && mcus_done % restart == 0 && mcus_done < total_mcus) {       \
            if (s.fill > 0) put_bits(&s, (unsigned)(8 - s.fill), 0xffu);               \
            put16(&s, 0xffd0 + (unsigned)((mcus_done / restart - 1) & 7));             \
            pred[0] = pred[1] = pred[2] = 0;                                           \
        }                                                                              \
    } while (0)

#define EMIT(comp)                                                                     \
    do {                                                                               \
        const int lum = (comp) == 0;                                                   \
        transform_block(blk, lum ? pql : pqc, du);                                     \
        unsigned nbits_ = code_block(&s, du, &pred[comp],                              \
                                     hl[lum ? HT_LUMA_DC : HT_CHROMA_DC], hc[lum ? HT_LUMA_DC : HT_CHROMA_DC], \
                                     hl[lum ? HT_LUMA_AC : HT_CHROMA_AC], hc[lum ? HT_LUMA_AC : HT_CHROMA_AC]); \
        if (coef_dump) for (int i_ = 0; i_ < 64; ++i_) coef_dump[nb * 64 + i_] = (int16_t)du[i_]; \
        if (bits_dump) bits_dump[nb] = nbits_;                                         \
        ++nb;                                                                          \
    } while (0)

    if (ncomp == 1) {
        /* extended: grayscale, sample = v - 128, luma tables only */
        for (int y0 = 0; y0 < h; y0 += 8)
            for (int x0 = 0; x0 < w; x0 += 8) {
                for (int j = 0; j < 8; ++j)
                    for (int i = 0; i < 8; ++i)
                        blk[j * 8 + i] = (float)pix(&im, x0 + i, y0 + j)[0] - 128;
                EMIT(0);
                MCU_DONE();
            }
    } else if (sub == ORC_SUB_444) {
        /* native: jpeg_enc.h:1094-1158 */
        for (int y0 = 0; y0 < h; y0 += 8)
            for (int x0 = 0; x0 < w; x0 += 8) {
                for (int c = 0; c < 3; ++c) {
                    for (int j = 0; j < 8; ++j)
                        for (int i = 0; i < 8; ++i) {
                            const uint8_t* p = pix(&im, x0 + i, y0 + j);
                            blk[j * 8 + i] = c == 0 ? rgb_y(p) : c == 1 ? rgb_cb(p) : rgb_cr(p);
                        }
                    EMIT(c);
                }
                MCU_DONE();
            }
    } else {
        /* extended: 4:2:0, 16x16 MCUs, chroma = ((a+b)+(c+d))*0.25f of float Cb/Cr */
        for (int y0 = 0; y0 < h; y0 += 16)
            for (int x0 = 0; x0 < w; x0 += 16) {
                for (int q = 0; q < 4; ++q) {
                    int bx = x0 + 8 * (q & 1), by = y0 + 8 * (q >> 1);
                    for (int j = 0; j < 8; ++j)
                        for (int i = 0; i < 8; ++i)
                            blk[j * 8 + i] = rgb_y(pix(&im, bx + i, by + j));
                    EMIT(0);
                }
                for (int c = 1; c < 3; ++c) {
                    for (int j = 0; j < 8; ++j)
                        for (int i = 0; i < 8; ++i) {
                            int x = x0 + 2 * i, y = y0 + 2 * j;
                            const uint8_t *pa = pix(&im, x, y), *pb = pix(&im, x + 1, y);
                            const uint8_t *pc = pix(&im, x, y + 1), *pd = pix(&im, x + 1, y + 1);
                            float a = c == 1 ? rgb_cb(pa) : rgb_cr(pa), b = c == 1 ? rgb_cb(pb) : rgb_cr(pb);
                            float cc = c == 1 ? rgb_cb(pc) : rgb_cr(pc), d = c == 1 ? rgb_cb(pd) : rgb_cr(pd);
                            blk[j * 8 + i] = ((a + b) + (cc + d)) * 0.25f;
                        }
                    EMIT(c);
                }
                MCU_DONE();
            }
    }
#undef EMIT
#undef MCU_DONE

    if (raw_bits) *raw_bits = s.raw_bits;
    /* jpeg_enc.h:1161-1167: pad the last byte with ZERO bits, then EOI */
    if (s.fill > 0) {
        uint8_t* keep = s.raw; uint64_t keepbits = s.raw_bits;
        s.raw = NULL;
        put_bits(&s, (unsigned)(8 - s.fill), restart > 0 ? 0xffu : 0u);   /* restart mode pads every interval with 1-bits */
        s.raw = keep; s.raw_bits = keepbits;
    }
    put16(&s, 0xffd9);
    if (out_size) *out_size = s.n;
    return 1;
}

/* Number of 8x8 blocks orc_encode() will emit for a given geometry. */
size_t orc_num_blocks(int w, int h, int ncomp, int sub)
{
    if (ncomp == 1) return (size_t)((w + 7) / 8) * ((h + 7) / 8);
    if (sub == ORC_SUB_420) return (size_t)((w + 15) / 16) * ((h + 15) / 16) * 6;
    return (size_t)((w + 7) / 8) * ((h + 7) / 8) * 3;
}

/* ------------------------------------------------------------------------- */
/* BMP reader exactly as the codecs.h path sees it (codecs.cpp:255-320):      */
/* 24-bit BITMAPINFOHEADER, rows bottom-up -> top-down, bytes left B,G,R.     */
/* ------------------------------------------------------------------------- */
int orc_read_bmp_mem(const uint8_t* file, size_t n, uint8_t* px, size_t cap, int* w, int* h)
{
    if (n < 54 || file[0] != 'B' || file[1] != 'M') return 0;   /* magic 19778, codecs.cpp:257,286 */
    int32_t bw, bh;
    memcpy(&bw, file + 18, 4);
    memcpy(&bh, file + 22, 4);
    int ah = bh < 0 ? -bh : bh;
    size_t row = (size_t)bw * 3, pad = (size_t)(bw % 4);
    if (w) *w = bw;
    if (h) *h = bh;
    if (bw <= 0 || ah == 0 || row * ah > cap || 54 + (row + pad) * ah > n + pad) return 0;
    const int off = bh > 0 ? 0 : ah - 1;
    const uint8_t* src = file + 54;                  /* the reader never seeks to bfOffBits */
    for (int y = ah - 1; y >= 0; --y) {
        int dst = y - off; if (dst < 0) dst = -dst;
        memcpy(px + (size_t)dst * row, src, row);
        src += row + pad;
    }
    return 1;
}
