// jpeg_host.cpp -- host-side half of the encoder: quantiser / Huffman table construction
// and marker emission.  Everything the reference does once per image before its block
// loop (jpeg_enc.h:962-1077, :1230-1266) lives here and stays on the CPU.
#include <algorithm>
#include <vector>
#include "jpeg_tables.h"

#include <string.h>

namespace jg {

// jpeg_enc.h:266-276
static const uint8_t kBaseLuma[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};

// jpeg_enc.h:294-305 ("example QT from JPEG paper" -- what the reference uses for chroma)
static const uint8_t kBaseChroma[64] = {
    16,  12,  14,  14, 18, 24,  49,  72,  11,  10,  16, 24, 40, 51,  61,  12,
    13,  17,  22,  35, 64, 92,  14,  16,  22,  37,  55, 78, 95, 19,  24,  29,
    56,  64,  87,  98, 26, 40,  51,  68,  81,  103, 112, 58, 57, 87,  109, 104,
    121, 100, 60,  69, 80, 103, 113, 120, 103, 55,  56, 62, 77, 92,  101, 99};

static const uint8_t kZigzag[64] = JG_ZZ_INIT;

// Annex K.3.3 (jpeg_enc.h:310-368): BITS then HUFFVAL per table, enum order of HuffTable.
static const uint8_t kBits[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}};

static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};

static const uint8_t kLumaAcVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3,
    0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
    0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};

static const uint8_t kChromaAcVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1,
    0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A,
    0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA,
    0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};

static const uint8_t* huff_vals(int t)
{
    switch (t) {
        case HT_LUMA_AC: return kLumaAcVals;
        case HT_CHROMA_AC: return kChromaAcVals;
        default: return kDcVals;
    }
}

static int huff_count(int t)
{
    int n = 0;
    for (int i = 0; i < 16; ++i) n += kBits[t][i];
    return n;
}

bool build_qt(int quality_mode, int quality, uint8_t qt_luma[64], uint8_t qt_chroma[64])
{
    auto clamp255 = [](int v) { return (uint8_t)(v < 1 ? 1 : (v > 255 ? 255 : v)); };
    if (quality_mode == 0) {  // TJE: jpeg_enc.h:1223, :1230-1256
        if (quality < 1 || quality > 3) return false;
        const int divisor = quality == 2 ? 10 : 1;
        for (int i = 0; i < 64; ++i) {
            qt_luma[i] = quality == 3 ? 1 : clamp255(kBaseLuma[i] / divisor);
            qt_chroma[i] = quality == 3 ? 1 : clamp255(kBaseChroma[i] / divisor);
        }
        return true;
    }
    if (quality_mode == 1) {  // IJG scaling of the same base tables (extended mode)
        if (quality < 1 || quality > 100) return false;
        const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
        for (int i = 0; i < 64; ++i) {
            qt_luma[i] = clamp255((kBaseLuma[i] * scale + 50) / 100);
            qt_chroma[i] = clamp255((kBaseChroma[i] * scale + 50) / 100);
        }
        return true;
    }
    return false;
}

void build_pqt(const uint8_t qt[64], float pqt[64])
{
    // jpeg_enc.h:974-977
    static const float aan[8] = {1.0f,          1.387039845f, 1.306562965f, 1.175875602f,
                                 1.0f,          0.785694958f, 0.541196100f, 0.275899379f};
    for (int i = 0; i < 64; ++i) {
        const int x = i & 7, y = i >> 3;
        // jpeg_enc.h:983: 1.0f / (8 * aan[x] * aan[y] * qt[zigzag[i]]), left to right in float
        float d = 8 * aan[x];
        d = d * aan[y];
        d = d * (float)qt[kZigzag[i]];
        pqt[i] = 1.0f / d;
    }
}

void build_huff_lut(HuffLut* lut)
{
    memset(lut, 0, sizeof(*lut));
    for (int t = 0; t < 4; ++t) {
        const uint8_t* vals = huff_vals(t);
        uint32_t code = 0;
        int k = 0;
        for (int len = 1; len <= 16; ++len, code <<= 1) {  // Annex C.2, jpeg_enc.h:546-592
            for (int j = 0; j < kBits[t][len - 1]; ++j) {
                const uint8_t sym = vals[k++];
                const uint32_t e = (code++ << 8) | (uint32_t)len;
                if (t == HT_LUMA_DC) lut->dc[0][sym & 15] = e;
                else if (t == HT_CHROMA_DC) lut->dc[1][sym & 15] = e;
                else if (t == HT_LUMA_AC) lut->ac[0][sym] = e;
                else lut->ac[1][sym] = e;
            }
        }
    }
}

namespace {
struct ByteWriter {
    uint8_t* p;
    size_t cap, n;
    void u8(unsigned v) { if (n < cap) p[n] = (uint8_t)v; ++n; }
    void be16(unsigned v) { u8(v >> 8); u8(v); }
    void bytes(const void* s, size_t k) { for (size_t i = 0; i < k; ++i) u8(((const uint8_t*)s)[i]); }
};
}  // namespace

size_t emit_headers(int w, int h, int ncomp_out, int subsampling, const uint8_t qt_luma[64],
                    const uint8_t qt_chroma[64], uint8_t* out, size_t cap, int restart_interval)
{
    ByteWriter bw{out, cap, 0};
    const bool color = ncomp_out == 3;

    // SOI + APP0/JFIF 1.02, 96x96 dpi, no thumbnail (jpeg_enc.h:989-1005)
    bw.be16(0xFFD8);
    bw.be16(0xFFE0); bw.be16(16); bw.bytes("JFIF", 5); bw.be16(0x0102); bw.u8(1);
    bw.be16(96); bw.be16(96); bw.u8(0); bw.u8(0);
    // COM (jpeg_enc.h:1006-1014)
    static const char kComment[] = "Created by Tiny JPEG Encoder";
    bw.be16(0xFFFE); bw.be16(2 + sizeof(kComment) - 1); bw.bytes(kComment, sizeof(kComment) - 1);
    // DQT x2: table bytes in stored order (jpeg_enc.h:498-509, :1017-1018)
    bw.be16(0xFFDB); bw.be16(67); bw.u8(0); bw.bytes(qt_luma, 64);
    if (color) { bw.be16(0xFFDB); bw.be16(67); bw.u8(1); bw.bytes(qt_chroma, 64); }
    // SOF0: height before width (jpeg_enc.h:1020-1045)
    bw.be16(0xFFC0); bw.be16(8 + 3 * ncomp_out); bw.u8(8); bw.be16(h); bw.be16(w); bw.u8(ncomp_out);
    for (int c = 0; c < ncomp_out; ++c) {
        bw.u8(c + 1);
        bw.u8(c == 0 && subsampling == 1 ? 0x22 : 0x11);
        bw.u8(c ? 1 : 0);
    }
    // DHT: luma DC, luma AC, chroma DC, chroma AC (jpeg_enc.h:1047-1050)
    for (int t = 0; t < (color ? 4 : 2); ++t) {
        const int nv = huff_count(t);
        bw.be16(0xFFC4); bw.be16(19 + nv);
        bw.u8(((t & 1) << 4) | (t >> 1));  // class = AC?, id = chroma?
        bw.bytes(kBits[t], 16); bw.bytes(huff_vals(t), nv);
    }
    if (restart_interval > 0) { bw.be16(0xFFDD); bw.be16(4); bw.be16((unsigned)restart_interval); }   // DRI (extended, opt-in)
    // SOS (jpeg_enc.h:1052-1077)
    bw.be16(0xFFDA); bw.be16(6 + 2 * ncomp_out); bw.u8(ncomp_out);
    for (int c = 0; c < ncomp_out; ++c) { bw.u8(c + 1); bw.u8(c ? 0x11 : 0x00); }
    bw.u8(0); bw.u8(63); bw.u8(0);
    return bw.n <= cap ? bw.n : 0;
}

size_t schedule_words(int n_images) { return 1 + (size_t)(n_images + 1) + 2 * (size_t)n_images + (size_t)n_images; }

size_t build_schedule(const int* tiles, int n, uint32_t* out)
{
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return tiles[a] < tiles[b]; });
    std::vector<uint32_t> cum, lt0, first;
    uint32_t tickets = 0;
    int lt = 0;
    for (int a = 0; a < n;) {
        // images order[a..n) are active for rounds [lt, tiles[order[a]])
        cum.push_back(tickets); lt0.push_back((uint32_t)lt); first.push_back((uint32_t)a);
        const int until = tiles[order[a]];
        tickets += (uint32_t)(until - lt) * (uint32_t)(n - a);
        lt = until;
        while (a < n && tiles[order[a]] == until) ++a;
    }
    const size_t D = cum.size();
    cum.push_back(tickets);
    size_t w = 0;
    out[w++] = (uint32_t)D;
    for (size_t k = 0; k <= D; ++k) out[w++] = cum[k];
    for (size_t k = 0; k < D; ++k) out[w++] = lt0[k];
    for (size_t k = 0; k < D; ++k) out[w++] = first[k];
    for (int i = 0; i < n; ++i) out[w++] = (uint32_t)order[i];
    return w;
}

}  // namespace jg
