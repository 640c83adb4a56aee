// jpeg_tables.h -- constant data of the encode path and the host-side table builders.
//
// The numbers are the reference's (jpeg_enc.h:266-386): Annex K.1 luminance table, the
// "from paper" chrominance table the reference actually uses (:294-305, NOT Annex K.2),
// the Annex K.3.3 Huffman specifications and the zigzag permutation.  They have to be
// these values for the output to be byte-identical.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace jg {

// zz[natural index] = position in zigzag scan order (jpeg_enc.h:376-386)
#define JG_ZZ_INIT                                                                              \
    { 0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43,    \
      9, 11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51,    \
      55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 }

enum HuffTable { HT_LUMA_DC = 0, HT_LUMA_AC = 1, HT_CHROMA_DC = 2, HT_CHROMA_AC = 3 };  // jpeg_enc.h:891-896

enum Layout { LAYOUT_444 = 0, LAYOUT_420 = 1, LAYOUT_GRAY = 2 };

// Reciprocal, AAN-scaled quantiser (natural order).  Passed BY VALUE as a kernel
// parameter so that every multiply reads its factor straight from the constant bank.
struct QuantSet {
    float luma[64];
    float chroma[64];
};

// Huffman lookup as the kernels want it: entry = (code << 8) | length, 0 if undefined.
//   ac[0] luma AC, ac[1] chroma AC, indexed by (run << 4) | category
//   dc[0] luma DC, dc[1] chroma DC, indexed by category
struct HuffLut {
    uint32_t ac[2][256];
    uint32_t dc[2][16];
};

// ---- host-side builders (jpeg_host.cpp) ---------------------------------------------
// quantiser bytes in the reference's stored order; returns false if (mode, quality) is invalid
bool build_qt(int quality_mode, int quality, uint8_t qt_luma[64], uint8_t qt_chroma[64]);
// 1.0f / (8 * aan[x] * aan[y] * qt[zz[i]]) evaluated like jpeg_enc.h:983-984
void build_pqt(const uint8_t qt[64], float pqt[64]);
void build_huff_lut(HuffLut* lut);
// header bytes up to and including SOS; ncomp_out is 3 or 1. Returns 0 if it does not fit.
// restart_interval > 0: a DRI segment (MCUs per restart interval) precedes SOS (extended, opt-in).
size_t emit_headers(int w, int h, int ncomp_out, int subsampling, const uint8_t qt_luma[64],
                    const uint8_t qt_chroma[64], uint8_t* out, size_t cap, int restart_interval = 0);
// Ticket schedule of a launch whose images have different tile counts (see draw_tile() in
// jpeg_kernel.cuh): tickets walk the images round-robin, tile 0 of every image, then tile 1 of
// every image that has one, ...  `out` (n_words of it, sized schedule_words(n)) receives
// D, cum[D+1], lt0[D], first[D], order[n]: while the round lt is in [lt0[k], lt0[k+1]) the images
// order[first[k]..n) are active and cum[k] tickets have been handed out before lt0[k].
size_t schedule_words(int n_images);
size_t build_schedule(const int* tiles_per_image, int n_images, uint32_t* out);

}  // namespace jg
