// Experiment (CPU harness, not part of the product): how often does SOME block-in-MCU variant of the guessed entry state of a
// subsequence end in the true state?  g++ -std=c++17 -O2 -Itests/emu -Iimagecodecs_b200/csrc tools/variant_probe.cpp imagecodecs_b200/csrc/jpeg_decode_host.cpp
// usage: variant_probe file.jpg sub_log2
#define JG_EMULATE 1
#include "cuda_emu.h"
#include "jpeg_decode.cuh"
#include "jpeg_decode.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(n); if (fread(buf.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
    int sub_log2 = atoi(argv[2]);
    jd::Info I; if (jd::parse(buf.data(), n, &I) != 0) return 1;
    jd::DevParams P; memset(&P, 0, sizeof P);
    std::vector<int16_t> coef(I.n_blocks * 64, 0); unsigned err = 0;
    P.data = buf.data(); P.vlc = I.vlc.data(); P.coef = coef.data(); P.error = &err; P.mbwidth = I.mbwidth; P.ncomp = I.ncomp; P.n_mcus = I.n_mcus;
    size_t scan_bytes = I.scan_end - I.scan_off;
    P.sub_log2 = sub_log2; P.n_sub = (int)((scan_bytes + ((size_t)1 << sub_log2) - 1) >> sub_log2);
    P.bpm = jd::mcu_block_map(I, P.blk); P.scan = buf.data() + I.scan_off; P.scan_bytes = (unsigned)scan_bytes; P.total_blocks = (unsigned long long)I.n_mcus * P.bpm;
    std::vector<unsigned long long> exits(P.n_sub, 0); std::vector<jd::SubStart> sums(P.n_sub), start(P.n_sub); std::vector<unsigned> la(P.n_sub), lb(P.n_sub); unsigned cnt[3] = {0,0,0};
    P.sub_exit = exits.data(); P.sub_sum = sums.data(); P.sub_start = start.data(); P.sub_list[0] = la.data(); P.sub_list[1] = lb.data(); P.sub_cnt = cnt;
    std::vector<uint16_t> l1(4 << jd::kL1Bits);
    for (int i = 0; i < (4 << jd::kL1Bits); ++i) l1[i] = jd::l1_entry(P.vlc, i >> jd::kL1Bits, i & ((1 << jd::kL1Bits) - 1));
    for (int r = 0;; ++r) { cnt[(r + 2) % 3] = 0; unsigned c = jd::sync_round_count(P, r), a = 0; for (unsigned k = 0; k < c; ++k) a += jd::sync_round_item(P, l1.data(), r, k, c); if (r >= 1 && !a) break; }
    // variants
    int any = 0, none = 0; std::vector<int> hist(P.bpm + 1, 0), rel(P.bpm, 0);
    const jd::SubStart zero = {0u, 0, 0, 0};
    for (int i = 1; i + 1 < P.n_sub; ++i) {
        unsigned tb = (unsigned)((exits[i - 1] & 0xFFFF) >> 8);      // block the true decoder is in when the piece starts
        int matches = 0;
        for (int v = 0; v < P.bpm; ++v) {
            unsigned long long e = 0; jd::SubStart s = zero;
            jd::decode_subsequence<false>(P, l1.data(), i, true, jd::guessed_entry_pos(P, i), (unsigned)v << 8, zero, &e, &s);
            if (e == exits[i]) { ++matches; rel[(v + P.bpm - tb) % P.bpm]++; }
        }
        hist[matches]++; if (matches) ++any; else ++none;
    }
    printf("%s S=%d pieces %d bpm %d: some variant ends in the true state %.1f%%; matches per piece:", argv[1], 1 << sub_log2, P.n_sub - 2, P.bpm, 100.0 * any / (any + none));
    for (int m = 0; m <= P.bpm; ++m) printf(" %d:%d", m, hist[m]);
    printf("; by (variant - true block) mod bpm:"); for (int v = 0; v < P.bpm; ++v) printf(" %d:%d", v, rel[v]);
    printf("\n");
}
