// jpeg_decode.cuh -- device half of the decoder (see jpeg_decode.h).  Every stage is written as a
// per-thread function of one work item, so that the CPU test harness (JG_EMULATE) can run the very
// same code in plain loops; the __global__ wrappers below only hand out indices.
//
//   decode_interval   one restart interval: Huffman decode (njGetVLC :643-656, njDecodeBlock :658-672
//                     without its IDCT) into quantised coefficients, natural order, int16
//   decode_subsequence  the same decode for a scan WITHOUT restart markers, cut into fixed-size subsequences
//                     (speculative rounds until neighbours agree, a scan, one writing pass; see "subsequences")
//   idct_block        dequantise + njRowIDCT x 8 + njColIDCT x 8 (:350-442) -> 8x8 bytes of the plane
//   upsample_h / _v   njUpsampleH / njUpsampleV (:736-790), one output pixel per thread
//   to_rgb / to_gray  the tail of njConvert (:817-866)
#pragma once
#include "jpeg_decode.h"
#include "jpeg_device.h"

namespace jd {

using jg::bswap32;
using jg::ldg_u32;
using jg::ldg_u8;
using jg::v_cmpeq4;

struct DevComponent {
    int ssx, ssy, bw;                 // blocks per MCU in x / y, blocks per row of the padded plane
    int dctab, actab;                 // rows of the VLC table
    int stride;                       // bytes per row of the padded plane
    unsigned long long n_blocks;      // blocks of the padded plane
    unsigned long long coef_off;      // first block in `coef`
    unsigned long long plane_off;     // first byte in `planes`
    int dq[64];                       // dequantisers in NATURAL order: dq[njZZ[k]] = qtab[k] (:666)
};

// per-subsequence record of the subsequence decode (see "subsequences" below): blocks completed and DC differences per component
struct SubStart { unsigned n; int dc0, dc1, dc2; };

struct DevParams {
    const uint8_t* data;              // the whole file
    const uint32_t* interval_off;     // [n_intervals + 1]
    int n_intervals, rstinterval, n_mcus, mbwidth, ncomp;
    const uint16_t* vlc;              // [4][65536]
    int16_t* coef;                    // [n_blocks][64], zeroed
    uint8_t* planes;
    unsigned* error;
    DevComponent comp[3];
    // self-synchronising decode of a restart-free scan (n_sub == 0: not used, the interval path decodes the image)
    int n_sub, sub_log2, bpm, sub_pad;            // subsequences, log2 of their size in bytes, blocks per MCU
    const uint8_t* scan;                          // first byte of the entropy-coded data
    unsigned scan_bytes, sub_pad2;
    unsigned long long total_blocks;              // n_mcus * bpm
    unsigned long long* sub_exit;                 // [n_sub] state at the end of each subsequence: position << 16 | block-in-MCU << 8 | s
    SubStart* sub_sum;                            // [n_sub] what the subsequence's last decode counted
    SubStart* sub_start;                          // [n_sub] exclusive sums of sub_sum
    unsigned* sub_list[2];                        // [n_sub] each: the subsequences to decode again in an even / odd round
    unsigned* sub_cnt;                            // [3] their number, slot = round % 3
    McuBlock blk[kMaxBlocksPerMcu];
};

// njZZ (jpeg_dec.h:332-337): natural index of zigzag position k
JG_CONST_TABLE unsigned char kZigzagToNatural[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                                                     28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54,
                                                     47, 55, 62, 63};
JG_DEV int zz_nat(int k) { return kZigzagToNatural[k]; }

// MSB-first bit reader over [p, end) with the byte rules of njShowBits (:447-482): FF 00 and FF FF
// yield one FF, past the end the stream continues with FF bytes.  Four bytes per load where the
// pointer is aligned and none of them is FF (the common case); byte by byte otherwise.
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    unsigned long long buf;
    int bits;
    unsigned ahead;            // the aligned word at `pa`, requested one refill early: its latency passes behind the decoding
    const uint8_t* pa;         // nullptr: nothing requested
};
JG_DEV void request_ahead(BitReader& r)
{
    const uint8_t* a = (const uint8_t*)(((size_t)r.p + 3u) & ~(size_t)3u);      // next aligned word at or after p
    if (a + 4 <= r.end) { r.ahead = ldg_u32(a); r.pa = a; } else r.pa = nullptr;
}
JG_DEV void refill(BitReader& r)       // called with fewer than 16 bits left; leaves at least 25
{
    if (r.pa == r.p) {                                     // p is aligned and its word is already here
        const unsigned w = r.ahead;
        if (v_cmpeq4(w, 0xffffffffu) == 0u) {
            r.buf = (r.buf << 32) | bswap32(w);
            r.bits += 32;
            r.p += 4;
            request_ahead(r);
            return;
        }
    }
    // byte by byte: until there are enough bits AND p is aligned again (the word path needs that), or the buffer is full
    do {
        unsigned b = 0xFF;
        if (r.p < r.end) {
            b = ldg_u8(r.p++);
            if (b == 0xFF && r.p < r.end) ++r.p;        // the stuffed 00 (or a fill FF): consumed, not data
        }
        r.buf = (r.buf << 8) | b;
        r.bits += 8;
    } while (r.bits <= 24 || ((((size_t)r.p) & 3u) != 0 && r.bits <= 48 && r.p < r.end));
    request_ahead(r);
}
JG_DEV unsigned show(BitReader& r, int n)
{
    if (r.bits < n) refill(r);
    return (unsigned)(r.buf >> (r.bits - n)) & ((1u << n) - 1u);
}
JG_DEV void skip(BitReader& r, int n) { r.bits -= n; }

// First-level lookup: the kL1Bits leading bits decide every code of at most that length (all but the
// rarest symbols); the table is the corresponding slice of the 16-bit table and lives in shared memory.
constexpr int kL1Bits = 9;
struct VlcTables {
    const uint16_t* full;      // [4][65536], global memory
    const uint16_t* l1;        // [4][1 << kL1Bits]: entry of the 16-bit table if its code length <= kL1Bits, else 0
};
JG_DEV uint16_t l1_entry(const uint16_t* full, int table, int i)
{
    const uint16_t e = full[(size_t)table * 65536 + ((size_t)i << (16 - kL1Bits))];
    return (e >> 8) <= kL1Bits ? e : (uint16_t)0;
}

// njGetVLC (:643-656)
JG_DEV int get_vlc(BitReader& r, const VlcTables& T, int table, unsigned* code_out, bool* bad)
{
    const unsigned peek = show(r, 16);
    unsigned e = T.l1[(table << kL1Bits) + (peek >> (16 - kL1Bits))];
    if (!e) e = T.full[(size_t)table * 65536 + peek];
    const int len = (int)(e >> 8);
    if (!len) { *bad = true; return 0; }
    skip(r, len);
    const unsigned code = e & 0xFFu;
    *code_out = code;
    const int nb = (int)(code & 15u);
    if (!nb) return 0;
    int v = (int)show(r, nb);
    skip(r, nb);
    if (v < (1 << (nb - 1))) v += (int)((0xFFFFFFFFu << nb) + 1u);
    return v;
}

JG_DEV void decode_interval(const DevParams& P, const uint16_t* l1, int iv)
{
    // Everything the loops need is read ONCE into locals: P lives in global memory, and after every
    // coefficient store the compiler would otherwise have to assume it changed and load it again.
    VlcTables T; T.full = P.vlc; T.l1 = l1;
    const int rst = P.rstinterval, n_mcus = P.n_mcus, mbwidth = P.mbwidth, ncomp = P.ncomp, n_iv = P.n_intervals;
    int16_t* const coef = P.coef;
    int ssx[3], ssy[3], bw[3], dctab[3], actab[3];
    unsigned long long coff[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ssx[c] = P.comp[c].ssx; ssy[c] = P.comp[c].ssy; bw[c] = P.comp[c].bw;
        dctab[c] = P.comp[c].dctab; actab[c] = P.comp[c].actab; coff[c] = P.comp[c].coef_off;
    }
    // Every lane of a warp walks the SAME number of blocks (a full interval) and meets the others after each
    // one: left alone, the lanes drift apart in the data-dependent symbol loops and the warp ends up running
    // them one after the other.  Lanes without an interval, and the image's shorter last interval, idle along.
    const bool have = iv < n_iv;
    BitReader r;
    r.p = r.end = P.data;
    if (have) {
        r.p = P.data + P.interval_off[iv];
        r.end = P.data + (iv + 1 < n_iv ? P.interval_off[iv + 1] - 2u : P.interval_off[n_iv]);   // minus the RSTm marker
    }
    r.buf = 0; r.bits = 0; r.ahead = 0; r.pa = nullptr;
    request_ahead(r);
    int dcpred[3] = {0, 0, 0};
    const int m0 = rst ? iv * rst : 0;
    const int m1 = !have ? m0 : (rst ? (m0 + rst < n_mcus ? m0 + rst : n_mcus) : n_mcus);
    const int m_end = m0 + (rst ? rst : n_mcus);
    bool bad = false;
    int mby = m0 / mbwidth, mbx = m0 - mby * mbwidth;
    for (int m = m0; m < m_end; ++m) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c >= ncomp) continue;
            for (int sby = 0; sby < ssy[c]; ++sby)
                for (int sbx = 0; sbx < ssx[c]; ++sbx) {
                    JG_RECONVERGE();
                    if (m >= m1 || bad) continue;
                    int16_t* blk = coef + (coff[c] + (unsigned long long)(mby * ssy[c] + sby) * bw[c] + (mbx * ssx[c] + sbx)) * 64ull;
                    unsigned code = 0;
                    dcpred[c] += get_vlc(r, T, dctab[c], &code, &bad);
                    blk[0] = (int16_t)dcpred[c];
                    int k = 0;
                    do {
                        const int v = get_vlc(r, T, actab[c], &code, &bad);
                        if (bad || !code) break;                                  // EOB
                        if (!(code & 0x0F) && code != 0xF0) { bad = true; break; }
                        k += (int)(code >> 4) + 1;
                        if (k > 63) { bad = true; break; }
                        blk[zz_nat(k)] = (int16_t)v;
                    } while (k < 63);
                }
        }
        if (++mbx >= mbwidth) { mbx = 0; ++mby; }
    }
    if (bad) *P.error = 5u;     // NJ_SYNTAX_ERROR
}

// ---- subsequences: parallel decode of a scan without restart markers -------------------------------------
// (the self-synchronisation scheme of Weissenberger & Schmidt, "Accelerating JPEG decompression on GPUs", restated
// for NanoJPEG's decode loop.)  The scan is cut every 2^sub_log2 bytes.  A decoder is in a known STATE between two
// symbols: the position of the next bit, the block of the MCU it is in, and how far into that block's zigzag
// sequence it is.  Thread i decodes the symbols that START inside subsequence i and records the state it ends in:
//   round 0      every subsequence from a guessed state (its first bit, start of an MCU) -- true only for i = 0;
//   round 1      every subsequence again, from the state its predecessor recorded;
//   round r > 1  only the subsequences whose predecessor recorded a DIFFERENT state in round r-1 than before (a work
//                list per image, appended to by the predecessor).  Huffman codes re-synchronise after a few symbols,
//                the block state at the next end-of-block that both decoders see, so the lists shrink geometrically.
//                The records are updated in place by single 64-bit stores: a thread may see its predecessor's old or
//                new state, and in the first case it is on the next list.  When a list comes out empty every record
//                is consistent with its predecessor's, and since record 0 is true, all are (induction over i).
//                A wrong state that runs into an impossible code records "bad"; its successor keeps its guess.
//   scan         exclusive sums over i of the blocks completed and of the DC differences per component;
//   write        once more from the true entry state, now storing coefficients (DC already predicted).
// Positions are raw bit offsets into the scan, canonical: the byte that holds the next unconsumed bit (never a
// stuffed 00) * 8 + the bit's index in it, so that two decoders at the same place hold the same number.
constexpr unsigned kBadState = 0xFFFFu;
// s: 0 = the block's DC symbol comes next; 1..63 = an AC symbol comes next and zigzag position s-1 was the last one filled
JG_DEV unsigned long long pack_state(unsigned pos, unsigned bs) { return bs == kBadState ? (unsigned long long)kBadState : ((unsigned long long)pos << 16) | bs; }

struct SyncReader {
    const uint8_t* p;          // next raw byte
    const uint8_t* end;
    unsigned long long buf;
    int bits;
    unsigned stuffed;          // bit j: the j-th newest byte of buf was followed by a stuffed byte
    unsigned ahead;            // the aligned word at `pa`, requested one refill early: its latency passes behind the decoding
    const uint8_t* pa;         // the next aligned word at or after p if it lies inside the scan, else nullptr
};
JG_DEV void sr_request(SyncReader& r)
{
    const uint8_t* a = (const uint8_t*)(((size_t)r.p + 3u) & ~(size_t)3u);
    if (a + 4 <= r.end) { r.ahead = ldg_u32(a); r.pa = a; } else r.pa = nullptr;
}
JG_DEV void sr_refill(SyncReader& r)       // called with fewer than 32 bits left; leaves 32 ... 63
{
    if (r.pa == r.p) {                                      // p is aligned and its word is already here
        const unsigned w = r.ahead;
        if (v_cmpeq4(w, 0xffffffffu) == 0u) {
            r.buf = (r.buf << 32) | bswap32(w);
            r.bits += 32;
            r.p += 4;
            r.stuffed <<= 4;
            sr_request(r);
            return;
        }
    }
    do {
        unsigned b = 0xFF, st = 0;
        if (r.p < r.end) {
            b = ldg_u8(r.p);
            if (b == 0xFF && r.p + 1 < r.end) st = 1;       // the stuffed 00: consumed, not data
        }
        r.p += 1 + st;                                      // past the end the stream continues with FF bytes (:452-456); p keeps counting
        r.buf = (r.buf << 8) | b;
        r.bits += 8;
        r.stuffed = (r.stuffed << 1) | st;
    } while (r.bits < 32 || ((((size_t)r.p) & 3u) != 0 && r.bits <= 48 && r.p < r.end));
    sr_request(r);
}
JG_DEV unsigned sr_pos(const SyncReader& r, const uint8_t* base)
{
    const int nb = (r.bits + 7) >> 3;                                        // buffered bytes with unconsumed bits
    const int st = jg::i_popc(r.stuffed & ((1u << nb) - 1u));                 // stuffed bytes behind them
    return (unsigned)((r.p - base) - nb - st) * 8u + (unsigned)((8 - (r.bits & 7)) & 7);
}
JG_DEV void sr_start(SyncReader& r, const uint8_t* base, const uint8_t* end, unsigned pos)
{
    r.p = base + (pos >> 3); r.end = end; r.buf = 0; r.bits = 0; r.stuffed = 0; r.ahead = 0;
    sr_request(r);
    const int off = (int)(pos & 7u);
    if (off) { sr_refill(r); r.bits -= off; }
}
// njGetVLC (:643-656) on ONE 32-bit window: a symbol is at most 16 code bits + 11 value bits, so after a top-up to 32 bits
// the code lookup and the value bits need no second look at the buffer (and no second refill check)
JG_DEV int sr_symbol(SyncReader& r, const VlcTables& T, int table, unsigned* code_out, bool* bad)
{
    if (r.bits < 32) sr_refill(r);
    const unsigned win = jg::funnel_r((unsigned)r.buf, (unsigned)(r.buf >> 32), (unsigned)(r.bits - 32));     // the next 32 bits
    unsigned e = T.l1[(table << kL1Bits) + (win >> (32 - kL1Bits))];
    if (!e) e = T.full[(size_t)table * 65536 + (win >> 16)];
    const int len = (int)(e >> 8);
    const unsigned code = e & 0xFFu;
    const int nb = (int)(code & 15u);
    *code_out = code;
    if (!len) { *bad = true; return 0; }
    r.bits -= len + nb;
    const unsigned raw = ((win << len) >> 1) >> (31 - nb);                    // the nb bits behind the code; nb = 0: 0 (a shift by 32 is undefined)
    return nb && raw < (1u << (nb - 1)) ? (int)raw + (int)((0xFFFFFFFFu << nb) + 1u) : (int)raw;
}

// the state a subsequence is entered with when its predecessor has none to offer
JG_DEV unsigned guessed_entry_pos(const DevParams& P, int i)
{
    unsigned at = (unsigned)i << P.sub_log2;
    if (i > 0 && P.scan[at - 1] == 0xFF) ++at;           // the subsequence begins with a stuffed byte
    return at << 3;
}

// Thread i, all symbols that start in subsequence i, from (entry_pos, entry_bs).  WRITE = false: only the state at
// the end and the sums (a round); WRITE = true: the coefficients, `first` holding what precedes the subsequence.
// Every lane of a warp runs the loop until the last one is done, one symbol per trip, so the lanes stay together.
template <bool WRITE>
JG_DEV void decode_subsequence(const DevParams& P, const uint16_t* l1, int i, bool have, unsigned entry_pos, unsigned entry_bs, SubStart first,
                               unsigned long long* exit_state, SubStart* sums)
{
    VlcTables T; T.full = P.vlc; T.l1 = l1;
    const uint8_t* const base = P.scan;
    const bool last = i + 1 >= P.n_sub;
    const unsigned limit = last ? P.scan_bytes << 3 : (unsigned)(i + 1) << (P.sub_log2 + 3);
    const int bpm = P.bpm, mbwidth = P.mbwidth;
    const unsigned long long total = P.total_blocks;
    SyncReader r;
    sr_start(r, base, P.scan + P.scan_bytes, have ? entry_pos : 0u);
    int b = (int)(entry_bs >> 8), s = (int)(entry_bs & 0xFFu);
    if (!have || b >= bpm) b = 0;
    McuBlock K = P.blk[b];
    unsigned long long g = first.n;
    int dc0 = first.dc0, dc1 = first.dc1, dc2 = first.dc2;
    unsigned n = 0, pos = entry_pos;
    int mbx = 0, mby = 0;
    int16_t* blk = nullptr;
    if (WRITE) {
        const unsigned long long m = g / (unsigned)bpm;
        mby = (int)(m / (unsigned)mbwidth); mbx = (int)(m - (unsigned long long)mby * (unsigned)mbwidth);
        blk = P.coef + (K.off + (unsigned long long)mby * K.row + (unsigned long long)mbx * K.sx) * 64ull;
    }
    bool bad = false;
    for (;;) {
        const bool go = have && !bad && (WRITE ? (g < total && (last || pos < limit)) : pos < limit);
        if (!JG_WARP_ANY(go)) break;
        if (!go) continue;
        // one path for DC and AC symbols: the table row is selected, not the code
        const bool is_dc = s == 0;
        unsigned code = 0;
        const int v = sr_symbol(r, T, is_dc ? K.dctab : K.actab, &code, &bad);
        const int k = s + (int)(code >> 4);                                           // AC: (s - 1) + run + 1
        const bool eob = !is_dc && !code;
        if (!is_dc && !eob && ((!(code & 0x0F) && code != 0xF0) || k > 63)) bad = true;
        if (is_dc) { dc0 += K.comp == 0 ? v : 0; dc1 += K.comp == 1 ? v : 0; dc2 += K.comp == 2 ? v : 0; }
        if (WRITE && !bad && !eob) blk[is_dc ? 0 : zz_nat(k)] = (int16_t)(is_dc ? (K.comp == 0 ? dc0 : K.comp == 1 ? dc1 : dc2) : v);
        const bool done = !bad && (eob || (!is_dc && k == 63));
        s = is_dc ? 1 : k + 1;
        if (done) {
            s = 0; ++n;
            if (++b == bpm) { b = 0; if (WRITE && ++mbx == mbwidth) { mbx = 0; ++mby; } }
            K = P.blk[b];
            if (WRITE) { ++g; blk = P.coef + (K.off + (unsigned long long)mby * K.row + (unsigned long long)mbx * K.sx) * 64ull; }
        }
        pos = sr_pos(r, base);
    }
    if (WRITE) {
        if (bad) *P.error = 5u;     // NJ_SYNTAX_ERROR
    } else {
        *exit_state = pack_state(pos, bad ? kBadState : ((unsigned)b << 8) | (unsigned)s);
        sums->n = n; sums->dc0 = dc0; sums->dc1 = dc1; sums->dc2 = dc2;
    }
}

// what subsequence i has to start from, given the state its predecessor recorded
JG_DEV void entry_of(const DevParams& P, int i, unsigned long long prev, unsigned* pos, unsigned* bs)
{
    if (i == 0 || (unsigned)(prev & 0xFFFFu) == kBadState) { *pos = guessed_entry_pos(P, i); *bs = 0; }
    else { *pos = (unsigned)(prev >> 16); *bs = (unsigned)(prev & 0xFFFFu); }
}

// Item t of round r for one image: decodes a subsequence, records its state and sums, and puts the successor on the
// next round's list if the state is not the one recorded before.  Returns whether it did.  Every lane of a warp calls it.
JG_DEV bool sync_round_item(const DevParams& P, const uint16_t* l1, int r, unsigned t, unsigned count)
{
    bool have = t < count;
    int i = 0;
    if (have) i = r < 2 ? (int)t : (int)P.sub_list[r & 1][t];
    if (r == 1 && i == 0) have = false;                                       // subsequence 0 was true in round 0
    unsigned pos = 0, bs = 0;
    if (have) {
        if (r == 0) pos = guessed_entry_pos(P, i);
        else entry_of(P, i, i ? jg::ld_flag64(&P.sub_exit[i - 1]) : 0ull, &pos, &bs);
    }
    const SubStart zero = {0u, 0, 0, 0};
    unsigned long long now = 0;
    SubStart sums = zero;
    decode_subsequence<false>(P, l1, i, have, pos, bs, zero, &now, &sums);
    if (!have) return false;
    const unsigned long long before = r ? jg::ld_flag64(&P.sub_exit[i]) : ~0ull;
    jg::st_flag64(&P.sub_exit[i], now);
    P.sub_sum[i] = sums;
    if (r == 0 || now == before || i + 1 >= P.n_sub) return false;
    P.sub_list[(r + 1) & 1][jg::gmem_atomic_add(&P.sub_cnt[(r + 1) % 3], 1u)] = (unsigned)(i + 1);
    return true;
}
// number of items of round r (read before any thread of the round appends: the appends go to another slot)
JG_DEV unsigned sync_round_count(const DevParams& P, int r) { return r < 2 ? (unsigned)P.n_sub : jg::ld_flag32(&P.sub_cnt[r % 3]); }

// the writing pass for subsequence i
JG_DEV void write_subsequence(const DevParams& P, const uint16_t* l1, int i)
{
    const bool have = i < P.n_sub;
    unsigned pos = 0, bs = 0;
    SubStart first = {0u, 0, 0, 0};
    if (have) {
        entry_of(P, i, i ? P.sub_exit[i - 1] : 0ull, &pos, &bs);
        first = P.sub_start[i];
    }
    decode_subsequence<true>(P, l1, i, have, pos, bs, first, nullptr, nullptr);
}

JG_DEV unsigned char clip8(int x) { return x < 0 ? 0 : (x > 0xFF ? 0xFF : (unsigned char)x); }   // njClip (:339-341)

// ---- the 8-point inverse DCT of jpeg_dec.h:343-442 (a fixed-point Chen-Wang) ------------------------------
// One routine for both passes; PASS 0 = row pass (input scaled by 2^11, result >> 8, stays int), PASS 1 = column
// pass (input scaled by 2^8, the three rotations rounded and pre-shifted by 3, result >> 14, + 128, clipped).
// The operation ORDER is the reference's: every product, sum and shift below has its twin there, so that
// the 32-bit intermediate values (and their wrap-around on hostile input) are the same.
constexpr int kC1 = 2841, kC2 = 2676, kC3 = 2408, kC5 = 1609, kC6 = 1108, kC7 = 565;   // 2048*sqrt(2)*cos(k*pi/16)

template <int PASS>
JG_DEV void idct8(int c0, int c1, int c2, int c3, int c4, int c5, int c6, int c7, int (&out)[8])
{
    constexpr int kIn = PASS == 0 ? 11 : 8;             // scale of the two even inputs that are not rotated
    constexpr int kBias = PASS == 0 ? 128 : 8192;       // rounding of the final shift, carried by the DC term
    constexpr int kRot = PASS == 0 ? 0 : 3;             // the column pass shifts every rotation down by 3 (after + 4)
    constexpr int kRnd = PASS == 0 ? 0 : 4;
    // odd half: two rotations (1,7) and (5,3)
    const int s17 = kC7 * (c1 + c7) + kRnd;
    const int o1 = (s17 + (kC1 - kC7) * c1) >> kRot;
    const int o7 = (s17 - (kC1 + kC7) * c7) >> kRot;
    const int s53 = kC3 * (c5 + c3) + kRnd;
    const int o5 = (s53 - (kC3 - kC5) * c5) >> kRot;
    const int o3 = (s53 - (kC3 + kC5) * c3) >> kRot;
    // even half: (0,4) butterfly and the (2,6) rotation
    const int e0 = (c0 << kIn) + kBias, e4 = c4 << kIn;
    const int sum04 = e0 + e4, dif04 = e0 - e4;
    const int s26 = kC6 * (c2 + c6) + kRnd;
    const int r6 = (s26 - (kC2 + kC6) * c6) >> kRot;
    const int r2 = (s26 + (kC2 - kC6) * c2) >> kRot;
    // second stage
    const int t15 = o1 + o5, d15 = o1 - o5, t73 = o7 + o3, d73 = o7 - o3;
    const int a = sum04 + r2, b = sum04 - r2, c = dif04 + r6, d = dif04 - r6;
    const int m = (181 * (d15 + d73) + 128) >> 8;       // 181/256 = 1/sqrt(2)
    const int n = (181 * (d15 - d73) + 128) >> 8;
    out[0] = a + t15; out[1] = c + m; out[2] = d + n; out[3] = b + t73;
    out[4] = b - t73; out[5] = d - n; out[6] = c - m; out[7] = a - t15;
}

JG_DEV void row_idct(int* blk)   // njRowIDCT (:350-396): in place, results keep 3 fraction bits
{
    if (!((blk[4] << 11) | blk[6] | blk[2] | blk[1] | blk[7] | blk[5] | blk[3])) {      // only DC: a flat row
        const int v = blk[0] << 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) blk[i] = v;
        return;
    }
    int o[8];
    idct8<0>(blk[0], blk[1], blk[2], blk[3], blk[4], blk[5], blk[6], blk[7], o);
#pragma unroll
    for (int i = 0; i < 8; ++i) blk[i] = o[i] >> 8;
}

// njColIDCT (:398-442); IS = distance between the column's elements (8 in a plain block, 9 in a padded shared-memory tile)
template <int IS>
JG_DEV void col_idct(const int* blk, unsigned char* out, int stride)
{
    if (!((blk[IS * 4] << 8) | blk[IS * 6] | blk[IS * 2] | blk[IS * 1] | blk[IS * 7] | blk[IS * 5] | blk[IS * 3])) {   // only DC: a flat column
        const unsigned char v = clip8(((blk[0] + 32) >> 6) + 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[(size_t)i * stride] = v;
        return;
    }
    int o[8];
    idct8<1>(blk[0], blk[IS * 1], blk[IS * 2], blk[IS * 3], blk[IS * 4], blk[IS * 5], blk[IS * 6], blk[IS * 7], o);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(size_t)i * stride] = clip8((o[i] >> 14) + 128);
}

// block b (raster index in its component's padded plane) of component c
JG_DEV void idct_block(const DevParams& P, int c, unsigned long long b)
{
    const DevComponent& K = P.comp[c];
    const int16_t* src = P.coef + (K.coef_off + b) * 64ull;
    int blk[64];
    for (int i = 0; i < 64; ++i) blk[i] = (int)src[i] * K.dq[i];
    for (int r = 0; r < 64; r += 8) row_idct(blk + r);
    const unsigned long long by = b / (unsigned long long)K.bw, bx = b - by * (unsigned long long)K.bw;
    unsigned char* out = P.planes + K.plane_off + ((by * (unsigned long long)K.stride + bx) << 3);
    for (int col = 0; col < 8; ++col) col_idct<8>(blk + col, out + col, K.stride);
}

// ---- chroma upsampling (jpeg_dec.h:722-790) --------------------------------------------------
#define JD_CF(x) clip8(((x) + 64) >> 7)

// output pixel (y, ox) of njUpsampleH: in = plane of width w (row pitch s), out has width 2w.
// The last three outputs read lin[-1..-3] AFTER lin += stride (:753-756): the end of the padded row.
JG_DEV unsigned char upsample_h(const unsigned char* in, int w, int s, int y, int ox)
{
    const unsigned char* l = in + (size_t)y * s;
    if (ox == 0) return JD_CF(139 * l[0] + -11 * l[1]);
    if (ox == 1) return JD_CF(104 * l[0] + 27 * l[1] + -3 * l[2]);
    if (ox == 2) return JD_CF(28 * l[0] + 109 * l[1] + -9 * l[2]);
    const int W2 = w << 1;
    if (ox >= W2 - 3) {
        const unsigned char* t = l + s;
        if (ox == W2 - 3) return JD_CF(28 * t[-1] + 109 * t[-2] + -9 * t[-3]);
        if (ox == W2 - 2) return JD_CF(104 * t[-1] + 27 * t[-2] + -3 * t[-3]);
        return JD_CF(139 * t[-1] + -11 * t[-2]);
    }
    if (ox & 1) { const int x = (ox - 3) >> 1; return JD_CF(-9 * l[x] + 111 * l[x + 1] + 29 * l[x + 2] + -3 * l[x + 3]); }
    const int x = (ox - 4) >> 1;
    return JD_CF(-3 * l[x] + 29 * l[x + 1] + 111 * l[x + 2] + -9 * l[x + 3]);
}

// output pixel (oy, x) of njUpsampleV: in = plane of height h (row pitch s), out has height 2h
JG_DEV unsigned char upsample_v(const unsigned char* in, int h, int s, int oy, int x)
{
    const unsigned char* c = in + x;
    auto r = [&](int row) { return (int)c[(size_t)row * s]; };
    if (oy == 0) return JD_CF(139 * r(0) + -11 * r(1));
    if (oy == 1) return JD_CF(104 * r(0) + 27 * r(1) + -3 * r(2));
    if (oy == 2) return JD_CF(28 * r(0) + 109 * r(1) + -9 * r(2));
    const int H2 = h << 1;
    if (oy == H2 - 3) return JD_CF(28 * r(h - 1) + 109 * r(h - 2) + -9 * r(h - 3));
    if (oy == H2 - 2) return JD_CF(104 * r(h - 1) + 27 * r(h - 2) + -3 * r(h - 3));
    if (oy == H2 - 1) return JD_CF(139 * r(h - 1) + -11 * r(h - 2));
    if (oy & 1) { const int i = (oy - 3) >> 1; return JD_CF(-9 * r(i) + 111 * r(i + 1) + 29 * r(i + 2) + -3 * r(i + 3)); }
    const int i = (oy - 4) >> 1;
    return JD_CF(-3 * r(i) + 29 * r(i + 1) + 111 * r(i + 2) + -9 * r(i + 3));
}

// njConvert's RGB loop (:838-852)
JG_DEV void to_rgb(unsigned char* rgb, int yv, int cbv, int crv)
{
    const int y = yv << 8, cb = cbv - 128, cr = crv - 128;
    rgb[0] = clip8((y + 359 * cr + 128) >> 8);
    rgb[1] = clip8((y - 88 * cb - 183 * cr + 128) >> 8);
    rgb[2] = clip8((y + 454 * cb + 128) >> 8);
}

// per-image work lists of the plane stages
struct PlaneOp { const unsigned char* in; unsigned char* out; int w, h, s; };                      // one upsampling pass of one plane
struct ColorOp {
    const unsigned char *py, *pcb, *pcr; int sy, scb, scr; unsigned char* out; int w, h, ncomp;
    int chroma_h;      // > 0: the chroma planes still have this many rows -- their last vertical filter pass runs inside the colour kernel
};

#if !defined(JG_EMULATE)
// Every kernel works on a BATCH: blockIdx.y (or .z) picks the image, the x dimension covers the largest
// image's work items and the others' surplus threads leave at once.
__global__ void decode_intervals_kernel(const DevParams* __restrict__ imgs)
{
    const DevParams& P = imgs[blockIdx.y];
    __shared__ uint16_t l1[4 << kL1Bits];
    if ((int)(blockIdx.x * blockDim.x) >= P.n_intervals) return;            // whole CTA surplus: no table needed
    for (int i = (int)threadIdx.x; i < (4 << kL1Bits); i += (int)blockDim.x) l1[i] = l1_entry(P.vlc, i >> kL1Bits, i & ((1 << kL1Bits) - 1));
    __syncthreads();
    decode_interval(P, l1, (int)(blockIdx.x * blockDim.x + threadIdx.x));      // every lane: the warp reconverges per block
}
// ---- subsequence decode: blockIdx.y = image, one thread per subsequence ----
constexpr int kSubThreads = 128;
JG_DEV void load_l1(const DevParams& P, uint16_t* l1)
{
    for (int i = (int)threadIdx.x; i < (4 << kL1Bits); i += (int)blockDim.x) l1[i] = l1_entry(P.vlc, i >> kL1Bits, i & ((1 << kL1Bits) - 1));
    __syncthreads();
}
// round r of the batch (see "subsequences"); *appended counts the subsequences put on the next round's lists.
// Rounds 0 and 1 cover every subsequence, later ones their image's list: CTAs beyond its end leave at once.
__global__ void __launch_bounds__(kSubThreads) sync_round_kernel(const DevParams* __restrict__ imgs, int r, unsigned* __restrict__ appended)
{
    const DevParams& P = imgs[blockIdx.y];
    __shared__ uint16_t l1[4 << kL1Bits];
    if (!P.n_sub) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.sub_cnt[(r + 2) % 3] = 0u;       // the list of round r-1: read by all, appended to again in round r+1
    const unsigned count = sync_round_count(P, r);
    if (blockIdx.x * blockDim.x >= count) return;
    load_l1(P, l1);
    const bool did = sync_round_item(P, l1, r, blockIdx.x * blockDim.x + threadIdx.x, count);
    const unsigned any = __ballot_sync(0xffffffffu, did);
    if (any && (threadIdx.x & 31) == 0) atomicAdd(appended, (unsigned)__popc(any));
}
// exclusive sums of the records per image: one CTA, a contiguous slice per thread
__global__ void __launch_bounds__(256) sync_scan_kernel(const DevParams* __restrict__ imgs)
{
    const DevParams& P = imgs[blockIdx.x];
    const int n = P.n_sub, t = (int)threadIdx.x;
    if (!n) return;
    const SubStart* S = P.sub_sum;
    const int per = (n + 255) / 256, lo = t * per < n ? t * per : n, hi = lo + per < n ? lo + per : n;
    __shared__ SubStart part[256];
    SubStart a = {0u, 0, 0, 0};
    for (int j = lo; j < hi; ++j) { a.n += S[j].n; a.dc0 += S[j].dc0; a.dc1 += S[j].dc1; a.dc2 += S[j].dc2; }
    part[t] = a;
    __syncthreads();
    if (t == 0) {
        SubStart run = {0u, 0, 0, 0};
        for (int j = 0; j < 256; ++j) {
            const SubStart v = part[j];
            part[j] = run;
            run.n += v.n; run.dc0 += v.dc0; run.dc1 += v.dc1; run.dc2 += v.dc2;
        }
    }
    __syncthreads();
    a = part[t];
    for (int j = lo; j < hi; ++j) {
        P.sub_start[j] = a;
        a.n += S[j].n; a.dc0 += S[j].dc0; a.dc1 += S[j].dc1; a.dc2 += S[j].dc2;
    }
}
__global__ void __launch_bounds__(kSubThreads) sync_write_kernel(const DevParams* __restrict__ imgs)
{
    const DevParams& P = imgs[blockIdx.y];
    __shared__ uint16_t l1[4 << kL1Bits];
    if ((int)(blockIdx.x * blockDim.x) >= P.n_sub) return;
    load_l1(P, l1);
    write_subsequence(P, l1, (int)(blockIdx.x * blockDim.x + threadIdx.x));
}
// IDCT: eight threads per block.  Thread r loads row r (one 16-byte load), dequantises and runs njRowIDCT
// in registers; the 8x8 goes through a padded shared-memory tile; thread c then runs njColIDCT on column c
// and the eight threads write one 8-byte row segment at a time.  (idct_block above is the same arithmetic
// with one thread per block; the CPU harness uses that one.)
constexpr int kIdctThreads = 128;
__global__ void idct_kernel(const DevParams* __restrict__ imgs, int c)
{
    const DevParams& P = imgs[blockIdx.y];
    if (c >= P.ncomp) return;
    __shared__ int tile[kIdctThreads / 8][72];
    __shared__ int dq[64];
    const int t = (int)threadIdx.x, l = t & 7, bi = t >> 3;
    if (t < 64) dq[t] = P.comp[c].dq[t];
    __syncthreads();
    const unsigned long long b = (unsigned long long)blockIdx.x * (kIdctThreads / 8) + bi;
    const bool valid = b < P.comp[c].n_blocks;
    if (valid) {
        const uint4 v = *reinterpret_cast<const uint4*>(P.coef + (P.comp[c].coef_off + b) * 64ull + 8 * l);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        int row[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) row[i] = (int)(short)((w[i >> 1] >> (16 * (i & 1))) & 0xffffu) * dq[8 * l + i];
        row_idct(row);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[bi][9 * l + i] = row[i];
    }
    __syncwarp();
    if (valid) {
        const int bw = P.comp[c].bw, stride = P.comp[c].stride;
        const unsigned long long by = b / (unsigned long long)bw, bx = b - by * (unsigned long long)bw;
        unsigned char* out = P.planes + P.comp[c].plane_off + ((by * (unsigned long long)stride + bx) << 3);
        col_idct<9>(&tile[bi][l], out + l, stride);
    }
}
// The plane stages handle FOUR neighbouring output pixels per thread and store them as one 32-bit word
// (three for RGB) where the address allows it: a quarter of the threads and store instructions.
__device__ __forceinline__ void store4(unsigned char* dst, const unsigned char (&v)[4], int n)
{
    if (n == 4 && (((size_t)dst) & 3u) == 0) *reinterpret_cast<unsigned*>(dst) = v[0] | (v[1] << 8) | (v[2] << 16) | ((unsigned)v[3] << 24);
    else for (int i = 0; i < n; ++i) dst[i] = v[i];
}
__global__ void upsample_h_kernel(const PlaneOp* __restrict__ ops)
{
    const PlaneOp o = ops[blockIdx.z];
    const int ox = 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, W2 = 2 * o.w;
    if (ox >= W2 || y >= o.h) return;
    const int n = W2 - ox < 4 ? W2 - ox : 4;
    unsigned char v[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) v[i] = upsample_h(o.in, o.w, o.s, y, ox + i);
    store4(o.out + (size_t)y * W2 + ox, v, n);
}
// taps of njUpsampleV for output row oy (:773-785): up to four input rows and their weights (the same for every column)
__device__ __forceinline__ int v_taps(int h, int oy, int (&row)[4], int (&k)[4])
{
    const int H2 = h << 1;
    if (oy == 0) { row[0] = 0; row[1] = 1; k[0] = 139; k[1] = -11; return 2; }
    if (oy == 1) { row[0] = 0; row[1] = 1; row[2] = 2; k[0] = 104; k[1] = 27; k[2] = -3; return 3; }
    if (oy == 2) { row[0] = 0; row[1] = 1; row[2] = 2; k[0] = 28; k[1] = 109; k[2] = -9; return 3; }
    if (oy == H2 - 3) { row[0] = h - 1; row[1] = h - 2; row[2] = h - 3; k[0] = 28; k[1] = 109; k[2] = -9; return 3; }
    if (oy == H2 - 2) { row[0] = h - 1; row[1] = h - 2; row[2] = h - 3; k[0] = 104; k[1] = 27; k[2] = -3; return 3; }
    if (oy == H2 - 1) { row[0] = h - 1; row[1] = h - 2; k[0] = 139; k[1] = -11; return 2; }
    const int i = (oy & 1) ? (oy - 3) >> 1 : (oy - 4) >> 1;
    row[0] = i; row[1] = i + 1; row[2] = i + 2; row[3] = i + 3;
    if (oy & 1) { k[0] = -9; k[1] = 111; k[2] = 29; k[3] = -3; } else { k[0] = -3; k[1] = 29; k[2] = 111; k[3] = -9; }
    return 4;
}
__global__ void upsample_v_kernel(const PlaneOp* __restrict__ ops)
{
    const PlaneOp o = ops[blockIdx.z];
    const int x = 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= o.w) return;
    const int n = o.w - x < 4 ? o.w - x : 4;
    // (2 h output rows: one more than gridDim.y can hold for the tallest planes, so the rows are strided over the grid)
    for (int oy = (int)blockIdx.y; oy < 2 * o.h; oy += (int)gridDim.y) {
    unsigned char v[4] = {0, 0, 0, 0};
    if (n == 4 && ((((size_t)o.in) | (size_t)o.s) & 3u) == 0) {
        // four columns at once: one word per input row instead of four bytes (same integer sums as upsample_v)
        int row[4] = {0, 0, 0, 0}, k[4] = {0, 0, 0, 0};
        const int taps = v_taps(o.h, oy, row, k);
        int sum[4] = {0, 0, 0, 0};
        for (int j = 0; j < taps; ++j) {
            const unsigned w = *reinterpret_cast<const unsigned*>(o.in + (size_t)row[j] * o.s + x);
#pragma unroll
            for (int i = 0; i < 4; ++i) sum[i] += k[j] * (int)((w >> (8 * i)) & 0xffu);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = JD_CF(sum[i]);
    } else {
        for (int i = 0; i < n; ++i) v[i] = upsample_v(o.in, o.h, o.s, oy, x + i);
    }
    store4(o.out + (size_t)oy * o.w + x, v, n);
    }
}
__global__ void color_kernel(const ColorOp* __restrict__ ops)
{
    const ColorOp o = ops[blockIdx.z];
    const int x = 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y;
    if (x >= o.w || y >= o.h) return;
    const int n = o.w - x < 4 ? o.w - x : 4;
    if (o.ncomp != 3) {
        unsigned char v[4] = {0, 0, 0, 0};
        for (int i = 0; i < n; ++i) v[i] = o.py[(size_t)y * o.sy + x + i];
        store4(o.out + (size_t)y * o.w + x, v, n);
        return;
    }
    unsigned char rgb[12];
    const unsigned char* py = o.py + (size_t)y * o.sy + x;
    if (o.chroma_h > 0) {
        // 4:2:0 and friends: njUpsampleV of Cb and Cr for output row y, computed here instead of written to a plane
        // and read back (same taps, same integer sums, same clip as upsample_v)
        unsigned char cb[4] = {0, 0, 0, 0}, cr[4] = {0, 0, 0, 0};
        if (n == 4 && ((((size_t)o.pcb) | ((size_t)o.pcr) | (size_t)o.scb | (size_t)o.scr) & 3u) == 0) {
            int row[4] = {0, 0, 0, 0}, k[4] = {0, 0, 0, 0};
            const int taps = v_taps(o.chroma_h, y, row, k);
            int sb[4] = {0, 0, 0, 0}, sr[4] = {0, 0, 0, 0};
            for (int j = 0; j < taps; ++j) {
                const unsigned wb = *reinterpret_cast<const unsigned*>(o.pcb + (size_t)row[j] * o.scb + x);
                const unsigned wr = *reinterpret_cast<const unsigned*>(o.pcr + (size_t)row[j] * o.scr + x);
#pragma unroll
                for (int i = 0; i < 4; ++i) { sb[i] += k[j] * (int)((wb >> (8 * i)) & 0xffu); sr[i] += k[j] * (int)((wr >> (8 * i)) & 0xffu); }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { cb[i] = JD_CF(sb[i]); cr[i] = JD_CF(sr[i]); }
        } else {
            for (int i = 0; i < n; ++i) { cb[i] = upsample_v(o.pcb, o.chroma_h, o.scb, y, x + i); cr[i] = upsample_v(o.pcr, o.chroma_h, o.scr, y, x + i); }
        }
        for (int i = 0; i < n; ++i) to_rgb(rgb + 3 * i, py[i], cb[i], cr[i]);
    } else {
        const unsigned char *pcb = o.pcb + (size_t)y * o.scb + x, *pcr = o.pcr + (size_t)y * o.scr + x;
        if (n == 4 && ((((size_t)py) | ((size_t)pcb) | ((size_t)pcr)) & 3u) == 0) {       // a word per plane instead of four bytes
            const unsigned wy = *reinterpret_cast<const unsigned*>(py), wb = *reinterpret_cast<const unsigned*>(pcb), wr = *reinterpret_cast<const unsigned*>(pcr);
#pragma unroll
            for (int i = 0; i < 4; ++i) to_rgb(rgb + 3 * i, (int)((wy >> (8 * i)) & 0xffu), (int)((wb >> (8 * i)) & 0xffu), (int)((wr >> (8 * i)) & 0xffu));
        } else {
            for (int i = 0; i < n; ++i) to_rgb(rgb + 3 * i, py[i], pcb[i], pcr[i]);
        }
    }
    unsigned char* dst = o.out + ((size_t)y * o.w + x) * 3;
    if (n == 4 && (((size_t)dst) & 3u) == 0) {
        unsigned* d = reinterpret_cast<unsigned*>(dst);
        for (int k = 0; k < 3; ++k) d[k] = rgb[4 * k] | (rgb[4 * k + 1] << 8) | (rgb[4 * k + 2] << 16) | ((unsigned)rgb[4 * k + 3] << 24);
    } else {
        for (int i = 0; i < 3 * n; ++i) dst[i] = rgb[i];
    }
}
#endif

}  // namespace jd
