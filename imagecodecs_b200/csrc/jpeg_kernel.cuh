// jpeg_kernel.cuh -- the whole JPEG entropy-coded segment in ONE persistent sm_100a kernel.
//
// Replaces the reference's serial block loop, jpeg_enc.h:1094-1158 (tjei_encode_main) and
// everything it calls: tjei_encode_and_write_MCU (:786-889), tjei_fdct (:656-763),
// tjei_calculate_variable_length_int (:598-610), tjei_write_bits (:613-643) and the tail
// (:1160-1167).  Output bytes are identical to the reference's for its native modes.
//
// Work decomposition
//   tile   = 192 consecutive 8x8 data units (blocks) of ONE image in stream order
//            (4:4:4: 64 MCUs x {Y,Cb,Cr}; 4:2:0: 32 MCUs x {Y00,Y01,Y10,Y11,Cb,Cr}; gray: 192).
//   CTA    = 192 threads, one thread per block, warps are component-uniform.  CTAs are
//            persistent and draw tiles from an atomic ticket, so a tile's predecessors are
//            always resident or finished (what makes the look-back below deadlock-free).
//   thread = colour conversion + AAN FDCT + quantise + zigzag entirely in registers
//            (64 coefficients, statically indexed), then two walks over them: one that
//            only sizes the block, one that packs its bits at the block's exact offset.
//
// Per tile:
//   1. stage   pixels of the tile's MCUs (+ the MCU before it) -> shared memory, edge
//              replication applied here (jpeg_enc.h:1106-1111)
//   2. xform   P2+P3+P4 of SURVEY 8a in registers; DC of the preceding MCU is recomputed
//              (DC-only transform) instead of being communicated between CTAs
//   3. size    per-block bit count (P5-P7), CTA exclusive scan in stream order
//   4. pack    every thread ORs/stores its bits into the tile's window in shared memory
//   5. chain   decoupled look-back over tiles (#1): exclusive BIT offset of the tile in its
//              image; the descriptor also carries the tile's last 7 bits so the successor can
//              complete the byte the two tiles share
//   6. stuff   count 0xFF bytes this tile owns, decoupled look-back (#2) over those counts:
//              exclusive BYTE offset in the stuffed stream; emit 0xFF00 (jpeg_enc.h:634-638)
//   7. write   coalesced copy of the stuffed bytes to the image's scan; the last tile pads
//              with zero bits (jpeg_enc.h:1161-1164) and appends EOI (:1166-1167)
// HBM traffic is therefore: every pixel read once (+1/64 for the predecessor MCU), every
// output byte written once.  Nothing else touches DRAM except two 8-byte descriptors per tile.
//
// A tile whose bits do not fit the shared-memory window (pathological content) is processed
// as several "groups" of blocks with the walks repeated; same bytes, lower speed.
#pragma once
#include "jpeg_device.h"
#include "jpeg_launch.h"
#include "jpeg_tables.h"

namespace jg {

template <int LAYOUT, int NC>
struct Geo {
    static constexpr int BPM = LAYOUT == LAYOUT_444 ? 3 : (LAYOUT == LAYOUT_420 ? 6 : 1);  // blocks per MCU
    static constexpr int MCU_W = LAYOUT == LAYOUT_420 ? 16 : 8;
    static constexpr int MCU_H = MCU_W;
    static constexpr int M = kBlocksPerTile / BPM;  // MCUs per tile
    static constexpr int SLOTS = M + 1;             // slot 0 = the MCU preceding the tile
    static constexpr int ROWB = MCU_W * NC;         // bytes of one MCU pixel row
    static constexpr int WPR = ROWB / 4;            // ... in 32-bit words
    static constexpr int STAGE_WORDS = MCU_H * SLOTS * WPR;  // layout [row][slot][WPR]
};

template <int LAYOUT, int NC>
struct Smem {
    using G = Geo<LAYOUT, NC>;
    static constexpr int A_WORDS = (G::STAGE_WORDS > kWinWordsMax + 8 ? G::STAGE_WORDS : kWinWordsMax + 8);
    alignas(16) uint32_t a[A_WORDS];               // pixel staging, then the unstuffed window
    alignas(16) uint8_t sbuf[kWinWordsMax * 8 + 64];  // stuffed bytes of one group (worst case 2x)
    uint32_t huff_ac[2][256];
    uint32_t huff_dc[2][16];
    uint32_t bits_s[kBlocksPerTile];      // bits per block, stream order
    uint32_t off_s[kBlocksPerTile + 1];   // exclusive scan of bits_s; [192] = total
    int dc_s[kBlocksPerTile];             // quantised DC per block, stream order
    uint16_t gstart[kBlocksPerTile + 2];  // group boundaries (block indices)
    uint32_t warp_tmp[kWarps];
    // tile-wide scalars (written by one thread, read after a barrier)
    int tile;
    int abort;
    int n_groups;
    unsigned pred_tail;                   // last 7 bits of the preceding tile
    unsigned long long bit_base;          // exclusive bit offset of the tile in its image
    unsigned long long ff_base;           // exclusive stuffed-FF count
};

// ------------------------------------------------------------------------------------------
// zigzag: position in scan order of natural index i (jpeg_enc.h:376-386)
// ------------------------------------------------------------------------------------------
JG_DEV constexpr int zz_of(int i)
{
    constexpr unsigned char t[64] = JG_ZZ_INIT;
    return t[i];
}

// ------------------------------------------------------------------------------------------
// sample fetch + colour conversion (jpeg_enc.h:1114-1124), operation order preserved
// ------------------------------------------------------------------------------------------
JG_DEV unsigned byte_of(const uint32_t* w, int idx) { return (w[idx >> 2] >> ((idx & 3) * 8)) & 0xffu; }

// Chroma weights of one component.  Cb and Cr share ONE code path (keeps the kernel small);
// x - y == x + (-y) and (-c)*g == -(c*g) exactly in IEEE arithmetic, so
//   Cb = (-0.1687f*r - 0.3313f*g) + 0.5f*b   and   Cr = (0.5f*r - 0.4187f*g) - 0.0813f*b
// are both ((k0*r + k1*g) + k2*b) with the signs folded into the constants.
struct ChromaK { float k0, k1, k2; };
JG_DEV ChromaK chroma_k(int comp)
{
    ChromaK k;
    if (comp == 1) { k.k0 = -0.1687f; k.k1 = -0.3313f; k.k2 = 0.5f; }
    else { k.k0 = 0.5f; k.k1 = -0.4187f; k.k2 = -0.0813f; }
    return k;
}

template <int CLS>   // 0 = luma, 1 = chroma
JG_DEV float ycc(float r, float g, float b, const ChromaK& k)
{
    if (CLS == 0) return f_sub(f_add(f_add(f_mul(0.299f, r), f_mul(0.587f, g)), f_mul(0.114f, b)), 128.0f);
    return f_add(f_add(f_mul(k.k0, r), f_mul(k.k1, g)), f_mul(k.k2, b));
}

template <int NC, int CLS>
JG_DEV float sample_of(const uint32_t* w, int i, const ChromaK& k)
{
    float r, g, b;
    if (NC == 4) {
        const uint32_t v = w[i];
        r = u8_to_f(v & 0xffu); g = u8_to_f((v >> 8) & 0xffu); b = u8_to_f((v >> 16) & 0xffu);
    } else {
        r = u8_to_f(byte_of(w, 3 * i)); g = u8_to_f(byte_of(w, 3 * i + 1)); b = u8_to_f(byte_of(w, 3 * i + 2));
    }
    return ycc<CLS>(r, g, b, k);
}

template <int N>
JG_DEV void lds_words(const uint32_t* p, uint32_t (&w)[N])
{
    // p is 8-byte aligned for every caller (row segments are multiples of 8 bytes)
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const uint2 v = reinterpret_cast<const uint2*>(p)[i];
        w[2 * i] = v.x; w[2 * i + 1] = v.y;
    }
}

// Row r (0..7) of the thread's 8x8 block, as 8 level-shifted / colour-converted samples.
// q: quadrant of the Y block inside a 4:2:0 MCU (ignored otherwise).
template <int LAYOUT, int NC, int CLS>
JG_DEV void fetch_row(const uint32_t* stage, int slot, int q, int r, const ChromaK& ck, float (&s)[8])
{
    using G = Geo<LAYOUT, NC>;
    if (LAYOUT == LAYOUT_GRAY) {
        uint32_t w[2];
        lds_words<2>(stage + (r * G::SLOTS + slot) * G::WPR, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = f_sub(u8_to_f(byte_of(w, i)), 128.0f);
    } else if (LAYOUT == LAYOUT_444 || CLS == 0) {
        constexpr int W8 = 8 * NC / 4;  // words of 8 pixels
        const int row = LAYOUT == LAYOUT_420 ? r + 8 * (q >> 1) : r;
        const int xoff = LAYOUT == LAYOUT_420 ? (q & 1) * W8 : 0;
        uint32_t w[W8];
        lds_words<W8>(stage + (row * G::SLOTS + slot) * G::WPR + xoff, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = sample_of<NC, CLS>(w, i, ck);
    } else {
        // 4:2:0 chroma: ((a+b)+(c+d))*0.25f over the 2x2 float Cb/Cr values (extended mode)
        constexpr int W16 = 16 * NC / 4;
        uint32_t w0[W16], w1[W16];
        lds_words<W16>(stage + ((2 * r) * G::SLOTS + slot) * G::WPR, w0);
        lds_words<W16>(stage + ((2 * r + 1) * G::SLOTS + slot) * G::WPR, w1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a = sample_of<NC, CLS>(w0, 2 * i, ck), b = sample_of<NC, CLS>(w0, 2 * i + 1, ck);
            const float c = sample_of<NC, CLS>(w1, 2 * i, ck), d = sample_of<NC, CLS>(w1, 2 * i + 1, ck);
            s[i] = f_mul(f_add(f_add(a, b), f_add(c, d)), 0.25f);
        }
    }
}

// ------------------------------------------------------------------------------------------
// AAN forward DCT, one 8-point pass (jpeg_enc.h:668-709 rows, :718-759 columns)
// ------------------------------------------------------------------------------------------
JG_DEV void aan8(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5, float& d6, float& d7)
{
    const float c4 = 0.707106781f, c6 = 0.382683433f, c2m6 = 0.541196100f, c2p6 = 1.306562965f;
    const float t0 = f_add(d0, d7), t7 = f_sub(d0, d7);
    const float t1 = f_add(d1, d6), t6 = f_sub(d1, d6);
    const float t2 = f_add(d2, d5), t5 = f_sub(d2, d5);
    const float t3 = f_add(d3, d4), t4 = f_sub(d3, d4);

    const float e0 = f_add(t0, t3), e3 = f_sub(t0, t3);
    const float e1 = f_add(t1, t2), e2 = f_sub(t1, t2);
    d0 = f_add(e0, e1);
    d4 = f_sub(e0, e1);
    const float z1 = f_mul(f_add(e2, e3), c4);
    d2 = f_add(e3, z1);
    d6 = f_sub(e3, z1);

    const float o0 = f_add(t4, t5), o1 = f_add(t5, t6), o2 = f_add(t6, t7);
    const float z5 = f_mul(f_sub(o0, o2), c6);
    const float z2 = f_add(f_mul(c2m6, o0), z5);
    const float z4 = f_add(f_mul(c2p6, o2), z5);
    const float z3 = f_mul(o1, c4);
    const float z11 = f_add(t7, z3), z13 = f_sub(t7, z3);
    d5 = f_add(z13, z2);
    d3 = f_sub(z13, z2);
    d1 = f_add(z11, z4);
    d7 = f_sub(z11, z4);
}

// the DC output of aan8 only: ((d0+d7)+(d3+d4)) + ((d1+d6)+(d2+d5))
JG_DEV float aan8_dc(const float (&d)[8])
{
    const float t0 = f_add(d[0], d[7]), t1 = f_add(d[1], d[6]), t2 = f_add(d[2], d[5]), t3 = f_add(d[3], d[4]);
    return f_add(f_add(t0, t3), f_add(t1, t2));
}

// jpeg_enc.h:808-816: v*pqt, floorf((v + 1024) + 0.5f) - 1024, (int)
JG_DEV int quantise(float v, float pq)
{
    v = f_mul(v, pq);
    v = f_add(f_add(v, 1024.0f), 0.5f);
    return f_floor_i(v) - 1024;
}

// samples -> 64 quantised coefficients in ZIGZAG order, all in registers
template <int LAYOUT, int NC, int CLS>
JG_DEV void transform_block(const uint32_t* stage, int slot, int q, const ChromaK& ck, const QuantSet& Q, int (&c)[64])
{
    float d[64];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float s[8];
        fetch_row<LAYOUT, NC, CLS>(stage, slot, q, r, ck, s);
        aan8(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[8 * r + i] = s[i];
    }
#pragma unroll
    for (int x = 0; x < 8; ++x)
        aan8(d[x], d[8 + x], d[16 + x], d[24 + x], d[32 + x], d[40 + x], d[48 + x], d[56 + x]);
#pragma unroll
    for (int i = 0; i < 64; ++i) c[zz_of(i)] = quantise(d[i], CLS == 0 ? Q.luma[i] : Q.chroma[i]);
}

// quantised DC of a block without the other 63 outputs (same roundings as transform_block)
template <int LAYOUT, int NC, int CLS>
JG_DEV int transform_dc_only(const uint32_t* stage, int slot, int q, const ChromaK& ck, const QuantSet& Q)
{
    float col[8];
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        float s[8];
        fetch_row<LAYOUT, NC, CLS>(stage, slot, q, r, ck, s);
        col[r] = aan8_dc(s);
    }
    return quantise(aan8_dc(col), CLS == 0 ? Q.luma[0] : Q.chroma[0]);
}

// ------------------------------------------------------------------------------------------
// entropy coding of one block (jpeg_enc.h:831-888)
// ------------------------------------------------------------------------------------------
JG_DEV unsigned category(int v)  // jpeg_enc.h:598-608; v != 0
{
    const unsigned a = (unsigned)(v < 0 ? -v : v);
    return 32u - (unsigned)i_clz(a);
}
JG_DEV unsigned amplitude(int v, unsigned cat)  // jpeg_enc.h:601-609: (v<0 ? v-1 : v) & mask
{
    return (unsigned)(v + (v >> 31)) & ((1u << cat) - 1u);
}

// walk 1: size only
JG_DEV unsigned block_bits(const int (&c)[64], int diff, const uint32_t* ac, const uint32_t* dc)
{
    const unsigned zrl_len = ac[0xF0] & 0xffu;
    const unsigned dcat = diff ? category(diff) : 0u;
    unsigned n = (dc[dcat] & 0xffu) + dcat;
    int last = 0;
#pragma unroll
    for (int k = 1; k < 64; ++k) {
        const int v = c[k];
        if (v != 0) {
            unsigned run = (unsigned)(k - 1 - last);
            last = k;
            n += (run >> 4) * zrl_len;   // one ZRL per 16 zeros (jpeg_enc.h:863-867)
            run &= 15u;
            const unsigned cat = category(v);
            n += (ac[(run << 4) | cat] & 0xffu) + cat;
        }
    }
    if (last != 63) n += ac[0] & 0xffu;  // EOB (jpeg_enc.h:884-887)
    return n;
}

// MSB-first bit writer into a shared-memory word array that several threads fill
// concurrently: a thread's first and last (partial) words are OR-ed atomically, the
// words in between belong to it alone.  Replaces the serial cursor of jpeg_enc.h:613-643.
struct BitPacker {
    uint32_t* buf;
    unsigned long long acc;  // pending bits, left-aligned
    int fill;                // bits pending in acc (including the foreign bits of the head word)
    int wi;                  // index of the word `acc`'s top half goes to
    bool head;

    JG_DEV void init(uint32_t* b, unsigned bitpos)
    {
        buf = b; acc = 0; fill = (int)(bitpos & 31u); wi = (int)(bitpos >> 5); head = true;
    }
    JG_DEV void flush()
    {
        const unsigned w = (unsigned)(acc >> 32);
        if (head) { smem_atomic_or(buf + wi, w); head = false; }
        else buf[wi] = w;
        ++wi; acc <<= 32; fill -= 32;
    }
    JG_DEV void put(unsigned val, unsigned len)  // 1 <= len <= 26, val < 2^len
    {
        acc |= (unsigned long long)val << (64 - fill - (int)len);
        fill += (int)len;
        if (fill >= 32) flush();
    }
    JG_DEV void finish()
    {
        if (fill > 0) smem_atomic_or(buf + wi, (unsigned)(acc >> 32));
    }
};

// walk 2: emit the block's bits at `bitpos` of `buf`
JG_DEV void block_pack(const int (&c)[64], int diff, const uint32_t* ac, const uint32_t* dc, uint32_t* buf, unsigned bitpos)
{
    BitPacker bp;
    bp.init(buf, bitpos);
    const unsigned zrl = ac[0xF0];
    {
        const unsigned cat = diff ? category(diff) : 0u;
        const unsigned e = dc[cat];
        const unsigned amp = diff ? amplitude(diff, cat) : 0u;
        bp.put(((e >> 8) << cat) | amp, (e & 0xffu) + cat);
    }
    int last = 0;
#pragma unroll
    for (int k = 1; k < 64; ++k) {
        const int v = c[k];
        if (v != 0) {
            unsigned run = (unsigned)(k - 1 - last);
            last = k;
            while (run >= 16u) { bp.put(zrl >> 8, zrl & 0xffu); run -= 16u; }
            const unsigned cat = category(v);
            const unsigned e = ac[(run << 4) | cat];
            bp.put(((e >> 8) << cat) | amplitude(v, cat), (e & 0xffu) + cat);
        }
    }
    if (last != 63) { const unsigned e = ac[0]; bp.put(e >> 8, e & 0xffu); }
    bp.finish();
}

// ------------------------------------------------------------------------------------------
// CTA-wide helpers
// ------------------------------------------------------------------------------------------
// exclusive scan over the 192 threads; contains two barriers
JG_DEV unsigned cta_scan_excl(unsigned v, uint32_t* warp_tmp, unsigned& total)
{
    const int lane = JG_TID & 31, wid = JG_TID >> 5;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned n = warp_shfl_up_u32(inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) warp_tmp[wid] = inc;
    cta_sync();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const unsigned x = warp_tmp[w];
        if (w < wid) base += x;
        tot += x;
    }
    cta_sync();
    total = tot;
    return base + inc - v;
}

JG_DEV unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += warp_shfl_xor_u64(v, m);
    return v;
}

// Decoupled look-back (Merrill & Garland) executed by one full warp.  desc[i] holds
// status[63:62] | payload.  Tiles before `first` do not exist (image start): they count
// as PREFIX 0.  Returns the exclusive prefix of tile g; *nearest receives desc[g-1].
// On timeout sets *err and returns 0 (all lanes agree).
JG_DEV unsigned long long lookback(const unsigned long long* desc, int g, int first, unsigned long long value_mask,
                                   unsigned long long* nearest, unsigned* err_flag, int* timed_out)
{
    const int lane = JG_TID & 31;
    unsigned long long running = 0;
    bool first_round = true;
    *timed_out = 0;
    for (int base = g - 1;; base -= 32) {
        const int idx = base - lane;
        const bool in_range = idx >= first;
        unsigned long long w = kStatusPrefix;  // virtual tile before the image: PREFIX 0
        unsigned spins = 0;
        for (;;) {
            if (in_range) w = ld_flag64(desc + idx);
            if (warp_ballot(in_range && (w >> 62) == 0) == 0u) break;
            if (++spins > kSpinLimit || ld_flag32(err_flag) != 0u) spins = 0xffffffffu;
            if (warp_ballot(spins == 0xffffffffu) != 0u) { *timed_out = 1; return 0; }
            backoff();
        }
        if (first_round) { *nearest = warp_shfl_u64(w, 0); first_round = false; }
        const unsigned pmask = warp_ballot((w >> 62) == 2u);
        const int stop = pmask ? i_ffs(pmask) - 1 : 32;   // nearest tile that already knows its prefix
        running += warp_sum_u64(lane <= stop ? (w & value_mask) : 0ull);
        if (pmask) return running;
    }
}

// n (<= 8) bits starting at bit `pos` of the MSB-first word array L
JG_DEV unsigned peek_bits(const uint32_t* L, unsigned pos, unsigned n)
{
    if (n == 0) return 0;
    const unsigned i = pos >> 5, s = pos & 31u;
    const unsigned long long two = ((unsigned long long)L[i] << 32) | L[i + 1];
    return (unsigned)(two >> (64u - s - n)) & ((1u << n) - 1u);
}

// word i of the byte-aligned stream X = (k head bits) ++ (local stream L); L[-1] := head bits
JG_DEV unsigned xword(const uint32_t* L, int i, unsigned k, unsigned hb)
{
    if (k == 0) return L[i];
    const unsigned hi = i == 0 ? hb : L[i - 1];
    return (hi << (32u - k)) | (L[i] >> k);
}

// ------------------------------------------------------------------------------------------
// stage 1: pixels -> shared memory
// ------------------------------------------------------------------------------------------
template <int LAYOUT, int NC>
JG_DEV void stage_tile(uint32_t* stage, const ImageDesc& im, int m0, int nM)
{
    using G = Geo<LAYOUT, NC>;
    constexpr int COLS = G::SLOTS * G::WPR;
    for (int col = JG_TID; col < COLS; col += kThreads) {
        const int j = col / G::WPR, wd = col - j * G::WPR;
        const int m = m0 - 1 + j;
        if (m < 0 || j > nM) continue;
        const int my = m / im.mcus_x, mx = m - my * im.mcus_x;
        const int y0 = my * G::MCU_H;
        const int xb = mx * G::ROWB + wd * 4;  // byte offset of this word inside a pixel row
        uint32_t* dst = stage + j * G::WPR + wd;
        if (im.aligned4 && (mx + 1) * G::MCU_W <= im.w) {
            uint32_t v[G::MCU_H];
#pragma unroll
            for (int r = 0; r < G::MCU_H; ++r) {
                const int y = (y0 + r < im.h) ? y0 + r : im.h - 1;   // replicate the last row
                v[r] = ldg_u32(im.px + (size_t)y * (size_t)im.stride + (size_t)xb);
            }
#pragma unroll
            for (int r = 0; r < G::MCU_H; ++r) dst[r * (G::SLOTS * G::WPR)] = v[r];
        } else {
            for (int r = 0; r < G::MCU_H; ++r) {
                const int y = (y0 + r < im.h) ? y0 + r : im.h - 1;
                const uint8_t* row = im.px + (size_t)y * (size_t)im.stride;
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int bo = xb + b;
                    int x = bo / NC;
                    const int ch = bo - x * NC;
                    if (x >= im.w) x = im.w - 1;                      // replicate the last column
                    v |= (uint32_t)ldg_u8(row + (size_t)x * NC + ch) << (8 * b);
                }
                dst[r * (G::SLOTS * G::WPR)] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-thread compute for one table class (0 = luma, 1 = chroma); leaves the coefficients in `c`
// ------------------------------------------------------------------------------------------
template <int LAYOUT, int NC, int CLS>
JG_DEV void thread_transform(Smem<LAYOUT, NC>& S, const QuantSet& Q, int comp, bool active, int slot, int q, int s,
                             bool pred_outside, bool have_prev_mcu, int pred_q, int (&c)[64], int& outside_dc)
{
    outside_dc = 0;
    if (active) {
        const ChromaK ck = chroma_k(comp);
        transform_block<LAYOUT, NC, CLS>(S.a, slot, q, ck, Q, c);
        S.dc_s[s] = c[0];
        if (pred_outside && have_prev_mcu) outside_dc = transform_dc_only<LAYOUT, NC, CLS>(S.a, 0, pred_q, ck, Q);
    }
}

// byte offset `o` of a little-endian byte array held as aligned words
JG_DEV unsigned word_at_byte(const uint32_t* w, unsigned o)
{
    const unsigned i = o >> 2, sh = (o & 3u) * 8u;
    return sh ? (w[i] >> sh) | (w[i + 1] << (32u - sh)) : w[i];
}

template <int LAYOUT, int NC>
JG_DEV void encode_tile(const LaunchParams& P, const QuantSet& Q, Smem<LAYOUT, NC>& S, const int g)
{
    using G = Geo<LAYOUT, NC>;
    const int t = JG_TID;

    // ---- which image, which MCUs -------------------------------------------------------
    int img_idx;
    if (P.tiles_per_image > 0) {
        img_idx = g / P.tiles_per_image;
    } else {
        int lo = 0, hi = P.n_images - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.images[mid].first_tile <= g) lo = mid; else hi = mid - 1;
        }
        img_idx = lo;
    }
    const ImageDesc im = P.images[img_idx];
    const int lt = g - im.first_tile;
    const int m0 = lt * G::M;
    const int nM = (im.n_mcus - m0 < G::M) ? im.n_mcus - m0 : G::M;
    const int nblk = nM * G::BPM;
    const bool first_tile = lt == 0, last_tile = lt == im.n_tiles - 1;

    // ---- 1. stage ----------------------------------------------------------------------
    stage_tile<LAYOUT, NC>(S.a, im, m0, nM);
    cta_sync();

    // ---- 2. transform: thread -> (MCU slot, component, quadrant), stream index s ---------
    int comp, ml, q = 0, s, pred_s, pred_q = 0;
    bool pred_outside;
    if (LAYOUT == LAYOUT_444) {
        comp = t >> 6; ml = t & 63; s = ml * 3 + comp;
        pred_s = s - 3; pred_outside = ml == 0;
    } else if (LAYOUT == LAYOUT_420) {
        if (t < 128) {
            comp = 0; ml = (t & 63) >> 1; q = ((t >> 6) << 1) | (t & 1); s = ml * 6 + q;
            pred_s = q ? s - 1 : s - 3; pred_outside = (q == 0) && (ml == 0); pred_q = 3;
        } else {
            comp = 1 + ((t - 128) >> 5); ml = (t - 128) & 31; s = ml * 6 + 3 + comp;
            pred_s = s - 6; pred_outside = ml == 0;
        }
    } else {
        comp = 0; ml = t; s = t; pred_s = s - 1; pred_outside = ml == 0;
    }
    const bool active = ml < nM;
    const int cls = comp ? 1 : 0;
    int c[64];
    int outside_dc;
    if (cls == 0) thread_transform<LAYOUT, NC, 0>(S, Q, comp, active, ml + 1, q, s, pred_outside, m0 > 0, pred_q, c, outside_dc);
    else thread_transform<LAYOUT, NC, 1>(S, Q, comp, active, ml + 1, q, s, pred_outside, m0 > 0, pred_q, c, outside_dc);
    cta_sync();   // dc_s complete; staging area dead from here on

    // ---- 3. size ------------------------------------------------------------------------
    const uint32_t* hac = S.huff_ac[cls];
    const uint32_t* hdc = S.huff_dc[cls];
    int diff = 0;
    if (active) {
        const int pred = pred_outside ? outside_dc : S.dc_s[pred_s];   // jpeg_enc.h:834-835
        diff = c[0] - pred;
        const unsigned my_bits = block_bits(c, diff, hac, hdc);
        S.bits_s[s] = my_bits;
        if (P.dbg_bits) P.dbg_bits[im.first_block + (unsigned long long)(m0 * G::BPM + s)] = my_bits;
        if (P.dbg_coefs) {
            int16_t* dst = P.dbg_coefs + (im.first_block + (unsigned long long)(m0 * G::BPM + s)) * 64ull;
#pragma unroll
            for (int i = 0; i < 64; ++i) dst[i] = (int16_t)c[i];
        }
    }
    cta_sync();
    unsigned T;
    {
        const unsigned v = t < nblk ? S.bits_s[t] : 0u;
        const unsigned ex = cta_scan_excl(v, S.warp_tmp, T);
        S.off_s[t] = ex;
        if (t == 0) S.off_s[kBlocksPerTile] = T;
    }
    // Publish the tile's bit count NOW, before packing: successors can then resolve their
    // look-back while we are still busy (the look-back needs every predecessor's count).
    const unsigned cap_bits = (unsigned)P.win_words * 32u - 64u;
    if (t == 0) {
        st_flag64(P.desc_bits + g, (first_tile ? kStatusPrefix : kStatusAgg) | (unsigned long long)T);
        S.n_groups = 1;
    }
    cta_sync();   // off_s visible
    if (T > cap_bits) {   // pathological tile: split into groups that fit the window
        if (t == 0) {
            int ng = 0;
            unsigned gbase = 0;
            S.gstart[0] = 0;
            for (int b = 0; b < nblk; ++b) {
                const unsigned end = S.off_s[b] + S.bits_s[b];
                if (end - gbase > cap_bits) { ++ng; S.gstart[ng] = (uint16_t)b; gbase = S.off_s[b]; }
            }
            ++ng;
            S.gstart[ng] = (uint16_t)nblk;
            S.n_groups = ng;
        }
        cta_sync();
    }
    const int n_groups = S.n_groups;

    // ---- 4..7 as a list of jobs with ONE pack call site ------------------------------------
    //   one group   : [ALL]
    //   many groups : [TAIL, COUNT_0..COUNT_{n-1}, EMIT_0..EMIT_{n-1}]   (walk 2 is repeated)
    const int n_jobs = n_groups == 1 ? 1 : 1 + 2 * n_groups;
    unsigned long long bit_base = 0, pos = 0;
    unsigned k0 = 0, hb0 = 0, k = 0, hb = 0;
    unsigned ff_tile = 0;
    bool overflow = false, have_pos = false;

    for (int job = 0; job < n_jobs; ++job) {
        const bool is_all = n_groups == 1;
        const bool is_tail = !is_all && job == 0;
        const bool is_count = !is_all && job >= 1 && job <= n_groups;
        const bool is_emit = !is_all && job > n_groups;
        const int j = is_all ? 0 : (is_count ? job - 1 : (is_emit ? job - 1 - n_groups : 0));
        int b0, b1;
        if (is_all) { b0 = 0; b1 = nblk; }
        else if (is_tail) { b0 = nblk >= 2 ? nblk - 2 : 0; b1 = nblk; }   // only to learn the last 7 bits
        else { b0 = S.gstart[j]; b1 = S.gstart[j + 1]; }

        // ---- 4. pack blocks [b0,b1) at their offsets relative to block b0 ----------------------
        const unsigned base = S.off_s[b0];
        const unsigned tg = S.off_s[b1] - base;
        for (int i = t; i < (int)(tg >> 5) + 3; i += kThreads) S.a[i] = 0u;
        cta_sync();
        if (active && s >= b0 && s < b1) block_pack(c, diff, hac, hdc, S.a, S.off_s[s] - base);
        cta_sync();

        // ---- 5. chain #1: publish our last 7 bits, learn our bit offset --------------------------
        if (is_all || is_tail) {
            if (t < 32) {
                const unsigned tail = tg >= 7u ? peek_bits(S.a, tg - 7u, 7u) : peek_bits(S.a, 0u, tg);
                if (t == 0) st_flag64(P.desc_tail + g, kStatusAgg | (unsigned long long)tail);
                unsigned long long excl = 0, nearest = 0, ptail = 0;
                int timed_out = 0;
                if (!first_tile) {
                    excl = lookback(P.desc_bits, g, im.first_tile, kCountMask, &nearest, P.error, &timed_out);
                    if (t == 0 && !timed_out) st_flag64(P.desc_bits + g, kStatusPrefix | (excl + T));
                    // the byte we share with the predecessor needs its last bits
                    unsigned spins = 0;
                    while (!timed_out && ((ptail = ld_flag64(P.desc_tail + g - 1)) >> 62) == 0) {
                        if (++spins > kSpinLimit || ld_flag32(P.error) != 0u) timed_out = 1;
                        backoff();
                    }
                    timed_out = warp_ballot(timed_out) != 0u;
                }
                if (t == 0) {
                    S.bit_base = excl;
                    S.pred_tail = (unsigned)ptail & 0x7fu;
                    S.abort = timed_out;
                    if (timed_out) gmem_atomic_or(P.error, 1u);
                }
            }
            cta_sync();
            if (S.abort) return;
            bit_base = S.bit_base;
            k0 = (unsigned)(bit_base & 7ull);            // bits of our first byte owned by the predecessor
            hb0 = S.pred_tail & ((1u << k0) - 1u);
            k = k0; hb = hb0;
            if (is_tail) continue;
        }

        // ---- 6. geometry of the byte-aligned stream X = (k head bits) ++ (group bits) ---------------
        const bool final_group = last_tile && j == n_groups - 1;
        unsigned n_bytes = (k + tg) >> 3, k_out = (k + tg) & 7u, hb_out = 0;
        if (k_out) {
            if (final_group) { n_bytes += 1; k_out = 0; }             // zero padding, jpeg_enc.h:1161-1164
            else if (tg >= k_out) hb_out = peek_bits(S.a, tg - k_out, k_out);
            else hb_out = ((hb << tg) | peek_bits(S.a, 0u, tg)) & ((1u << k_out) - 1u);
        }
        // per-thread chunk of X words; the odd stride keeps the shared-memory banks apart
        const unsigned cw = (((n_bytes + 3u) / 4u + kThreads - 1u) / kThreads) | 1u;
        const unsigned w_lo = (unsigned)t * cw;
        unsigned cnt = 0;
        for (unsigned i = w_lo; i < w_lo + cw && i * 4u < n_bytes; ++i) {
            unsigned m = v_cmpeq4(xword(S.a, (int)i, k, hb), 0xffffffffu);
            const unsigned valid = n_bytes - i * 4u;
            if (valid < 4u) m &= 0xffffffffu << (8u * (4u - valid));
            cnt += (unsigned)i_popc(m) >> 3;
        }
        unsigned ff_group;
        const unsigned ff_ex = cta_scan_excl(cnt, S.warp_tmp, ff_group);
        if (!is_emit) ff_tile += ff_group;

        // publish the stuffed-byte count as soon as it is complete
        if ((is_all || (is_count && j == n_groups - 1)) && t == 0)
            st_flag64(P.desc_ff + g, (first_tile ? kStatusPrefix : kStatusAgg) | (unsigned long long)ff_tile);

        if (is_count) {
            if (j == n_groups - 1) { k = k0; hb = hb0; } else { k = k_out; hb = hb_out; }
            continue;
        }

        // ---- 7a. emit stuffed bytes into sbuf (position independent) ----------------------------------
        {
            unsigned o = w_lo * 4u + ff_ex;
            for (unsigned i = w_lo; i < w_lo + cw && i * 4u < n_bytes; ++i) {
                const unsigned x = xword(S.a, (int)i, k, hb);
                const unsigned valid = n_bytes - i * 4u < 4u ? n_bytes - i * 4u : 4u;
                for (unsigned b = 0; b < valid; ++b) {
                    const unsigned byte = (x >> (24u - 8u * b)) & 0xffu;
                    S.sbuf[o++] = (uint8_t)byte;
                    if (byte == 0xffu) S.sbuf[o++] = 0;               // jpeg_enc.h:634-638
                }
            }
        }
        // ---- 6b. chain #2 (after the emit, so predecessors had time to publish) ---------------------------
        if (!have_pos) {
            if (t < 32) {
                unsigned long long excl = 0, nearest = 0;
                int timed_out = 0;
                if (!first_tile) {
                    excl = lookback(P.desc_ff, g, im.first_tile, kCountMask, &nearest, P.error, &timed_out);
                    if (t == 0 && !timed_out) st_flag64(P.desc_ff + g, kStatusPrefix | (excl + ff_tile));
                }
                if (t == 0) {
                    S.ff_base = excl;
                    S.abort = timed_out;
                    if (timed_out) gmem_atomic_or(P.error, 1u);
                }
            }
            cta_sync();   // also orders the sbuf writes above
            if (S.abort) return;
            pos = (bit_base >> 3) + S.ff_base;   // byte position of the tile's first owned byte
            have_pos = true;
        } else {
            cta_sync();
        }

        // ---- 7b. copy out: bytes to the first 16B boundary, aligned 16B stores, tail bytes ------------------
        const unsigned out_bytes = n_bytes + ff_group;
        if (pos + out_bytes + (last_tile ? 2u : 0u) > im.out_cap) {
            overflow = true;
        } else {
            uint8_t* dst = im.out + pos;
            unsigned head = (16u - (unsigned)((size_t)dst & 15u)) & 15u;
            if (head > out_bytes) head = out_bytes;
            const unsigned nvec = (out_bytes - head) >> 4;
            const unsigned tail0 = head + (nvec << 4);
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(S.sbuf);
            if ((unsigned)t < head) dst[t] = S.sbuf[t];
            uint4* dst4 = reinterpret_cast<uint4*>(dst + head);
            for (unsigned i = (unsigned)t; i < nvec; i += kThreads) {
                const unsigned o = head + (i << 4);
                uint4 v;
                v.x = word_at_byte(sw, o); v.y = word_at_byte(sw, o + 4u);
                v.z = word_at_byte(sw, o + 8u); v.w = word_at_byte(sw, o + 12u);
                dst4[i] = v;
            }
            if (tail0 + (unsigned)t < out_bytes) dst[tail0 + t] = S.sbuf[tail0 + t];
        }
        pos += out_bytes;
        k = k_out; hb = hb_out;
        if (job + 1 < n_jobs) cta_sync();   // sbuf / window are reused by the next group
    }
    if (t == 0) {
        if (last_tile) {
            if (!overflow) { im.out[pos] = 0xFF; im.out[pos + 1] = 0xD9; }    // EOI, jpeg_enc.h:1166-1167
            P.scan_bytes[img_idx] = pos + 2;
        }
        if (overflow) gmem_atomic_or(P.img_status + img_idx, 1u);
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int LAYOUT, int NC>
JG_KERNEL(kThreads, 2)
void encode_tiles_kernel(const JG_GRID_CONSTANT LaunchParams P, const JG_GRID_CONSTANT QuantSet Q)
{
    JG_DYNAMIC_SMEM(smem_raw);
    Smem<LAYOUT, NC>& S = *reinterpret_cast<Smem<LAYOUT, NC>*>(smem_raw);
    const int t = JG_TID;
    for (int i = t; i < 512; i += kThreads) (&S.huff_ac[0][0])[i] = (&P.huff->ac[0][0])[i];
    if (t < 32) (&S.huff_dc[0][0])[t] = (&P.huff->dc[0][0])[t];
    for (;;) {
        cta_sync();   // everyone is done with the previous tile's shared state
        if (t == 0) {
            S.tile = (int)gmem_atomic_add(P.ticket, 1u);
            S.abort = ld_flag32(P.error) != 0u;
        }
        cta_sync();
        const int g = S.tile;
        if (g >= P.n_tiles || S.abort) break;
        encode_tile<LAYOUT, NC>(P, Q, S, g);
    }
}

}  // namespace jg
