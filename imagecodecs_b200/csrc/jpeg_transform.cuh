// jpeg_transform.cuh -- pass A of the split pipeline: pixels -> quantised coefficients in HBM.
//
// Replaces, for every 8x8 block of the batch, the reference's block gather + edge replication
// (jpeg_enc.h:1094-1112), RGB -> YCbCr (:1114-1124), tjei_fdct (:656-763) and the quantiser of
// tjei_encode_and_write_MCU (:799-817): the result is `du[64]` of that function, as int16 in ZIGZAG
// order, 128 bytes per block, blocks in the order the scan codes them (MCU after MCU, inside an MCU
// Y.. Cb Cr, :1128-1154).  The entropy pass (jpeg_entropy.cuh) reads nothing else.
//
// Why its own kernel: no chain, no tickets, no queue -- a warp's whole state is one MCU row of floats,
// the hot loop is ~6 KB of straight-line FADD2/FMUL, and the arithmetic is issued two values at a time
// wherever two identically shaped computations exist:
//   4:4:4  8 lanes own TWO MCUs (m and m+4 of the warp's eight); lane u holds pixel row u of both, so
//          Y, Cb and Cr of the pair go through the packed passes together (the fused kernel could
//          pack only Cb with Cr).
//   4:2:0  16 lanes own one MCU; the left and the right luma block of a row travel packed.
//   gray   8 lanes own two consecutive blocks.
// The 8x8 transposition between the row and the column pass goes through shared memory as float2
// (STS.64 / LDS.64: both members of a pair in one access), rows padded to 9 -- conflict-free both ways.
// Multiplications stay scalar (ptxas fuses FMUL2 into a following FADD2, see jpeg_device.h).
// Coefficients are scattered to zigzag order in shared memory (16-bit stores) and leave as 16-byte
// vectors: 128 contiguous bytes per block, 3 KB per warp iteration.
#pragma once
#include "jpeg_kernel.cuh"

namespace jg {

template <int LAYOUT>
struct TGeo {
    static constexpr int BPM = LAYOUT == LAYOUT_444 ? 3 : (LAYOUT == LAYOUT_420 ? 6 : 1);
    static constexpr int ITER_MCUS = LAYOUT == LAYOUT_420 ? 2 : 8;          // MCUs one warp iteration transforms
    static constexpr int ITERS = LAYOUT == LAYOUT_444 ? 4 : 8;              // iterations per item
    static constexpr int ITEM_MCUS = ITER_MCUS * ITERS;                     // 32 / 16 / 64
    static constexpr int ITER_BLOCKS = ITER_MCUS * BPM;                     // 24 / 12 / 8
    // exchange space of one lane group, in floats: 4:4:4 three pair tiles, 4:2:0 two pair tiles + two single tiles, gray one pair tile
    static constexpr int GROUP_FLOATS = LAYOUT == LAYOUT_444 ? 3 * 2 * kTileFloats : (LAYOUT == LAYOUT_420 ? 6 * kTileFloats : 2 * kTileFloats);
    static constexpr int GROUPS = LAYOUT == LAYOUT_420 ? 2 : 4;
    // Staging area: block b of the iteration at int16 offset b * kCoefStride + (b / BPM) * MCU_PAD.  A 16-bit scatter store
    // covers FOUR blocks (8 lanes each) whose 8 zigzag positions fall on pseudo-random banks; the four patterns interfere
    // least when the blocks sit 8 banks apart (2.5 wavefronts per store; any placement: >= 2.5, brute force).  With 144-byte
    // blocks alone the four blocks of a 4:2:0 store are 0, 8, 24 and 0 banks apart (3.75 per store, measured 4), those of a
    // 4:4:4 store 0, 12, 24, 4 (2.75): a pad of 96 / 48 bytes per MCU puts them at 0, 8, 16, 24.
    static constexpr int MCU_PAD = LAYOUT == LAYOUT_420 ? 48 : (LAYOUT == LAYOUT_444 ? 24 : 0);
    static constexpr int STAGE_INT16 = ITER_BLOCKS * kCoefStride + ITER_MCUS * MCU_PAD;
    static JG_DEV int stage_off(int b) { return b * kCoefStride + (b / BPM) * MCU_PAD; }
    static JG_DEV int stage_off(int mcu, int j) { return (mcu * BPM + j) * kCoefStride + mcu * MCU_PAD; }   // block j of MCU mcu
};
static_assert(TGeo<LAYOUT_444>::ITEM_MCUS == transform_item_mcus(LAYOUT_444) && TGeo<LAYOUT_420>::ITEM_MCUS == transform_item_mcus(LAYOUT_420) &&
              TGeo<LAYOUT_GRAY>::ITEM_MCUS == transform_item_mcus(LAYOUT_GRAY), "host and kernel agree on the item size");

template <int LAYOUT>
struct TWarp {
    using G = TGeo<LAYOUT>;
    alignas(16) float xch[G::GROUPS * G::GROUP_FLOATS];                 // row pass -> column pass
    alignas(16) int16_t stage[G::STAGE_INT16];                           // zigzag scatter -> 16-byte vectors
};
template <int LAYOUT>
struct TSmem { TWarp<LAYOUT> wm[kWarps]; };

struct TLane {
    float pq_l[8], pq_c[8];      // reciprocal quantisers of column u = lane & 7: [v] = pqt[8v+u]
    unsigned zz2[8];             // byte offset of (v, u) inside a zigzag-ordered block
};

// (v*pq + 1024) + 0.5, floor, - 1024 (jpeg_enc.h:808-816) for a pair; the results are the LOW 16 BITS of kx / ky.
// b = (t + 1024) + 0.5 as the reference rounds it; then ONE more packed add, rounded toward minus infinity, of
// 1.5 * 2^23 - 1024: the sum lies in [2^23, 2^24) where the ulp is 1, so it equals 1.5 * 2^23 + floor(b) - 1024 exactly
// and, 1.5 * 2^23 being 0x4B400000, the low 16 bits of its bit pattern are floor(b) - 1024 in two's complement.
// No F2I, no integer subtraction: the XU pipe (16 lanes per clock) stays out of the quantiser.
constexpr float kFloorMagic = 12582912.0f - 1024.0f;
JG_DEV void quantise2(f32x2 c, float pq, unsigned& kx, unsigned& ky)
{
    const f32x2 b = f2_add(f2_add(f2_mul(c, f2(pq, pq)), f2(1024.0f, 1024.0f)), f2(0.5f, 0.5f));
    const f32x2 r = f2_add_rd(b, f2(kFloorMagic, kFloorMagic));
    kx = f_bits(r.x);
    ky = f_bits(r.y);
}
JG_DEV unsigned quantise1(float c, float pq)
{
    return f_bits(f_add_rd(f_add(f_add(f_mul(c, pq), 1024.0f), 0.5f), kFloorMagic));
}
// Channels of pixel i of a loaded row segment as 2^23 + value (see u8_biased): one PRMT each
template <int NC>
JG_DEV void rgb_biased(const uint32_t* w, int i, float& r, float& g, float& b)
{
    if (NC == 4) { r = u8_biased(w[i], 0); g = u8_biased(w[i], 1); b = u8_biased(w[i], 2); }
    else {
        r = u8_biased(w[(3 * i) >> 2], (3 * i) & 3);
        g = u8_biased(w[(3 * i + 1) >> 2], (3 * i + 1) & 3);
        b = u8_biased(w[(3 * i + 2) >> 2], (3 * i + 2) & 3);
    }
}
#ifndef JG_U8_VIA_PRMT
#define JG_U8_VIA_PRMT 1      // 0: I2F.U8 (XU pipe) -- kept for the A/B in DESIGN.md
#endif
// R, G, B of pixel i of two row segments, as pairs
template <int NC>
JG_DEV void rgb_pair(const uint32_t* wa, const uint32_t* wb, int ia, int ib, f32x2& R, f32x2& G, f32x2& B)
{
    float r0, g0, b0, r1, g1, b1;
#if JG_U8_VIA_PRMT
    rgb_biased<NC>(wa, ia, r0, g0, b0);
    rgb_biased<NC>(wb, ib, r1, g1, b1);
    const f32x2 bias = f2(kU8Bias, kU8Bias);
    R = f2_sub(f2(r0, r1), bias); G = f2_sub(f2(g0, g1), bias); B = f2_sub(f2(b0, b1), bias);
#else
    rgb_of<NC>(wa, ia, r0, g0, b0);
    rgb_of<NC>(wb, ib, r1, g1, b1);
    R = f2(r0, r1); G = f2(g0, g1); B = f2(b0, b1);
#endif
}

// Shared-memory addresses (32-bit, see jpeg_device.h) a lane works with: fixed for the whole kernel
struct TAddr {
    unsigned xch;        // the lane group's exchange space
    unsigned stage;      // the warp's staging area
};

// copy `nblk` staged blocks (at TGeo::stage_off) to global memory, 16 bytes per lane and step
template <int LAYOUT>
JG_DEV void stage_to_global(unsigned stage, int16_t* gout, int nblk)
{
    using G = TGeo<LAYOUT>;
    const int t = JG_TID & 31;
#pragma unroll
    for (int q = 0; q < (G::ITER_BLOCKS * 8 + 31) / 32; ++q) {
        const int c = q * 32 + t, blk = c >> 3, part = c & 7;
        if (blk < nblk) *reinterpret_cast<uint4*>(gout + blk * 64 + part * 8) = lds_v4(stage + (unsigned)(G::stage_off(blk) + part * 8) * 2u);
    }
}

// The pixels one lane needs for one iteration, as loaded words.  They are fetched ONE ITERATION AHEAD (right after the
// previous iteration has turned its own pixels into floats), so the global-memory latency runs under the row and column
// passes instead of in front of them.
template <int LAYOUT, int NC>
struct TPixels {
    static constexpr int WORDS = LAYOUT == LAYOUT_420 ? 16 * NC / 4 : (LAYOUT == LAYOUT_444 ? 8 * NC / 4 : 2);
    uint32_t a[WORDS];          // 4:2:0: the lane's 16-pixel row; 4:4:4 / gray: its row of MCU (block) A
    uint32_t b[LAYOUT == LAYOUT_420 ? 1 : WORDS];   // 4:4:4 / gray: its row of MCU (block) B
};

template <int LAYOUT, int NC>
JG_DEV void fetch_pixels(const ImageDesc& im, int my0, int mx0, int nM, TPixels<LAYOUT, NC>& px)
{
    const int t = JG_TID & 31, u = t & 7;
#pragma unroll
    for (int i = 0; i < TPixels<LAYOUT, NC>::WORDS; ++i) px.a[i] = 0u;
    if (LAYOUT == LAYOUT_420) {
        const int grp = t >> 4, r16 = t & 15;
        if (grp < nM) {
            int my, mx;
            mcu_pos(im, my0, mx0, grp, my, mx);
            int y = my * 16 + r16; if (y >= im.h) y = im.h - 1;            // replicate the last row (jpeg_enc.h:1106-1111)
            load_segment<NC, 16>(im, mx * 16, y, reinterpret_cast<uint32_t (&)[16 * NC / 4]>(px.a));
        }
    } else {
        constexpr int C = LAYOUT == LAYOUT_444 ? NC : 1;
        const int g = t >> 3;
        const int sa = LAYOUT == LAYOUT_444 ? g : 2 * g, sb = LAYOUT == LAYOUT_444 ? g + 4 : 2 * g + 1;   // the lane group's two MCUs / blocks
#pragma unroll
        for (int i = 0; i < TPixels<LAYOUT, NC>::WORDS; ++i) px.b[i] = 0u;
        if (sa < nM) {
            int my, mx;
            mcu_pos(im, my0, mx0, sa, my, mx);
            int y = my * 8 + u; if (y >= im.h) y = im.h - 1;
            load_segment<C, 8>(im, mx * 8, y, reinterpret_cast<uint32_t (&)[8 * C / 4]>(px.a));
        }
        if (sb < nM) {
            int my, mx;
            mcu_pos(im, my0, mx0, sb, my, mx);
            int y = my * 8 + u; if (y >= im.h) y = im.h - 1;
            load_segment<C, 8>(im, mx * 8, y, reinterpret_cast<uint32_t (&)[8 * C / 4]>(px.b));
        }
    }
}

// ---- 4:4:4: eight MCUs per iteration, lane group g owns MCUs g and g + 4 -----------------------------------------
template <int NC, class Prefetch>
JG_DEV void transform_iter_444(const TAddr& A, const TPixels<LAYOUT_444, NC>& px, int nM, int16_t* gout, const TLane& LC, Prefetch&& prefetch)
{
    const int t = JG_TID & 31, u = t & 7, g = t >> 3;
    f32x2 sy[8], sb[8], sr[8];       // .x = MCU A, .y = MCU B
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f32x2 R, Gc, B;
        rgb_pair<NC>(px.a, px.b, i, i, R, Gc, B);
        // jpeg_enc.h:1118-1120, additions packed over the two MCUs
        sy[i] = f2_sub(f2_add(f2_add(f2_mul(f2(0.299f, 0.299f), R), f2_mul(f2(0.587f, 0.587f), Gc)), f2_mul(f2(0.114f, 0.114f), B)),
                       f2(128.0f, 128.0f));
        sb[i] = f2_add(f2_sub(f2_mul(f2(-0.1687f, -0.1687f), R), f2_mul(f2(0.3313f, 0.3313f), Gc)), f2_mul(f2(0.5f, 0.5f), B));
        sr[i] = f2_sub(f2_sub(f2_mul(f2(0.5f, 0.5f), R), f2_mul(f2(0.4187f, 0.4187f), Gc)), f2_mul(f2(0.0813f, 0.0813f), B));
    }
    prefetch();                       // the pixel registers are free: the next iteration's loads go out now
    const unsigned row = A.xch + (unsigned)(u * 9) * 8u;
    aan8x2(sy);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_v2f(row + 8u * i, sy[i]);
    aan8x2(sb);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_v2f(row + 8u * (kTileFloats + i), sb[i]);
    aan8x2(sr);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_v2f(row + 8u * (2 * kTileFloats + i), sr[i]);
    warp_sync();
    const unsigned colA = A.xch + 8u * (unsigned)u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        f32x2 col[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) col[v] = lds_v2f(colA + 8u * (unsigned)(c * kTileFloats + v * 9));
        aan8x2(col);
        const unsigned dA = A.stage + 2u * (unsigned)TGeo<LAYOUT_444>::stage_off(g, c), dB = A.stage + 2u * (unsigned)TGeo<LAYOUT_444>::stage_off(g + 4, c);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            unsigned ka, kb;
            quantise2(col[v], c ? LC.pq_c[v] : LC.pq_l[v], ka, kb);
            sts_u16(dA + LC.zz2[v], ka);
            sts_u16(dB + LC.zz2[v], kb);   // (an absent MCU B computes on zeros into its own staging slot, never copied out)
        }
    }
    warp_sync();
    stage_to_global<LAYOUT_444>(A.stage, gout, nM * 3);
}

// ---- 4:2:0: two MCUs per iteration, 16 lanes per MCU, lane r16 owns pixel row r16 ---------------------------------
template <int NC, class Prefetch>
JG_DEV void transform_iter_420(const TAddr& A, const TPixels<LAYOUT_420, NC>& px, int nM, int16_t* gout, const TLane& LC, Prefetch&& prefetch)
{
    const int t = JG_TID & 31, u = t & 7, grp = t >> 4, r16 = t & 15, h = r16 >> 3;
    const unsigned ypair = A.xch;                                    // pair tile 0: (Y00, Y01), pair tile 1: (Y10, Y11)
    const unsigned ctile = A.xch + 4u * (4 * kTileFloats);           // Cb tile, Cr tile
    const bool valid = grp < nM;
    f32x2 cbs[4], crs[4];            // horizontal pair sums (a+b) of this row: .x samples 0-3, .y samples 4-7
    f32x2 sy[8];                     // .x = pixel i (left luma block), .y = pixel i + 8 (right luma block)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f32x2 cb[2], cr[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * j + e;
            f32x2 R, Gc, B;
            rgb_pair<NC>(px.a, px.a, i, i + 8, R, Gc, B);
            sy[i] = f2_sub(f2_add(f2_add(f2_mul(f2(0.299f, 0.299f), R), f2_mul(f2(0.587f, 0.587f), Gc)),
                                  f2_mul(f2(0.114f, 0.114f), B)), f2(128.0f, 128.0f));
            cb[e] = f2_add(f2_sub(f2_mul(f2(-0.1687f, -0.1687f), R), f2_mul(f2(0.3313f, 0.3313f), Gc)),
                           f2_mul(f2(0.5f, 0.5f), B));
            cr[e] = f2_sub(f2_sub(f2_mul(f2(0.5f, 0.5f), R), f2_mul(f2(0.4187f, 0.4187f), Gc)),
                           f2_mul(f2(0.0813f, 0.0813f), B));
        }
        cbs[j] = f2_add(cb[0], cb[1]);
        crs[j] = f2_add(cr[0], cr[1]);
    }
    prefetch();                       // the pixel registers are free: the next iteration's loads go out now
    aan8x2(sy);
    if (valid) {
        const unsigned ty = ypair + 8u * (unsigned)(h * kTileFloats + (r16 & 7) * 9);
#pragma unroll
        for (int i = 0; i < 8; ++i) sts_v2f(ty + 8u * i, sy[i]);
    }
    // vertical pairs live in neighbouring lanes: the even lane finishes Cb, the odd lane Cr;
    // sample = ((a+b) + (c+d)) * 0.25f with (a+b) from the even row (extended mode); the addition commutes
    const bool even = (r16 & 1) == 0;
    float samp[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const f32x2 mine = even ? cbs[i] : crs[i], give = even ? crs[i] : cbs[i];
        const f32x2 other = f2(warp_shfl_xor_f32(give.x, 1), warp_shfl_xor_f32(give.y, 1));
        const f32x2 q = f2_mul(f2_add(mine, other), f2(0.25f, 0.25f));
        samp[i] = q.x; samp[4 + i] = q.y;
    }
    aan8(samp);
    if (valid) {
        const unsigned tc = ctile + 4u * (unsigned)((r16 & 1) * kTileFloats + (r16 >> 1) * 9);
#pragma unroll
        for (int i = 0; i < 8; ++i) sts_f32(tc + 4u * i, samp[i]);
    }
    warp_sync();
    if (valid) {
        // lanes 0-7 finish (Y00, Y01) and Cb, lanes 8-15 (Y10, Y11) and Cr
        f32x2 col[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) col[v] = lds_v2f(ypair + 8u * (unsigned)(h * kTileFloats + v * 9 + u));
        aan8x2(col);
        const unsigned dL = A.stage + 2u * (unsigned)TGeo<LAYOUT_420>::stage_off(grp, 2 * h), dR = dL + 2u * kCoefStride;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            unsigned kl, kr;
            quantise2(col[v], LC.pq_l[v], kl, kr);
            sts_u16(dL + LC.zz2[v], kl);
            sts_u16(dR + LC.zz2[v], kr);
        }
        float c[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) c[v] = lds_f32(ctile + 4u * (unsigned)(h * kTileFloats + v * 9 + u));
        aan8(c);
        const unsigned dC = A.stage + 2u * (unsigned)TGeo<LAYOUT_420>::stage_off(grp, 4 + h);
#pragma unroll
        for (int v = 0; v < 8; ++v) sts_u16(dC + LC.zz2[v], quantise1(c[v], LC.pq_c[v]));
    }
    warp_sync();
    stage_to_global<LAYOUT_420>(A.stage, gout, nM * 6);
}

// ---- gray: eight blocks per iteration, lane group g owns blocks 2g and 2g + 1 ------------------------------------
template <class Prefetch>
JG_DEV void transform_iter_gray(const TAddr& A, const TPixels<LAYOUT_GRAY, 1>& px, int nM, int16_t* gout, const TLane& LC, Prefetch&& prefetch)
{
    const int t = JG_TID & 31, u = t & 7, g = t >> 3;
    f32x2 s[8];
#if JG_U8_VIA_PRMT
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s[i] = f2_sub(f2_sub(f2(u8_biased(px.a[i >> 2], i & 3), u8_biased(px.b[i >> 2], i & 3)), f2(kU8Bias, kU8Bias)), f2(128.0f, 128.0f));
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = f2_sub(f2(u8_to_f(byte_of(px.a, i)), u8_to_f(byte_of(px.b, i))), f2(128.0f, 128.0f));
#endif
    prefetch();
    aan8x2(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_v2f(A.xch + 8u * (unsigned)(u * 9 + i), s[i]);
    warp_sync();
    f32x2 col[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) col[v] = lds_v2f(A.xch + 8u * (unsigned)(v * 9 + u));
    aan8x2(col);
    const unsigned dA = A.stage + 2u * (unsigned)((2 * g) * kCoefStride), dB = dA + 2u * kCoefStride;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        unsigned ka, kb;
        quantise2(col[v], LC.pq_l[v], ka, kb);
        sts_u16(dA + LC.zz2[v], ka);
        sts_u16(dB + LC.zz2[v], kb);
    }
    warp_sync();
    stage_to_global<LAYOUT_GRAY>(A.stage, gout, nM);
}

// ------------------------------------------------------------------------------------------
// kernel A: one warp per item (ITEM_MCUS consecutive MCUs of one image)
// ------------------------------------------------------------------------------------------
// Resident CTAs per SM the register allocation aims at: 4 (128 registers) for 4:2:0 and gray, 3 (168) for 4:4:4, whose lanes
// hold two MCUs of three components.  Measured (tools/build_variants.py): 6 -> 5 -> 4 CTAs 1.31 / 1.27 / 1.06 ms for the
// 4:2:0 batch, 4 -> 3 CTAs 0.886 -> 0.840 ms for the 4:4:4 one (and 1.06 -> 1.10 for 4:2:0): registers beat occupancy here.
#ifndef JG_TR_MINB
#define JG_TR_MINB (LAYOUT == LAYOUT_444 ? 3 : 4)
#endif
template <int LAYOUT, int NC>
JG_KERNEL(kThreads, JG_TR_MINB)
void transform_kernel(const JG_GRID_CONSTANT TransformParams P, const JG_GRID_CONSTANT QuantSet Q)
{
    using G = TGeo<LAYOUT>;
    JG_DYNAMIC_SMEM(smem_raw);
    TSmem<LAYOUT>& S = *reinterpret_cast<TSmem<LAYOUT>*>(smem_raw);
    const int t = JG_TID;
    const unsigned item = (unsigned)JG_CTA_ID * (unsigned)kWarps + (unsigned)(t >> 5);
    if (item >= (unsigned)P.n_items) return;             // whole warps leave; no CTA barrier in this kernel
    TLane LC;
    {
        const int u = t & 7;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            LC.pq_l[v] = Q.luma[8 * v + u];
            LC.pq_c[v] = Q.chroma[8 * v + u];
            LC.zz2[v] = 2u * (unsigned)zz_of(8 * v + u);
        }
    }
    unsigned img, local;
    if (P.items_per_image > 0) {
        img = item / (unsigned)P.items_per_image;
        local = item - img * (unsigned)P.items_per_image;
    } else {
        unsigned lo = 0, hi = (unsigned)P.n_images - 1u;       // last image whose first item is <= item
        while (lo < hi) {
            const unsigned mid = (lo + hi + 1u) >> 1;
            if (ldg_u32(P.first_item + mid) <= item) lo = mid; else hi = mid - 1u;
        }
        img = lo;
        local = item - ldg_u32(P.first_item + lo);
    }
    const ImageDesc im = P.images[img];
    TWarp<LAYOUT>& W = S.wm[t >> 5];
    TAddr A;
    A.xch = smem_addr(W.xch) + 4u * (unsigned)(((t & 31) / (32 / G::GROUPS)) * G::GROUP_FLOATS);
    A.stage = smem_addr(W.stage);
    const int m_begin = (int)local * G::ITEM_MCUS;
    const int m_end = (im.n_mcus - m_begin < G::ITEM_MCUS) ? im.n_mcus : m_begin + G::ITEM_MCUS;
    int16_t* gout = P.coefs + (im.first_block + (unsigned long long)m_begin * G::BPM) * 64ull;
    int my0 = m_begin / im.mcus_x, mx0 = m_begin - my0 * im.mcus_x;
    TPixels<LAYOUT, NC> px;
    fetch_pixels<LAYOUT, NC>(im, my0, mx0, m_end - m_begin < G::ITER_MCUS ? m_end - m_begin : G::ITER_MCUS, px);
#pragma unroll 1
    for (int m0 = m_begin; m0 < m_end; m0 += G::ITER_MCUS) {
        const int nM = m_end - m0 < G::ITER_MCUS ? m_end - m0 : G::ITER_MCUS;
        // position and size of the NEXT iteration, whose pixels are requested in the middle of this one
        int mx1 = mx0 + G::ITER_MCUS, my1 = my0;
        while (mx1 >= im.mcus_x) { mx1 -= im.mcus_x; ++my1; }
        const int left = m_end - m0 - G::ITER_MCUS;
        const int nM1 = left < 0 ? 0 : (left < G::ITER_MCUS ? left : G::ITER_MCUS);
        TPixels<LAYOUT, NC> nx;
        auto prefetch = [&]() { fetch_pixels<LAYOUT, NC>(im, my1, mx1, nM1, nx); };
        if constexpr (LAYOUT == LAYOUT_444) transform_iter_444<NC>(A, px, nM, gout, LC, prefetch);
        else if constexpr (LAYOUT == LAYOUT_420) transform_iter_420<NC>(A, px, nM, gout, LC, prefetch);
        else transform_iter_gray(A, px, nM, gout, LC, prefetch);
        px = nx;
        gout += G::ITER_BLOCKS * 64;
        mx0 = mx1; my0 = my1;
    }
}

}  // namespace jg
