// jpeg_kernel.cuh -- the whole JPEG entropy-coded segment in ONE persistent sm_100a kernel.
//
// Replaces the reference's serial block loop, jpeg_enc.h:1094-1158 (tjei_encode_main) and
// everything it calls: tjei_encode_and_write_MCU (:786-889), tjei_fdct (:656-763),
// tjei_calculate_variable_length_int (:598-610), tjei_write_bits (:613-643) and the tail
// (:1160-1167).  Output bytes are identical to the reference's for its native modes.
//
// Work decomposition
//   tile   = 24 blocks (8x8 data units) = 8 / 4 / 24 consecutive MCUs of ONE image
//            (4:4:4 / 4:2:0 / gray) in stream order.  A tile belongs to ONE WARP from its pixels
//            to its bytes: after the tables are loaded no CTA barrier is executed any more, warps
//            never wait for their siblings.
//   CTA    = 128 threads (4 warps, 6 CTAs per SM), persistent; every warp draws its tiles from an
//            atomic ticket, so a tile's predecessors are always resident or finished (what makes
//            the look-back and the DC hand-over below safe).
//   DC prediction across tiles: the lanes that compute the last DCs of a tile publish them
//            (desc_dc) right after their column pass; the next tile reads them after its own
//            transform.  Nothing is recomputed and nobody waits on anybody's entropy coding.
//
// Per tile (one warp):
//   1. transform  lane groups of 8 (16 for 4:2:0) own one MCU: each lane loads ONE pixel row
//                 straight from global memory (edge replication jpeg_enc.h:1106-1111 applied
//                 on the way), converts it to Y/Cb/Cr once for all components (:1118-1120),
//                 runs the AAN row pass (:668-709), exchanges the 8x8 through a padded
//                 shared-memory tile, runs the column pass (:718-759), quantises (:808-816)
//                 and scatters int16 coefficients to shared memory in zigzag order.
//   2. write the PREVIOUS tile of this warp (see 5): its offset is resolved while our own
//                 predecessors are still coding, one whole transform after its count was published.
//   3. entropy    warp-cooperative, one pass (P5-P8 of SURVEY 8a).  Four blocks at a time the warp
//                 compacts the nonzero coefficients (+ DC + EOB) into a dense symbol queue, then
//                 codes 64 symbols per round (two per lane), every lane busy: category / run /
//                 Huffman lookup, warp scan of the code lengths, atomicOr of the bits into the
//                 warp's region.  No per-block size walk, no divergence between sparse and dense blocks.
//   4. publish    ONE 64-bit descriptor per tile: status | the tile's last 7 bits | its bit count.
//   5. chain + write (one iteration later): decoupled look-back over the descriptors gives the
//                 exclusive BIT offset of the tile in its image and, from the nearest descriptor,
//                 the bits that complete the byte the two tiles share.  The region goes to the
//                 image's UNSTUFFED scan (every byte written once, by the tile that holds its last
//                 bit); the last tile pads with zero bits (jpeg_enc.h:1161-1164).
//                 0xFF00 stuffing + EOI are the second pass (jpeg_stuff.cuh).
// HBM traffic of this kernel: every pixel read once, the unstuffed scan written once, one
// 8-byte and three 4-byte descriptors per tile.
//
// A tile whose bits overflow the region (pathological content) is redone in six groups of four
// blocks; same bytes, lower speed.
#pragma once
#include "jpeg_device.h"
#include "jpeg_launch.h"
#include "jpeg_tables.h"

namespace jg {

constexpr int kTileFloats = 72;   // one 8x8 float tile with rows padded to 9 (bank-conflict free both ways)
constexpr int kCoefStride = 72;   // int16 per block in shared memory (144 B: 16-byte aligned 8-coefficient reads)
constexpr int kQueueEntries = 424;                     // 2 pad + 63 carried + 32 lanes x (8 coefficients + EOB) + 64 read-ahead
constexpr int kGroupBlocks = 4;                        // blocks per group on the slow path (always fit kWinWordsMin)
constexpr int kHalfWords = kWinWordsMax / 2;           // a tile that fits half the region leaves the other half to its successor

template <int LAYOUT>
struct Geo {
    static constexpr int BPM = LAYOUT == LAYOUT_444 ? 3 : (LAYOUT == LAYOUT_420 ? 6 : 1);  // blocks per MCU
    static constexpr int MCU = LAYOUT == LAYOUT_420 ? 16 : 8;                               // MCU edge in pixels
    static constexpr int M = mcus_per_tile(LAYOUT);        // MCUs per tile
    static constexpr int LANES = LAYOUT == LAYOUT_420 ? 16 : 8;   // lanes that share one MCU
    static constexpr int GROUPS = 32 / LANES;                     // lane groups of the warp
    static constexpr int PER_GROUP = LAYOUT == LAYOUT_GRAY ? 2 : 1;   // MCUs a lane group handles per iteration
    static constexpr int ITERS = M / (GROUPS * PER_GROUP);
    static constexpr int TILES = LAYOUT == LAYOUT_420 ? 6 : (LAYOUT == LAYOUT_444 ? 3 : 2);    // exchange tiles per lane group
    static constexpr int NCOMP = LAYOUT == LAYOUT_GRAY ? 1 : 3;
    static_assert(M % (GROUPS * PER_GROUP) == 0, "MCUs must divide evenly among the lane groups");
    static_assert(M * BPM == kBlocksPerTile, "a full tile holds kBlocksPerTile blocks");
};

// what the pending (coded, not yet written) tile of a warp needs one iteration later
struct Pending {
    unsigned long long raw;       // the image's unstuffed scan
    unsigned long long raw_cap;
    int img_idx;
    int first_tile_of_img;
    unsigned T;                   // bits of the tile
    unsigned tail;                // its last 7 bits
    int last;                     // last tile of its image
    unsigned base;                // first word of its bits in the region (0 or kHalfWords)
};

// everything one warp owns
template <int LAYOUT>
struct WarpMem {
    using G = Geo<LAYOUT>;
    static constexpr int SCRATCH = G::GROUPS * (G::TILES * kTileFloats + 8);
    static constexpr int R1_WORDS = SCRATCH > kQueueEntries ? SCRATCH : kQueueEntries;
    alignas(16) uint32_t r1[R1_WORDS];                        // transform exchange tiles; then the symbol queue
    alignas(16) int16_t coef[(kBlocksPerTile + 1) * kCoefStride];   // quantised coefficients, zigzag order (+1: scratch slot)
    alignas(16) uint32_t region[kWinWordsMax + 8];            // the tile's packed bits (MSB-first words); survives into the next iteration
    int pred_dc[4];                                           // DCs of the MCU preceding the tile, per component
    Pending pend[3];                                          // iteration mod 3: the tile being coded and the (up to) two waiting to be written
};

struct CodeTables {
    uint32_t huff[2][272];                // [class][(run<<4)|cat] AC, [class][256+cat] DC; entry = code<<8 | length
    unsigned long long zrl[2][4];         // [class][n]: n ZRL codes back to back (<= 33 bits), bits<<8 | length
};

template <int LAYOUT, int NC>
struct Smem {
    WarpMem<LAYOUT> wm[kWarps];
    CodeTables tab;
};

// per-lane constants: the lane's column u = lane & 7 of every coefficient matrix it finishes
struct LaneConst {
    float pq_l[8], pq_c[8];   // reciprocal quantisers of column u: [v] = pqt[8v+u], luma / chroma
    unsigned zz_lo, zz_hi;    // zigzag positions of (v,u), v = 0..7, one byte each
};

JG_DEV int zz_of(int i)   // jpeg_enc.h:376-386: position in scan order of natural index i
{
    const unsigned char t[64] = JG_ZZ_INIT;
    return t[i];
}

// ------------------------------------------------------------------------------------------
// pixels: one row segment of NPX pixels starting at column x0 of row y, as little-endian words
// ------------------------------------------------------------------------------------------
template <int NC, int NPX>
JG_DEV void load_segment(const ImageDesc& im, int x0, int y, uint32_t (&w)[NPX * NC / 4])
{
    constexpr int BYTES = NPX * NC;
    const uint8_t* row = im.px + (long long)y * (long long)im.stride;
    if (x0 + NPX <= im.w && im.align >= 4) {
        const uint8_t* p = row + (size_t)x0 * NC;
        // widest load the segment size and the image's alignment allow: fewer, fatter requests
        if (BYTES % 16 == 0 && im.align >= 16) {
#pragma unroll
            for (int i = 0; i < BYTES / 16; ++i) {
                const uint4 v = ldg_u128(p + 16 * i);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
        } else if (BYTES % 8 == 0 && im.align >= 8) {
#pragma unroll
            for (int i = 0; i < BYTES / 8; ++i) {
                const uint2 v = ldg_u64(p + 8 * i);
                w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < BYTES / 4; ++i) w[i] = ldg_u32(p + 4 * i);
        }
    } else {
        // JPEG_GPU_FLAG_SWAP_RB: channels 0 and 2 change places.  Only this byte loader knows the flag (the
        // host gives flagged images align = 1), as two per-channel base pointers: the vector path of everybody
        // else stays as it is -- a 20-instruction swizzle block behind a never-taken branch there cost 3.5 %.
        const int swz = (NC >= 3 && (im.flags & 1)) ? 2 : 0;
        const uint8_t* base[4] = {row + swz, row, row - swz, row};
#pragma unroll
        for (int i = 0; i < BYTES / 4; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int bo = 4 * i + b;
                const int px = bo / NC, ch = bo - px * NC;
                int x = x0 + px;
                if (x >= im.w) x = im.w - 1;                     // replicate the last column
                v |= ldg_u8(base[ch] + (size_t)x * NC + ch) << (8 * b);
            }
            w[i] = v;
        }
    }
}

JG_DEV unsigned byte_of(const uint32_t* w, int idx) { return (w[idx >> 2] >> ((idx & 3) * 8)) & 0xffu; }

template <int NC>
JG_DEV void rgb_of(const uint32_t* w, int i, float& r, float& g, float& b)
{
    if (NC == 4) {
        const uint32_t v = w[i];
        r = u8_to_f(v & 0xffu); g = u8_to_f((v >> 8) & 0xffu); b = u8_to_f((v >> 16) & 0xffu);
    } else {
        r = u8_to_f(byte_of(w, 3 * i)); g = u8_to_f(byte_of(w, 3 * i + 1)); b = u8_to_f(byte_of(w, 3 * i + 2));
    }
}

// jpeg_enc.h:1118-1120, C's left-to-right evaluation made explicit
JG_DEV float rgb_y(float r, float g, float b) { return f_sub(f_add(f_add(f_mul(0.299f, r), f_mul(0.587f, g)), f_mul(0.114f, b)), 128.0f); }

// ------------------------------------------------------------------------------------------
// AAN forward DCT, one 8-point pass (jpeg_enc.h:668-709 rows, :718-759 columns)
// ------------------------------------------------------------------------------------------
JG_DEV void aan8(float (&d)[8])
{
    const float c4 = 0.707106781f, c6 = 0.382683433f, c2m6 = 0.541196100f, c2p6 = 1.306562965f;
    const float t0 = f_add(d[0], d[7]), t7 = f_sub(d[0], d[7]);
    const float t1 = f_add(d[1], d[6]), t6 = f_sub(d[1], d[6]);
    const float t2 = f_add(d[2], d[5]), t5 = f_sub(d[2], d[5]);
    const float t3 = f_add(d[3], d[4]), t4 = f_sub(d[3], d[4]);

    const float e0 = f_add(t0, t3), e3 = f_sub(t0, t3);
    const float e1 = f_add(t1, t2), e2 = f_sub(t1, t2);
    d[0] = f_add(e0, e1);
    d[4] = f_sub(e0, e1);
    const float half = f_mul(f_add(e2, e3), c4);
    d[2] = f_add(e3, half);
    d[6] = f_sub(e3, half);

    const float o0 = f_add(t4, t5), o1 = f_add(t5, t6), o2 = f_add(t6, t7);
    const float rot = f_mul(f_sub(o0, o2), c6);
    const float lo = f_add(f_mul(c2m6, o0), rot);
    const float hi = f_add(f_mul(c2p6, o2), rot);
    const float mid = f_mul(o1, c4);
    const float sum7 = f_add(t7, mid), dif7 = f_sub(t7, mid);
    d[5] = f_add(dif7, lo);
    d[3] = f_sub(dif7, lo);
    d[1] = f_add(sum7, hi);
    d[7] = f_sub(sum7, hi);
}

// The same pass over TWO independent 8-point vectors at once (.x and .y): the 29 additions are
// packed (FADD2, one issue slot for two), the 5 multiplications stay scalar (jpeg_device.h says why).
JG_DEV void aan8x2(f32x2 (&d)[8])
{
    const f32x2 c4 = f2(0.707106781f, 0.707106781f), c6 = f2(0.382683433f, 0.382683433f);
    const f32x2 c2m6 = f2(0.541196100f, 0.541196100f), c2p6 = f2(1.306562965f, 1.306562965f);
    const f32x2 t0 = f2_add(d[0], d[7]), t7 = f2_sub(d[0], d[7]);
    const f32x2 t1 = f2_add(d[1], d[6]), t6 = f2_sub(d[1], d[6]);
    const f32x2 t2 = f2_add(d[2], d[5]), t5 = f2_sub(d[2], d[5]);
    const f32x2 t3 = f2_add(d[3], d[4]), t4 = f2_sub(d[3], d[4]);

    const f32x2 e0 = f2_add(t0, t3), e3 = f2_sub(t0, t3);
    const f32x2 e1 = f2_add(t1, t2), e2 = f2_sub(t1, t2);
    d[0] = f2_add(e0, e1);
    d[4] = f2_sub(e0, e1);
    const f32x2 half = f2_mul(f2_add(e2, e3), c4);
    d[2] = f2_add(e3, half);
    d[6] = f2_sub(e3, half);

    const f32x2 o0 = f2_add(t4, t5), o1 = f2_add(t5, t6), o2 = f2_add(t6, t7);
    const f32x2 rot = f2_mul(f2_sub(o0, o2), c6);
    const f32x2 lo = f2_add(f2_mul(c2m6, o0), rot);
    const f32x2 hi = f2_add(f2_mul(c2p6, o2), rot);
    const f32x2 mid = f2_mul(o1, c4);
    const f32x2 sum7 = f2_add(t7, mid), dif7 = f2_sub(t7, mid);
    d[5] = f2_add(dif7, lo);
    d[3] = f2_sub(dif7, lo);
    d[1] = f2_add(sum7, hi);
    d[7] = f2_sub(sum7, hi);
}

// jpeg_enc.h:808-816: v*pqt, floorf((v + 1024) + 0.5f) - 1024, (int)
JG_DEV int quantise(float v, float pq)
{
    v = f_mul(v, pq);
    v = f_add(f_add(v, 1024.0f), 0.5f);
    return f_floor_i(v) - 1024;
}

// row pass of one 8-sample row, result into row `r` of an exchange tile
JG_DEV void row_pass_store(float (&s)[8], float* tile, int r)
{
    aan8(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[r * 9 + i] = s[i];
}
// two rows at once: .x into row r of tile_x, .y into row r of tile_y
JG_DEV void row_pass_store_x2(f32x2 (&s)[8], float* tile_x, float* tile_y, int r)
{
    aan8x2(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) { tile_x[r * 9 + i] = s[i].x; tile_y[r * 9 + i] = s[i].y; }
}

// MCU `slot` of a tile whose first MCU sits at (my0, mx0): one division per tile (by the caller), not one per MCU
JG_DEV void mcu_pos(const ImageDesc& im, int my0, int mx0, int slot, int& my, int& mx)
{
    my = my0; mx = mx0 + slot;
    while (mx >= im.mcus_x) { mx -= im.mcus_x; ++my; }
}

JG_DEV unsigned zz_at(const LaneConst& LC, int v) { return ((v < 4 ? LC.zz_lo : LC.zz_hi) >> (8 * (v & 3))) & 0xffu; }
JG_DEV void publish_dc(unsigned* dc_out, int k) { st_flag32(dc_out, 0x80000000u | ((unsigned)k & 0xffffu)); }

// column u of an exchange tile: column pass, quantise, scatter to zigzag order.
// dc_out != nullptr: the block is the last of its component in the tile; its DC is published for
// the next tile's DC prediction (the lane that computed it stores it, nobody waits for anybody).
template <int LAYOUT>
JG_DEV void column_pass(WarpMem<LAYOUT>& W, const float* tile, int u, bool chroma, int blk, unsigned* dc_out, const LaneConst& LC)
{
    float c[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) c[v] = tile[v * 9 + u];
    aan8(c);
    int16_t* dst = W.coef + blk * kCoefStride;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int k = quantise(c[v], chroma ? LC.pq_c[v] : LC.pq_l[v]);
        dst[zz_at(LC, v)] = (int16_t)k;
        if (v == 0 && u == 0 && dc_out != nullptr) publish_dc(dc_out, k);
    }
}
// the same for column u of TWO tiles that use the same quantiser table (.x -> blk_x, .y -> blk_y)
template <int LAYOUT>
JG_DEV void column_pass_x2(WarpMem<LAYOUT>& W, const float* tile_x, const float* tile_y, int u, bool chroma, int blk_x, int blk_y,
                           unsigned* dc_x, unsigned* dc_y, const LaneConst& LC)
{
    f32x2 c[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) c[v] = f2(tile_x[v * 9 + u], tile_y[v * 9 + u]);
    aan8x2(c);
    int16_t* dst_x = W.coef + blk_x * kCoefStride;
    int16_t* dst_y = W.coef + blk_y * kCoefStride;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float pq = chroma ? LC.pq_c[v] : LC.pq_l[v];
        const f32x2 q = f2_add(f2_add(f2_mul(c[v], f2(pq, pq)), f2(1024.0f, 1024.0f)), f2(0.5f, 0.5f));   // jpeg_enc.h:808-813
        const int kx = f_floor_i(q.x) - 1024, ky = f_floor_i(q.y) - 1024;
        const unsigned zz = zz_at(LC, v);
        dst_x[zz] = (int16_t)kx;
        dst_y[zz] = (int16_t)ky;
        if (v == 0 && u == 0) {
            if (dc_x != nullptr) publish_dc(dc_x, kx);
            if (dc_y != nullptr) publish_dc(dc_y, ky);
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 1, grayscale: 8 lanes per PAIR of consecutive blocks (both go through the packed passes),
// 3 pairs per lane group.  Gray has registers to spare, so the next pair's pixels are requested
// while the current column pass runs.  (The colour paths sit at the register cap: there the same
// prefetch cost more in spills than the hidden latency gained -- measured -- so they load at the
// top of each iteration.)
// ------------------------------------------------------------------------------------------
template <int LAYOUT, int NC>
JG_DEV void transform_tile_gray(WarpMem<LAYOUT>& W, const ImageDesc& im, int my0, int mx0, int nM, unsigned* dc_out, const LaneConst& LC)
{
    using G = Geo<LAYOUT>;
    const int t = JG_TID & 31, u = t & 7, grp = t >> 3;
    // group stride 2 tiles + 8 floats: the four groups start 24 banks apart, all 32 lanes hit different banks
    float* tile_x = reinterpret_cast<float*>(W.r1) + grp * (2 * kTileFloats + 8);
    float* tile_y = tile_x + kTileFloats;
    auto fetch = [&](int it, uint32_t (&w)[4]) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int slot = 2 * (it * G::GROUPS + grp) + half;
            uint32_t ww[2] = {0u, 0u};
            if (slot < nM) {
                int my, mx;
                mcu_pos(im, my0, mx0, slot, my, mx);
                int y = my * 8 + u; if (y >= im.h) y = im.h - 1;          // replicate the last row
                load_segment<1, 8>(im, mx * 8, y, ww);
            }
            w[2 * half] = ww[0]; w[2 * half + 1] = ww[1];
        }
    };
    uint32_t w[4];
    fetch(0, w);
#pragma unroll 1
    for (int it = 0; it < G::ITERS; ++it) {
        const int slot = 2 * (it * G::GROUPS + grp);
        const bool valid = slot < nM, valid_y = slot + 1 < nM;
        if (valid) {
            f32x2 s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = f2_sub(f2(u8_to_f(byte_of(w, i)), u8_to_f(byte_of(w, 8 + i))), f2(128.0f, 128.0f));
            row_pass_store_x2(s, tile_x, tile_y, u);
        }
        if (it + 1 < G::ITERS) fetch(it + 1, w);
        warp_sync();
        if (valid) {
            // an absent second block computes on zeros into the coefficient slot after the tile's last block
            // (inside the array, never read)
            column_pass_x2(W, tile_x, tile_y, u, false, slot, valid_y ? slot + 1 : kBlocksPerTile,
                           slot == nM - 1 ? dc_out : nullptr, slot + 1 == nM - 1 ? dc_out : nullptr, LC);
        }
        warp_sync();
    }
}

// ------------------------------------------------------------------------------------------
// stage 1: transform all MCUs of the tile
// ------------------------------------------------------------------------------------------
template <int LAYOUT, int NC>
JG_DEV void transform_tile(WarpMem<LAYOUT>& W, const ImageDesc& im, int my0, int mx0, int nM, unsigned* dc_out, const LaneConst& LC)
{
    using G = Geo<LAYOUT>;
    const int t = JG_TID & 31;
    float* scratch = reinterpret_cast<float*>(W.r1);
    const int u = t & 7;
    const int grp = t / G::LANES;

#pragma unroll 1
    for (int it = 0; it < G::ITERS; ++it) {
        const int slot = it * G::GROUPS + grp;
        const bool valid = slot < nM;
        unsigned* dcs = (slot == nM - 1) ? dc_out : nullptr;     // the tile's last MCU publishes its DCs
        int my = 0, mx = 0;
        if (valid) mcu_pos(im, my0, mx0, slot, my, mx);

        if (LAYOUT == LAYOUT_444) {
            // 8 lanes per MCU, lane u owns pixel row u.  Y goes through the scalar passes, Cb and Cr
            // together through the packed ones (same structure, same quantiser table).
            float* base = scratch + grp * (3 * kTileFloats);
            if (valid) {
                int y = my * 8 + u; if (y >= im.h) y = im.h - 1;
                uint32_t w[8 * NC / 4];
                load_segment<NC, 8>(im, mx * 8, y, w);
                float sy[8];
                f32x2 sc[8];                 // .x = Cb, .y = Cr
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float r, g, b;
                    rgb_of<NC>(w, i, r, g, b);
                    sy[i] = rgb_y(r, g, b);
                    // jpeg_enc.h:1119-1120 with x - k*y written as x + (-k)*y where the two halves differ
                    const f32x2 p1 = f2(f_mul(-0.1687f, r), f_mul(0.5f, r));
                    const f32x2 p2 = f2(f_mul(0.3313f, g), f_mul(0.4187f, g));
                    const f32x2 p3 = f2(f_mul(0.5f, b), f_mul(-0.0813f, b));
                    sc[i] = f2_add(f2_sub(p1, p2), p3);
                }
                row_pass_store(sy, base, u);
                row_pass_store_x2(sc, base + kTileFloats, base + 2 * kTileFloats, u);
            }
            warp_sync();
            if (valid) {
                column_pass(W, base, u, false, slot * 3, dcs, LC);
                column_pass_x2(W, base + kTileFloats, base + 2 * kTileFloats, u, true, slot * 3 + 1, slot * 3 + 2,
                               dcs ? dcs + 1 : nullptr, dcs ? dcs + 2 : nullptr, LC);
            }
            warp_sync();
        } else {
            // 4:2:0: 16 lanes per MCU, lane r16 owns pixel row r16 (16 pixels).  Pixels i and i + 8 (the
            // left and the right 8x8 luma block of the row) travel as one packed value.
            const int r16 = t & 15;
            float* base = scratch + grp * (6 * kTileFloats);
            f32x2 cbs[4], crs[4];            // horizontal pair sums (a+b) of this row: .x samples 0-3, .y samples 4-7
#pragma unroll
            for (int i = 0; i < 4; ++i) { cbs[i] = f2(0.0f, 0.0f); crs[i] = f2(0.0f, 0.0f); }
            if (valid) {
                int y = my * 16 + r16; if (y >= im.h) y = im.h - 1;
                uint32_t w[16 * NC / 4];
                load_segment<NC, 16>(im, mx * 16, y, w);
                f32x2 sy[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    f32x2 cb[2], cr[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = 2 * j + e;
                        float r0, g0, b0, r1, g1, b1;
                        rgb_of<NC>(w, i, r0, g0, b0);
                        rgb_of<NC>(w, i + 8, r1, g1, b1);
                        const f32x2 R = f2(r0, r1), Gc = f2(g0, g1), B = f2(b0, b1);
                        // jpeg_enc.h:1118-1120, additions packed over the two pixels
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
                float* ty = base + ((r16 >> 3) * 2) * kTileFloats;
                row_pass_store_x2(sy, ty, ty + kTileFloats, r16 & 7);
            }
            // vertical pairs live in neighbouring lanes: the even lane finishes Cb, the odd lane Cr;
            // sample = ((a+b) + (c+d)) * 0.25f with (a+b) from the even row (extended mode);
            // the addition commutes, so both lanes compute (mine + other)
            const bool even = (r16 & 1) == 0;
            float samp[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const f32x2 mine = even ? cbs[i] : crs[i], give = even ? crs[i] : cbs[i];
                const f32x2 other = f2(warp_shfl_xor_f32(give.x, 1), warp_shfl_xor_f32(give.y, 1));
                const f32x2 q = f2_mul(f2_add(mine, other), f2(0.25f, 0.25f));
                samp[i] = q.x; samp[4 + i] = q.y;
            }
            if (valid) row_pass_store(samp, base + (4 + (r16 & 1)) * kTileFloats, r16 >> 1);
            warp_sync();
            if (valid) {
                // lanes 0-7 finish tiles {0,2,4}, lanes 8-15 tiles {1,3,5}; the next tile predicts from
                // Y11 (tile 3), Cb and Cr of our last MCU
                const int h = r16 >> 3;
                column_pass_x2(W, base + h * kTileFloats, base + (h + 2) * kTileFloats, u, false, slot * 6 + h, slot * 6 + h + 2,
                               nullptr, (dcs != nullptr && h == 1) ? dcs : nullptr, LC);
                column_pass(W, base + (4 + h) * kTileFloats, u, true, slot * 6 + 4 + h, dcs != nullptr ? dcs + 1 + h : nullptr, LC);
            }
            warp_sync();
        }
    }
}

// ------------------------------------------------------------------------------------------
// entropy coding of one block (jpeg_enc.h:831-888); cz = 64 coefficients in zigzag order
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

// Component of block `blk` (j = blk mod blocks-per-MCU) and its stream-order predecessor with the same
// component (negative: it lies before the tile).
template <int LAYOUT>
JG_DEV void block_kind(int blk, int j, int& comp, int& pred_blk)
{
    if (LAYOUT == LAYOUT_444) { comp = j; pred_blk = blk - 3; }
    else if (LAYOUT == LAYOUT_420) {
        comp = j < 4 ? 0 : j - 3;
        pred_blk = j >= 4 ? blk - 6 : (j == 0 ? blk - 3 : blk - 1);     // Y00 follows the previous MCU's Y11
    } else { comp = 0; pred_blk = blk - 1; }
}

// One queue entry -> its Huffman code + amplitude bits (sym, slen <= 27 bits), the number of ZRL
// codes that precede it (nz) and its table class.  `ep` is the previous entry of the queue.
JG_DEV void decode_symbol(const CodeTables& T, unsigned e, unsigned ep, unsigned& sym, unsigned& slen, unsigned& nz, unsigned& cls)
{
    const int v = (int)e >> 16;
    cls = (e >> 13) & 1u;
    const unsigned cat = category(v) & 15u;                      // 0 for EOB and for a zero DC difference
    const unsigned run = ((e & 63u) - (ep & 63u) - 1u) & 63u;    // zeros since the previous symbol of the block
    unsigned idx = ((run & 15u) << 4) | cat;
    if (e & 0x4000u) idx = 256u + cat;                           // DC
    if (e & 0x8000u) idx = 0u;                                   // EOB
    nz = (e & 0xC000u) ? 0u : run >> 4;                          // one ZRL per 16 zeros; none for DC / EOB
    const unsigned h = T.huff[cls][idx];
    slen = (h & 0xffu) + cat;
    sym = ((h >> 8) << cat) | amplitude(v, cat);
}

// OR `len` (1..64) bits of `sym` into the MSB-first word array at bit `start` (nothing if they
// would not fit the region: the tile is then redone on the slow path).
JG_DEV void put_bits64(uint32_t* region, unsigned start, unsigned long long sym, unsigned len, unsigned cap_bits)
{
    if (start + len + 64u > cap_bits) return;
    const unsigned long long al = sym << (64u - len);
    const unsigned hi = (unsigned)(al >> 32), lo = (unsigned)al;
    const unsigned sh = start & 31u, wi = start >> 5;
    smem_atomic_or(region + wi, hi >> sh);
    if (sh + len > 32u) {
        smem_atomic_or(region + wi + 1, sh ? (hi << (32u - sh)) | (lo >> sh) : lo);
        if (sh + len > 64u) smem_atomic_or(region + wi + 2, lo << (32u - sh));
    }
}

// The warp codes blocks [first, end) of its tile into `region` (zeroed, MSB-first words).
// Symbol queue entry: value<<16 | EOB<<15 | DC<<14 | chroma<<13 | block-in-warp<<8 | zigzag position.
// queue[-1] must be readable (one pad word).  Returns the bits emitted; sets `overflow` if they
// did not fit region_words (the count stays right).
template <int LAYOUT, bool DBG>
JG_DEV unsigned encode_blocks_warp(WarpMem<LAYOUT>& W, const CodeTables& T, int first, int end, uint32_t* region, unsigned region_words,
                                   uint32_t* queue, uint32_t* dbg_bits, bool& overflow)
{
    const int lane = JG_TID & 31, L = lane & 7, b4 = lane >> 3;
    const unsigned cap_bits = region_words * 32u;
    unsigned carry = 0;
    unsigned left = 0;      // symbols queued but not yet coded (an incomplete round is carried to the next step)
    constexpr int BPM = Geo<LAYOUT>::BPM;
    int j = (first + b4) % BPM;                             // block-in-MCU index of my block, stepped along with b0
#pragma unroll 1
    for (int b0 = first; b0 < end; b0 += 4, j = (j + 4 % BPM >= BPM) ? j + 4 % BPM - BPM : j + 4 % BPM) {
        const int blk = b0 + b4;
        const bool valid = blk < end;
        // ---- compaction: my 8 coefficients (zigzag positions 8L..8L+7 of block blk) -----------------
        uint4 w = {0u, 0u, 0u, 0u};
        int comp = 0, pred_blk = -1;
        if (valid) {
            w = *reinterpret_cast<const uint4*>(W.coef + blk * kCoefStride + 8 * L);
            block_kind<LAYOUT>(blk, j, comp, pred_blk);
        }
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
        unsigned m8 = 0;                                    // bit i: coefficient i of my eight is coded
#pragma unroll
        for (int q = 0; q < 4; ++q) {                       // two compares per word; the bits assemble through predicates
            if ((ww[q] & 0xffffu) != 0u) m8 |= 1u << (2 * q);
            if (ww[q] > 0xffffu) m8 |= 2u << (2 * q);
        }
        int diff = 0;
        if (valid && L == 0) {                              // DC: always coded, as a difference (jpeg_enc.h:834-844)
            const int pred = pred_blk >= 0 ? (int)W.coef[pred_blk * kCoefStride] : W.pred_dc[comp];
            diff = (int)(int16_t)(ww[0] & 0xffffu) - pred;
            m8 |= 1u;
        }
        const unsigned eob = (valid && L == 7 && (ww[3] >> 16) == 0u) ? 1u : 0u;   // jpeg_enc.h:884-887
        const unsigned cnt = (unsigned)i_popc(m8) + eob;
        const unsigned inc = warp_scan_incl_u32(cnt);       // (a bit-sliced ballot prefix -- 4 independent votes -- was 2.5 % slower)
        const unsigned N = warp_shfl_u32(inc, 31);
        uint32_t* qp = queue + left + (inc - cnt);
        const unsigned common = (comp ? 1u << 13 : 0u) | ((unsigned)(blk - first) << 8) | (unsigned)(8 * L);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (m8 & (1u << i)) {
                unsigned hv = (i & 1) ? (ww[i >> 1] & 0xffff0000u) : (ww[i >> 1] << 16);     // value in the top half
                unsigned e = common + (unsigned)i;
                if (i == 0 && L == 0) { hv = (unsigned)diff << 16; e |= 1u << 14; }
                *qp++ = hv | e;
            }
        }
        if (eob) *qp = (1u << 15) | common | 7u;
        warp_sync();

        // ---- 64 symbols per round, two per lane, branch-free; only full rounds except at the very end ----
        const unsigned M = left + N;
        const unsigned full = (b0 + 4 >= end) ? M : (M & ~63u);
#pragma unroll 1
        for (unsigned j0 = 0; j0 < full; j0 += 64) {
            const unsigned a = j0 + 2u * (unsigned)lane;                   // this lane codes symbols a and a+1
            const uint2 ee = *reinterpret_cast<const uint2*>(queue + a);
            const unsigned ep = queue[(int)a - 1];
            unsigned symA, lenA, nzA, clsA, symB, lenB, nzB, clsB;
            decode_symbol(T, ee.x, ep, symA, lenA, nzA, clsA);
            decode_symbol(T, ee.y, ee.x, symB, lenB, nzB, clsB);
            if (a >= full) { lenA = 0u; symA = 0u; nzA = 0u; }
            if (a + 1u >= full) { lenB = 0u; symB = 0u; nzB = 0u; }
            if (DBG) {
                if (lenA) gmem_atomic_add(dbg_bits + ((ee.x >> 8) & 31u), lenA + nzA * (T.huff[clsA][0xF0] & 0xffu));
                if (lenB) gmem_atomic_add(dbg_bits + ((ee.y >> 8) & 31u), lenB + nzB * (T.huff[clsB][0xF0] & 0xffu));
            }
            unsigned tot;
            if (warp_ballot((nzA | nzB) != 0u) == 0u) {
                // common case: both codes (<= 27 bits each) travel as one string
                const unsigned len = lenA + lenB;
                const unsigned endb = warp_scan_incl_u32(len);
                tot = warp_shfl_u32(endb, 31);
                if (len) put_bits64(region, carry + endb - len, ((unsigned long long)symA << lenB) | symB, len, cap_bits);
            } else {
                // some lane skipped 16+ zeros: its ZRL codes (jpeg_enc.h:863-867) go in front of the symbol
                const unsigned long long zA = T.zrl[clsA][nzA], zB = T.zrl[clsB][nzB];
                const unsigned tA = lenA ? lenA + (unsigned)(zA & 0xffull) : 0u, tB = lenB ? lenB + (unsigned)(zB & 0xffull) : 0u;
                const unsigned endb = warp_scan_incl_u32(tA + tB);
                tot = warp_shfl_u32(endb, 31);
                const unsigned start = carry + endb - (tA + tB);
                if (tA) put_bits64(region, start, ((zA >> 8) << lenA) | symA, tA, cap_bits);
                if (tB) put_bits64(region, start + tA, ((zB >> 8) << lenB) | symB, tB, cap_bits);
            }
            carry += tot;
        }
        // move the incomplete round to the front; queue[-1] keeps the symbol before it (for its run)
        left = M - full;
        if (left && full) {
            const unsigned keep0 = queue[full + (unsigned)lane], keep1 = queue[full + 32u + (unsigned)lane];
            const unsigned before = queue[full - 1u];
            warp_sync();
            if ((unsigned)lane < left) queue[lane] = keep0;
            if ((unsigned)lane + 32u < left) queue[lane + 32] = keep1;
            if (lane == 0) queue[-1] = before;
        }
        warp_sync();   // the queue is appended to by the next four blocks
    }
    overflow = carry + 64u > cap_bits;
    return carry;
}


// ------------------------------------------------------------------------------------------
// scans and look-back
// ------------------------------------------------------------------------------------------
// exclusive scan over the CTA's threads; contains two barriers (second pass only)
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
// status[63:62] | payload[61:55] | value[54:0].  Tiles before `first` do not exist (image start):
// they count as PREFIX 0.  Returns the exclusive prefix of tile g; *nearest (if asked for) is
// the descriptor of tile g-1 as read.  On timeout sets *timed_out (all lanes agree).
JG_DEV unsigned long long lookback(const unsigned long long* desc, int g, int first, unsigned* err_flag, int* timed_out,
                                   unsigned long long* nearest = nullptr)
{
    const int lane = JG_TID & 31;
    unsigned long long running = 0;
    *timed_out = 0;
    for (int base = g - 1;; base -= 32) {
        const int idx = base - lane;
        const bool in_range = idx >= first;
        unsigned long long w = kStatusPrefix;  // virtual tile before the image: PREFIX 0
        unsigned spins = 0, pmask;
        for (;;) {
            if (in_range) w = ld_flag64(desc + idx);
            // what is needed: every descriptor from g-1 back to the NEAREST one that knows its prefix; older ones may lag
            pmask = warp_ballot((w >> 62) == 2u);
            const unsigned waiting = warp_ballot(in_range && (w >> 62) == 0);
            if ((waiting & (pmask ? (pmask & (0u - pmask)) - 1u : 0xffffffffu)) == 0u) break;
            // the error flag is ONE address for every spinning warp of the launch: looked at every 16th round only
            // (a hot line in L2 makes every round of every spinning warp slower, which makes more warps spin)
            ++spins;
            if (spins > kSpinLimit || ((spins & 15u) == 0u && ld_flag32(err_flag) != 0u)) spins = 0xffffffffu;
            if (warp_ballot(spins == 0xffffffffu) != 0u) { *timed_out = 1; return 0; }
            backoff();
        }
        if (nearest != nullptr && base == g - 1) *nearest = warp_shfl_u64(w, 0);
        const int stop = pmask ? i_ffs(pmask) - 1 : 32;   // nearest tile that already knows its prefix
        running += warp_sum_u64(lane <= stop ? (w & kCountMask) : 0ull);
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

// The warp copies bytes [0, n_bytes) of the byte-aligned stream X = (k head bits) ++ L to dst
// (any alignment).  Every byte is written exactly once and nothing outside the range is touched:
// the bytes up to the second 16-byte boundary and after the last one go out singly, the rest
// as aligned 16-byte vectors (X words are MSB-first, memory wants them byte-swapped).
JG_DEV void copy_stream_out(const uint32_t* L, unsigned k, unsigned hb, unsigned n_bytes, uint8_t* dst)
{
    const unsigned t = (unsigned)(JG_TID & 31);
    const unsigned d = (unsigned)((size_t)dst & 15u);
    uint8_t* A = dst - d;                         // 16-byte aligned; V = d pad bytes ++ X starts here
    const unsigned total = d + n_bytes;
    auto xbyte = [&](unsigned j) { return (xword(L, (int)(j >> 2), k, hb) >> (24u - 8u * (j & 3u))) & 0xffu; };
    // head: V bytes [d, min(total, 32))
    if (t >= d && t < total) A[t] = (uint8_t)xbyte(t - d);
    // body: whole vectors v >= 2 (V byte 16v is X byte 16v - d, bit 8(16v-d) - k of L: never negative here)
    const unsigned nvec = total >> 4;
    for (unsigned v = 2u + t; v < nvec; v += 32u) {
        const unsigned p = 8u * (16u * v - d) - k;      // bit position in L of the vector's first bit
        const unsigned i = p >> 5, sft = p & 31u;
        uint32_t w[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) w[q] = L[i + q];
        uint4 o;
        o.x = bswap32(sft ? (w[0] << sft) | (w[1] >> (32u - sft)) : w[0]);
        o.y = bswap32(sft ? (w[1] << sft) | (w[2] >> (32u - sft)) : w[1]);
        o.z = bswap32(sft ? (w[2] << sft) | (w[3] >> (32u - sft)) : w[2]);
        o.w = bswap32(sft ? (w[3] << sft) | (w[4] >> (32u - sft)) : w[3]);
        *reinterpret_cast<uint4*>(A + 16u * v) = o;
    }
    // tail: V bytes [max(32, 16*nvec), total)
    const unsigned tail0 = nvec >= 2u ? nvec << 4 : 32u;
    if (tail0 + t < total) A[tail0 + t] = (uint8_t)xbyte(tail0 + t - d);
}

// Write the region (tg bits) as the next piece of the image's unstuffed scan.
// k/hb: bits of the first byte that precede the piece (carry in); updated to the carry out.
JG_DEV void flush_region(const uint32_t* region, unsigned tg, bool final_piece, uint8_t* raw, unsigned long long raw_cap,
                         unsigned& k, unsigned& hb, unsigned long long& pos, bool& overflow)
{
    unsigned n_bytes = (k + tg) >> 3, k_out = (k + tg) & 7u, hb_out = 0;
    if (k_out) {
        if (final_piece) { n_bytes += 1; k_out = 0; }                 // zero padding, jpeg_enc.h:1161-1164
        else if (tg >= k_out) hb_out = peek_bits(region, tg - k_out, k_out);
        else hb_out = ((hb << tg) | peek_bits(region, 0u, tg)) & ((1u << k_out) - 1u);
    }
    if (pos + n_bytes > raw_cap) overflow = true;
    else copy_stream_out(region, k, hb, n_bytes, raw + pos);
    pos += n_bytes;
    k = k_out; hb = hb_out;
}

// zero words [0, n) of the region (the warp; followed by a warp barrier)
JG_DEV void clear_region(uint32_t* region, unsigned n)
{
    for (unsigned i = (unsigned)(JG_TID & 31); i < n; i += 32u) region[i] = 0u;
    warp_sync();
}

// n (< 8) 1-bits at bit `pos` of the MSB-first word array (they may straddle two words); one lane
JG_DEV void set_ones(uint32_t* region, unsigned pos, unsigned n)
{
    const unsigned long long m = ((1ull << n) - 1ull) << (64u - (pos & 31u) - n);
    region[pos >> 5] |= (unsigned)(m >> 32);
    region[(pos >> 5) + 1] |= (unsigned)m;
}
JG_DEV unsigned tail_bits(const uint32_t* region, unsigned T) { return T >= 7u ? peek_bits(region, T - 7u, 7u) : peek_bits(region, 0u, T); }

JG_DEV unsigned long long make_desc(unsigned long long status, unsigned tail, unsigned long long count)
{
    return status | ((unsigned long long)tail << kTailShift) | count;
}

// Draw the warp's next tile: its launch-wide index g (>= n_tiles: none left) and its image.
// Tickets are handed out so that a tile's predecessor always holds a smaller ticket (it is then
// resident or finished), and they walk the images ROUND-ROBIN: tile 0 of every image, tile 1 of
// every image that has one, ...  The thousands of tiles in flight then spread over all images, each
// image has only a handful in flight, and the look-back along an image stays within one step.
// (In image-major order a 256 x 1080p batch is 1.7x slower: every stall of one tile sends the
// hundreds of tiles drawn after it on long walks.)  Images of different sizes: the schedule of
// build_schedule() tells which images are still active in which round.
JG_DEV unsigned tile_of_ticket(const LaunchParams& P, unsigned v, unsigned& img)
{
    if (P.tiles_per_image > 0) {
        const unsigned lt = v / (unsigned)P.n_images;
        img = v - lt * (unsigned)P.n_images;
        return img * (unsigned)P.tiles_per_image + lt;
    }
    const uint32_t* S = P.sched;
    const unsigned D = ldg_u32(S);
    const uint32_t *cum = S + 1, *lt0 = cum + D + 1, *first = lt0 + D, *order = first + D;
    unsigned lo = 0, hi = D - 1;           // last segment whose first ticket is <= v
    while (lo < hi) {
        const unsigned mid = (lo + hi + 1) >> 1;
        if (ldg_u32(cum + mid) <= v) lo = mid; else hi = mid - 1;
    }
    const unsigned rem = v - ldg_u32(cum + lo), a = ldg_u32(first + lo), c = (unsigned)P.n_images - a;
    const unsigned r = rem / c;
    img = ldg_u32(order + a + (rem - r * c));
    return (unsigned)P.images[img].first_tile + ldg_u32(lt0 + lo) + r;
}

JG_DEV int draw_tile(const LaunchParams& P, int& img_out)
{
    unsigned g = 0, img = 0;
    if ((JG_TID & 31) == 0) {
        g = gmem_atomic_add(P.ticket, 1u);
        if (g < (unsigned)P.n_tiles) g = tile_of_ticket(P, g, img);
    }
    img_out = (int)warp_shfl_u32(img, 0);
    return (int)warp_shfl_u32(g, 0);
}

// Learn a tile's bit offset and the bits that complete the byte it shares with its predecessor:
// one look-back over the descriptors (the nearest one carries the predecessor's last 7 bits).
// Lane 0 publishes the inclusive prefix.  Returns false on a timeout (error flag raised).
JG_DEV bool chain_bits(const LaunchParams& P, int g, int first_tile_of_img, unsigned T, unsigned tail,
                       unsigned long long& bit_base, unsigned& pred_tail)
{
    int timed_out = 0;
    unsigned long long nearest = 0;
    const unsigned long long excl = lookback(P.desc_bits, g, first_tile_of_img, P.error, &timed_out, &nearest);
    if (timed_out) {
        if ((JG_TID & 31) == 0) gmem_atomic_or(P.error, 1u);
        return false;
    }
    if ((JG_TID & 31) == 0) st_flag64(P.desc_bits + g, make_desc(kStatusPrefix, tail, excl + T));
    bit_base = excl;
    pred_tail = (unsigned)(nearest >> kTailShift) & 0x7fu;
    return true;
}

// ---- back half of a tile, run one iteration after it was coded: offset + write -----------------
template <int LAYOUT>
JG_DEV bool tile_back(const LaunchParams& P, WarpMem<LAYOUT>& W, int g, int slot)
{
    const Pending pd = W.pend[slot];
    const bool first = g == pd.first_tile_of_img;
    unsigned long long bit_base = 0;
    unsigned pred_tail = 0;
    if (!first && !chain_bits(P, g, pd.first_tile_of_img, pd.T, pd.tail, bit_base, pred_tail)) return false;
    unsigned k = (unsigned)(bit_base & 7ull);          // bits of our first byte owned by the predecessor
    unsigned hb = pred_tail & ((1u << k) - 1u);
    unsigned long long pos = bit_base >> 3;
    bool overflow = false;
    flush_region(W.region + pd.base, pd.T, pd.last != 0, reinterpret_cast<uint8_t*>(pd.raw), pd.raw_cap, k, hb, pos, overflow);
    if ((JG_TID & 31) == 0) {
        if (pd.last) P.raw_bytes[pd.img_idx] = pos;
        if (overflow) gmem_atomic_or(P.img_status + pd.img_idx, 1u);
    }
    warp_sync();                                        // every lane has read the region
    clear_region(W.region + pd.base, (pd.T >> 5) + 2u);
    return true;
}

// Stage-dump variant of the coder (parity tests only), kept out of line so the hot loop stays small.
template <int LAYOUT>
JG_DEV_NOINLINE unsigned encode_blocks_dbg(WarpMem<LAYOUT>& W, const CodeTables& T, int nblk, unsigned cap_words, uint32_t* queue,
                                           uint32_t* dbg_bits, bool& overflow)
{
    return encode_blocks_warp<LAYOUT, true>(W, T, 0, nblk, W.region, cap_words, queue, dbg_bits, overflow);
}

// Pathological tile (its `bits` did not fit the region): groups of four blocks (always fit), coded
// again and written out right away, not pipelined.  The last group goes first, only to learn the
// tile's last 7 bits: successors must not wait for our whole slow pass.  Out of line: rare.
// Returns false on a look-back timeout.
template <int LAYOUT, bool restart>
JG_DEV_NOINLINE bool tile_slow(const LaunchParams& P, WarpMem<LAYOUT>& W, const CodeTables& T, uint32_t* queue, unsigned cap_words,
                               int g, int nblk, unsigned bits, int slot)
{
    const int lane = JG_TID & 31;
    const Pending pd = W.pend[slot];
    const bool first = g == pd.first_tile_of_img;
    const int n_groups = (nblk + kGroupBlocks - 1) / kGroupBlocks;
    bool ovf2;
    clear_region(W.region, cap_words + 8u);
    unsigned tl = encode_blocks_warp<LAYOUT, false>(W, T, (n_groups - 1) * kGroupBlocks, nblk, W.region, cap_words, queue, nullptr, ovf2);
    // restart interval: the tile ends with 1-bits up to the byte boundary (the groups are contiguous in
    // the bit stream, so the tile's total decides the padding of its last group)
    const unsigned pad = restart ? (0u - bits) & 7u : 0u;
    if (pad) {
        if (lane == 0) set_ones(W.region, tl, pad);
        warp_sync();
        tl += pad; bits += pad;
    }
    const unsigned tail = tail_bits(W.region, tl);
    if (lane == 0) st_flag64(P.desc_bits + g, make_desc(first ? kStatusPrefix : kStatusAgg, tail, bits));
    unsigned long long bit_base = 0;
    unsigned pred_tail = 0;
    if (!first && !chain_bits(P, g, pd.first_tile_of_img, bits, tail, bit_base, pred_tail)) return false;
    unsigned k = (unsigned)(bit_base & 7ull);
    unsigned hb = pred_tail & ((1u << k) - 1u);
    unsigned long long pos = bit_base >> 3;
    bool cap_overflow = false;
    for (int gi = 0; gi < n_groups; ++gi) {
        const int b_lo = gi * kGroupBlocks, b_hi = b_lo + kGroupBlocks < nblk ? b_lo + kGroupBlocks : nblk;
        warp_sync();
        clear_region(W.region, cap_words + 8u);
        unsigned tg = encode_blocks_warp<LAYOUT, false>(W, T, b_lo, b_hi, W.region, cap_words, queue, nullptr, ovf2);
        if (pad && gi == n_groups - 1) {
            if (lane == 0) set_ones(W.region, tg, pad);      // (the group does not start on a byte boundary: the bits may straddle words)
            warp_sync();
            tg += pad;
        }
        flush_region(W.region, tg, pd.last && gi == n_groups - 1, reinterpret_cast<uint8_t*>(pd.raw), pd.raw_cap, k, hb, pos, cap_overflow);
    }
    if (lane == 0) {
        if (pd.last) P.raw_bytes[pd.img_idx] = pos;
        if (cap_overflow) gmem_atomic_or(P.img_status + pd.img_idx, 1u);
    }
    warp_sync();
    clear_region(W.region, cap_words + 8u);
    return true;
}

// A tile did not fit its half of the region: write out the tile that occupies the other half (if
// any), then code the tile again into the whole region.  Out of line: happens when the content
// turns dense, after which the warp stops using halves for a while.  Returns false on a timeout.
template <int LAYOUT>
JG_DEV_NOINLINE bool tile_recode(const LaunchParams& P, WarpMem<LAYOUT>& W, const CodeTables& T, uint32_t* queue, unsigned cap_words,
                                 int other_g, int other_slot, int nblk, unsigned& bits, bool& overflow)
{
    if (other_g >= 0 && !tile_back<LAYOUT>(P, W, other_g, other_slot)) return false;
    warp_sync();
    clear_region(W.region, cap_words + 8u);
    bits = encode_blocks_warp<LAYOUT, false>(W, T, 0, nblk, W.region, cap_words, queue, nullptr, overflow);
    return true;
}

// ------------------------------------------------------------------------------------------
// kernel 1: pixels -> unstuffed entropy-coded bits
// ------------------------------------------------------------------------------------------
// DEEP: tiles wait up to two iterations for their write-out (see the loop).  Chosen by the host for
// launches with few images: there the thousands of tiles in flight belong to the same image, every
// look-back depends on tiles drawn nanoseconds earlier, and one iteration of slack is not enough
// (16384^2 gray: 23 % of all warp time was spent waiting in the look-back; DEEP: 1.17 -> 1.00 ms).
// With many images in flight the round-robin ticket order already provides the slack and the
// simpler loop is ~3 % faster.
// MODE 2 = restart intervals (JPEG_GPU_FLAG_RESTART; one-iteration pipeline): every tile predicts its DCs
// from 0 and ends on a byte boundary, padded with 1-bits (T.81 F.1.2.3); the markers between the tiles are
// inserted by the stuffing pass.  Its own instantiation because even a few never-taken branches in the hot
// loop cost 2.5 % (measured): the host groups restart images into their own launches.
constexpr int kModePlain = 0, kModeDeep = 1, kModeRestart = 2;
template <int LAYOUT, int NC, int MODE>
JG_KERNEL(kThreads, 6)
void encode_tiles_kernel(const JG_GRID_CONSTANT LaunchParams P, const JG_GRID_CONSTANT QuantSet Q)
{
    constexpr bool DEEP = MODE == kModeDeep;
    constexpr bool restart = MODE == kModeRestart;
    using G = Geo<LAYOUT>;
    JG_DYNAMIC_SMEM(smem_raw);
    Smem<LAYOUT, NC>& S = *reinterpret_cast<Smem<LAYOUT, NC>*>(smem_raw);
    const int t = JG_TID, lane = t & 31;
    for (int i = t; i < 2 * 272; i += kThreads) {
        const int cls = i / 272, k = i - cls * 272;
        S.tab.huff[cls][k] = k < 256 ? P.huff->ac[cls][k] : P.huff->dc[cls][k - 256];
    }
    if (t < 8) {      // n = t & 3 ZRL codes back to back (the luma ZRL has 11 bits: up to 33 bits)
        const int cls = t >> 2, n = t & 3;
        const unsigned z = P.huff->ac[cls][0xF0];
        unsigned long long bits = 0, len = 0;
        for (int q = 0; q < n; ++q) { bits = (bits << (z & 0xffu)) | (z >> 8); len += z & 0xffu; }
        S.tab.zrl[cls][n] = (bits << 8) | len;
    }
    LaneConst LC;
    {
        const int u = t & 7;
        LC.zz_lo = 0; LC.zz_hi = 0;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            LC.pq_l[v] = Q.luma[8 * v + u];
            LC.pq_c[v] = Q.chroma[8 * v + u];
            const unsigned z = (unsigned)zz_of(8 * v + u);
            if (v < 4) LC.zz_lo |= z << (8 * v); else LC.zz_hi |= z << (8 * (v - 4));
        }
    }
    WarpMem<LAYOUT>& W = S.wm[t >> 5];
    const CodeTables& T = S.tab;
    uint32_t* const queue = W.r1 + 2;                   // 8-byte aligned; [-1] is a pad word
    const unsigned cap_words = (unsigned)P.win_words;
    clear_region(W.region, kWinWordsMax + 8);
    cta_sync();       // the tables are loaded: the ONLY CTA barrier of the kernel; from here every warp is on its own

    // Software pipeline: tile g is transformed and coded now, its bits wait in the region, and it is
    // chained + written one iteration later (after the next tile's transform) -- by then its count
    // and those of its predecessors were published long ago, so the look-back does not wait.
    // DEEP adds a second iteration of slack, which needs room for two coded tiles: a tile that fits
    // half the region (192 words = 256 bits per block: every BASELINE quality below ~90) is coded
    // into the half its predecessor does not occupy.  When a tile does not fit a half the warp codes
    // whole-region tiles for a while (`dense`), each written one iteration later.
    int p1_g = -1, p2_g = -1;        // tiles coded one / two iterations ago that are still to be written
    bool p1_full = false;            // p1 occupies the whole region
    unsigned p1_base = 0;
    unsigned dense = (DEEP && cap_words == (unsigned)kWinWordsMax && P.dbg_bits == nullptr) ? 0u : 0xffffffffu;
    for (int slot = 0;; slot = slot == 2 ? 0 : slot + 1) {
        // Tiles must START in ticket order (a tile started long before its predecessor would sit in
        // the DC / look-back waits -- measured: drawing the ticket one iteration ahead is 30 % slower):
        // the ticket is drawn when the warp is ready for it, not earlier.
        int img_idx;
        const int g = draw_tile(P, img_idx);
        const bool have = g < P.n_tiles;
        // ---- front: pixels -> coefficients ----
        int nblk = 0;
        bool first = false;
        if (have) {
            const ImageDesc im = P.images[img_idx];
            const int lt = g - im.first_tile;
            const int m0 = lt * G::M;
            const int nM = (im.n_mcus - m0 < G::M) ? im.n_mcus - m0 : G::M;
            nblk = nM * G::BPM;
            first = lt == 0;
            const bool last = lt == im.n_tiles - 1;
            if (lane == 0) {         // what the write-out needs later; T, tail and base follow after the coding
                Pending& pd = W.pend[slot];
                pd.raw = reinterpret_cast<unsigned long long>(im.raw); pd.raw_cap = im.raw_cap;
                pd.img_idx = img_idx; pd.first_tile_of_img = im.first_tile; pd.last = last ? 1 : 0;
            }
            const int my0 = m0 / im.mcus_x, mx0 = m0 - my0 * im.mcus_x;
            if (LAYOUT == LAYOUT_GRAY) transform_tile_gray<LAYOUT, NC>(W, im, my0, mx0, nM, P.desc_dc + 3 * (size_t)g, LC);
            else transform_tile<LAYOUT, NC>(W, im, my0, mx0, nM, P.desc_dc + 3 * (size_t)g, LC);
        }
        // DC predictors of the tile's first blocks: the last DCs of the previous tile (published by the
        // lanes that computed them, right after their column pass -- that tile was drawn before ours), or 0
        // at the start of the image (jpeg_enc.h:1085-1087).  Requested here, looked at after the write-outs:
        // the round trip to L2 costs nothing.
        unsigned dcv = 0x80000000u;
        const bool dc_wanted = have && !first && !restart && lane < G::NCOMP;   // a restart interval predicts from 0
        if (dc_wanted) dcv = ld_flag32(P.desc_dc + 3 * (size_t)(g - 1) + lane);
        // ---- write-outs that are due: the tile coded two iterations ago; last iteration's too if the
        //      coming tile needs the whole region (or nothing comes any more) ----
        const int slot1 = slot == 0 ? 2 : slot - 1, slot2 = slot == 2 ? 0 : slot + 1;   // slots of p1 / p2
        if (DEEP) {
            const bool p1_due = p1_g >= 0 && (p1_full || dense != 0u || !have);
            bool ok = true;
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {                  // oldest first
                const int fg = q == 0 ? p2_g : (p1_due ? p1_g : -1);
                if (fg >= 0 && !tile_back<LAYOUT>(P, W, fg, q == 0 ? slot2 : slot1)) { ok = false; break; }
            }
            if (!ok) break;
            p2_g = -1;
            if (p1_due) p1_g = -1;
        } else if (p1_g >= 0) {
            if (!tile_back<LAYOUT>(P, W, p1_g, slot1)) break;
            p1_g = -1;
        }
        if (!have) break;
        bool bad = false;
        if (dc_wanted) {
            unsigned spins = 0;
            while ((dcv >> 31) == 0u) {
                if (++spins > kSpinLimit || ld_flag32(P.error) != 0u) { bad = true; break; }
                backoff();
                dcv = ld_flag32(P.desc_dc + 3 * (size_t)(g - 1) + lane);
            }
        }
        if (lane < G::NCOMP) W.pred_dc[lane] = (int)(int16_t)(dcv & 0xffffu);
        if (warp_ballot(bad) != 0u) {
            if (lane == 0) gmem_atomic_or(P.error, 4u);
            break;
        }
        warp_sync();   // coefficients + predictors complete; the exchange tiles are dead, their space becomes the queue

        unsigned long long dbg_base = 0;
        if (P.dbg_coefs || P.dbg_bits) {                 // stage dumps for the parity tests
            const int di = W.pend[slot].img_idx;
            dbg_base = P.images[di].first_block + (unsigned long long)((g - P.images[di].first_tile) * kBlocksPerTile);
        }
        if (P.dbg_coefs) {
            for (int i = lane; i < nblk * 64; i += 32)
                P.dbg_coefs[dbg_base * 64ull + (unsigned long long)i] = W.coef[(i >> 6) * kCoefStride + (i & 63)];
        }
        // ---- entropy: coefficients -> bits in the region (the half p1 does not occupy, or all of it) ----
        const bool half = DEEP && dense == 0u;
        const unsigned base = (half && p1_g >= 0 && p1_base == 0u) ? (unsigned)kHalfWords : 0u;
        bool overflow = false;
        unsigned bits;
        if (P.dbg_bits) bits = encode_blocks_dbg<LAYOUT>(W, T, nblk, cap_words, queue, P.dbg_bits + dbg_base, overflow);
        else bits = encode_blocks_warp<LAYOUT, false>(W, T, 0, nblk, W.region + base, half ? (unsigned)kHalfWords : cap_words, queue, nullptr, overflow);
        bool full = !half;
        if (DEEP && half && overflow) {
            // the content turned dense: p1 goes out now, the tile is coded again into the whole region
            if (!tile_recode<LAYOUT>(P, W, T, queue, cap_words, p1_g, slot1, nblk, bits, overflow)) break;
            p1_g = -1;
            full = true;
            dense = 16u;
        } else if (DEEP && !half && dense != 0xffffffffu) {
            if (bits <= (unsigned)kHalfWords * 32u - 512u) --dense;      // 16 tiles in a row that would have fit a half: use halves again
            else dense = 16u;
        }

        if (!overflow) {
            // Publish the tile's bit count and last bits NOW: they are consumed (by us and by every
            // successor) at least one iteration later.
            const unsigned cbase = full ? 0u : base;
            unsigned tail = tail_bits(W.region + cbase, bits);
            if (restart && (bits & 7u) != 0u) {      // restart interval: 1-bits up to the byte boundary (T.81 F.1.2.3)
                const unsigned pad = 8u - (bits & 7u);
                if (lane == 0) set_ones(W.region + cbase, bits, pad);   // read again only by the write-out, barriers later
                tail = ((tail << pad) | ((1u << pad) - 1u)) & 0x7fu;
                bits += pad;
            }
            if (lane == 0) {
                st_flag64(P.desc_bits + g, make_desc(first ? kStatusPrefix : kStatusAgg, tail, bits));
                W.pend[slot].T = bits; W.pend[slot].tail = tail; W.pend[slot].base = cbase;
            }
            warp_sync();
            if (DEEP) p2_g = p1_g;   // (still waiting only if it sits in the other half)
            p1_g = g; p1_full = full; p1_base = cbase;
        } else {
            // does not even fit the whole region (p1 was written out above: the region is ours)
            if (!tile_slow<LAYOUT, restart>(P, W, T, queue, cap_words, g, nblk, bits, slot)) break;
            p2_g = -1; p1_g = -1;
        }
    }
}

}  // namespace jg
