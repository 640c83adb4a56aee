// jpeg_entropy.cuh -- pass B of the split pipeline: quantised coefficients -> the unstuffed entropy-coded bits.
//
// Replaces the serial half of the reference's block loop: DC / AC coding of tjei_encode_and_write_MCU
// (jpeg_enc.h:831-888), tjei_calculate_variable_length_int (:598-610), the bit cursor of tjei_write_bits
// (:613-643, without its 0xFF00 stuffing -- that is pass 2, jpeg_stuff.cuh) and the tail (:1160-1164).
// Input: `du[64]` of every block as pass A (jpeg_transform.cuh) left it in HBM, zigzag order, blocks in scan
// order.  Output: identical to pass 1 of the fused kernel (jpeg_kernel.cuh), whose chain / write-out code it shares.
//
// Work decomposition
//   tile   = 32 consecutive blocks of ONE image in scan order (24 with restart intervals, so that a tile is
//            whole MCUs), whatever the MCU shape: the DC predictor of a block is the DC of an earlier block
//            of the same component, which is simply read (from the tile, or from HBM when it lies before it),
//            so nothing is handed over between tiles except the bit offset.
//   warp   owns a tile:
//            1. stage   ONE 2-D TMA copy (cp.async.bulk.tensor, mbarrier completion) brings the tile's 32 x 128
//                       bytes into shared memory while the warp writes out its previous tile.  The box is 72
//                       elements wide over a tensor that is 64 wide: the out-of-bounds columns pad every block
//                       to a 144-byte row (with 128-byte rows the 32 lanes of the map phase would all sit on the
//                       same four banks: measured, 256 wavefronts per tile for those eight loads alone).
//            2. map     lane l reads block l once (8 x LDS.128) and builds the 64-bit map of what the block
//                       codes: bit 0 the DC difference, bits 1..62 the non-zero coefficients, bit 63 either the
//                       last coefficient or -- where that is zero -- the end-of-block code.  A symbol is a set bit.
//            3. share   a warp scan of the 32 symbol counts cuts the tile's symbols into 32 EQUAL pieces: lane l codes
//                       symbols [l*q, (l+1)*q) -- whatever blocks they belong to.  It finds its first block with five
//                       shuffles over the scanned counts and drops the symbols of that block that belong to the lanes
//                       before it with a popcount search in the block's map.  (Until late in round 2 a list phase
//                       wrote every symbol as a 16-bit entry first -- a lane per block, so the busiest block of the
//                       tile set the pace: 22 % of the kernel's instructions.)
//            4. code    the lane walks the maps (count-leading-zeros over the bit-reversed map gives the next
//                       position, position 63 ends a block and loads the next block's map) and codes its symbols
//                       serially into its own word stream: coefficient, run from the previous position, category (clz),
//                       {code, length} from ONE 8-byte table load, append to a 64-bit register accumulator, full words
//                       to shared memory.  Every lane does the same amount of work however the symbols are spread over
//                       luma / chroma, busy / flat blocks; no scan, no atomics, no vote inside the loop.
//            5. merge   a warp scan of the 32 stream lengths places them; every lane shifts its words to their
//                       final bit position in the tile's region (plain stores for words it owns alone, atomicOr
//                       for the two it shares with its neighbours).
//            6. publish + (one iteration later) chain + write, exactly as in jpeg_kernel.cuh.
// A tile with more than 24576 bits, or one whose halves still overflow a lane's 768-bit stream, goes the slow way:
// four blocks at a time (always fit), written piecewise; same bytes.
#pragma once
#include "jpeg_kernel.cuh"

namespace jg {

constexpr int kEntBlocks = 32;                              // blocks per tile at most: one per lane in the map phase
constexpr int kEntCoefStride = 72;                          // int16 per staged block: 144-byte rows (the TMA box is 72 wide over a 64-wide tensor,
                                                            // the out-of-bounds columns arrive as zeros) keep 32 lanes off each other's banks
#ifndef JG_ENT_SUBWORDS
#define JG_ENT_SUBWORDS 24
#endif
constexpr int kSubWords = JG_ENT_SUBWORDS;                  // words of a lane's private stream (768 bits)
constexpr int kEntRegionWords = 768;                        // a tile's merged bits on the fast path: 24576 at most
#ifndef JG_ENT_PIECE
#define JG_ENT_PIECE 96
#endif
constexpr int kOnePieceSymbols = JG_ENT_PIECE * kSubWords;     // a tile with more symbols is coded in two halves right away (3 symbols per stream word: ~10.7 bits each)
constexpr int kSlowBlocks = 4;                              // slow path: 4 blocks at a time (<= 256 symbols, <= 8 x 59 bits per lane: always fit)
constexpr int kEntModePlain = 0, kEntModeRestart = 2;
constexpr unsigned kEntStageBytes = kEntBlocks * kEntCoefStride * 2;   // what one TMA box delivers (out-of-bounds rows count)
constexpr unsigned kTabRunBytes = 136u, kTabAcBytes = 16u * kTabRunBytes;    // see EntTables
constexpr unsigned kTabDc = 0u, kTabAc = 2u, kTabChroma = 1u, kTabAcChroma = kTabAcBytes / 128u;   // luma DC 0, chroma DC 1, luma AC 2, chroma AC 19
static_assert(kTabAcBytes % 128u == 0u, "table offsets are multiples of 128 bytes");

struct EntWarp {
    alignas(128) int16_t coef[kEntBlocks * kEntCoefStride];  // TMA destination
    alignas(16) uint32_t sub[kEntBlocks * kSubWords];        // word k of lane l at [k * 32 + l] (conflict-free)
    alignas(16) uint32_t region[kEntRegionWords + 8];        // the tile's merged bits (MSB-first words); survives into the next iteration
    alignas(8) uint2 rec[kEntBlocks + 4];                    // per block: {map of positions 0..31, map of 32..63} bit-reversed (position p at bit 31 - p mod 32);
                                                             // bit 0 of .y (position 63, always coded) carries the block's class instead.  The last 4 stay zero
    alignas(8) unsigned long long mbar;                      // completion of the staged coefficients
    Pending pend[2];
};
// Huffman tables as the symbol loop wants them: entry (run, category) = {code, length + category} at byte 136 * run +
// 8 * category of its table -- 17 entries per run, so that symbols of one category and different runs do not share a bank
// (with 16 the bank depended on the category alone: 9.5 wavefronts per lookup, measured).  The two DC tables (run 0) at
// byte 0 and 128, the two AC tables (16 runs x 136 bytes) at byte 256 and 2432.
struct EntTables { uint2 e[(256 + 2 * kTabAcBytes) / 8]; };
struct EntSmem {
    EntWarp wm[kEntWarps];
    EntTables tab;
};

// Component class and DC predecessor of block b of an image (bpm blocks per MCU: 1 gray, 3 4:4:4, 6 4:2:0 with
// Y00 Y01 Y10 Y11 Cb Cr): j = b mod bpm.
JG_DEV void block_role(int bpm, int j, unsigned& cls, int& delta)
{
    if (bpm == 6) { cls = j >= 4 ? 1u : 0u; delta = j >= 4 ? 6 : (j == 0 ? 3 : 1); }   // Y00 follows the previous MCU's Y11
    else { cls = (bpm == 3 && j != 0) ? 1u : 0u; delta = bpm; }
}

// 8 coefficients (one 16-byte vector) -> 8 flag bits
JG_DEV unsigned nonzero8(uint4 v)
{
    const unsigned m0 = v_minu2(v.x, 0x00010001u), m1 = v_minu2(v.y, 0x00010001u);
    const unsigned m2 = v_minu2(v.z, 0x00010001u), m3 = v_minu2(v.w, 0x00010001u);
    const unsigned a = (m0 + 4u * m1) + 16u * (m2 + 4u * m3);     // even coefficients at bits 0,2,4,6; odd ones at 16,18,20,22
    return (a | (a >> 15)) & 0xffu;
}

// Map phase, lane = block `slot` of the staged tile (active: slot < nblk): returns the block's symbol map (zero if inactive); its
// DC difference REPLACES the DC in the staged block (after every lane has read its predictor: the warp barrier inside).
JG_DEV uint2 map_block(EntWarp& W, int slot, int slot_j, int nblk, int jb, int bpm, int b0, int pred_outside, bool restart)
{
    const bool active = slot < nblk;
    int16_t* cz = W.coef + slot * kEntCoefStride;
    unsigned mlo = 0, mhi = 0, cls_of = 0;
    int diff = 0;
    if (active) {
        const uint4* row = reinterpret_cast<const uint4*>(cz);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mlo |= nonzero8(row[q]) << (8 * q);
            mhi |= nonzero8(row[4 + q]) << (8 * q);
        }
        mlo |= 1u;                     // the DC difference is always coded (jpeg_enc.h:834-844)
        mhi |= 0x80000000u;            // position 63: the last coefficient, or the end-of-block code where it is zero (:884-887)
        int j = jb + slot_j; if (j >= bpm) j -= bpm;                 // slot_j = slot mod bpm
        unsigned cls; int delta;
        block_role(bpm, j, cls, delta);
        cls_of = cls;
        const int ps = slot - delta;                                 // predecessor of the same component, tile-relative
        int pred = 0;
        if (ps >= 0) pred = W.coef[ps * kEntCoefStride];
        else if (!restart && b0 + ps >= 0) pred = pred_outside;      // before the tile: fetched from HBM by the caller
        diff = (int)cz[0] - pred;                                    // image start / restart interval: predictor 0 (jpeg_enc.h:1085-1087)
    }
    warp_sync();                       // every predictor has been read
    if (active) cz[0] = (int16_t)diff;
    fence_proxy_async();               // ... a generic store into the TMA's destination: ordered before the next tile's copy
    uint2 m; m.x = mlo; m.y = mhi;
    uint2 r; r.x = bit_reverse(mlo); r.y = (bit_reverse(mhi) & ~1u) | cls_of;
    W.rec[slot] = r;                   // (an inactive lane: zeros)
    warp_sync();
    return m;
}

// Code phase: the lane codes n symbols, starting with symbol r of block `blk` of the mapped tile, into its private word
// stream (word k at shared address sub + 128 * k); returns the bits.  Words beyond kSubWords pile up on the last slot (the
// count tells: the caller takes another route).  Everything inside the loop works on 32-bit shared-memory addresses.
JG_DEV unsigned code_symbols(const EntTables& T, const int16_t* coef, const uint2* rec, unsigned blk, unsigned r, unsigned n, const uint32_t* sub)
{
    unsigned alo = 0, ahi = 0;       // the last 64 bits appended
    unsigned t = 0;                  // bits so far
    unsigned wa = smem_addr(sub);    // where the next full word goes
    const unsigned wlast = wa + (kSubWords - 1) * 128u;
    auto put = [&](unsigned code, unsigned len) {          // len < 32
        ahi = funnel_l(alo, ahi, len);
        alo = (alo << len) | code;
        const unsigned t2 = t + len;
        if ((t ^ t2) >= 32u) {                             // a 32-bit boundary was crossed: the word that ends there is complete
            sts_u32(wa, funnel_r(alo, ahi, t2));           // (shift count taken mod 32)
            wa = add_min_u32(wa, 128u, wlast);
        }
        t = t2;
    };
    const unsigned tabs = pinned(smem_addr(&T.e[0]));
    // the walk: what is left of the current block's map (wh: positions 0..31, wl: 32..63, most significant bit first), its
    // record and coefficients, the position of the symbol before, the block's two tables
    unsigned ra = smem_addr(rec) + 8u * blk;
    unsigned cb = smem_addr(coef) + 2u * (unsigned)kEntCoefStride * blk;
    unsigned wh, wl, prevp = 0xffffffffu, dc_tb, ac_tb;
    auto open_block = [&]() {
        const uint2 q = lds_u64(ra);
        wh = q.x; wl = q.y | 1u;
        const unsigned cls = q.y & 1u;
        dc_tb = tabs + cls * 128u;
        ac_tb = tabs + 256u + cls * kTabAcBytes;
    };
    open_block();
    if (r) {
        // the first r symbols of the block belong to the lanes before me
        const unsigned ch = (unsigned)i_popc(wh);
        unsigned w = wh, base = 0u;
        if (r >= ch) {                                     // all of the first word (never empty: the DC is there)
            r -= ch;
            prevp = 32u - (unsigned)i_ffs(wh);
            w = wl; base = 32u; wh = 0u;
        }
        if (r) {
            unsigned pos = 0u;                             // the longest run of leading bits of w that holds r set bits
#pragma unroll
            for (unsigned sh = 16u; sh; sh >>= 1) {
                if ((unsigned)i_popc(w >> (32u - pos - sh)) <= r) pos += sh;
            }
            const unsigned keep = 0xffffffffu >> pos;
            prevp = base + 32u - (unsigned)i_ffs(w & ~keep);
            w &= keep;
        }
        if (base) wl = w; else wh = w;
    }
    // next symbol of the walk: address of its coefficient, zeros since the symbol before it in the block
    // (jpeg_enc.h:856-862; the DC at position 0 opens the block), its table
    struct Pos { unsigned ca, run, tb; };
    auto step = [&](Pos& o) {
        const bool inhi = wh != 0u;
        const unsigned w = inhi ? wh : wl;
        const unsigned f = (unsigned)i_clz(w);             // (w == 0 only behind the lane's last symbol: those are never used)
        const unsigned p = inhi ? f : f + 32u;
        const unsigned w2 = w & ~(0x80000000u >> (f & 31u));
        wh = inhi ? w2 : 0u;
        wl = inhi ? wl : w2;
        o.ca = cb + 2u * p;
        o.run = p - prevp - 1u;
        o.tb = p ? ac_tb : dc_tb;
        {
            // Position 63 is the block's last symbol (its last coefficient or the end-of-block code, :884-887): the walk moves
            // on to the next block.  Without a branch: some lane of the warp finishes a block in nine iterations of ten, so a
            // branch was taken (divergently) almost every time -- selects are 6 % faster (measured).
            const bool last = p == 63u;
            const uint2 q = lds_u64(ra + 8u);
            ra = last ? ra + 8u : ra;
            cb = last ? cb + 2u * (unsigned)kEntCoefStride : cb;
            prevp = last ? 0xffffffffu : p;
            wh = last ? q.x : wh;
            wl = last ? (q.y | 1u) : wl;
            const unsigned cls = q.y & 1u;
            dc_tb = last ? tabs + cls * 128u : dc_tb;
            ac_tb = last ? tabs + 256u + cls * kTabAcBytes : ac_tb;
        }
    };
    // one symbol: its Huffman code + amplitude bits, the ZRL codes in front of it, its table
    struct Sym { unsigned val, len, tb, nz; };
    auto lookup = [&](const Pos& q, int v, Sym& y) {
        const unsigned run = v == 0 ? 0u : q.run;              // the end-of-block code (and a zero DC difference): entry 0 of its table
        y.tb = q.tb;
        y.nz = run >> 4;                                       // one ZRL per 16 zeros (:863-867)
        const unsigned lz = (unsigned)i_clz((unsigned)(v < 0 ? -v : v));          // category = 32 - lz (jpeg_enc.h:598-608)
        const uint2 h = lds_u64(q.tb + 256u + (run & 15u) * kTabRunBytes - 8u * lz);   // entry (run, category): {code, length + category}
        const unsigned x = funnel_lc(0u, (unsigned)(v + (v >> 31)), lz);           // amplitude bits (:601-609), left-aligned; none for category 0
        y.val = funnel_rc(x, h.x, lz);                                             // (code << category) | amplitude
        y.len = h.y;
    };
    auto emit = [&](const Sym& y) {
        if (y.nz) {
            const uint2 z = lds_u64(y.tb + 15u * kTabRunBytes);
#pragma unroll 1
            for (unsigned k = y.nz; k; --k) put(z.x, z.y);
        }
        put(y.val, y.len);
    };
    // TWO symbols per iteration: their loads, category and table lookups are independent and overlap (the loop is bound by
    // the latency of that chain, not by issue slots); positions and coefficients are requested one iteration ahead.
    // (The walk runs up to four symbols past the lane's last one -- into the next blocks' records, or the zero records
    // behind the tile -- and those are never used.)
    Pos q0, q1;
    step(q0); step(q1);
    int v0 = lds_s16(q0.ca), v1 = lds_s16(q1.ca);
#pragma unroll 1
    for (unsigned i = 0; i < n; i += 2u) {
        Pos q2, q3;
        step(q2); step(q3);
        const int v2 = lds_s16(q2.ca), v3 = lds_s16(q3.ca);
        Sym a, b;
        lookup(q0, v0, a);
        lookup(q1, v1, b);
        if (i + 1u >= n) { b.val = 0u; b.len = 0u; b.nz = 0u; }     // an odd count: nothing is appended for the missing symbol
        emit(a);
        emit(b);
        q0 = q2; q1 = q3; v0 = v2; v1 = v3;
    }
    if (t & 31u) sts_u32(wa, alo << (32u - (t & 31u)));                     // the rest, left-aligned
    return t;
}

// zero words [0, n) of a 16-byte aligned word array, 16 bytes per lane and step (may zero up to 3 words more; followed by a warp barrier)
JG_DEV void clear_words16(uint32_t* region, unsigned n)
{
    const uint4 z = {0u, 0u, 0u, 0u};
    for (unsigned i = 4u * (unsigned)(JG_TID & 31); i < n; i += 128u) *reinterpret_cast<uint4*>(region + i) = z;
    warp_sync();
}

// The lanes' streams (lane l: nbits bits at sub[k * 32 + l], ending at bit `incl` of the tile: the inclusive scan of the
// lengths) -> one contiguous bit string in `region` (zeroed; the caller has checked that everything fits).  A word of the region that holds bits of ONE lane only is stored, the others (at most two per
// lane) are OR-ed in.
JG_DEV void merge_streams(const uint32_t* sub, unsigned nbits, unsigned incl, uint32_t* region)
{
    const unsigned lane = (unsigned)(JG_TID & 31);
    const unsigned o = incl - nbits, end = incl;
    const unsigned nwords = (nbits + 31u) >> 5, s = o & 31u, d = o >> 5;
    const unsigned kmax = warp_max_u32(nwords);
    unsigned prev = 0;
    for (unsigned k = 0; k <= kmax; ++k) {
        const unsigned cur = k < nwords ? sub[k * 32u + lane] : 0u;
        if (k <= nwords) {
            const unsigned w = funnel_r(cur, prev, s);           // (prev:cur) >> s
            const unsigned bit0 = (d + k) << 5;
            if (bit0 >= o && bit0 + 32u <= end) region[d + k] = w;
            else if (w) smem_atomic_or(region + d + k, w);
        }
        prev = cur;
    }
    warp_sync();
}

// The tile this warp coded one iteration ago: its bits still sit in the region and leave right before the region is needed
// again -- for the first merge of the next tile.  By then a whole map + code phase has passed since its size was
// published, and the sizes of its predecessors are almost always there: the look-back does not wait.
struct PendingOut {
    int g = -1;          // tile (launch-wide index), -1: none
    int slot = 0;        // its Pending record
    bool failed = false; // the look-back timed out: the kernel gives up
};
JG_DEV_NOINLINE bool ent_tile_back(const LaunchParams& P, EntWarp& W, int g, int slot);
JG_DEV void flush_pending(const LaunchParams& P, EntWarp& W, PendingOut& po)
{
    if (po.g < 0) return;
    if (!ent_tile_back(P, W, po.g, po.slot)) po.failed = true;
    po.g = -1;
}

// Blocks [lo, hi) of the staged + mapped tile -> their bits appended at bit `bit_base` of the region (zero from there on).
// The previous tile, whose bits still sit in the region, leaves after the symbols are coded, right before the merge.
// Returns the bits of the blocks; `fits` = every lane's stream and the total fit.
JG_DEV unsigned code_blocks(const LaunchParams& P, EntWarp& W, const EntTables& T, PendingOut& po, uint2 own, int lo, int hi, unsigned bit_base, bool& fits)
{
    const int lane = JG_TID & 31;
    const unsigned cnt = (lane >= lo && lane < hi) ? (unsigned)(i_popc(own.x) + i_popc(own.y)) : 0u;      // lane = block
    const unsigned incl = warp_scan_incl_u32(cnt);
    const unsigned S = warp_shfl_u32(incl, 31);
    const unsigned q = (S + 31u) >> 5;                   // symbols per lane
    const unsigned s0 = (unsigned)lane * q < S ? (unsigned)lane * q : S;
    const unsigned n = S - s0 < q ? S - s0 : q;
    // the block that holds symbol s0: the number of blocks whose inclusive count is <= s0 (31 for a lane without symbols)
    unsigned b = 0;
#pragma unroll
    for (unsigned st = 16u; st; st >>= 1) {
        const unsigned v = warp_shfl_u32(incl, (int)(b + st - 1u));
        if (v <= s0) b += st;
    }
    const unsigned before = warp_shfl_u32(incl - cnt, (int)b);
    const unsigned nbits = code_symbols(T, W.coef, W.rec, b, n ? s0 - before : 0u, n, W.sub + lane);
    warp_sync();                                         // the streams are complete
    const unsigned incl_bits = warp_scan_incl_u32(nbits);
    const unsigned bits = warp_shfl_u32(incl_bits, 31);
    fits = warp_ballot(nbits > (unsigned)kSubWords * 32u) == 0u && bit_base + bits <= (unsigned)kEntRegionWords * 32u;
    flush_pending(P, W, po);                             // the region is needed now: the previous tile leaves
    if (po.failed) fits = false;
    if (fits) merge_streams(W.sub, nbits, bit_base + incl_bits, W.region);
    return bits;
}

// The whole staged + mapped tile on the fast path.  A tile with more than 48 symbols per lane (or one where a lane's
// stream went beyond its 768 bits) is coded in two halves of 16 blocks: half the symbols and half the bits per lane.
JG_DEV unsigned code_tile(const LaunchParams& P, EntWarp& W, const EntTables& T, PendingOut& po, uint2 own, int nblk, bool& fits)
{
    const unsigned cnt = (unsigned)(i_popc(own.x) + i_popc(own.y));
    const unsigned S = warp_shfl_u32(warp_scan_incl_u32(cnt), 31);
    if (S <= (unsigned)kOnePieceSymbols || nblk <= 16) {
        const unsigned bits = code_blocks(P, W, T, po, own, 0, nblk, 0u, fits);
        if (fits || po.failed || nblk <= 16 || bits > (unsigned)kEntRegionWords * 32u) return bits;
        // a lane's stream overflowed: once more, in halves
    }
    unsigned bits = code_blocks(P, W, T, po, own, 0, 16, 0u, fits);
    if (!fits) return bits;
    bits += code_blocks(P, W, T, po, own, 16, nblk, bits, fits);
    return bits;
}

// Bits of every block of the mapped tile, for the stage dumps of the parity tests (lane = block; not on the product's path).
JG_DEV_NOINLINE unsigned count_block_bits(const EntWarp& W, const EntTables& T, uint2 m, int slot, int nblk, int jb, int bpm)
{
    if (slot >= nblk) return 0u;
    unsigned cls; int delta;
    block_role(bpm, (jb + slot) % bpm, cls, delta);
    const int16_t* cz = W.coef + slot * kEntCoefStride;
    unsigned bits = 0;
    int prev = -1;
    for (int pos = 0; pos < 64; ++pos) {
        if (!(((pos < 32 ? m.x : m.y) >> (pos & 31)) & 1u)) continue;
        const int v = cz[pos];
        unsigned run = (unsigned)(pos - prev - 1);
        prev = pos;
        if (v == 0) run = 0;
        const uint2* tb = T.e + 16u * ((pos ? kTabAc + (cls ? kTabAcChroma : 0u) : (cls ? kTabChroma : 0u)));
        bits += (run >> 4) * tb[15 * 17].y;
        const unsigned cat = v ? 32u - (unsigned)i_clz((unsigned)(v < 0 ? -v : v)) : 0u;
        bits += tb[(run & 15u) * 17u + cat].y;
    }
    return bits;
}

JG_DEV_NOINLINE bool ent_tile_back(const LaunchParams& P, EntWarp& W, int g, int slot)
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
    flush_region(W.region, pd.T, pd.last != 0, reinterpret_cast<uint8_t*>(pd.raw), pd.raw_cap, k, hb, pos, overflow);
    if ((JG_TID & 31) == 0) {
        if (pd.last) P.raw_bytes[pd.img_idx] = pos;
        if (overflow) gmem_atomic_or(P.img_status + pd.img_idx, 1u);
    }
    warp_sync();                                        // every lane has read the region
    clear_words16(W.region, (pd.T >> 5) + 2u);
    return true;
}

// A tile that does not fit the fast path: four blocks at a time, written right away.  The groups are coded once to
// learn the tile's size and last bits (successors must not wait for the whole slow pass), then again to be written.
// Returns false on a look-back timeout.
template <bool restart>
JG_DEV_NOINLINE bool ent_tile_slow(const LaunchParams& P, EntWarp& W, const EntTables& T, uint2 own, int g, int nblk, int slot)
{
    const int lane = JG_TID & 31;
    PendingOut po;                    // (nothing pending: the caller has written the previous tile out)
    const Pending pd = W.pend[slot];
    const bool first = g == pd.first_tile_of_img;
    bool fits;
    unsigned bits = 0, tail = 0;
    for (int lo = 0; lo < nblk; lo += kSlowBlocks) {
        warp_sync();
        clear_region(W.region, kEntRegionWords + 8);
        const unsigned tg = code_blocks(P, W, T, po, own, lo, lo + kSlowBlocks < nblk ? lo + kSlowBlocks : nblk, 0u, fits);
        bits += tg;
        // running last-7-bits of the tile (a group may hold fewer than 7)
        tail = tg >= 7u ? tail_bits(W.region, tg) : (((tail << tg) | peek_bits(W.region, 0u, tg)) & 0x7fu);
    }
    const unsigned pad = restart ? (0u - bits) & 7u : 0u;       // restart interval: 1-bits up to the byte boundary
    if (pad) { tail = ((tail << pad) | ((1u << pad) - 1u)) & 0x7fu; bits += pad; }
    if (lane == 0) st_flag64(P.desc_bits + g, make_desc(first ? kStatusPrefix : kStatusAgg, tail, bits));
    unsigned long long bit_base = 0;
    unsigned pred_tail = 0;
    if (!first && !chain_bits(P, g, pd.first_tile_of_img, bits, tail, bit_base, pred_tail)) return false;
    unsigned k = (unsigned)(bit_base & 7ull);
    unsigned hb = pred_tail & ((1u << k) - 1u);
    unsigned long long pos = bit_base >> 3;
    bool cap_overflow = false;
    for (int lo = 0; lo < nblk; lo += kSlowBlocks) {
        const int hi = lo + kSlowBlocks < nblk ? lo + kSlowBlocks : nblk;
        warp_sync();
        clear_region(W.region, kEntRegionWords + 8);
        unsigned tg = code_blocks(P, W, T, po, own, lo, hi, 0u, fits);
        if (pad && hi == nblk) {
            if (lane == 0) set_ones(W.region, tg, pad);
            warp_sync();
            tg += pad;
        }
        flush_region(W.region, tg, pd.last && hi == nblk, reinterpret_cast<uint8_t*>(pd.raw), pd.raw_cap, k, hb, pos, cap_overflow);
    }
    if (lane == 0) {
        if (pd.last) P.raw_bytes[pd.img_idx] = pos;
        if (cap_overflow) gmem_atomic_or(P.img_status + pd.img_idx, 1u);
    }
    warp_sync();
    clear_region(W.region, kEntRegionWords + 8);
    return true;
}

// ------------------------------------------------------------------------------------------
// kernel B: coefficients -> unstuffed entropy-coded bits (raw), bits chained over the tiles of an image
// ------------------------------------------------------------------------------------------
// The previous tile is written out in the MIDDLE of the iteration -- after the current tile's symbols are coded, right
// before its merge needs the region -- instead of at the top.  With one image all ~2600 tiles in flight are consecutive
// tiles of that image; a write-out at the top of the iteration asks for the sizes of tiles that were drawn nanoseconds
// before ours and are published at the same moment as ours: every warp waits for the slowest of its 32 predecessors
// (16384^2 gray: 1.03 ms, ~100 polling rounds per tile); with most of an iteration in between the sizes are there (0.60 ms).
template <int MODE>
JG_KERNEL(kEntThreads, JG_ENT_MINB)
void entropy_kernel(const JG_GRID_CONSTANT LaunchParams P, const JG_GRID_CONSTANT CoefMap cmap)
{
    constexpr bool restart = MODE == kEntModeRestart;
    JG_DYNAMIC_SMEM(smem_raw);
    EntSmem& S = *reinterpret_cast<EntSmem*>(smem_raw);
    const int t = JG_TID, lane = t & 31;
    for (int i = t; i < 32 + 512; i += kEntThreads) {
        const int cls = i < 32 ? i >> 4 : (i - 32) >> 8, k = i < 32 ? i & 15 : (i - 32) & 255;
        const unsigned h = i < 32 ? P.huff->dc[cls][k] : P.huff->ac[cls][k];      // code << 8 | length
        uint2 e;
        e.x = h >> 8;
        e.y = (h & 0xffu) + (unsigned)(k & 15);
        S.tab.e[i < 32 ? i : (256u + (unsigned)cls * kTabAcBytes) / 8u + (unsigned)(k >> 4) * 17u + (unsigned)(k & 15)] = e;
    }
    EntWarp& W = S.wm[t >> 5];
    const EntTables& T = S.tab;
    if (lane == 0) mbar_init(&W.mbar, 1u);
    clear_region(W.region, kEntRegionWords + 8);
    if (lane < 4) { uint2 z; z.x = 0u; z.y = 0u; W.rec[kEntBlocks + lane] = z; }
    mbar_fence_init();
    cta_sync();       // tables + barriers are set up: the only CTA barrier; from here every warp is on its own

    const int bpm = P.bpm, bpt = P.blocks_per_tile;
    const int lane_j = lane % bpm;
    unsigned phase = 0;
    PendingOut po;                    // the tile coded one iteration ago, still to be written
    for (int slot = 0;; slot ^= 1) {
        int img_idx;
        const int g = draw_tile(P, img_idx);
        const bool have = g < P.n_tiles;
        int nblk = 0, b0 = 0, jb = 0, pred_outside = 0;
        bool first = false;
        if (have) {
            const ImageDesc im = P.images[img_idx];
            const int lt = g - im.first_tile;
            const int n_blocks = im.n_mcus * bpm;
            b0 = lt * bpt;
            nblk = n_blocks - b0 < bpt ? n_blocks - b0 : bpt;
            first = lt == 0;
            jb = b0 % bpm;
            // ---- stage: one TMA box (32 blocks from row first_block + b0 of the coefficient plane; rows past the
            //      plane's end arrive as zeros and are never looked at).  The previous tile's coefficients were
            //      overwritten by generic stores (DC differences): order them before the async-proxy write. ----
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&W.mbar, kEntStageBytes);
                tma_load_2d(W.coef, &cmap, 0, (int)(im.first_block + (unsigned long long)b0), &W.mbar);
            }
            // DC predecessor that lies before the tile (not with restart intervals: they predict from 0)
            if (!restart && lane < nblk) {
                int j = jb + lane_j; if (j >= bpm) j -= bpm;
                unsigned cls; int delta;
                block_role(bpm, j, cls, delta);
                const int ps = lane - delta;
                if (ps < 0 && b0 + ps >= 0)
                    pred_outside = (int)ldg_s16(P.coefs + (im.first_block + (unsigned long long)(b0 + ps)) * 64ull);
            }
            if (lane == 0) {         // what the write-out needs later; T and tail follow after the coding
                Pending& pd = W.pend[slot];
                pd.raw = reinterpret_cast<unsigned long long>(im.raw); pd.raw_cap = im.raw_cap;
                pd.img_idx = img_idx; pd.first_tile_of_img = im.first_tile; pd.last = (lt == im.n_tiles - 1) ? 1 : 0;
                pd.base = 0;
            }
        }
        if (!have) {
            flush_pending(P, W, po);                     // the last tile of this warp
            break;
        }
        mbar_wait(&W.mbar, phase);
        phase ^= 1u;

        // ---- map (lane = block), then code (lane = an equal share of the tile's symbols) ----
        const uint2 own = map_block(W, lane, lane_j, nblk, jb, bpm, b0, pred_outside, restart);
        if (P.dbg_bits) {
            const unsigned nb = count_block_bits(W, T, own, lane, nblk, jb, bpm);
            if (lane < nblk) P.dbg_bits[P.images[img_idx].first_block + (unsigned long long)(b0 + lane)] = nb;
        }
        bool fits;
        unsigned bits = code_tile(P, W, T, po, own, nblk, fits);
        if (po.failed) break;
        if (P.win_words < kWinWordsMax && bits > 32u * (unsigned)P.win_words) {                // parity tests: force the slow path
            fits = false;
            warp_sync();
            clear_region(W.region, kEntRegionWords + 8);
        }
        if (fits) {
            // Publish the tile's bit count and last bits NOW: they are consumed (by us and by every successor) later.
            unsigned tail = tail_bits(W.region, bits);
            if (restart && (bits & 7u) != 0u) {      // restart interval: 1-bits up to the byte boundary (T.81 F.1.2.3)
                const unsigned pad = 8u - (bits & 7u);
                if (lane == 0) set_ones(W.region, bits, pad);
                tail = ((tail << pad) | ((1u << pad) - 1u)) & 0x7fu;
                bits += pad;
            }
            if (lane == 0) {
                st_flag64(P.desc_bits + g, make_desc(first ? kStatusPrefix : kStatusAgg, tail, bits));
                W.pend[slot].T = bits; W.pend[slot].tail = tail;
            }
            warp_sync();
            po.g = g; po.slot = slot;
        } else {
            flush_pending(P, W, po);
            if (po.failed) break;
            if (!ent_tile_slow<restart>(P, W, T, own, g, nblk, slot)) break;
        }
    }
}

}  // namespace jg
