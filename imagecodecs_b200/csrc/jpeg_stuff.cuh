// jpeg_stuff.cuh -- second pass: 0xFF00 byte stuffing (jpeg_enc.h:634-638) + EOI (:1166-1167).
//
// Pass 1 leaves every image's entropy-coded bits unstuffed and byte-aligned in `raw`.  Stuffing moves every byte by the
// number of 0xFF bytes before it: a scan over the whole image.  Round 1 (and the first half of round 2) ran it as ONE
// kernel with a decoupled look-back over 8 KB chunks; ncu showed that kernel waiting, not working: the 1184 chunks in
// flight are neighbours in one image, every look-back walks hundreds of descriptors back to the nearest known prefix
// (44 % of the instructions, half of the stall samples in the back-off) and the pass sat at 1.5 TB/s whatever the data.
// Now the scan is taken out of the copy:
//
//   plan_chunks_kernel   one CTA: first_chunk[i] = sum over images j < i of ceil(raw_bytes[j] / 8192)
//   count_ff_kernel      a CTA per 32 consecutive chunks, a warp per chunk: counts 0xFF (__vcmpeq4), finds the chunk's
//                        image, writes desc_ff[c] = image << 32 | extra bytes of the chunks before c in its group of 32,
//                        and the group's total
//   scan_groups_kernel   one CTA: exclusive scan of the group totals (n_chunks / 32 values)
//   stuff_kernel         chunk c of image i starts at  c's prefix - prefix of the image's first chunk: no CTA waits for
//                        another.  A warp owns 2 KB of the chunk as 16 rows of 32 consecutive words (one word per lane:
//                        shared-memory accesses free of bank conflicts); positions inside a row from four ballots; the
//                        stuffed bytes are laid down in shared memory AT THE DESTINATION'S 16-byte phase and leave by
//                        LDS.128 / STG.128.
//                        Restart mode keeps the round-1 shape (64 contiguous bytes per thread, markers merged in).
// raw is read twice (the second time partly from L2); in exchange nothing is serial.
#pragma once
#include "jpeg_device.h"
#include "jpeg_kernel.cuh"
#include "jpeg_launch.h"

namespace jg {

// restart mode: a tile is at least 24 blocks x 4 bits = 12 bytes, so a chunk holds at most this many tile ends
constexpr int kMaxMarks = kChunkBytes / 12 + 2;

struct StuffSmem {
    alignas(16) uint8_t sbuf[2 * kChunkBytes + 2 * kMaxMarks + 64];   // every byte 0xFF + a marker per tile end
    uint32_t warp_tmp[kWarps];
    uint32_t mark[kMaxMarks + 1];      // restart mode: raw offsets (relative to the chunk) of the tile ends inside the chunk
    unsigned warp_ff[kWarps];
    unsigned n_marks, first_mark;      // tile ends inside the chunk / index of the first one in the image
};
constexpr int kGroupChunks = 32;       // chunks per group of the two-level scan
struct CountSmem { unsigned cnt[kGroupChunks]; int img[kGroupChunks]; };
struct ScanSmem { uint32_t warp_tmp[kWarps]; };

// Restart mode (JPEG_GPU_FLAG_RESTART): after pass 1 desc_bits[first_tile + j] holds the inclusive bit
// count of tile j of the image, a multiple of 8: the raw byte offset e_j at which restart interval j
// ends.  RST(j mod 8) goes in front of raw byte e_j for every j but the last tile.
JG_DEV unsigned long long tile_end_byte(const LaunchParams& P, const ImageDesc& im, unsigned j)
{
    return (ld_flag64(P.desc_bits + im.first_tile + j) & kCountMask) >> 3;
}
// number of tile ends e_j (j < n) with e_j < x
JG_DEV unsigned tile_ends_below(const LaunchParams& P, const ImageDesc& im, unsigned n, unsigned long long x)
{
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (tile_end_byte(P, im, mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

JG_KERNEL(kThreads, 1)
void plan_chunks_kernel(const JG_GRID_CONSTANT LaunchParams P)
{
    JG_DYNAMIC_SMEM(smem_raw);
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(smem_raw);
    const int t = JG_TID;
    unsigned carry = 0;
    for (int i0 = 0; i0 < P.n_images; i0 += kThreads) {
        const int i = i0 + t;
        unsigned n = 0;
        if (i < P.n_images) {
            const unsigned long long rb = P.raw_bytes[i];
            if (P.img_status[i] & 1u) P.scan_bytes[i] = 2ull * rb + 2ull;   // did not fit: report a sufficient size
            else n = (unsigned)((rb + kChunkBytes - 1) / kChunkBytes);
        }
        unsigned tot;
        const unsigned ex = cta_scan_excl(n, S.warp_tmp, tot);
        if (i < P.n_images) P.first_chunk[i] = carry + ex;
        carry += tot;
    }
    if (t == 0) P.first_chunk[P.n_images] = carry;
}

// The image of chunk c: the last one whose first chunk is <= c (images without chunks are skipped).  The whole warp calls
// it: a 32-way search over first_chunk -- two rounds of loads for 1024 images.
JG_DEV int image_of_chunk(const LaunchParams& P, unsigned c, int lane)
{
    int lo = 0, hi = P.n_images - 1;
    while (lo < hi) {
        const int step = (hi - lo) / 32 + 1;
        const int probe = lo + (lane + 1) * step;
        const int ok = probe <= hi && P.first_chunk[probe <= hi ? probe : hi] <= c;
        const int k = i_popc(warp_ballot(ok));                     // first_chunk is sorted: the lanes that say yes are a prefix
        lo += k * step;
        if (lo + step - 1 < hi) hi = lo + step - 1;
    }
    return lo;
}

// extra bytes the stuffing pass adds to the nb raw bytes at `src` (offset `off` of the image): a zero per 0xFF, two marker
// bytes per restart interval that ends inside.  The warp reads the chunk as uint4 rows; every lane returns the total.
JG_DEV unsigned count_extra_bytes(const LaunchParams& P, const ImageDesc& im, unsigned long long off, unsigned nb, int lane)
{
    const uint8_t* src = im.raw + off;                 // raw is 256-byte aligned and a chunk starts at a multiple of its size
    unsigned cnt = 0;
    constexpr int NV = kChunkBytes / 512;
#pragma unroll 8
    for (int k = 0; k < NV; ++k) {
        const unsigned b = (unsigned)(k * 32 + lane) * 16u;
        if (b < nb) {
            uint4 v = ldg_u128(src + b);
            if (nb - b < 16u) {                        // the image's last vector: bytes past the end do not count
                const unsigned v4 = nb - b;
                unsigned* w = &v.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int have = (int)v4 - 4 * j;
                    if (have <= 0) w[j] = 0u; else if (have < 4) w[j] &= 0xffffffffu >> (8 * (4 - have));
                }
            }
            cnt += (unsigned)(i_popc(v_cmpeq4(v.x, 0xffffffffu)) + i_popc(v_cmpeq4(v.y, 0xffffffffu)) +
                              i_popc(v_cmpeq4(v.z, 0xffffffffu)) + i_popc(v_cmpeq4(v.w, 0xffffffffu)));
        }
    }
    unsigned total = warp_sum_u32(cnt >> 3);
    if ((im.flags & kFlagRestart) != 0 && im.n_tiles > 1) {
        const unsigned n_ends = (unsigned)im.n_tiles - 1u;          // no marker after the last tile
        total += 2u * (tile_ends_below(P, im, n_ends, off + nb) - tile_ends_below(P, im, n_ends, off));
    }
    return total;
}

constexpr int kCountThreads = 32 * kGroupChunks;     // a warp per chunk of the group
JG_KERNEL(kCountThreads, 2)
void count_ff_kernel(const JG_GRID_CONSTANT LaunchParams P)
{
    JG_DYNAMIC_SMEM(smem_raw);
    CountSmem& S = *reinterpret_cast<CountSmem*>(smem_raw);
    const int t = JG_TID, lane = t & 31, wq = t >> 5;
    const unsigned n_chunks = P.first_chunk[P.n_images];
    const unsigned n_groups = (n_chunks + kGroupChunks - 1) / kGroupChunks;
    for (unsigned g = (unsigned)JG_CTA_ID; g < n_groups; g += (unsigned)JG_GRID_DIM) {
        const unsigned c = g * kGroupChunks + (unsigned)wq;
        unsigned extra = 0;
        int img = 0;
        if (c < n_chunks) {
            img = image_of_chunk(P, c, lane);
            const ImageDesc im = P.images[img];
            const unsigned long long raw_n = P.raw_bytes[img];
            const unsigned long long off = (unsigned long long)(c - P.first_chunk[img]) * kChunkBytes;
            const unsigned nb = raw_n - off < (unsigned long long)kChunkBytes ? (unsigned)(raw_n - off) : (unsigned)kChunkBytes;
            extra = count_extra_bytes(P, im, off, nb, lane);
        }
        if (lane == 0) { S.cnt[wq] = extra; S.img[wq] = img; }
        cta_sync();
        if (wq == 0) {
            const unsigned v = S.cnt[lane];
            const unsigned inc = warp_scan_incl_u32(v);
            const unsigned cl = g * kGroupChunks + (unsigned)lane;
            if (cl < n_chunks) P.desc_ff[cl] = ((unsigned long long)(unsigned)S.img[lane] << 32) | (unsigned long long)(inc - v);
            if (lane == 31) P.ff_groups[g] = inc;
        }
        cta_sync();
    }
}

// exclusive scan (mod 2^32: only differences inside one image are used) of the group totals, in place
JG_KERNEL(kThreads, 1)
void scan_groups_kernel(const JG_GRID_CONSTANT LaunchParams P)
{
    JG_DYNAMIC_SMEM(smem_raw);
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(smem_raw);
    const int t = JG_TID;
    const unsigned n_chunks = P.first_chunk[P.n_images];
    const unsigned n_groups = (n_chunks + kGroupChunks - 1) / kGroupChunks;
    const unsigned per = (n_groups + kThreads - 1) / kThreads;       // a contiguous run per thread: one CTA scan however many groups
    const unsigned lo = (unsigned)t * per, hi = lo + per < n_groups ? lo + per : n_groups;
    unsigned sum = 0;
    for (unsigned i = lo; i < hi; ++i) sum += P.ff_groups[i];
    unsigned tot;
    unsigned run = cta_scan_excl(sum, S.warp_tmp, tot);
    for (unsigned i = lo; i < hi; ++i) { const unsigned v = P.ff_groups[i]; P.ff_groups[i] = run; run += v; }
}

// extra bytes in front of chunk c, counted from the start of the launch (mod 2^32)
JG_DEV unsigned extra_before(const LaunchParams& P, unsigned c, unsigned long long d)
{
    return P.ff_groups[c / kGroupChunks] + (unsigned)d;
}

// Plain (restart-free) chunk.  Warp `wq` owns bytes [2048 wq, 2048 wq + 2048) of the chunk as ROWS rows of 32 words, lane l
// the word l of each row.
JG_DEV unsigned stuff_plain_chunk(const LaunchParams& P, StuffSmem& S, const ImageDesc& im, int img, unsigned ff_base,
                              unsigned long long off, unsigned nb, bool last_chunk, int t)
{
    constexpr int ROWS = kChunkBytes / (kWarps * 128);
    const int lane = t & 31, wq = t >> 5;
    const unsigned wbase = (unsigned)wq * (ROWS * 128u) + 4u * (unsigned)lane;      // my byte of row 0
    const uint8_t* src = im.raw + off + wbase;
    unsigned w[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) w[q] = wbase + 128u * q < nb ? ldg_u32(src + 128 * q) : 0u;
    if (nb < (unsigned)kChunkBytes) {                 // the image's last chunk: bytes past the end read as zero (never 0xFF)
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const int v = (int)nb - (int)(wbase + 128u * q);
            if (v > 0 && v < 4) w[q] &= 0xffffffffu >> (8 * (4 - v));
        }
    }
    unsigned cnt = 0;
#pragma unroll
    for (int q = 0; q < ROWS; ++q) cnt += (unsigned)i_popc(v_cmpeq4(w[q], 0xffffffffu));
    const unsigned wtot = warp_sum_u32(cnt >> 3);
    if (lane == 0) S.warp_ff[wq] = wtot;
    cta_sync();
    unsigned ff_chunk = 0, ff_mine = 0;               // stuffed zeros of the chunk / of the warps before mine
#pragma unroll
    for (int k = 0; k < kWarps; ++k) { const unsigned f = S.warp_ff[k]; ff_chunk += f; if (k < wq) ff_mine += f; }

    const unsigned long long pos = off + ff_base;
    const unsigned out_bytes = nb + ff_chunk;
    const bool fits = pos + out_bytes + (last_chunk ? 2u : 0u) <= im.out_cap;
    uint8_t* dst = im.out + pos;
    const unsigned a = (unsigned)((size_t)dst & 15u);          // sbuf[a + i] is output byte i: sbuf and dst share the 16-byte phase
    if (fits) {
        unsigned o = a + wbase + ff_mine;                      // where my byte of the row goes when no 0xFF precedes it in the row
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const unsigned x = w[q];
            const unsigned m = v_cmpeq4(x, 0xffffffffu);
            if (warp_ballot(m != 0u) == 0u) {                  // six rows of ten: no 0xFF in these 128 bytes
                S.sbuf[o] = (uint8_t)x; S.sbuf[o + 1] = (uint8_t)(x >> 8); S.sbuf[o + 2] = (uint8_t)(x >> 16); S.sbuf[o + 3] = (uint8_t)(x >> 24);
                o += 128u;
            } else {
                const unsigned e0 = warp_ballot(m & 0xffu), e1 = warp_ballot(m & 0xff00u);
                const unsigned e2 = warp_ballot(m & 0xff0000u), e3 = warp_ballot(m & 0xff000000u);
                const unsigned before = (unsigned)(i_popc(e0 & lt) + i_popc(e1 & lt) + i_popc(e2 & lt) + i_popc(e3 & lt));
                unsigned p = o + before;
                S.sbuf[p++] = (uint8_t)x;         if (m & 0xffu) S.sbuf[p++] = 0;          // jpeg_enc.h:634-638
                S.sbuf[p++] = (uint8_t)(x >> 8);  if (m & 0xff00u) S.sbuf[p++] = 0;
                S.sbuf[p++] = (uint8_t)(x >> 16); if (m & 0xff0000u) S.sbuf[p++] = 0;
                S.sbuf[p++] = (uint8_t)(x >> 24); if (m & 0xff000000u) S.sbuf[p] = 0;
                o += 128u + (unsigned)(i_popc(e0) + i_popc(e1) + i_popc(e2) + i_popc(e3));
            }
        }
    }
    cta_sync();
    if (fits) {
        uint8_t* dstv = dst - a;                               // 16-byte aligned
        const unsigned end = a + out_bytes, vbeg = (a + 15u) & ~15u, vend = end & ~15u;
        if (vbeg >= vend) {
            for (unsigned i = a + (unsigned)t; i < end; i += kThreads) dstv[i] = S.sbuf[i];
        } else {
            if (a + (unsigned)t < vbeg) dstv[a + t] = S.sbuf[a + t];
            const uint4* s4 = reinterpret_cast<const uint4*>(S.sbuf);
            uint4* d4 = reinterpret_cast<uint4*>(dstv);
            for (unsigned i = (vbeg >> 4) + (unsigned)t; i < (vend >> 4); i += kThreads) d4[i] = s4[i];
            if (vend + (unsigned)t < end) dstv[vend + t] = S.sbuf[vend + t];
        }
    }
    if (t == 0) {
        if (!fits) gmem_atomic_or(P.img_status + img, 1u);
        if (last_chunk) {
            if (fits) { dst[out_bytes] = 0xFF; dst[out_bytes + 1] = 0xD9; }   // EOI (jpeg_enc.h:1166-1167)
            P.scan_bytes[img] = pos + out_bytes + 2;
        }
    }
    return ff_chunk;
}

// Restart-mode chunk: 64 contiguous bytes per thread, RSTn markers merged in at the tile ends (S.mark), shifted copy-out.
JG_DEV unsigned stuff_restart_chunk(const LaunchParams& P, StuffSmem& S, const ImageDesc& im, int img, unsigned ff_base,
                                unsigned long long off, unsigned nb, bool last_chunk, int t)
{
    const unsigned n_ends = (unsigned)im.n_tiles - 1u;          // none after the last tile
    if (t == 0) {
        S.first_mark = tile_ends_below(P, im, n_ends, off);
        S.n_marks = tile_ends_below(P, im, n_ends, off + nb) - S.first_mark;
    }
    cta_sync();
    for (unsigned i = (unsigned)t; i < S.n_marks; i += kThreads) S.mark[i] = (uint32_t)(tile_end_byte(P, im, S.first_mark + i) - off);
    if (t == 0) S.mark[S.n_marks] = 0xffffffffu;
    cta_sync();

    constexpr int NV = kChunkBytes / kThreads / 16;
    const unsigned b0 = (unsigned)t * (16u * NV);
    const unsigned mine = b0 < nb ? (nb - b0 < 16u * NV ? nb - b0 : 16u * NV) : 0u;
    unsigned w[4 * NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        uint4 v = {0u, 0u, 0u, 0u};
        if (b0 + 16u * q < nb) v = ldg_u128(im.raw + off + b0 + 16u * q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    unsigned cnt = 0;
#pragma unroll
    for (int q = 0; q < 4 * NV; ++q) {
        unsigned m = v_cmpeq4(w[q], 0xffffffffu);
        const int valid = (int)mine - 4 * q;                    // bytes of this word that exist
        if (valid <= 0) m = 0; else if (valid < 4) m &= 0xffffffffu >> (8 * (4 - valid));
        cnt += (unsigned)i_popc(m) >> 3;
    }
    unsigned mi;                          // first tile end at or after my first byte
    {
        unsigned lo = 0, hi = S.n_marks;
        while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (S.mark[mid] < b0) lo = mid + 1; else hi = mid; }
        mi = lo;
        unsigned k = mi;
        while (S.mark[k] < b0 + mine) ++k;                     // sentinel-terminated
        cnt += 2u * (k - mi);                                   // two marker bytes per tile end in my range
    }
    unsigned ff_chunk;                    // extra bytes of the chunk: stuffed zeros + markers
    const unsigned ff_ex = cta_scan_excl(cnt, S.warp_tmp, ff_chunk);
    {
        unsigned o = b0 + ff_ex;
#pragma unroll 4
        for (int j = 0; j < 16 * NV; ++j) {
            if ((unsigned)j < mine) {
                if (S.mark[mi] == b0 + (unsigned)j) {              // a restart interval ended in front of this byte
                    S.sbuf[o++] = 0xff;
                    S.sbuf[o++] = (uint8_t)(0xd0u + ((S.first_mark + mi) & 7u));
                    ++mi;
                }
                const unsigned byte = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                S.sbuf[o++] = (uint8_t)byte;
                if (byte == 0xffu) S.sbuf[o++] = 0;
            }
        }
    }
    cta_sync();

    const unsigned long long pos = off + ff_base;
    const unsigned out_bytes = nb + ff_chunk;
    const bool fits = pos + out_bytes + (last_chunk ? 2u : 0u) <= im.out_cap;
    if (fits) {
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
            const unsigned wi = o >> 2, sh = (o & 3u) * 8u;
            uint4 r;
            if (sh == 0) { r.x = sw[wi]; r.y = sw[wi + 1]; r.z = sw[wi + 2]; r.w = sw[wi + 3]; }
            else {
                const unsigned a0 = sw[wi], a1 = sw[wi + 1], a2 = sw[wi + 2], a3 = sw[wi + 3], a4 = sw[wi + 4];
                r.x = (a0 >> sh) | (a1 << (32u - sh)); r.y = (a1 >> sh) | (a2 << (32u - sh));
                r.z = (a2 >> sh) | (a3 << (32u - sh)); r.w = (a3 >> sh) | (a4 << (32u - sh));
            }
            dst4[i] = r;
        }
        if (tail0 + (unsigned)t < out_bytes) dst[tail0 + t] = S.sbuf[tail0 + t];
    }
    if (t == 0) {
        if (!fits) gmem_atomic_or(P.img_status + img, 1u);
        if (last_chunk) {
            if (fits) { im.out[pos + out_bytes] = 0xFF; im.out[pos + out_bytes + 1] = 0xD9; }   // EOI
            P.scan_bytes[img] = pos + out_bytes + 2;
        }
    }
    return ff_chunk;
}

JG_KERNEL(kThreads, 8)
void stuff_kernel(const JG_GRID_CONSTANT LaunchParams P)
{
    JG_DYNAMIC_SMEM(smem_raw);
    StuffSmem& S = *reinterpret_cast<StuffSmem*>(smem_raw);
    const int t = JG_TID;
    const unsigned n_chunks = P.first_chunk[P.n_images];
    if (ld_flag32(P.error) != 0u) return;             // pass 1 gave up (a look-back timed out): nothing to stuff
    // a contiguous run of chunks per CTA: the image, its sizes and the running count of extra bytes stay in registers
    // from one chunk to the next; only a new image costs the chain of dependent loads
    const unsigned per = (n_chunks + (unsigned)JG_GRID_DIM - 1u) / (unsigned)JG_GRID_DIM;
    const unsigned c_begin = (unsigned)JG_CTA_ID * per, c_end = c_begin + per < n_chunks ? c_begin + per : n_chunks;
    int img = 0;
    unsigned first = 0, next_first = 0, ff_base = 0;
    unsigned long long raw_n = 0;
    ImageDesc im;
    for (unsigned c = c_begin; c < c_end; ++c) {
        if (c == c_begin || c >= next_first) {
            const unsigned long long d = P.desc_ff[c];
            img = (int)(d >> 32);
            first = P.first_chunk[img];
            next_first = P.first_chunk[img + 1];
            im = P.images[img];
            raw_n = P.raw_bytes[img];
            ff_base = extra_before(P, c, d) - extra_before(P, first, P.desc_ff[first]);
        }
        const unsigned long long off = (unsigned long long)(c - first) * kChunkBytes;
        const unsigned nb = raw_n - off < (unsigned long long)kChunkBytes ? (unsigned)(raw_n - off) : (unsigned)kChunkBytes;
        const bool last_chunk = off + nb == raw_n;
        if ((im.flags & kFlagRestart) != 0 && im.n_tiles > 1) ff_base += stuff_restart_chunk(P, S, im, img, ff_base, off, nb, last_chunk, t);
        else ff_base += stuff_plain_chunk(P, S, im, img, ff_base, off, nb, last_chunk, t);
        cta_sync();                                   // sbuf is read until here
    }
}

}  // namespace jg
