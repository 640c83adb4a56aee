// jpeg_stuff.cuh -- second pass: 0xFF00 byte stuffing (jpeg_enc.h:634-638) + EOI (:1166-1167).
//
// The encode kernel leaves every image's entropy-coded bits unstuffed and byte-aligned in
// `raw`.  Stuffing moves every byte by the number of 0xFF bytes before it, so it is a scan
// over the whole image; doing it as its own streaming pass keeps that chain away from the
// heavy kernel (where tiles take ~100 us and differ) and puts it where a step is a uniform
// 4 KB copy.
//
//   plan_chunks_kernel  one CTA: first_chunk[i] = sum over images j < i of ceil(raw_bytes[j] / 4096)
//   stuff_kernel        persistent CTAs draw 4 KB chunks from a ticket: 16 bytes per thread,
//                       count 0xFF (__vcmpeq4), CTA scan, publish the chunk's count, emit the
//                       stuffed bytes into shared memory, THEN decoupled look-back over the
//                       chunks of the image for the byte offset, aligned 16-byte copy-out.
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
    int chunk;
    int abort;
    unsigned n_marks, first_mark;      // tile ends inside the chunk / index of the first one in the image
    unsigned long long ff_base;
};

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
    StuffSmem& S = *reinterpret_cast<StuffSmem*>(smem_raw);
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

JG_KERNEL(kThreads, 8)
void stuff_kernel(const JG_GRID_CONSTANT LaunchParams P)
{
    JG_DYNAMIC_SMEM(smem_raw);
    StuffSmem& S = *reinterpret_cast<StuffSmem*>(smem_raw);
    const int t = JG_TID;
    const int n_chunks = (int)P.first_chunk[P.n_images];
    for (;;) {
        cta_sync();
        if (t == 0) {
            S.chunk = (int)gmem_atomic_add(P.ticket2, 1u);
            S.abort = ld_flag32(P.error) != 0u;
        }
        cta_sync();
        const int c = S.chunk;
        if (c >= n_chunks || S.abort) break;

        // chunk -> image: last image whose first chunk is <= c (images without chunks are skipped)
        int lo = 0, hi = P.n_images - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int)P.first_chunk[mid] <= c) lo = mid; else hi = mid - 1;
        }
        const int img = lo;
        const ImageDesc im = P.images[img];
        const int first = (int)P.first_chunk[img];
        const unsigned long long raw_n = P.raw_bytes[img];
        const unsigned long long off = (unsigned long long)(c - first) * kChunkBytes;
        const unsigned nb = raw_n - off < (unsigned long long)kChunkBytes ? (unsigned)(raw_n - off) : (unsigned)kChunkBytes;
        const bool last_chunk = off + nb == raw_n;

        // restart mode: the tile ends e_j with off <= e_j < off + nb get a marker in front of raw byte e_j
        const bool rst = (im.flags & kFlagRestart) != 0 && im.n_tiles > 1;
        if (rst) {
            const unsigned n_ends = (unsigned)im.n_tiles - 1u;          // none after the last tile
            if (t == 0) {
                S.first_mark = tile_ends_below(P, im, n_ends, off);
                S.n_marks = tile_ends_below(P, im, n_ends, off + nb) - S.first_mark;
            }
            cta_sync();
            for (unsigned i = (unsigned)t; i < S.n_marks; i += kThreads) S.mark[i] = (uint32_t)(tile_end_byte(P, im, S.first_mark + i) - off);
            if (t == 0) S.mark[S.n_marks] = 0xffffffffu;
            cta_sync();
        }

        // kStuffPerThread contiguous bytes per thread, as 16-byte vectors (raw is 256-byte aligned and
        // chunk offsets are multiples of the chunk size: always legal uint4 loads)
        constexpr int NV = kChunkBytes / kThreads / 16;
        const unsigned b0 = (unsigned)t * (16u * NV);
        const unsigned mine = b0 < nb ? (nb - b0 < 16u * NV ? nb - b0 : 16u * NV) : 0u;
        unsigned w[4 * NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            uint4 v = {0u, 0u, 0u, 0u};
            if (b0 + 16u * q < nb) v = *reinterpret_cast<const uint4*>(im.raw + off + b0 + 16u * q);
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
        unsigned mi = 0;                      // first tile end at or after my first byte
        if (rst) {
            unsigned lo = 0, hi = S.n_marks;
            while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (S.mark[mid] < b0) lo = mid + 1; else hi = mid; }
            mi = lo;
            unsigned k = mi;
            while (S.mark[k] < b0 + mine) ++k;                     // sentinel-terminated
            cnt += 2u * (k - mi);                                   // two marker bytes per tile end in my range
        }
        unsigned ff_chunk;                    // extra bytes of the chunk: stuffed zeros (+ markers)
        const unsigned ff_ex = cta_scan_excl(cnt, S.warp_tmp, ff_chunk);
        const bool first_chunk_of_img = c == first;
        if (t == 0) st_flag64(P.desc_ff + c, (first_chunk_of_img ? kStatusPrefix : kStatusAgg) | (unsigned long long)ff_chunk);

        // emit (position independent) while the predecessors publish
        if (!rst) {
            unsigned o = b0 + ff_ex;
#pragma unroll
            for (int j = 0; j < 16 * NV; ++j) {
                if ((unsigned)j < mine) {
                    const unsigned byte = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                    S.sbuf[o++] = (uint8_t)byte;
                    if (byte == 0xffu) S.sbuf[o++] = 0;               // jpeg_enc.h:634-638
                }
            }
        } else {
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
        if (t < 32) {
            unsigned long long excl = 0;
            int timed_out = 0;
            if (!first_chunk_of_img) {
                excl = lookback(P.desc_ff, c, first, P.error, &timed_out);
                if (t == 0 && !timed_out) st_flag64(P.desc_ff + c, kStatusPrefix | (excl + ff_chunk));
            }
            if (t == 0) {
                S.ff_base = excl;
                S.abort = timed_out;
                if (timed_out) gmem_atomic_or(P.error, 2u);
            }
        }
        cta_sync();   // also orders the sbuf writes
        if (S.abort) break;

        const unsigned long long pos = off + S.ff_base;
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
    }
}

}  // namespace jg
