// jpeg_launch.h -- what the host side (jpeg_gpu_api.cu) and the kernel agree on: tile
// constants, the per-image record, the launch parameter block, and the per-specialisation
// launchers defined by jpeg_kernel_inst.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if !defined(JG_EMULATE)
#include <cuda_runtime.h>
#endif

#include "jpeg_tables.h"

namespace jg {

constexpr int kThreads = 128;
constexpr int kBlocksPerTile = 24;           // blocks of a (full) tile; one warp owns a tile from pixels to bytes
constexpr int kWarps = kThreads / 32;
constexpr int kChunkBytes = 64 * kThreads;   // unstuffed bytes one stuffing step handles (64 per thread)

// MCUs per tile: 24 blocks = 8 MCUs (4:4:4), 4 MCUs (4:2:0), 24 MCUs (gray)
constexpr int mcus_per_tile(int layout) { return layout == LAYOUT_444 ? 8 : (layout == LAYOUT_420 ? 4 : 24); }
constexpr int kWinWordsMax = 384;    // a warp's region (unstuffed bits of one tile): 512 bits per block before the tile goes slow
constexpr int kWinWordsMin = 216;    // must hold four worst-case blocks (4 x 1658 bits) + slack: the slow path's group
constexpr unsigned kSpinLimit = 1u << 24;
#ifndef JG_DEEP_MAX_IMAGES
#define JG_DEEP_MAX_IMAGES 48
#endif
constexpr int kDeepMaxImages = JG_DEEP_MAX_IMAGES;   // launches with fewer images use the two-iteration pipeline (DEEP kernels)

constexpr unsigned long long kStatusAgg = 1ull << 62;
constexpr unsigned long long kStatusPrefix = 2ull << 62;
constexpr int kFlagSwapRB = 1, kFlagRestart = 2;   // ImageDesc.flags == jpeg_gpu_image.flags
constexpr int kTailShift = 55;
constexpr unsigned long long kCountMask = (1ull << kTailShift) - 1;  // descriptors: status[63:62] | payload[61:55] | value[54:0]

struct ImageDesc {
    const uint8_t* px;               // device pixels
    uint8_t* raw;                    // device scratch: the image's entropy-coded bits BEFORE 0xFF00 stuffing
    uint8_t* out;                    // device destination of the entropy-coded segment (+EOI)
    unsigned long long raw_cap;      // bytes available at `raw`
    unsigned long long out_cap;      // bytes available at `out`
    unsigned long long first_block;  // index of the image's first block in the debug dumps
    int w, h;
    int stride;      // bytes from one pixel row to the next (negative: bottom-up storage, px = top row)
    int mcus_x;      // MCUs per MCU row
    int n_mcus;
    int first_tile;  // launch-local index of the image's first tile
    int n_tiles;
    int align;       // largest power of two (<= 16) dividing both px and stride: widest legal vector load
    int flags;       // JPEG_GPU_FLAG_* (bit 0: exchange channels 0 and 2 on load; bit 1: every tile is a restart interval)
};

// One parameter block for the three kernels of a launch group:
//   encode_tiles_kernel  pixels -> unstuffed bits (raw), bits chained over tiles
//   plan_chunks_kernel   raw sizes -> chunk table of the stuffing pass
//   stuff_kernel         raw -> final scan with 0xFF00 stuffing + EOI, bytes chained over chunks
struct LaunchParams {
    const ImageDesc* images;
    int n_images;
    int n_tiles;
    int tiles_per_image;             // > 0 when every image of the launch has this many tiles
    const uint32_t* sched;           // otherwise: the ticket schedule (build_schedule, jpeg_tables.h)
    int win_words;                   // region words actually used (kWinWordsMin..kWinWordsMax)
    unsigned* ticket;                // zeroed before the launch (encode)
    unsigned* error;                 // OUT: non-zero if a look-back timed out
    unsigned long long* desc_bits;   // [n_tiles], zeroed: status | the tile's last 7 bits | bits of the tile -> inclusive bit prefix
    unsigned* desc_dc;               // [3 * n_tiles], zeroed: valid<<31 | quantised DC of the tile's last block per component
    unsigned long long* desc_ff;     // [max_chunks] chunk -> image << 32 | extra bytes (stuffed zeros, markers) of the chunks before it in its group of 32
    unsigned* ff_groups;             // [max_chunks / 32 + 1] extra bytes of a group of chunks -> exclusive prefix over the launch
    unsigned long long* raw_bytes;   // [n_images] bytes of unstuffed scan (encode -> plan/stuff)
    unsigned* first_chunk;           // [n_images + 1] chunk table (plan -> stuff)
    unsigned long long* scan_bytes;  // [n_images] OUT: bytes of scan + EOI
    unsigned* img_status;            // [n_images] OUT: bit0 = capacity exceeded
    const HuffLut* huff;
    int16_t* dbg_coefs;              // optional [blocks*64], zigzag order
    uint32_t* dbg_bits;              // optional [blocks]
    // split pipeline (jpeg_transform.cuh -> jpeg_entropy.cuh): the coefficient plane between pass A and pass B
    const int16_t* coefs;            // [blocks of the plan][64] int16, zigzag order; an image starts at ImageDesc.first_block
    int bpm;                         // blocks per MCU of the launch: 1 gray, 3 4:4:4, 6 4:2:0
    int blocks_per_tile;             // pass B tile: 32 blocks (24 = whole MCUs with restart intervals)
};

// pass B (jpeg_entropy.cuh): tiles, CTA shape, and the TMA descriptor of the coefficient plane
constexpr int kEntTileBlocks = 32, kEntTileBlocksRestart = 24;
#ifndef JG_ENT_WARPS
#define JG_ENT_WARPS 20
#endif
#ifndef JG_ENT_MINB
#define JG_ENT_MINB 1
#endif
constexpr int kEntWarps = JG_ENT_WARPS, kEntThreads = 32 * kEntWarps;      // ONE CTA of 20 warps per SM: 20 x 10.9 KB + the tables once (three CTAs of 6 warps: 18 warps; measured 7 % slower)
// A CUtensorMap (cuTensorMapEncodeTiled, filled in by the host: 2-D, int16, {64, blocks} with a {72, 32} box), passed
// by value as a __grid_constant__ kernel parameter.  Under the CPU emulation: q[0] = base pointer, q[1] = rows.
struct alignas(64) CoefMap { unsigned long long q[16]; };

// pass A of the split pipeline (jpeg_transform.cuh): one warp per item = transform_item_mcus(layout) consecutive MCUs of one image
struct TransformParams {
    const ImageDesc* images;
    int n_images;
    int n_items;                 // work items of the launch
    int items_per_image;         // > 0 when every image of the launch has this many items
    const uint32_t* first_item;  // otherwise [n_images + 1]: first item of every image
    int16_t* coefs;              // [blocks of the plan][64], zigzag order
};
constexpr int transform_item_mcus(int layout) { return layout == LAYOUT_444 ? 32 : (layout == LAYOUT_420 ? 16 : 64); }

#if !defined(JG_EMULATE)
// one set per (layout, channels) specialisation; see jpeg_kernel_inst.cu
#define JG_DECLARE_SPEC(L, N)                                                                   \
    size_t smem_bytes_##L##_##N();                                                              \
    cudaError_t prepare_##L##_##N(int* ctas_per_sm);                                            \
    cudaError_t launch_##L##_##N(int grid, cudaStream_t stream, const LaunchParams& P, const QuantSet& Q, int mode);      \
    cudaError_t transform_prepare_##L##_##N();                                                                            \
    cudaError_t transform_launch_##L##_##N(cudaStream_t stream, const struct TransformParams& P, const QuantSet& Q);
JG_DECLARE_SPEC(0, 3)
JG_DECLARE_SPEC(0, 4)
JG_DECLARE_SPEC(1, 3)
JG_DECLARE_SPEC(1, 4)
JG_DECLARE_SPEC(2, 1)
#undef JG_DECLARE_SPEC
// layout-independent pass B of the split pipeline (jpeg_entropy.cu); mode 0 plain, 2 restart intervals
cudaError_t entropy_prepare(int* ctas_per_sm);
cudaError_t entropy_launch(int grid, cudaStream_t stream, const LaunchParams& P, const CoefMap& cmap, bool restart);
// layout-independent second pass (jpeg_stuff.cu)
size_t stuff_smem_bytes();
cudaError_t stuff_prepare(int* ctas_per_sm);
cudaError_t stuff_launch(int grid, cudaStream_t stream, const LaunchParams& P);
#endif

}  // namespace jg
