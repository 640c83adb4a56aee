// jpeg_gpu_api.cpp -- host side of the C ABI declared in include/jpeg_gpu.h.
//
// What stays on the host (as in the reference): argument validation (jpeg_enc.h:954-960,
// :1223-1226), quantiser / Huffman table construction (:1230-1266, :962-987) and marker
// emission (:989-1077).  What goes to the GPU: the block loop (:1094-1172), as ONE kernel
// launch per (layout, channels, quantiser) group of the batch -- see jpeg_kernel.cuh.
//
// There is no CPU encode path in this library: if CUDA is unusable every entry point fails.
#include "jpeg_gpu.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <chrono>
#include <thread>
#include <tuple>
#include <vector>

#include "jpeg_launch.h"
#include "jpeg_tables.h"

namespace {

using namespace jg;

thread_local std::string g_last_error;

void set_error(const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

#define JG_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return false;                                                                   \
        }                                                                                   \
    } while (0)

// ---- kernel specialisations -------------------------------------------------------------
// the group totals of the stuffing pass's two-level scan (jpeg_stuff.cuh), padded to keep what follows 16-byte aligned
size_t ff_groups_bytes(size_t chunks) { return ((chunks / 32 + 2) * 4 + 15) / 16 * 16; }

struct Spec {
    int layout, nc;
    cudaError_t (*prepare)(int*);
    cudaError_t (*launch)(int, cudaStream_t, const LaunchParams&, const QuantSet&, int);
    cudaError_t (*transform_prepare)();
    cudaError_t (*transform_launch)(cudaStream_t, const TransformParams&, const QuantSet&);
};
const Spec kSpecs[5] = {
    {LAYOUT_444, 3, prepare_0_3, launch_0_3, transform_prepare_0_3, transform_launch_0_3},
    {LAYOUT_444, 4, prepare_0_4, launch_0_4, transform_prepare_0_4, transform_launch_0_4},
    {LAYOUT_420, 3, prepare_1_3, launch_1_3, transform_prepare_1_3, transform_launch_1_3},
    {LAYOUT_420, 4, prepare_1_4, launch_1_4, transform_prepare_1_4, transform_launch_1_4},
    {LAYOUT_GRAY, 1, prepare_2_1, launch_2_1, transform_prepare_2_1, transform_launch_2_1}};

// Which pipeline new plans use.  Split (default): transform kernel -> coefficient plane in HBM -> entropy kernel
// (jpeg_transform.cuh, jpeg_entropy.cuh).  Fused: the single pass-1 kernel of round 1 (jpeg_kernel.cuh), kept for the
// A/B numbers in DESIGN.md; selected with JPEG_GPU_PIPELINE=fused in the environment.  Same bytes either way.
bool use_fused_pipeline()
{
    const char* e = getenv("JPEG_GPU_PIPELINE");
    return e && strcmp(e, "fused") == 0;
}

// cuTensorMapEncodeTiled comes from the driver (the library links the static runtime only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

int spec_index(int layout, int nc)
{
    for (int i = 0; i < 5; ++i)
        if (kSpecs[i].layout == layout && kSpecs[i].nc == nc) return i;
    return -1;
}

// ---- devices ----------------------------------------------------------------------------
struct Device {
    int id = -1;
    int sm_count = 0;
    int ctas_per_sm[5] = {0, 0, 0, 0, 0};
    int stuff_ctas_per_sm = 0;
    int entropy_ctas_per_sm = 0;
    HuffLut* d_huff = nullptr;
    cudaStream_t stream = nullptr;   // used when the caller gives none
    cudaStream_t pipe[4] = {nullptr, nullptr, nullptr, nullptr};   // chunk pipeline of jpeg_gpu_encode_batch
    // Caching allocator: plans come and go per call, cudaMalloc / cudaMallocHost must not.
    std::mutex* pool_mutex = nullptr;
    std::multimap<size_t, void*>* pool_dev = nullptr;    // free device blocks by size
    std::multimap<size_t, void*>* pool_host = nullptr;   // free pinned host blocks by size
    size_t pooled_bytes = 0, pooled_host_bytes = 0;
};

std::mutex g_mutex;
std::vector<Device> g_devices;

bool init_device(Device& d, int id)
{
    d.id = id;
    JG_CUDA(cudaSetDevice(id));
    cudaDeviceProp prop;
    JG_CUDA(cudaGetDeviceProperties(&prop, id));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only", id, prop.major, prop.minor);
        return false;
    }
    d.sm_count = prop.multiProcessorCount;
    for (int i = 0; i < 5; ++i) {
        JG_CUDA(kSpecs[i].prepare(&d.ctas_per_sm[i]));
        if (d.ctas_per_sm[i] < 1) { set_error("kernel spec %d does not fit an SM", i); return false; }
    }
    for (int i = 0; i < 5; ++i) JG_CUDA(kSpecs[i].transform_prepare());
    JG_CUDA(entropy_prepare(&d.entropy_ctas_per_sm));
    if (d.entropy_ctas_per_sm < 1) { set_error("entropy kernel does not fit an SM"); return false; }
    if (!g_encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        JG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("the driver has no cuTensorMapEncodeTiled"); return false; }
        g_encode_tiled = (EncodeTiledFn)fn;
    }
    JG_CUDA(stuff_prepare(&d.stuff_ctas_per_sm));
    if (d.stuff_ctas_per_sm < 1) { set_error("stuffing kernel does not fit an SM"); return false; }
    HuffLut lut;
    build_huff_lut(&lut);
    JG_CUDA(cudaMalloc(&d.d_huff, sizeof(HuffLut)));
    JG_CUDA(cudaMemcpy(d.d_huff, &lut, sizeof(HuffLut), cudaMemcpyHostToDevice));
    JG_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    for (auto& ps : d.pipe) JG_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    d.pool_mutex = new std::mutex();
    d.pool_dev = new std::multimap<size_t, void*>();
    d.pool_host = new std::multimap<size_t, void*>();
    return true;
}

// size classes: powers of two from 64 KB, so that a freed block fits the next request of its kind
size_t pool_class(size_t bytes)
{
    size_t c = 64 * 1024;
    while (c < bytes) c <<= 1;
    return c;
}

// Idle memory the caching allocator keeps per GPU: device blocks up to JPEG_GPU_POOL_LIMIT_MB (default 24 GB: a config-5 sized
// plan comes back without a cudaMalloc), pinned host blocks up to JPEG_GPU_PINNED_POOL_LIMIT_MB (default 1 GB); beyond that
// blocks go back to the driver.  (Blocks are recycled only after the freeing plan's stream has drained: plan_free.)
size_t pool_limit(bool host)
{
    static const size_t dev_limit = [] { const char* e = getenv("JPEG_GPU_POOL_LIMIT_MB"); return (e ? (size_t)strtoull(e, nullptr, 10) : (size_t)24576) << 20; }();
    static const size_t host_limit = [] { const char* e = getenv("JPEG_GPU_PINNED_POOL_LIMIT_MB"); return (e ? (size_t)strtoull(e, nullptr, 10) : (size_t)1024) << 20; }();
    return host ? host_limit : dev_limit;
}

bool pool_alloc(Device& d, size_t bytes, bool host, void** out)
{
    const size_t c = pool_class(bytes);
    {
        std::lock_guard<std::mutex> lk(*d.pool_mutex);
        auto& m = host ? *d.pool_host : *d.pool_dev;
        auto f = m.find(c);
        if (f != m.end()) {
            *out = f->second;
            m.erase(f);
            (host ? d.pooled_host_bytes : d.pooled_bytes) -= c;
            return true;
        }
    }
    cudaError_t e = host ? cudaMallocHost(out, c) : cudaMalloc(out, c);
    if (e != cudaSuccess && !host) {
        // out of memory: drop the cache and try once more
        std::lock_guard<std::mutex> lk(*d.pool_mutex);
        for (auto& kv : *d.pool_dev) cudaFree(kv.second);
        d.pool_dev->clear();
        d.pooled_bytes = 0;
        (void)cudaGetLastError();
        e = cudaMalloc(out, c);
    }
    if (e != cudaSuccess) {
        set_error("%s of %zu bytes failed: %s", host ? "cudaMallocHost" : "cudaMalloc", c, cudaGetErrorString(e));
        *out = nullptr;
        return false;
    }
    return true;
}

void pool_free(Device& d, void* p, size_t bytes, bool host)
{
    if (!p) return;
    const size_t c = pool_class(bytes);
    std::lock_guard<std::mutex> lk(*d.pool_mutex);
    size_t& pooled = host ? d.pooled_host_bytes : d.pooled_bytes;
    if (pooled + c > pool_limit(host)) { if (host) cudaFreeHost(p); else cudaFree(p); return; }
    (host ? *d.pool_host : *d.pool_dev).emplace(c, p);
    pooled += c;
}

bool ensure_init()
{
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        if (!g_devices.empty()) return true;
    }
    return jpeg_gpu_init(nullptr, 0) > 0;
}

// ---- geometry ---------------------------------------------------------------------------
struct Geometry {
    int layout, nc_in, ncomp_out, mcu, bpm, mcus_x, mcus_y, n_mcus, n_tiles;
    size_t n_blocks;
};

// tiles of pass 1: 24 blocks = whole MCUs (fused kernel; split pipeline with restart intervals) or 32 blocks (split)
int tiles_of(const Geometry& g, bool fused, bool restart)
{
    if (fused) { const int M = mcus_per_tile(g.layout); return (g.n_mcus + M - 1) / M; }
    const size_t bpt = restart ? kEntTileBlocksRestart : kEntTileBlocks;
    return (int)((g.n_blocks + bpt - 1) / bpt);
}

bool geometry_of(const jpeg_gpu_image& im, Geometry* g)
{
    // jpeg_enc.h:954-960 (+ the extended 1-channel input)
    if (im.ncomp != 1 && im.ncomp != 3 && im.ncomp != 4) return false;
    if (im.width <= 0 || im.height <= 0 || im.width > 0xffff || im.height > 0xffff) return false;
    if (im.subsampling != JPEG_GPU_SUB_444 && im.subsampling != JPEG_GPU_SUB_420) return false;
    if (im.ncomp == 1 && im.subsampling != JPEG_GPU_SUB_444) return false;
    if (im.stride != 0 && std::abs(im.stride) < im.width * im.ncomp) return false;   // negative: bottom-up rows
    if (im.flags & ~(JPEG_GPU_FLAG_SWAP_RB | JPEG_GPU_FLAG_RESTART)) return false;
    g->layout = im.ncomp == 1 ? LAYOUT_GRAY : (im.subsampling == JPEG_GPU_SUB_420 ? LAYOUT_420 : LAYOUT_444);
    g->nc_in = im.ncomp;
    g->ncomp_out = im.ncomp == 1 ? 1 : 3;
    g->mcu = g->layout == LAYOUT_420 ? 16 : 8;
    g->bpm = g->layout == LAYOUT_444 ? 3 : (g->layout == LAYOUT_420 ? 6 : 1);
    g->mcus_x = (im.width + g->mcu - 1) / g->mcu;
    g->mcus_y = (im.height + g->mcu - 1) / g->mcu;
    g->n_mcus = g->mcus_x * g->mcus_y;
    const int M = mcus_per_tile(g->layout);
    g->n_tiles = (g->n_mcus + M - 1) / M;
    g->n_blocks = (size_t)g->n_mcus * g->bpm;
    return true;
}

// worst case of one block: DC 11+11 bits, 63 x (16+10) AC bits = 1660 bits -> 208 bytes,
// doubled if every byte were 0xFF; + EOI
size_t worst_scan_bytes(size_t n_blocks) { return n_blocks * 416 + 16; }

// what the arena reserves per image: generous for real content (uniform noise at all-ones
// quantisers needs ~4.1 B/px colour, ~1.4 B/px gray); anything beyond that is re-encoded
// with the worst-case bound by jpeg_gpu_encode_batch.
size_t default_scan_bytes(const jpeg_gpu_image& im, const Geometry& g)
{
    const size_t px = (size_t)im.width * im.height;
    const size_t guess = (g.ncomp_out == 3 ? 6 * px : 5 * px / 2) + 4096;
    return std::min(guess, worst_scan_bytes(g.n_blocks));
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bottom-up storage (negative stride): byte offset of the top row inside the image's memory block
size_t top_row_offset(const jpeg_gpu_image& im)
{
    return im.stride < 0 ? (size_t)(-(long long)im.stride) * (size_t)(im.height - 1) : 0;
}

// largest power of two <= 16 dividing both the base address and the row pitch
int alignment_of(const void* px, int stride)
{
    const size_t v = (size_t)px | (size_t)stride | 16u;
    return (int)(v & (~v + 1));
}

}  // namespace

// =========================================================================================
// plan
// =========================================================================================
struct jpeg_gpu_plan {
    struct Item {
        jpeg_gpu_image img;
        Geometry geo;
        bool valid = false;
        int group = -1, index_in_group = -1;
        std::vector<uint8_t> header;
        size_t scan_cap = 0, arena_off = 0;
        size_t pixel_off = 0, pixel_bytes = 0;   // in the plan-owned upload buffer
        const uint8_t* d_pixels = nullptr;
        size_t first_block = 0;
    };
    struct Group {
        int spec = -1;
        QuantSet quant;
        std::vector<int> items;
        int n_tiles = 0, tiles_per_image = 0;
        size_t state_off = 0;     // offset of {ticket, error, pad, desc_bits[], desc_ff[], ff_groups[], desc_dc[]} in d_state
        size_t max_chunks = 0;    // upper bound of stuffing chunks (from the capacities)
        unsigned long long* d_raw_bytes = nullptr;    // into d_aux
        unsigned* d_first_chunk = nullptr;
        ImageDesc* d_images = nullptr;
        bool restart = false;     // JPEG_GPU_FLAG_RESTART images (their own kernel instantiation)
        size_t sched_off = 0;     // words into d_sched (groups with images of different tile counts)
        bool has_sched = false;
        unsigned long long* d_scan_bytes = nullptr;   // into d_results
        unsigned* d_status = nullptr;
        size_t result_off = 0;    // index of first image of the group in the results arrays
        int n_items = 0, items_per_image = 0;     // pass A work items (split pipeline)
        size_t item_off = 0;      // words into d_sched: first_item[n + 1] when the images differ
        bool has_items = false;
    };
    bool fused = false;           // pipeline of this plan (use_fused_pipeline() at creation)
    int16_t* d_coefs = nullptr;      size_t coefs_bytes = 0;    // split pipeline: the coefficient plane between pass A and pass B
    CoefMap cmap;                    // its TMA descriptor
    cudaStream_t run_stream = nullptr;   // stream of the last run / upload: results are fetched on it, the plan's memory is released after it
    bool ran = false;
    int dev_index = 0;
    int win_words = kWinWordsMax;
    bool worst_case = false;
    std::vector<Item> items;
    std::vector<Group> groups;
    uint8_t* d_arena = nullptr;      size_t arena_bytes = 0;
    uint8_t* d_raw = nullptr;        // unstuffed scans, same layout as the arena
    uint8_t* d_aux = nullptr;        size_t aux_bytes = 0;      // raw_bytes u64[n] + first_chunk u32[n + groups]
    size_t images_bytes = 0;
    uint8_t* d_pixels = nullptr;     size_t pixels_bytes = 0;
    uint8_t* d_state = nullptr;      size_t state_bytes = 0;    // zeroed before every run
    uint8_t* d_results = nullptr;    size_t results_bytes = 0;  // scan_bytes[n] u64, status[n] u32
    uint8_t* h_results = nullptr;                               // pinned mirror
    ImageDesc* d_images = nullptr;
    std::vector<ImageDesc> h_images;
    uint32_t* d_sched = nullptr;     size_t sched_bytes = 0;    // ticket schedules (jpeg_tables.h build_schedule)
    std::vector<uint32_t> h_sched;
    bool images_dirty = true;
    int16_t* dbg_coefs = nullptr;
    uint32_t* dbg_bits = nullptr;
    size_t n_blocks = 0;
    int n_valid = 0;
    bool fetched_results = false;
    bool timing = false;
    std::vector<cudaEvent_t> events;   // per group: before pass A, after pass A, after pass B (fused: same as after A), after stuff
};

namespace {

bool plan_build(jpeg_gpu_plan* p, const jpeg_gpu_image* images, int n, bool worst_case)
{
    Device& dev = g_devices[p->dev_index];
    JG_CUDA(cudaSetDevice(dev.id));
    p->worst_case = worst_case;
    p->fused = use_fused_pipeline();
    p->items.resize(n);
    std::map<std::tuple<int, int, int, int>, int> group_of;   // (spec, qmode, quality, restart) -> group
    size_t arena = 0, pixels = 0, blocks = 0;
    for (int i = 0; i < n; ++i) {
        jpeg_gpu_plan::Item& it = p->items[i];
        it.img = images[i];
        uint8_t ql[64], qc[64];
        if (!geometry_of(it.img, &it.geo) || !build_qt(it.img.quality_mode, it.img.quality, ql, qc)) continue;
        it.geo.n_tiles = tiles_of(it.geo, p->fused, (it.img.flags & JPEG_GPU_FLAG_RESTART) != 0);
        it.valid = true;
        ++p->n_valid;
        if (it.img.stride == 0) it.img.stride = it.img.width * it.img.ncomp;
        it.header.resize(1024);
        const bool restart = (it.img.flags & JPEG_GPU_FLAG_RESTART) != 0;
        it.header.resize(emit_headers(it.img.width, it.img.height, it.geo.ncomp_out, it.img.subsampling, ql, qc,
                                      it.header.data(), it.header.size(), restart ? mcus_per_tile(it.geo.layout) : 0));
        const int spec = spec_index(it.geo.layout, it.geo.nc_in);
        auto key = std::make_tuple(spec, it.img.quality_mode, it.img.quality, restart ? 1 : 0);   // restart images launch on their own kernels
        auto f = group_of.find(key);
        if (f == group_of.end()) {
            jpeg_gpu_plan::Group g;
            g.spec = spec;
            g.restart = restart;
            build_pqt(ql, g.quant.luma);
            build_pqt(qc, g.quant.chroma);
            f = group_of.emplace(key, (int)p->groups.size()).first;
            p->groups.push_back(g);
        }
        jpeg_gpu_plan::Group& g = p->groups[f->second];
        it.group = f->second;
        it.index_in_group = (int)g.items.size();
        g.items.push_back(i);
        // restart mode: + 1 pad byte and 2 marker bytes per tile
        it.scan_cap = align_up((worst_case ? worst_scan_bytes(it.geo.n_blocks) : default_scan_bytes(it.img, it.geo)) +
                                   (restart ? 3 * (size_t)it.geo.n_tiles : 0), 256);
        it.arena_off = arena;
        arena += it.scan_cap;
        it.pixel_bytes = (size_t)std::abs(it.img.stride) * it.img.height;
        if (!it.img.pixels_on_device) {
            it.pixel_off = pixels;
            pixels += align_up(it.pixel_bytes, 256);
        }
        it.first_block = blocks;
        blocks += it.geo.n_blocks;
    }
    p->n_blocks = blocks;
    p->arena_bytes = arena;
    p->pixels_bytes = pixels;

    // device state: per group {ticket u32, error u32, pad} + desc_bits + desc_ff + ff_groups + desc_dc
    size_t state = 0, res_index = 0;
    for (auto& g : p->groups) {
        int tiles = 0;
        bool uniform = true;
        const int t0 = p->items[g.items[0]].geo.n_tiles;
        for (int idx : g.items) {
            tiles += p->items[idx].geo.n_tiles;
            uniform = uniform && p->items[idx].geo.n_tiles == t0;
        }
        g.n_tiles = tiles;
        g.tiles_per_image = uniform ? t0 : 0;
        if (!uniform) {
            std::vector<int> counts;
            for (int idx : g.items) counts.push_back(p->items[idx].geo.n_tiles);
            g.sched_off = p->h_sched.size();
            g.has_sched = true;
            p->h_sched.resize(g.sched_off + schedule_words((int)counts.size()));
            build_schedule(counts.data(), (int)counts.size(), p->h_sched.data() + g.sched_off);
        }
        if (!p->fused) {
            // pass A: one warp per item of transform_item_mcus(layout) MCUs
            const int im = transform_item_mcus(kSpecs[g.spec].layout);
            int items = 0;
            bool same = true;
            const int i0 = (p->items[g.items[0]].geo.n_mcus + im - 1) / im;
            for (int idx : g.items) {
                const int k = (p->items[idx].geo.n_mcus + im - 1) / im;
                items += k;
                same = same && k == i0;
            }
            g.n_items = items;
            g.items_per_image = same ? i0 : 0;
            if (!same) {
                g.item_off = p->h_sched.size();
                g.has_items = true;
                uint32_t run = 0;
                for (int idx : g.items) { p->h_sched.push_back(run); run += (uint32_t)((p->items[idx].geo.n_mcus + im - 1) / im); }
                p->h_sched.push_back(run);
            }
        }
        g.state_off = state;
        size_t chunks = 1;
        for (int idx : g.items) chunks += (p->items[idx].scan_cap + kChunkBytes - 1) / kChunkBytes;
        g.max_chunks = chunks;
        state += 16 + (size_t)tiles * 8 + chunks * 8 + ff_groups_bytes(chunks) + (((size_t)tiles * 12 + 15) / 16) * 16;
        g.result_off = res_index;
        res_index += g.items.size();
    }
    p->state_bytes = state;
    const size_t nres = res_index;
    p->results_bytes = nres * 8 + nres * 4;
    if (nres == 0) return true;

    p->arena_bytes = std::max<size_t>(arena, 256);
    p->aux_bytes = nres * 8 + (nres + p->groups.size()) * 4 + 16;
    p->images_bytes = nres * sizeof(ImageDesc);
    if (!pool_alloc(dev, p->arena_bytes, false, (void**)&p->d_arena)) return false;
    if (!pool_alloc(dev, p->arena_bytes, false, (void**)&p->d_raw)) return false;
    if (!pool_alloc(dev, p->aux_bytes, false, (void**)&p->d_aux)) return false;
    if (pixels && !pool_alloc(dev, pixels, false, (void**)&p->d_pixels)) return false;
    if (!pool_alloc(dev, state, false, (void**)&p->d_state)) return false;
    if (!pool_alloc(dev, p->results_bytes, false, (void**)&p->d_results)) return false;
    if (!pool_alloc(dev, p->results_bytes, true, (void**)&p->h_results)) return false;
    if (!pool_alloc(dev, p->images_bytes, false, (void**)&p->d_images)) return false;
    if (!p->fused) {
        p->coefs_bytes = std::max<size_t>(p->n_blocks, 1) * 128;
        if (!pool_alloc(dev, p->coefs_bytes, false, (void**)&p->d_coefs)) return false;
        // 2-D tensor {64 int16, blocks}; the box is 72 x 32 (one pass-B tile): its 8 out-of-bounds columns arrive as zeros
        // and pad every block to a 144-byte row in shared memory (jpeg_entropy.cuh)
        CUtensorMap tm;
        const cuuint64_t dims[2] = {64, (cuuint64_t)std::max<size_t>(p->n_blocks, 1)};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {72, 32}, estr[2] = {1, 1};
        const CUresult r = g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, p->d_coefs, dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
        static_assert(sizeof(CoefMap) == sizeof(CUtensorMap), "CoefMap carries a CUtensorMap");
        memcpy(&p->cmap, &tm, sizeof tm);
    }
    p->sched_bytes = p->h_sched.size() * sizeof(uint32_t);
    if (p->sched_bytes && !pool_alloc(dev, p->sched_bytes, false, (void**)&p->d_sched)) return false;
    p->h_images.resize(nres);
    for (auto& g : p->groups) {
        g.d_images = p->d_images + g.result_off;
        g.d_scan_bytes = reinterpret_cast<unsigned long long*>(p->d_results) + g.result_off;
        g.d_status = reinterpret_cast<unsigned*>(p->d_results + nres * 8) + g.result_off;
        g.d_raw_bytes = reinterpret_cast<unsigned long long*>(p->d_aux) + g.result_off;
        g.d_first_chunk = reinterpret_cast<unsigned*>(p->d_aux + nres * 8) + g.result_off + (&g - &p->groups[0]);
        int tile = 0;
        for (size_t k = 0; k < g.items.size(); ++k) {
            jpeg_gpu_plan::Item& it = p->items[g.items[k]];
            // `pixels` is the image's top row; with bottom-up storage that is the LAST row of the memory block
            if (!it.img.pixels_on_device) it.d_pixels = p->d_pixels + it.pixel_off + top_row_offset(it.img);
            else it.d_pixels = it.img.pixels;
            ImageDesc& d = p->h_images[g.result_off + k];
            d.px = it.d_pixels;
            d.raw = p->d_raw + it.arena_off;
            d.raw_cap = it.scan_cap;
            d.out = p->d_arena + it.arena_off;
            d.out_cap = it.scan_cap;
            d.first_block = it.first_block;
            d.w = it.img.width; d.h = it.img.height; d.stride = it.img.stride;
            d.mcus_x = it.geo.mcus_x; d.n_mcus = it.geo.n_mcus;
            d.first_tile = tile; d.n_tiles = it.geo.n_tiles;
            d.flags = it.img.flags;
            d.align = (d.flags & JPEG_GPU_FLAG_SWAP_RB) ? 1 : alignment_of(d.px, d.stride);   // the swizzle lives in the byte loader
            tile += it.geo.n_tiles;
        }
    }
    p->images_dirty = true;
    return true;
}

bool plan_sync_images(jpeg_gpu_plan* p, cudaStream_t s)
{
    if (!p->images_dirty || p->h_images.empty()) return true;
    // h_images is pageable: the copy is staged by the runtime before the call returns
    JG_CUDA(cudaMemcpyAsync(p->d_images, p->h_images.data(), p->h_images.size() * sizeof(ImageDesc),
                            cudaMemcpyHostToDevice, s));
    if (p->sched_bytes)
        JG_CUDA(cudaMemcpyAsync(p->d_sched, p->h_sched.data(), p->sched_bytes, cudaMemcpyHostToDevice, s));
    p->images_dirty = false;
    return true;
}

bool plan_run(jpeg_gpu_plan* p, cudaStream_t s)
{
    Device& dev = g_devices[p->dev_index];
    JG_CUDA(cudaSetDevice(dev.id));
    if (p->groups.empty()) return true;
    p->run_stream = s;
    p->ran = true;
    if (!plan_sync_images(p, s)) return false;
    JG_CUDA(cudaMemsetAsync(p->d_state, 0, p->state_bytes, s));
    JG_CUDA(cudaMemsetAsync(p->d_results, 0, p->results_bytes, s));
    for (auto& g : p->groups) {
        LaunchParams P;
        memset(&P, 0, sizeof P);
        P.images = g.d_images;
        P.n_images = (int)g.items.size();
        P.n_tiles = g.n_tiles;
        P.tiles_per_image = g.tiles_per_image;
        P.sched = g.has_sched ? p->d_sched + g.sched_off : nullptr;
        P.win_words = p->win_words;
        uint8_t* st = p->d_state + g.state_off;
        P.ticket = reinterpret_cast<unsigned*>(st);
        P.error = reinterpret_cast<unsigned*>(st + 4);
        P.desc_bits = reinterpret_cast<unsigned long long*>(st + 16);
        P.desc_ff = P.desc_bits + g.n_tiles;
        P.ff_groups = reinterpret_cast<unsigned*>(P.desc_ff + g.max_chunks);
        P.desc_dc = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(P.ff_groups) + ff_groups_bytes(g.max_chunks));
        P.raw_bytes = g.d_raw_bytes;
        P.first_chunk = g.d_first_chunk;
        P.scan_bytes = g.d_scan_bytes;
        P.img_status = g.d_status;
        P.huff = dev.d_huff;
        P.dbg_coefs = p->dbg_coefs;
        P.dbg_bits = p->dbg_bits;
        const size_t gi = (size_t)(&g - &p->groups[0]);
        if (p->timing) JG_CUDA(cudaEventRecord(p->events[4 * gi], s));
        if (p->fused) {
            const int grid = std::min((g.n_tiles + kWarps - 1) / kWarps, dev.sm_count * dev.ctas_per_sm[g.spec]);   // one tile per warp at a time
            const int mode = g.restart ? 2 : (P.n_images < kDeepMaxImages ? 1 : 0);   // kModeRestart / kModeDeep / kModePlain
            JG_CUDA(kSpecs[g.spec].launch(grid, s, P, g.quant, mode));
            if (p->timing) JG_CUDA(cudaEventRecord(p->events[4 * gi + 1], s));
        } else {
            // pass A: pixels -> coefficient plane; pass B: coefficient plane -> unstuffed bits
            TransformParams TP;
            TP.images = g.d_images;
            TP.n_images = P.n_images;
            TP.n_items = g.n_items;
            TP.items_per_image = g.items_per_image;
            TP.first_item = g.has_items ? p->d_sched + g.item_off : nullptr;
            TP.coefs = p->d_coefs;
            JG_CUDA(kSpecs[g.spec].transform_launch(s, TP, g.quant));
            if (p->timing) JG_CUDA(cudaEventRecord(p->events[4 * gi + 1], s));
            P.coefs = p->d_coefs;
            P.bpm = kSpecs[g.spec].layout == LAYOUT_444 ? 3 : (kSpecs[g.spec].layout == LAYOUT_420 ? 6 : 1);
            P.blocks_per_tile = g.restart ? kEntTileBlocksRestart : kEntTileBlocks;
            P.dbg_coefs = nullptr;
            // persistent CTAs of kEntWarps warps that draw tiles from a ticket; with few tiles still one CTA per SM, so that the
            // tiles spread over the SMs instead of filling a handful of them (warps that draw no ticket leave at once)
            const int grid = std::min(g.n_tiles, dev.sm_count * dev.entropy_ctas_per_sm);
            JG_CUDA(entropy_launch(grid, s, P, p->cmap, g.restart));
        }
        if (p->timing) JG_CUDA(cudaEventRecord(p->events[4 * gi + 2], s));
        JG_CUDA(stuff_launch(dev.sm_count * dev.stuff_ctas_per_sm, s, P));
        if (p->timing) JG_CUDA(cudaEventRecord(p->events[4 * gi + 3], s));
    }
    // stage dump of the split pipeline: the coefficient plane IS the dump
    if (!p->fused && p->dbg_coefs)
        JG_CUDA(cudaMemcpyAsync(p->dbg_coefs, p->d_coefs, p->n_blocks * 128, cudaMemcpyDeviceToDevice, s));
    p->fetched_results = false;
    return true;
}

// Bring scan sizes / status / error flags to the host.  The copies are ordered behind the kernels: they go to the
// stream of the last run (whatever stream the caller passes for the output copies), which is then synchronised.
bool plan_results(jpeg_gpu_plan* p, cudaStream_t caller)
{
    if (p->fetched_results || p->groups.empty()) return true;
    cudaStream_t s = p->ran ? p->run_stream : caller;
    JG_CUDA(cudaMemcpyAsync(p->h_results, p->d_results, p->results_bytes, cudaMemcpyDeviceToHost, s));
    std::vector<unsigned> errs(p->groups.size());
    for (size_t k = 0; k < p->groups.size(); ++k)
        JG_CUDA(cudaMemcpyAsync(&errs[k], p->d_state + p->groups[k].state_off + 4, 4, cudaMemcpyDeviceToHost, s));
    JG_CUDA(cudaStreamSynchronize(s));
    for (unsigned e : errs)
        if (e) { set_error("encode kernel aborted (look-back timeout, flag %u)", e); return false; }
    p->fetched_results = true;
    return true;
}

void plan_free(jpeg_gpu_plan* p)
{
    if (!p) return;
    if (p->dev_index < (int)g_devices.size()) {
        Device& dev = g_devices[p->dev_index];
        cudaSetDevice(dev.id);
        // The pool hands these blocks to the next plan at once (no cudaFree, so nothing waits for the device):
        // kernels or copies of this plan that are still in flight must finish first.
        if (p->ran) cudaStreamSynchronize(p->run_stream);
        if (p->d_coefs) pool_free(dev, p->d_coefs, p->coefs_bytes, false);
        pool_free(dev, p->d_arena, p->arena_bytes, false);
        pool_free(dev, p->d_raw, p->arena_bytes, false);
        pool_free(dev, p->d_aux, p->aux_bytes, false);
        pool_free(dev, p->d_pixels, p->pixels_bytes, false);
        pool_free(dev, p->d_state, p->state_bytes, false);
        pool_free(dev, p->d_results, p->results_bytes, false);
        pool_free(dev, p->d_images, p->images_bytes, false);
        if (p->d_sched) pool_free(dev, p->d_sched, p->sched_bytes, false);
        pool_free(dev, p->h_results, p->results_bytes, true);
    }
    for (cudaEvent_t e : p->events) cudaEventDestroy(e);
    delete p;
}

jpeg_gpu_plan* plan_create(const jpeg_gpu_image* images, int n, int device, int win_words, bool worst_case)
{
    if (!ensure_init()) return nullptr;
    if (n < 0 || (n > 0 && !images)) { set_error("bad image list"); return nullptr; }
    if (device < 0 || device >= (int)g_devices.size()) { set_error("device index %d out of range", device); return nullptr; }
    jpeg_gpu_plan* p = new jpeg_gpu_plan();
    memset(&p->cmap, 0, sizeof p->cmap);
    p->dev_index = device;
    if (win_words) p->win_words = std::max(kWinWordsMin, std::min(kWinWordsMax, win_words));
    if (!plan_build(p, images, n, worst_case)) { plan_free(p); return nullptr; }
    return p;
}

// per-image result after plan_results()
void item_result(jpeg_gpu_plan* p, int i, size_t* file_size, int* status)
{
    const jpeg_gpu_plan::Item& it = p->items[i];
    if (!it.valid) { *file_size = 0; *status = JPEG_GPU_ERR_ARG; return; }
    const jpeg_gpu_plan::Group& g = p->groups[it.group];
    const size_t nres = p->h_images.size();
    const unsigned long long scan = reinterpret_cast<unsigned long long*>(p->h_results)[g.result_off + it.index_in_group];
    const unsigned st = reinterpret_cast<unsigned*>(p->h_results + nres * 8)[g.result_off + it.index_in_group];
    *file_size = it.header.size() + (size_t)scan;
    *status = (st & 1u) ? JPEG_GPU_ERR_CAPACITY : JPEG_GPU_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
// hooks for jpeg_decode_api.cpp
namespace jg {
int cuda_device_of(int index)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    return index >= 0 && index < (int)g_devices.size() ? g_devices[index].id : 0;
}
void set_error_text(const char* text) { set_error("%s", text); }
}  // namespace jg

extern "C" {

int jpeg_gpu_init(const int* device_ids, int n_devices)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!g_devices.empty()) return (int)g_devices.size();
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0) {
        set_error("no usable CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 0;
    }
    std::vector<int> ids;
    if (device_ids && n_devices > 0) ids.assign(device_ids, device_ids + n_devices);
    else for (int i = 0; i < visible; ++i) ids.push_back(i);
    std::vector<Device> devs(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        if (ids[i] < 0 || ids[i] >= visible) { set_error("device id %d not visible", ids[i]); return 0; }
        if (!init_device(devs[i], ids[i])) return 0;
    }
    g_devices = devs;
    return (int)g_devices.size();
}

void jpeg_gpu_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    for (auto& d : g_devices) {
        cudaSetDevice(d.id);
        cudaFree(d.d_huff);
        if (d.stream) cudaStreamDestroy(d.stream);
        for (auto& ps : d.pipe) if (ps) cudaStreamDestroy(ps);
        if (d.pool_dev) { for (auto& kv : *d.pool_dev) cudaFree(kv.second); delete d.pool_dev; }
        if (d.pool_host) { for (auto& kv : *d.pool_host) cudaFreeHost(kv.second); delete d.pool_host; }
        delete d.pool_mutex;
    }
    g_devices.clear();
}

int jpeg_gpu_device_count(void)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    return (int)g_devices.size();
}

const char* jpeg_gpu_last_error(void) { return g_last_error.c_str(); }

size_t jpeg_gpu_max_encoded_size(int width, int height, int ncomp, int subsampling)
{
    jpeg_gpu_image im;
    memset(&im, 0, sizeof im);
    im.width = width; im.height = height; im.ncomp = ncomp; im.subsampling = subsampling;
    Geometry g;
    if (!geometry_of(im, &g)) return 0;
    return 1024 + worst_scan_bytes(g.n_blocks);
}

size_t jpeg_gpu_emit_headers(int width, int height, int ncomp, int quality_mode, int quality, int subsampling,
                             uint8_t* out, size_t capacity)
{
    jpeg_gpu_image im;
    memset(&im, 0, sizeof im);
    im.width = width; im.height = height; im.ncomp = ncomp; im.subsampling = subsampling;
    Geometry g;
    uint8_t ql[64], qc[64];
    if (!geometry_of(im, &g) || !build_qt(quality_mode, quality, ql, qc)) return 0;
    return emit_headers(width, height, g.ncomp_out, subsampling, ql, qc, out, capacity);
}

size_t jpeg_gpu_emit_headers_for(const jpeg_gpu_image* image, uint8_t* out, size_t capacity)
{
    if (!image) return 0;
    Geometry g;
    uint8_t ql[64], qc[64];
    if (!geometry_of(*image, &g) || !build_qt(image->quality_mode, image->quality, ql, qc)) return 0;
    return emit_headers(image->width, image->height, g.ncomp_out, image->subsampling, ql, qc, out, capacity,
                        (image->flags & JPEG_GPU_FLAG_RESTART) ? mcus_per_tile(g.layout) : 0);
}

jpeg_gpu_plan* jpeg_gpu_plan_create(const jpeg_gpu_image* images, int n, int device, int debug_window_words)
{
    return plan_create(images, n, device, debug_window_words, false);
}

int jpeg_gpu_plan_set_pixels(jpeg_gpu_plan* p, int i, const uint8_t* device_pixels)
{
    if (!p || i < 0 || i >= (int)p->items.size() || !p->items[i].valid) return 0;
    jpeg_gpu_plan::Item& it = p->items[i];
    it.d_pixels = device_pixels;
    ImageDesc& d = p->h_images[p->groups[it.group].result_off + it.index_in_group];
    d.px = device_pixels;
    d.align = (d.flags & JPEG_GPU_FLAG_SWAP_RB) ? 1 : alignment_of(d.px, d.stride);
    p->images_dirty = true;
    return 1;
}

int jpeg_gpu_plan_upload(jpeg_gpu_plan* p, int i, const uint8_t* host_pixels, void* stream)
{
    if (!p || i < 0 || i >= (int)p->items.size() || !p->items[i].valid) return 0;
    jpeg_gpu_plan::Item& it = p->items[i];
    if (it.img.pixels_on_device) { set_error("image %d was declared device-resident", i); return 0; }
    if (p->dev_index >= (int)g_devices.size()) { set_error("library shut down"); return 0; }
    cudaStream_t s = stream ? (cudaStream_t)stream : g_devices[p->dev_index].stream;
    if (cudaSetDevice(g_devices[p->dev_index].id) != cudaSuccess) return 0;
    // host_pixels is the top row; the memory block starts top_row_offset before it when the rows are bottom-up
    cudaError_t e = cudaMemcpyAsync(p->d_pixels + it.pixel_off, host_pixels - top_row_offset(it.img), it.pixel_bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { set_error("pixel upload failed: %s", cudaGetErrorString(e)); return 0; }
    if (!p->ran) { p->run_stream = s; p->ran = true; }     // plan_free waits for it
    return 1;
}

int jpeg_gpu_plan_run(jpeg_gpu_plan* p, void* stream)
{
    if (!p || p->dev_index >= (int)g_devices.size()) { set_error("no plan / library shut down"); return 0; }
    cudaStream_t s = stream ? (cudaStream_t)stream : g_devices[p->dev_index].stream;
    return plan_run(p, s) ? 1 : 0;
}

int jpeg_gpu_plan_enable_timing(jpeg_gpu_plan* p, int enable)
{
    if (!p || g_devices.empty()) return 0;
    if (enable && p->events.empty()) {
        if (cudaSetDevice(g_devices[p->dev_index].id) != cudaSuccess) return 0;
        p->events.resize(4 * p->groups.size());
        for (auto& e : p->events)
            if (cudaEventCreate(&e) != cudaSuccess) { set_error("cudaEventCreate failed"); return 0; }
    }
    p->timing = enable != 0;
    return 1;
}

int jpeg_gpu_plan_pass_times(jpeg_gpu_plan* p, float* transform_ms, float* entropy_ms, float* stuff_ms)
{
    if (!p || !p->timing || p->events.empty()) return 0;
    float ta = 0.f, tb = 0.f, tc = 0.f;
    for (size_t gi = 0; gi < p->groups.size(); ++gi) {
        float a = 0.f, b = 0.f, c = 0.f;
        if (cudaEventSynchronize(p->events[4 * gi + 3]) != cudaSuccess) return 0;
        if (cudaEventElapsedTime(&a, p->events[4 * gi], p->events[4 * gi + 1]) != cudaSuccess) return 0;
        if (cudaEventElapsedTime(&b, p->events[4 * gi + 1], p->events[4 * gi + 2]) != cudaSuccess) return 0;
        if (cudaEventElapsedTime(&c, p->events[4 * gi + 2], p->events[4 * gi + 3]) != cudaSuccess) return 0;
        ta += a; tb += b; tc += c;
    }
    if (transform_ms) *transform_ms = ta;
    if (entropy_ms) *entropy_ms = tb;
    if (stuff_ms) *stuff_ms = tc;
    return 1;
}

int jpeg_gpu_plan_kernel_times(jpeg_gpu_plan* p, float* encode_ms, float* stuff_ms)
{
    float a = 0.f, b = 0.f, c = 0.f;
    if (!jpeg_gpu_plan_pass_times(p, &a, &b, &c)) return 0;
    if (encode_ms) *encode_ms = a + b;
    if (stuff_ms) *stuff_ms = c;
    return 1;
}

int jpeg_gpu_plan_is_fused(const jpeg_gpu_plan* p) { return p && p->fused ? 1 : 0; }

// pass 1 (split: transform + entropy; fused: one encode kernel) + pass 2 (plan_chunks + count_ff + scan_groups + stuff)
int jpeg_gpu_plan_launches(const jpeg_gpu_plan* p) { return p ? (p->fused ? 5 : 6) * (int)p->groups.size() : 0; }

size_t jpeg_gpu_plan_num_blocks(const jpeg_gpu_plan* p) { return p ? p->n_blocks : 0; }

int jpeg_gpu_plan_attach_debug(jpeg_gpu_plan* p, int16_t* dev_coefs, uint32_t* dev_block_bits)
{
    if (!p) return 0;
    p->dbg_coefs = dev_coefs;
    p->dbg_bits = dev_block_bits;
    return 1;
}

size_t jpeg_gpu_plan_encoded_size(jpeg_gpu_plan* p, int i)
{
    if (!p || i < 0 || i >= (int)p->items.size() || p->dev_index >= (int)g_devices.size()) return 0;
    if (!plan_results(p, g_devices[p->dev_index].stream)) return 0;
    size_t sz; int st;
    item_result(p, i, &sz, &st);
    return sz;
}

int jpeg_gpu_plan_fetch(jpeg_gpu_plan* p, jpeg_gpu_output* outs, int outputs_on_device, void* stream)
{
    if (!p || !outs || p->dev_index >= (int)g_devices.size()) return 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : g_devices[p->dev_index].stream;
    if (cudaSetDevice(g_devices[p->dev_index].id) != cudaSuccess) return 0;
    if (!plan_results(p, s)) {
        for (size_t i = 0; i < p->items.size(); ++i) { outs[i].size = 0; outs[i].status = JPEG_GPU_ERR_CUDA; }
        return 0;
    }
    int ok = 0;
    bool copies = false;
    for (size_t i = 0; i < p->items.size(); ++i) {
        size_t sz; int st;
        item_result(p, (int)i, &sz, &st);
        outs[i].size = sz;
        outs[i].status = st;
        if (st != JPEG_GPU_OK) continue;
        if (!outs[i].data || outs[i].capacity < sz) { outs[i].status = JPEG_GPU_ERR_CAPACITY; continue; }
        const jpeg_gpu_plan::Item& it = p->items[i];
        const size_t hdr = it.header.size();
        cudaError_t e;
        if (outputs_on_device) {
            e = cudaMemcpyAsync(outs[i].data, it.header.data(), hdr, cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(outs[i].data + hdr, p->d_arena + it.arena_off, sz - hdr, cudaMemcpyDeviceToDevice, s);
        } else {
            memcpy(outs[i].data, it.header.data(), hdr);
            e = cudaMemcpyAsync(outs[i].data + hdr, p->d_arena + it.arena_off, sz - hdr, cudaMemcpyDeviceToHost, s);
        }
        if (e != cudaSuccess) {
            set_error("output copy failed: %s", cudaGetErrorString(e));
            outs[i].status = JPEG_GPU_ERR_CUDA;
            continue;
        }
        copies = true;
        ++ok;
    }
    if (copies && cudaStreamSynchronize(s) != cudaSuccess) { set_error("stream sync failed"); return 0; }
    return ok;
}

void jpeg_gpu_plan_destroy(jpeg_gpu_plan* p) { plan_free(p); }

// -----------------------------------------------------------------------------------------
// One chunk of a batch on one stream: plan, uploads, kernels enqueued; fetched later.
static jpeg_gpu_plan* start_chunk(const jpeg_gpu_image* images, int n, int device, cudaStream_t s, int win_words)
{
    jpeg_gpu_plan* p = plan_create(images, n, device, win_words, false);
    if (!p) return nullptr;
    // Uploads: images that follow one another in host memory AND in the plan's pixel arena (a batch cut out of one tensor;
    // 1080p RGB is a multiple of the arena's 256-byte alignment) travel as ONE copy -- a cudaMemcpyAsync per 6 MB image
    // costs the link ~8 us each, 2 ms of the 31 ms a 256-image batch takes.
    if (cudaSetDevice(g_devices[device].id) != cudaSuccess) { plan_free(p); return nullptr; }
    for (int i = 0; i < n;) {
        jpeg_gpu_plan::Item& a = p->items[i];
        if (!a.valid || images[i].pixels_on_device) { ++i; continue; }
        const uint8_t* h0 = images[i].pixels - top_row_offset(a.img);
        size_t bytes = a.pixel_bytes;
        int j = i + 1;
        for (; j < n; ++j) {
            jpeg_gpu_plan::Item& b = p->items[j];
            if (!b.valid || images[j].pixels_on_device) break;
            if (images[j].pixels - top_row_offset(b.img) != h0 + bytes || b.pixel_off != a.pixel_off + bytes) break;
            bytes += b.pixel_bytes;
        }
        const cudaError_t e = cudaMemcpyAsync(p->d_pixels + a.pixel_off, h0, bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { set_error("pixel upload failed: %s", cudaGetErrorString(e)); plan_free(p); return nullptr; }
        if (!p->ran) { p->run_stream = s; p->ran = true; }     // plan_free waits for it
        i = j;
    }
    if (!plan_run(p, s)) { plan_free(p); return nullptr; }
    return p;
}

static int finish_chunk(jpeg_gpu_plan* p, const jpeg_gpu_image* images, int n, jpeg_gpu_output* outs, int device,
                        int outputs_on_device, cudaStream_t s, int win_words)
{
    if (!p) {
        for (int i = 0; i < n; ++i) { outs[i].size = 0; outs[i].status = JPEG_GPU_ERR_CUDA; }
        return 0;
    }
    int ok = jpeg_gpu_plan_fetch(p, outs, outputs_on_device, s);
    // content that outgrew the default reservation: once more, alone, with the worst-case bound
    for (int i = 0; i < n; ++i) {
        if (!p->items[i].valid || outs[i].status != JPEG_GPU_ERR_CAPACITY) continue;
        size_t sz; int st;
        item_result(p, i, &sz, &st);
        if (st != JPEG_GPU_ERR_CAPACITY) continue;   // the caller's buffer is what is too small
        jpeg_gpu_plan* q = plan_create(&images[i], 1, device, win_words, true);
        if (!q) continue;
        if ((images[i].pixels_on_device || jpeg_gpu_plan_upload(q, 0, images[i].pixels, s)) && plan_run(q, s))
            ok += jpeg_gpu_plan_fetch(q, &outs[i], outputs_on_device, s);
        plan_free(q);
    }
    plan_free(p);
    return ok;
}

// A batch on one device.  Large host batches are cut into chunks that travel through a small
// ring of streams: the upload of chunk c+1 overlaps the kernels of chunk c and the download of
// chunk c-1 (the PCIe link is the bottleneck of this path, the kernels are not).
static int encode_on_device(const jpeg_gpu_image* images, int n, jpeg_gpu_output* outs, int device,
                            int outputs_on_device, cudaStream_t stream, int win_words)
{
    Device& dev = g_devices[device];
    if (stream) {   // caller-ordered work: everything on the caller's stream, one plan
        jpeg_gpu_plan* p = start_chunk(images, n, device, stream, win_words);
        return finish_chunk(p, images, n, outs, device, outputs_on_device, stream, win_words);
    }
    // chunk size: large enough that the per-chunk host work (plan, launches, two syncs) hides behind the link, small enough
    // that the tail (the last chunk's kernels and download, which nothing overlaps) stays short
    static const size_t kChunkPixelBytes = [] { const char* e = getenv("JPEG_GPU_CHUNK_MB"); return (size_t)(e && atoi(e) > 0 ? atoi(e) : 64) << 20; }();   // 256 x 1080p: 192 MB 29.59 ms, 96 MB 29.38, 64 MB 29.20, 32 MB 29.25
    constexpr int kRing = 4;          // chunks in flight (2, 3 and 4 measure the same, on one GPU and on four)
    struct Chunk { int lo, hi; jpeg_gpu_plan* plan; cudaStream_t s; };
    std::vector<Chunk> chunks;
    for (int lo = 0; lo < n;) {
        size_t bytes = 0;
        int hi = lo;
        while (hi < n && (hi == lo || bytes < kChunkPixelBytes)) {
            bytes += (size_t)std::max(std::abs(images[hi].stride), images[hi].width * images[hi].ncomp) * (size_t)std::max(images[hi].height, 0);
            ++hi;
        }
        chunks.push_back({lo, hi, nullptr, dev.pipe[chunks.size() % kRing]});
        lo = hi;
    }
    int ok = 0;
    size_t started = 0;
    // JPEG_GPU_TRACE=1: where the host thread spends the call (enqueueing chunks / waiting for results), to stderr
    static const bool trace = getenv("JPEG_GPU_TRACE") != nullptr;
    double t_start = 0, t_finish = 0;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0;
    for (size_t c = 0; c < chunks.size(); ++c) {
        // keep at most kRing chunks in flight (bounds device memory), finishing them in order
        const double a = trace ? now() : 0;
        for (; started < chunks.size() && started < c + kRing; ++started) {
            Chunk& k = chunks[started];
            k.plan = start_chunk(images + k.lo, k.hi - k.lo, device, k.s, win_words);
        }
        const double b = trace ? now() : 0;
        Chunk& k = chunks[c];
        ok += finish_chunk(k.plan, images + k.lo, k.hi - k.lo, outs + k.lo, device, outputs_on_device, k.s, win_words);
        if (trace) { t_start += b - a; t_finish += now() - b; }
    }
    if (trace) fprintf(stderr, "jpeg_gpu trace: device %d, %zu chunks, call %.2f ms: enqueue %.2f ms, fetch (waits + downloads) %.2f ms\n", device, chunks.size(), now() - t0, t_start, t_finish);
    return ok;
}

int jpeg_gpu_encode_batch(const jpeg_gpu_image* images, int n, jpeg_gpu_output* outs, const jpeg_gpu_batch_opts* opts)
{
    if (n <= 0 || !images || !outs) { set_error("empty batch"); return 0; }
    if (!ensure_init()) {
        for (int i = 0; i < n; ++i) { outs[i].size = 0; outs[i].status = JPEG_GPU_ERR_CUDA; }
        return 0;
    }
    const int device = opts ? opts->device : -1;
    if (device >= (int)g_devices.size()) {
        set_error("device index %d out of range (%d initialised)", device, (int)g_devices.size());
        for (int i = 0; i < n; ++i) { outs[i].size = 0; outs[i].status = JPEG_GPU_ERR_ARG; }
        return 0;
    }
    const int on_dev = opts ? opts->outputs_on_device : 0;
    const int win = opts ? opts->debug_window_words : 0;
    if (device >= 0)
        return encode_on_device(images, n, outs, device, on_dev, opts ? (cudaStream_t)opts->stream : nullptr, win);

    // shard by image index: GPU g takes [g*n/G, (g+1)*n/G); no inter-GPU traffic
    const int G = std::min((int)g_devices.size(), n);
    if (G == 1) return encode_on_device(images, n, outs, 0, on_dev, nullptr, win);
    std::vector<int> oks(G, 0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; ++g) {
        workers.emplace_back([&, g] {
            const int lo = (int)((long long)g * n / G), hi = (int)((long long)(g + 1) * n / G);
            oks[g] = encode_on_device(images + lo, hi - lo, outs + lo, g, on_dev, nullptr, win);
            errs[g] = g_last_error;
        });
    }
    int ok = 0;
    for (int g = 0; g < G; ++g) {
        workers[g].join();
        ok += oks[g];
        if (!errs[g].empty()) g_last_error = errs[g];
    }
    return ok;
}

// ---- drop-in twins -------------------------------------------------------------------------
int jpeg_gpu_encode_with_func(jpeg_gpu_write_func* func, void* context, const int quality, const int width,
                              const int height, const int num_components, const unsigned char* src_data)
{
    if (quality < 1 || quality > 3) { set_error("valid quality values are 1, 2, 3"); return 0; }   // jpeg_enc.h:1223
    if (num_components != 3 && num_components != 4) { set_error("3 or 4 components only"); return 0; }  // :954
    if (!func || !src_data) { set_error("null argument"); return 0; }
    jpeg_gpu_image im;
    memset(&im, 0, sizeof im);
    im.pixels = src_data; im.width = width; im.height = height; im.ncomp = num_components;
    im.quality_mode = JPEG_GPU_QMODE_TJE; im.quality = quality; im.subsampling = JPEG_GPU_SUB_444;
    Geometry g;
    if (!geometry_of(im, &g)) { set_error("unsupported geometry"); return 0; }   // :958-960
    std::vector<uint8_t> buf(1024 + default_scan_bytes(im, g));
    jpeg_gpu_output out;
    out.data = buf.data(); out.capacity = buf.size(); out.size = 0; out.status = 0;
    jpeg_gpu_batch_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.device = 0;
    if (jpeg_gpu_encode_batch(&im, 1, &out, &opts) != 1) {
        if (out.status != JPEG_GPU_ERR_CAPACITY) return 0;
        buf.resize(out.size);                     // pathological content: retry with the exact size
        out.data = buf.data(); out.capacity = buf.size();
        if (jpeg_gpu_encode_batch(&im, 1, &out, &opts) != 1) return 0;
    }
    // hand the bytes over the way tjei_write does: 1023-byte chunks, then the rest (jpeg_enc.h:487-490, :1169-1172)
    const size_t chunk = 1023;
    for (size_t off = 0; off < out.size; off += chunk)
        func(context, buf.data() + off, (int)std::min(chunk, out.size - off));
    return 1;
}

static void file_sink(void* ctx, void* data, int size) { fwrite(data, (size_t)size, 1, (FILE*)ctx); }   // jpeg_enc.h:1187-1191

int jpeg_gpu_encode_to_file_at_quality(const char* dest_path, const int quality, const int width, const int height,
                                       const int num_components, const unsigned char* src_data)
{
    FILE* fd = fopen(dest_path, "wb");   // like jpeg_enc.h:1201: the file is created before anything is checked
    if (!fd) { set_error("could not open %s for writing", dest_path); return 0; }
    int result = jpeg_gpu_encode_with_func(file_sink, fd, quality, width, height, num_components, src_data);
    if (fclose(fd) != 0) result = 0;
    return result;
}

int jpeg_gpu_encode_to_file(const char* dest_path, const int width, const int height, const int num_components,
                            const unsigned char* src_data)
{
    return jpeg_gpu_encode_to_file_at_quality(dest_path, 3, width, height, num_components, src_data);   // jpeg_enc.h:1183
}

}  // extern "C"
