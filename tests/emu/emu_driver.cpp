// tests/emu/emu_driver.cpp -- TEST INFRASTRUCTURE.  Runs the product's kernel sources
// (imagecodecs_b200/csrc/jpeg_kernel.cuh, jpeg_stuff.cuh) on CPU threads through cuda_emu.h and
// exposes one C function the CPU test-suite calls.  See cuda_emu.h for why this exists.
#define JG_EMULATE 1
#include "jpeg_stuff.cuh"
#include "jpeg_transform.cuh"
#include "jpeg_entropy.cuh"
#include "jpeg_decode.cuh"
#include "jpeg_decode.h"

#include <stdlib.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>
#include <cstdlib>

namespace jg {
namespace emu {
thread_local Tls tls;
}
}  // namespace jg

namespace {

using namespace jg;

struct ThreadArg {
    emu::Cta* cta;
    int tid;
    const std::function<void()>* body;
};

void* thread_main(void* p)
{
    ThreadArg* a = (ThreadArg*)p;
    emu::tls.tid = a->tid;
    emu::tls.cta = a->cta;
    (*a->body)();
    return nullptr;
}

// "launch" n_ctas CTAs of kThreads threads that all run concurrently
void launch(int n_ctas, size_t smem_bytes, const std::function<void()>& body, int kThreads = jg::kThreads, int first_cta = 0, int grid = 0)
{
    std::vector<emu::Cta> ctas(n_ctas);
    std::vector<ThreadArg> args((size_t)n_ctas * kThreads);
    std::vector<pthread_t> th((size_t)n_ctas * kThreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int c = 0; c < n_ctas; ++c) {
        emu::Cta& cta = ctas[c];
        cta.nthreads = kThreads;
        cta.id = first_cta + c;
        cta.grid = grid > 0 ? grid : first_cta + n_ctas;
        cta.smem = (unsigned char*)aligned_alloc(64, (smem_bytes + 63) / 64 * 64);
        pthread_barrier_init(&cta.bar, nullptr, kThreads);
        for (int wv = 0; wv < kThreads / 32; ++wv) pthread_barrier_init(&cta.wbar[wv], nullptr, 32);
        for (int t = 0; t < kThreads; ++t) {
            ThreadArg& a = args[(size_t)c * kThreads + t];
            a.cta = &cta; a.tid = t; a.body = &body;
            pthread_create(&th[(size_t)c * kThreads + t], &attr, thread_main, &a);
        }
    }
    for (auto& x : th) pthread_join(x, nullptr);
    for (auto& cta : ctas) free(cta.smem);
}

size_t smem_bytes(int layout, int nc)
{
    if (layout == LAYOUT_444) return nc == 3 ? sizeof(Smem<LAYOUT_444, 3>) : sizeof(Smem<LAYOUT_444, 4>);
    if (layout == LAYOUT_420) return nc == 3 ? sizeof(Smem<LAYOUT_420, 3>) : sizeof(Smem<LAYOUT_420, 4>);
    return sizeof(Smem<LAYOUT_GRAY, 1>);
}

}  // namespace

extern "C" {

// Encode `n_images` images of identical geometry/quality with `n_ctas` emulated CTAs
// running concurrently.  scan_out: [n_images][scan_cap]; scan_bytes: [n_images].
// Returns 0 on success, otherwise the kernel's error flag / a negative setup error.
int emu_encode(const uint8_t* pixels, int n_images, int w, int h, int ncomp, int stride, int flags, int subsampling,
               int quality_mode, int quality, int win_words, int n_ctas,
               uint8_t* scan_out, size_t scan_cap, unsigned long long* scan_bytes, unsigned* img_status,
               int16_t* dbg_coefs, uint32_t* dbg_bits)
{
    const int layout = ncomp == 1 ? LAYOUT_GRAY : (subsampling ? LAYOUT_420 : LAYOUT_444);
    uint8_t ql[64], qc[64];
    if (!build_qt(quality_mode, quality, ql, qc)) return -1;
    QuantSet Q;
    build_pqt(ql, Q.luma);
    build_pqt(qc, Q.chroma);
    HuffLut lut;
    build_huff_lut(&lut);

    const int mcu = layout == LAYOUT_420 ? 16 : 8;
    const int bpm = layout == LAYOUT_444 ? 3 : (layout == LAYOUT_420 ? 6 : 1);
    const int mcus_x = (w + mcu - 1) / mcu, mcus_y = (h + mcu - 1) / mcu;
    const int n_mcus = mcus_x * mcus_y;
    const int M = mcus_per_tile(layout);
    const int tiles = (n_mcus + M - 1) / M;
    if (stride == 0) stride = w * ncomp;
    const int pitch = stride < 0 ? -stride : stride;       // negative stride: each image's rows are stored bottom-up

    const size_t raw_cap = (scan_cap + 255) / 256 * 256;
    uint8_t* raw = (uint8_t*)aligned_alloc(256, raw_cap * n_images);
    std::vector<ImageDesc> imgs(n_images);
    for (int i = 0; i < n_images; ++i) {
        ImageDesc& d = imgs[i];
        d.px = pixels + (size_t)i * pitch * h + (stride < 0 ? (size_t)pitch * (h - 1) : 0);
        d.flags = flags;
        d.raw = raw + (size_t)i * raw_cap;
        d.raw_cap = scan_cap;
        d.out = scan_out + (size_t)i * scan_cap;
        d.out_cap = scan_cap;
        d.first_block = (unsigned long long)i * n_mcus * bpm;
        d.w = w; d.h = h; d.stride = stride; d.mcus_x = mcus_x; d.n_mcus = n_mcus;
        d.first_tile = i * tiles; d.n_tiles = tiles;
        { const size_t v = (size_t)d.px | (size_t)stride | 16u; d.align = (flags & 1) ? 1 : (int)(v & (~v + 1)); }
    }
    const int n_tiles = tiles * n_images;
    const size_t max_chunks = (size_t)n_images * ((scan_cap + kChunkBytes - 1) / kChunkBytes) + 1;
    std::vector<unsigned long long> desc_bits(n_tiles, 0), desc_ff(max_chunks, 0), raw_bytes(n_images, 0);
    std::vector<unsigned> first_chunk(n_images + 1, 0), desc_dc(3 * (size_t)n_tiles, 0);
    unsigned ticket = 0, error = 0;
    std::vector<unsigned> ff_groups(max_chunks / 32 + 2, 0);
    for (int i = 0; i < n_images; ++i) { scan_bytes[i] = 0; img_status[i] = 0; }

    LaunchParams P;
    P.images = imgs.data(); P.n_images = n_images; P.n_tiles = n_tiles;
    P.tiles_per_image = (n_images % 2) ? tiles : 0;   // exercise both ticket->tile paths
    std::vector<int> counts(n_images, tiles);
    std::vector<uint32_t> sched(schedule_words(n_images));
    build_schedule(counts.data(), n_images, sched.data());
    P.sched = sched.data();
    P.win_words = win_words ? (win_words < kWinWordsMin ? kWinWordsMin : (win_words > kWinWordsMax ? kWinWordsMax : win_words)) : kWinWordsMax;
    P.ticket = &ticket; P.error = &error; P.ff_groups = ff_groups.data();
    P.desc_bits = desc_bits.data(); P.desc_ff = desc_ff.data(); P.desc_dc = desc_dc.data();
    P.raw_bytes = raw_bytes.data(); P.first_chunk = first_chunk.data();
    P.scan_bytes = scan_bytes; P.img_status = img_status; P.huff = &lut;
    P.dbg_coefs = dbg_coefs; P.dbg_bits = dbg_bits;

    const int nc = ncomp;
    launch(n_ctas > (n_tiles + kWarps - 1) / kWarps ? (n_tiles + kWarps - 1) / kWarps : n_ctas, smem_bytes(layout, nc), [&] {
        // the host's rule: restart images -> kModeRestart, else n_images < kDeepMaxImages -> kModeDeep; here every mode gets exercised
        const int mode = (flags & kFlagRestart) ? kModeRestart : (n_images < 2 ? kModeDeep : kModePlain);
#define JG_RUN(L, N)                                                                                   \
        do {                                                                                           \
            if (mode == kModeDeep) encode_tiles_kernel<L, N, kModeDeep>(P, Q);                         \
            else if (mode == kModeRestart) encode_tiles_kernel<L, N, kModeRestart>(P, Q);              \
            else encode_tiles_kernel<L, N, kModePlain>(P, Q);                                          \
        } while (0)
        if (layout == LAYOUT_444 && nc == 3) JG_RUN(LAYOUT_444, 3);
        else if (layout == LAYOUT_444 && nc == 4) JG_RUN(LAYOUT_444, 4);
        else if (layout == LAYOUT_420 && nc == 3) JG_RUN(LAYOUT_420, 3);
        else if (layout == LAYOUT_420 && nc == 4) JG_RUN(LAYOUT_420, 4);
        else JG_RUN(LAYOUT_GRAY, 1);
#undef JG_RUN
    });
    if (!error) {
        launch(1, sizeof(StuffSmem), [&] { plan_chunks_kernel(P); });
        launch(n_ctas, sizeof(StuffSmem), [&] { count_ff_kernel(P); }, kCountThreads);
        launch(1, sizeof(StuffSmem), [&] { scan_groups_kernel(P); });
        launch(n_ctas, sizeof(StuffSmem), [&] { stuff_kernel(P); });
    }
    free(raw);
    return (int)error;
}

// The same through the split pipeline: transform_kernel (pass A) -> entropy_kernel (pass B) -> plan_chunks / stuff.
// n_ctas bounds how many CTAs of pass B and of the stuffing pass run concurrently; pass A runs its grid in waves.
// dbg_coefs (optional) receives the coefficient plane, dbg_bits the bits of every block.
int emu_encode_split(const uint8_t* pixels, int n_images, int w, int h, int ncomp, int stride, int flags, int subsampling,
                     int quality_mode, int quality, int win_words, int n_ctas,
                     uint8_t* scan_out, size_t scan_cap, unsigned long long* scan_bytes, unsigned* img_status,
                     int16_t* dbg_coefs, uint32_t* dbg_bits)
{
    const int layout = ncomp == 1 ? LAYOUT_GRAY : (subsampling ? LAYOUT_420 : LAYOUT_444);
    uint8_t ql[64], qc[64];
    if (!build_qt(quality_mode, quality, ql, qc)) return -1;
    QuantSet Q;
    build_pqt(ql, Q.luma);
    build_pqt(qc, Q.chroma);
    HuffLut lut;
    build_huff_lut(&lut);

    const int mcu = layout == LAYOUT_420 ? 16 : 8;
    const int bpm = layout == LAYOUT_444 ? 3 : (layout == LAYOUT_420 ? 6 : 1);
    const int mcus_x = (w + mcu - 1) / mcu, mcus_y = (h + mcu - 1) / mcu;
    const int n_mcus = mcus_x * mcus_y;
    const bool restart = (flags & kFlagRestart) != 0;
    const int bpt = restart ? kEntTileBlocksRestart : kEntTileBlocks;
    const int n_blocks = n_mcus * bpm;
    const int tiles = (n_blocks + bpt - 1) / bpt;
    const int item_mcus = transform_item_mcus(layout);
    const int items = (n_mcus + item_mcus - 1) / item_mcus;
    if (stride == 0) stride = w * ncomp;
    const int pitch = stride < 0 ? -stride : stride;

    const size_t raw_cap = (scan_cap + 255) / 256 * 256;
    uint8_t* raw = (uint8_t*)aligned_alloc(256, raw_cap * n_images);
    std::vector<int16_t> coefs((size_t)n_images * n_blocks * 64 + 64, (int16_t)0x5555);
    std::vector<ImageDesc> imgs(n_images);
    for (int i = 0; i < n_images; ++i) {
        ImageDesc& d = imgs[i];
        d.px = pixels + (size_t)i * pitch * h + (stride < 0 ? (size_t)pitch * (h - 1) : 0);
        d.flags = flags;
        d.raw = raw + (size_t)i * raw_cap;
        d.raw_cap = scan_cap;
        d.out = scan_out + (size_t)i * scan_cap;
        d.out_cap = scan_cap;
        d.first_block = (unsigned long long)i * n_blocks;
        d.w = w; d.h = h; d.stride = stride; d.mcus_x = mcus_x; d.n_mcus = n_mcus;
        d.first_tile = i * tiles; d.n_tiles = tiles;
        { const size_t v = (size_t)d.px | (size_t)stride | 16u; d.align = (flags & 1) ? 1 : (int)(v & (~v + 1)); }
    }
    const int n_tiles = tiles * n_images;
    const size_t max_chunks = (size_t)n_images * ((scan_cap + kChunkBytes - 1) / kChunkBytes) + 1;
    std::vector<unsigned long long> desc_bits(n_tiles, 0), desc_ff(max_chunks, 0), raw_bytes(n_images, 0);
    std::vector<unsigned> first_chunk(n_images + 1, 0);
    unsigned ticket = 0, error = 0;
    std::vector<unsigned> ff_groups(max_chunks / 32 + 2, 0);
    for (int i = 0; i < n_images; ++i) { scan_bytes[i] = 0; img_status[i] = 0; }

    // ---- pass A ----
    std::vector<uint32_t> first_item(n_images + 1);
    for (int i = 0; i <= n_images; ++i) first_item[i] = (uint32_t)(i * items);
    TransformParams TP;
    TP.images = imgs.data(); TP.n_images = n_images; TP.n_items = items * n_images;
    TP.items_per_image = (n_images % 2) ? items : 0;     // exercise both item -> image paths
    TP.first_item = first_item.data();
    TP.coefs = coefs.data();
    const int gridA = (TP.n_items + kWarps - 1) / kWarps;
    const size_t smemA = layout == LAYOUT_444 ? sizeof(TSmem<LAYOUT_444>) : (layout == LAYOUT_420 ? sizeof(TSmem<LAYOUT_420>) : sizeof(TSmem<LAYOUT_GRAY>));
    for (int c0 = 0; c0 < gridA; c0 += 16) {
        launch(gridA - c0 < 16 ? gridA - c0 : 16, smemA, [&] {
            if (layout == LAYOUT_444 && ncomp == 3) transform_kernel<LAYOUT_444, 3>(TP, Q);
            else if (layout == LAYOUT_444) transform_kernel<LAYOUT_444, 4>(TP, Q);
            else if (layout == LAYOUT_420 && ncomp == 3) transform_kernel<LAYOUT_420, 3>(TP, Q);
            else if (layout == LAYOUT_420) transform_kernel<LAYOUT_420, 4>(TP, Q);
            else transform_kernel<LAYOUT_GRAY, 1>(TP, Q);
        }, kThreads, c0);
    }
    if (dbg_coefs) memcpy(dbg_coefs, coefs.data(), (size_t)n_images * n_blocks * 64 * sizeof(int16_t));

    // ---- pass B ----
    LaunchParams P;
    memset(&P, 0, sizeof P);
    P.images = imgs.data(); P.n_images = n_images; P.n_tiles = n_tiles;
    P.tiles_per_image = (n_images % 2) ? tiles : 0;
    std::vector<int> counts(n_images, tiles);
    std::vector<uint32_t> sched(schedule_words(n_images));
    build_schedule(counts.data(), n_images, sched.data());
    P.sched = sched.data();
    P.win_words = win_words ? win_words : kWinWordsMax;      // a small value sends every tile with more bits the slow way
    P.ticket = &ticket; P.error = &error; P.ff_groups = ff_groups.data();
    P.desc_bits = desc_bits.data(); P.desc_ff = desc_ff.data(); P.desc_dc = nullptr;
    P.raw_bytes = raw_bytes.data(); P.first_chunk = first_chunk.data();
    P.scan_bytes = scan_bytes; P.img_status = img_status; P.huff = &lut;
    P.dbg_coefs = nullptr; P.dbg_bits = dbg_bits;
    P.coefs = coefs.data(); P.bpm = bpm; P.blocks_per_tile = bpt;
    CoefMap cmap;
    memset(&cmap, 0, sizeof cmap);
    cmap.q[0] = (unsigned long long)(size_t)coefs.data();
    cmap.q[1] = (unsigned long long)n_images * n_blocks;
    const int gridB = n_ctas > (n_tiles + kEntWarps - 1) / kEntWarps ? (n_tiles + kEntWarps - 1) / kEntWarps : n_ctas;
    launch(gridB, sizeof(EntSmem), [&] {
        if (restart) entropy_kernel<kEntModeRestart>(P, cmap);
        else entropy_kernel<kEntModePlain>(P, cmap);
    }, kEntThreads);
    if (!error) {
        launch(1, sizeof(StuffSmem), [&] { plan_chunks_kernel(P); });
        launch(n_ctas, sizeof(StuffSmem), [&] { count_ff_kernel(P); }, kCountThreads);
        launch(1, sizeof(StuffSmem), [&] { scan_groups_kernel(P); });
        launch(n_ctas, sizeof(StuffSmem), [&] { stuff_kernel(P); });
    }
    free(raw);
    return (int)error;
}

// The ticket -> tile mapping of the kernel (tile_of_ticket) for images with the given tile counts:
// out_g / out_img receive the launch-wide tile index and the image of every ticket.
int emu_ticket_map(const int* tiles, int n_images, int force_schedule, unsigned* out_g, unsigned* out_img)
{
    std::vector<ImageDesc> imgs(n_images);
    int total = 0;
    bool uniform = true;
    for (int i = 0; i < n_images; ++i) {
        imgs[i].first_tile = total; imgs[i].n_tiles = tiles[i];
        total += tiles[i];
        uniform = uniform && tiles[i] == tiles[0];
    }
    std::vector<uint32_t> sched(schedule_words(n_images));
    build_schedule(tiles, n_images, sched.data());
    LaunchParams P;
    memset(&P, 0, sizeof P);
    P.images = imgs.data(); P.n_images = n_images; P.n_tiles = total;
    P.tiles_per_image = (uniform && !force_schedule) ? tiles[0] : 0;
    P.sched = sched.data();
    for (int v = 0; v < total; ++v) out_g[v] = tile_of_ticket(P, (unsigned)v, out_img[v]);
    return total;
}

// The decoder's per-item device functions (jpeg_decode.cuh) run in plain loops: same code as on the GPU,
// where one thread executes one call.  Returns the nj_result_t; out receives RGB / gray pixels.
// Order in which the emulated threads of a subsequence round run (on the GPU: any).  0 = ascending in even rounds and
// descending in odd ones, 1 = always descending (every thread sees its predecessor's OLD state: the slowest case),
// 2 = always ascending (every thread sees the new one), 3 = eight host threads share a round's items and race on the
// records like the GPU's threads do (the in-place update has to tolerate every interleaving).
static int g_round_order = 0;
void emu_set_round_order(int o) { g_round_order = o; }

// sub_log2: -1 = the library's policy (jd::subsequence_log2), 0 = interval path, > 0 = subsequences of that size where
// the stream allows them at all.  rounds (optional) receives the number of rounds the subsequence decode took (0: not used).
int emu_decode_sub(const uint8_t* jpeg, size_t size, uint8_t* out, size_t cap, int* w, int* h, int* ncomp, int sub_log2, int* rounds)
{
    if (rounds) *rounds = 0;
    jd::Info I;
    const int rc = jd::parse(jpeg, size, &I);
    if (rc != jd::kOk) return rc;
    *w = I.width; *h = I.height; *ncomp = I.ncomp;
    if (cap < (size_t)I.width * I.height * I.ncomp) return -1;
    std::vector<int16_t> coef(I.n_blocks * 64, 0);
    std::vector<uint8_t> planes(I.plane_bytes + 16, 0);
    unsigned err = 0;
    jd::DevParams P;
    memset(&P, 0, sizeof P);
    P.data = jpeg; P.interval_off = I.interval_off.data(); P.n_intervals = (int)I.interval_off.size() - 1;
    P.rstinterval = I.rstinterval; P.n_mcus = I.n_mcus; P.mbwidth = I.mbwidth; P.ncomp = I.ncomp;
    P.vlc = I.vlc.data(); P.coef = coef.data(); P.planes = planes.data(); P.error = &err;
    for (int c = 0; c < I.ncomp; ++c) {
        const jd::Component& k = I.comp[c];
        jd::DevComponent& d = P.comp[c];
        d.ssx = k.ssx; d.ssy = k.ssy; d.bw = k.bw; d.dctab = k.dctabsel; d.actab = k.actabsel; d.stride = k.stride;
        d.coef_off = k.coef_off; d.plane_off = k.plane_off;
        for (int i = 0; i < 64; ++i) d.dq[jd::zz_nat(i)] = I.qtab[k.qtsel][i];
    }
    std::vector<uint16_t> l1(4 << jd::kL1Bits);
    for (int i = 0; i < (4 << jd::kL1Bits); ++i) l1[i] = jd::l1_entry(P.vlc, i >> jd::kL1Bits, i & ((1 << jd::kL1Bits) - 1));
    const size_t scan_bytes = I.scan_end - I.scan_off;
    if (sub_log2 < 0) sub_log2 = jd::subsequence_log2(I, scan_bytes);
    else if (sub_log2 > 0 && (I.rstinterval || !I.clean_stuffing || !jd::mcu_block_map(I, P.blk) || scan_bytes >= ((size_t)1 << 28))) sub_log2 = 0;
    if (sub_log2 > 0) {
        // what jpeg_decode_api.cpp launches as kernels: rounds until nobody decodes again, the scan, the writing pass
        P.sub_log2 = sub_log2; P.n_sub = (int)((scan_bytes + ((size_t)1 << sub_log2) - 1) >> sub_log2);
        P.bpm = jd::mcu_block_map(I, P.blk); P.scan = jpeg + I.scan_off; P.scan_bytes = (unsigned)scan_bytes;
        P.total_blocks = (unsigned long long)I.n_mcus * P.bpm;
        std::vector<unsigned long long> exits(P.n_sub, 0);
        std::vector<jd::SubStart> sums(P.n_sub), start(P.n_sub);
        std::vector<unsigned> la(P.n_sub), lb(P.n_sub);
        unsigned cnt[3] = {0, 0, 0};
        P.sub_exit = exits.data(); P.sub_sum = sums.data(); P.sub_start = start.data(); P.sub_list[0] = la.data(); P.sub_list[1] = lb.data(); P.sub_cnt = cnt;
        int r = 0;
        for (;; ++r) {
            cnt[(r + 2) % 3] = 0;
            const unsigned count = jd::sync_round_count(P, r);
            unsigned appended = 0;
            const bool descending = g_round_order == 1 || (g_round_order == 0 && (r & 1));
            if (g_round_order == 3) {
                std::atomic<unsigned> total{0};
                std::vector<std::thread> pool;
                for (unsigned t = 0; t < 8; ++t)
                    pool.emplace_back([&, t] {
                        unsigned mine = 0;
                        for (unsigned k = t; k < count; k += 8) mine += jd::sync_round_item(P, l1.data(), r, k, count);
                        total += mine;
                    });
                for (std::thread& th : pool) th.join();
                appended = total;
            } else
            for (unsigned k = 0; k < count; ++k) appended += jd::sync_round_item(P, l1.data(), r, descending ? count - 1 - k : k, count);
            if (r >= 1 && !appended) break;
            if (r > P.n_sub + 2) return -2;       // cannot happen: every round settles at least one more subsequence
        }
        if (rounds) *rounds = r + 1;
        jd::SubStart run = {0u, 0, 0, 0};
        for (int i = 0; i < P.n_sub; ++i) { start[i] = run; run.n += sums[i].n; run.dc0 += sums[i].dc0; run.dc1 += sums[i].dc1; run.dc2 += sums[i].dc2; }
        for (int i = 0; i < P.n_sub; ++i) jd::write_subsequence(P, l1.data(), i);
    } else {
        for (int iv = 0; iv < P.n_intervals; ++iv) jd::decode_interval(P, l1.data(), iv);
    }
    if (err) return (int)err;
    for (int c = 0; c < I.ncomp; ++c)
        for (unsigned long long b = 0; b < (unsigned long long)I.comp[c].bw * I.comp[c].bh; ++b) jd::idct_block(P, c, b);
    std::vector<std::vector<uint8_t>> tmp;
    const uint8_t* plane[3]; int pw[3], ph[3], ps[3];
    for (int c = 0; c < I.ncomp; ++c) {
        plane[c] = planes.data() + I.comp[c].plane_off; pw[c] = I.comp[c].width; ph[c] = I.comp[c].height; ps[c] = I.comp[c].stride;
        while (pw[c] < I.width || ph[c] < I.height) {
            if (pw[c] < I.width) {
                tmp.emplace_back((size_t)pw[c] * ph[c] * 2);
                uint8_t* o = tmp.back().data();
                for (int y = 0; y < ph[c]; ++y) for (int ox = 0; ox < 2 * pw[c]; ++ox) o[(size_t)y * 2 * pw[c] + ox] = jd::upsample_h(plane[c], pw[c], ps[c], y, ox);
                plane[c] = o; pw[c] <<= 1; ps[c] = pw[c];
            }
            if (ph[c] < I.height) {
                tmp.emplace_back((size_t)pw[c] * ph[c] * 2);
                uint8_t* o = tmp.back().data();
                for (int oy = 0; oy < 2 * ph[c]; ++oy) for (int x = 0; x < pw[c]; ++x) o[(size_t)oy * pw[c] + x] = jd::upsample_v(plane[c], ph[c], ps[c], oy, x);
                plane[c] = o; ph[c] <<= 1; ps[c] = pw[c];
            }
        }
    }
    for (int y = 0; y < I.height; ++y)
        for (int x = 0; x < I.width; ++x) {
            if (I.ncomp == 3) jd::to_rgb(out + ((size_t)y * I.width + x) * 3, plane[0][(size_t)y * ps[0] + x], plane[1][(size_t)y * ps[1] + x], plane[2][(size_t)y * ps[2] + x]);
            else out[(size_t)y * I.width + x] = plane[0][(size_t)y * ps[0] + x];
        }
    return 0;
}
int emu_decode(const uint8_t* jpeg, size_t size, uint8_t* out, size_t cap, int* w, int* h, int* ncomp)
{
    return emu_decode_sub(jpeg, size, out, cap, w, h, ncomp, -1, nullptr);
}
}
