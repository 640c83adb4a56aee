// jpeg_decode_api.cpp -- C entry points of the decoder (include/jpeg_gpu.h, "decode" section):
// parse on the host, everything else on the GPU, pixels back to the caller.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <utility>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "jpeg_decode.cuh"
#include "jpeg_decode.h"
#include "jpeg_gpu.h"

namespace jg {
int cuda_device_of(int index);                 // jpeg_gpu_api.cpp
void set_error_text(const char* text);
}

namespace {

using namespace jd;

const char* result_text(int rc)
{
    switch (rc) {
        case kNoJpeg: return "not a JPEG file (NJ_NO_JPEG)";
        case kUnsupported: return "unsupported format (NJ_UNSUPPORTED)";
        case kOutOfMem: return "out of memory (NJ_OUT_OF_MEM)";
        case kSyntaxError: return "syntax error (NJ_SYNTAX_ERROR)";
        default: return "internal error (NJ_INTERNAL_ERR)";
    }
}

// The decoder's own stream-ordered memory pool (one per device, created on first use): it keeps its memory between calls
// -- the default pool hands everything back at every synchronisation -- without touching the release threshold of the
// device's DEFAULT pool, which other users of cudaMallocAsync in the process (torch) share.
cudaMemPool_t decoder_pool(int dev_id)
{
    static std::mutex m;
    static std::vector<std::pair<int, cudaMemPool_t>> pools;
    std::lock_guard<std::mutex> lk(m);
    for (auto& kv : pools) if (kv.first == dev_id) return kv.second;
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev_id;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { (void)cudaGetLastError(); pool = nullptr; }
    if (pool) {
        unsigned long long keep_all = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all);
    }
    pools.push_back({dev_id, pool});
    return pool;
}

struct DeviceBuffers {     // freed in order on every exit path
    cudaStream_t s = nullptr;
    cudaMemPool_t pool = nullptr;       // nullptr: the device's default pool
    std::vector<void*> ptrs;
    template <typename T>
    T* alloc(size_t n)
    {
        void* p = nullptr;
        const size_t bytes = std::max<size_t>(n * sizeof(T), 16);
        if ((pool ? cudaMallocFromPoolAsync(&p, bytes, pool, s) : cudaMallocAsync(&p, bytes, s)) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T*)p;
    }
    ~DeviceBuffers()
    {
        for (void* p : ptrs) cudaFreeAsync(p, s);
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    }
};

#define JD_CUDA(x)                                                                                  \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) {                                                                    \
            char b_[256]; snprintf(b_, sizeof b_, "decode: %s: %s", #x, cudaGetErrorString(e_));    \
            jg::set_error_text(b_);                                                                 \
            return 0;                                                                               \
        }                                                                                           \
    } while (0)

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct Job {                 // one image of a batch
    Info I;
    bool ok = false;
    size_t data_off = 0, iv_off = 0, coef_off = 0, plane_off = 0, out_off = 0, out_bytes = 0, sub_off = 0;
    int vlc_slot = 0;
    int sub_log2 = 0, n_sub = 0;         // > 0: the scan has no restart markers and is decoded as subsequences
};

// Decode n files.  outs[i].pixels are host buffers (or device buffers if pixels_on_device); kernel_ms (optional)
// receives the device time of all kernels of the batch.  Returns the number of images decoded.
int decode_batch(const jpeg_gpu_stream* in, int n, jpeg_gpu_decoded* outs, int pixels_on_device, float* kernel_ms)
{
    if (!in || !outs || n <= 0) { jg::set_error_text("decode: null argument"); return 0; }
    if (n > 32767) { jg::set_error_text("decode: at most 32767 images per call"); return 0; }   // two plane ops per image in grid.z
    std::vector<Job> jobs((size_t)n);
    std::vector<std::vector<uint8_t>> dht_sets;       // distinct Huffman table sets of the batch (usually one), by their DHT bytes
    std::vector<std::vector<uint16_t>> vlc_sets;
    size_t data_bytes = 0, iv_words = 0, coef_words = 0, plane_bytes = 0, out_total = 0;
    int n_ok = 0, max_iv = 0;
    unsigned long long max_blocks[3] = {0, 0, 0};
    // the marker loops of the files are independent of one another: a small pool of host threads walks them (a file of
    // the bench batch takes ~0.1 ms; 256 of them one after the other were a quarter of the whole call)
    std::vector<int> parse_rc((size_t)n, (int)kNoJpeg);
    {
        const int workers = std::max(1, std::min({n / 8, 16, (int)std::thread::hardware_concurrency()}));
        std::atomic<int> next{0};
        auto work = [&] {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1))
                parse_rc[i] = in[i].data ? parse(in[i].data, in[i].size, &jobs[i].I, false) : (int)kNoJpeg;
        };
        std::vector<std::thread> pool;
        for (int k = 1; k < workers; ++k) pool.emplace_back(work);
        work();
        for (std::thread& th : pool) th.join();
    }
    for (int i = 0; i < n; ++i) {
        Job& j = jobs[i];
        outs[i].width = outs[i].height = outs[i].ncomp = 0;
        int rc = parse_rc[i];
        // A header may announce far more blocks than the scan can hold (a 200-byte file with a 65535 x 65535 frame).  NanoJPEG
        // would try to allocate the planes and, at this size, report NJ_OUT_OF_MEM; so does this, BEFORE anything is
        // allocated: every block takes at least 4 bits of scan.
        if (rc == kOk && j.I.plane_bytes > ((size_t)1 << 31) && j.I.n_blocks > 2 * (j.I.scan_end - j.I.scan_off) + 1024) rc = kOutOfMem;
        size_t slot = 0;
        if (rc == kOk) {                                      // the 65536-entry tables are built once per distinct DHT content
            while (slot < dht_sets.size() && dht_sets[slot] != j.I.dht) ++slot;
            if (slot == dht_sets.size()) {
                std::vector<uint16_t> tables;
                rc = build_vlc_tables(j.I.dht, &tables);
                if (rc == kOk) { dht_sets.push_back(j.I.dht); vlc_sets.push_back(std::move(tables)); }
            }
        }
        outs[i].status = rc == kOk ? JPEG_GPU_OK : JPEG_GPU_ERR_ARG;
        if (rc != kOk) { jg::set_error_text(result_text(rc)); continue; }
        outs[i].width = j.I.width; outs[i].height = j.I.height; outs[i].ncomp = j.I.ncomp;
        j.out_bytes = (size_t)j.I.width * j.I.height * j.I.ncomp;
        if (!outs[i].pixels || outs[i].capacity < j.out_bytes) { outs[i].status = JPEG_GPU_ERR_CAPACITY; jg::set_error_text("decode: output buffer too small"); continue; }
        j.ok = true; ++n_ok;
        j.vlc_slot = (int)slot;
        j.data_off = data_bytes; data_bytes += align256(j.I.scan_end + 16);
        j.iv_off = iv_words; iv_words += j.I.interval_off.size();
        j.coef_off = coef_words; coef_words += j.I.n_blocks * 64;
        j.plane_off = plane_bytes; plane_bytes += align256(j.I.plane_bytes + 16);
        j.out_off = out_total; out_total += align256(j.out_bytes);
        for (int c = 0; c < j.I.ncomp; ++c) max_blocks[c] = std::max(max_blocks[c], (unsigned long long)j.I.comp[c].bw * j.I.comp[c].bh);
    }
    if (n_ok == 0) return 0;
    // restart-free scans: subsequences (their size depends on how much entropy-coded data the whole call has)
    size_t scan_total = 0, sub_total = 0;
    int max_sub = 0;
    for (const Job& j : jobs) if (j.ok) scan_total += j.I.scan_end - j.I.scan_off;
    for (Job& j : jobs) {
        if (!j.ok) continue;
        j.sub_log2 = subsequence_log2(j.I, scan_total);
        if (j.sub_log2) {
            const size_t bytes = j.I.scan_end - j.I.scan_off;
            j.n_sub = (int)((bytes + ((size_t)1 << j.sub_log2) - 1) >> j.sub_log2);
            j.sub_off = sub_total; sub_total += (size_t)j.n_sub;
            max_sub = std::max(max_sub, j.n_sub);
        } else {
            max_iv = std::max(max_iv, (int)j.I.interval_off.size() - 1);
        }
    }
    if (jpeg_gpu_device_count() == 0 && jpeg_gpu_init(nullptr, 0) <= 0) return 0;
    JD_CUDA(cudaSetDevice(jg::cuda_device_of(0)));

    DeviceBuffers B;
    B.pool = decoder_pool(jg::cuda_device_of(0));
    JD_CUDA(cudaStreamCreateWithFlags(&B.s, cudaStreamNonBlocking));
    uint8_t* d_data = B.alloc<uint8_t>(data_bytes);
    uint32_t* d_iv = B.alloc<uint32_t>(iv_words);
    uint16_t* d_vlc = B.alloc<uint16_t>(vlc_sets.size() * 4 * 65536);
    int16_t* d_coef = B.alloc<int16_t>(coef_words);
    uint8_t* d_planes = B.alloc<uint8_t>(plane_bytes);
    uint8_t* d_out = pixels_on_device ? nullptr : B.alloc<uint8_t>(out_total);
    unsigned* d_err = B.alloc<unsigned>((size_t)n);
    DevParams* d_params = B.alloc<DevParams>((size_t)n);
    constexpr int kMaxRoundsPerCheck = 256;
    unsigned long long* d_exit = B.alloc<unsigned long long>(sub_total);
    SubStart* d_sums = B.alloc<SubStart>(2 * sub_total);                       // per subsequence: its own sums, then their exclusive sums
    unsigned* d_lists = B.alloc<unsigned>(2 * sub_total);
    unsigned* d_cnt = B.alloc<unsigned>(3 * (size_t)n);
    unsigned* d_appended = B.alloc<unsigned>(kMaxRoundsPerCheck);
    if (!d_exit || !d_sums || !d_lists || !d_cnt || !d_appended) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
    JD_CUDA(cudaMemsetAsync(d_cnt, 0, 3 * (size_t)n * 4, B.s));
    if (!d_data || !d_iv || !d_vlc || !d_coef || !d_planes || (!pixels_on_device && !d_out) || !d_err || !d_params) {
        jg::set_error_text(result_text(kOutOfMem)); return 0;
    }
    for (size_t k = 0; k < vlc_sets.size(); ++k)
        JD_CUDA(cudaMemcpyAsync(d_vlc + k * 4 * 65536, vlc_sets[k].data(), 4 * 65536 * 2, cudaMemcpyHostToDevice, B.s));
    JD_CUDA(cudaMemsetAsync(d_coef, 0, coef_words * 2, B.s));
    JD_CUDA(cudaMemsetAsync(d_err, 0, (size_t)n * 4, B.s));

    static const unsigned char zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                                         28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54,
                                         47, 55, 62, 63};
    std::vector<DevParams> params((size_t)n);
    std::vector<uint32_t> iv_host(iv_words);
    // plane pipeline of njConvert (:817-836), as rounds of batch-wide passes: H before V, until the image size is reached
    struct PlaneState { const uint8_t* p; int w, h, s; };
    std::vector<PlaneState> st((size_t)n * 3);
    for (int i = 0; i < n; ++i) {
        Job& j = jobs[i];
        DevParams& P = params[i];
        memset(&P, 0, sizeof P);
        if (!j.ok) continue;
        const Info& I = j.I;
        JD_CUDA(cudaMemcpyAsync(d_data + j.data_off, in[i].data, I.scan_end, cudaMemcpyHostToDevice, B.s));
        std::copy(I.interval_off.begin(), I.interval_off.end(), iv_host.begin() + j.iv_off);
        P.data = d_data + j.data_off; P.interval_off = d_iv + j.iv_off; P.n_intervals = (int)I.interval_off.size() - 1;
        P.rstinterval = I.rstinterval; P.n_mcus = I.n_mcus; P.mbwidth = I.mbwidth; P.ncomp = I.ncomp;
        P.vlc = d_vlc + (size_t)j.vlc_slot * 4 * 65536; P.coef = d_coef + j.coef_off; P.planes = d_planes + j.plane_off; P.error = d_err + i;
        for (int c = 0; c < I.ncomp; ++c) {
            const Component& k = I.comp[c];
            DevComponent& d = P.comp[c];
            d.ssx = k.ssx; d.ssy = k.ssy; d.bw = k.bw; d.dctab = k.dctabsel; d.actab = k.actabsel; d.stride = k.stride;
            d.n_blocks = (unsigned long long)k.bw * k.bh; d.coef_off = k.coef_off; d.plane_off = k.plane_off;
            for (int q = 0; q < 64; ++q) d.dq[zz[q]] = I.qtab[k.qtsel][q];
            st[(size_t)i * 3 + c] = {P.planes + k.plane_off, k.width, k.height, k.stride};
        }
        if (j.n_sub) {
            P.n_intervals = 0;                        // the interval kernel passes this image by
            P.n_sub = j.n_sub; P.sub_log2 = j.sub_log2; P.bpm = mcu_block_map(I, P.blk);
            P.scan = P.data + I.scan_off; P.scan_bytes = (unsigned)(I.scan_end - I.scan_off);
            P.total_blocks = (unsigned long long)I.n_mcus * P.bpm;
            P.sub_exit = d_exit + j.sub_off; P.sub_sum = d_sums + j.sub_off; P.sub_start = d_sums + sub_total + j.sub_off;
            P.sub_list[0] = d_lists + j.sub_off; P.sub_list[1] = d_lists + sub_total + j.sub_off; P.sub_cnt = d_cnt + 3 * (size_t)i;
        }
    }
    JD_CUDA(cudaMemcpyAsync(d_iv, iv_host.data(), iv_words * 4, cudaMemcpyHostToDevice, B.s));
    JD_CUDA(cudaMemcpyAsync(d_params, params.data(), (size_t)n * sizeof(DevParams), cudaMemcpyHostToDevice, B.s));

    // rounds of batch-wide passes; the op lists stay alive until the final synchronisation (pageable copies read them then at the latest)
    // 3-component images whose two chroma planes need exactly one more vertical pass at full width: that pass is fused
    // into the colour kernel (ColorOp::chroma_h)
    std::vector<int> fused_h((size_t)n, 0);
    auto fuse_now = [&](int i) {
        const Info& I = jobs[i].I;
        if (I.ncomp != 3) return false;
        const PlaneState &a = st[(size_t)i * 3 + 1], &b = st[(size_t)i * 3 + 2];
        return a.w >= I.width && b.w >= I.width && a.h == b.h && a.h < I.height && 2 * a.h >= I.height && st[(size_t)i * 3].h >= I.height &&
               st[(size_t)i * 3].w >= I.width;
    };
    struct Round { int dir; PlaneOp* d_ops; dim3 grid; };
    std::vector<Round> rounds;
    std::vector<std::vector<PlaneOp>> keep;
    for (;;) {
        bool any = false;
        for (int dir = 0; dir < 2; ++dir) {                 // 0: horizontal, 1: vertical
            std::vector<PlaneOp> ops;
            size_t bytes = 0;
            int gw = 0, gh = 0;
            if (dir == 1)
                for (int i = 0; i < n; ++i)
                    if (jobs[i].ok && !fused_h[i] && fuse_now(i)) fused_h[i] = st[(size_t)i * 3 + 1].h;
            for (int i = 0; i < n; ++i) {
                if (!jobs[i].ok || fused_h[i]) continue;
                for (int c = 0; c < jobs[i].I.ncomp; ++c) {
                    const PlaneState& p = st[(size_t)i * 3 + c];
                    if (dir == 0 ? p.w < jobs[i].I.width : p.h < jobs[i].I.height) bytes += align256((size_t)p.w * p.h * 2);
                }
            }
            if (!bytes) continue;
            uint8_t* pool = B.alloc<uint8_t>(bytes);
            if (!pool) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
            size_t off = 0;
            for (int i = 0; i < n; ++i) {
                if (!jobs[i].ok || fused_h[i]) continue;
                for (int c = 0; c < jobs[i].I.ncomp; ++c) {
                    PlaneState& p = st[(size_t)i * 3 + c];
                    if (!(dir == 0 ? p.w < jobs[i].I.width : p.h < jobs[i].I.height)) continue;
                    uint8_t* o = pool + off;
                    off += align256((size_t)p.w * p.h * 2);
                    ops.push_back({p.p, o, p.w, p.h, p.s});
                    if (dir == 0) { gw = std::max(gw, 2 * p.w); gh = std::max(gh, p.h); p = {o, p.w << 1, p.h, p.w << 1}; }
                    else { gw = std::max(gw, p.w); gh = std::max(gh, 2 * p.h); p = {o, p.w, p.h << 1, p.w}; }
                }
            }
            PlaneOp* d_ops = B.alloc<PlaneOp>(ops.size());
            if (!d_ops) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
            keep.push_back(std::move(ops));
            const std::vector<PlaneOp>& kept = keep.back();
            JD_CUDA(cudaMemcpyAsync(d_ops, kept.data(), kept.size() * sizeof(PlaneOp), cudaMemcpyHostToDevice, B.s));
            rounds.push_back({dir, d_ops, dim3((unsigned)((gw + 511) / 512), (unsigned)std::min(gh, 65535), (unsigned)kept.size())});   // four pixels per thread; rows beyond the grid limit: strided (vertical pass)
            any = true;
        }
        if (!any) break;
    }
    std::vector<ColorOp> cops;
    int cw = 0, ch = 0;
    for (int i = 0; i < n; ++i) {
        if (!jobs[i].ok) continue;
        const Info& I = jobs[i].I;
        const PlaneState* p = &st[(size_t)i * 3];
        uint8_t* dst = pixels_on_device ? outs[i].pixels : d_out + jobs[i].out_off;
        cops.push_back({p[0].p, I.ncomp == 3 ? p[1].p : nullptr, I.ncomp == 3 ? p[2].p : nullptr, p[0].s, p[1].s, p[2].s, dst, I.width, I.height, I.ncomp,
                        fused_h[i]});
        cw = std::max(cw, I.width); ch = std::max(ch, I.height);
    }
    ColorOp* d_cops = B.alloc<ColorOp>(cops.size());
    if (!d_cops) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
    JD_CUDA(cudaMemcpyAsync(d_cops, cops.data(), cops.size() * sizeof(ColorOp), cudaMemcpyHostToDevice, B.s));
    // everything is allocated and uploaded: the kernels go out back to back (what kernel_ms measures)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) { JD_CUDA(cudaEventCreate(&e0)); JD_CUDA(cudaEventCreate(&e1)); JD_CUDA(cudaEventRecord(e0, B.s)); }
    if (max_iv) decode_intervals_kernel<<<dim3((unsigned)((max_iv + 63) / 64), (unsigned)n), 64, 0, B.s>>>(d_params);
    if (max_sub) {
        // rounds until no subsequence is put on a list any more; the host looks at the count of a group's last round only
        // (a round with empty lists costs a launch of CTAs that leave at once, so a group may overshoot)
        const dim3 grid((unsigned)((max_sub + kSubThreads - 1) / kSubThreads), (unsigned)n);
        int r = 0;
        for (int check = 0;; ++check) {
            const int group = check < 4 ? 8 : std::min(8 << (check - 3), kMaxRoundsPerCheck);     // 8 8 8 8 16 32 ... 256
            JD_CUDA(cudaMemsetAsync(d_appended, 0, (size_t)group * 4, B.s));
            for (int k = 0; k < group; ++k, ++r) sync_round_kernel<<<grid, kSubThreads, 0, B.s>>>(d_params, r, d_appended + k);
            unsigned last = 0;
            JD_CUDA(cudaMemcpyAsync(&last, d_appended + group - 1, 4, cudaMemcpyDeviceToHost, B.s));
            JD_CUDA(cudaStreamSynchronize(B.s));
            if (!last) break;
            if (r > max_sub + 2 * kMaxRoundsPerCheck) { jg::set_error_text("decode: the subsequence rounds did not settle"); return 0; }   // every round settles at least one more
        }
        sync_scan_kernel<<<(unsigned)n, 256, 0, B.s>>>(d_params);
        sync_write_kernel<<<grid, kSubThreads, 0, B.s>>>(d_params);
    }
    for (int c = 0; c < 3; ++c)
        if (max_blocks[c]) idct_kernel<<<dim3((unsigned)((max_blocks[c] + 15) / 16), (unsigned)n), kIdctThreads, 0, B.s>>>(d_params, c);   // 16 blocks per CTA
    for (const Round& r : rounds) {
        if (r.dir == 0) upsample_h_kernel<<<r.grid, 128, 0, B.s>>>(r.d_ops);
        else upsample_v_kernel<<<r.grid, 128, 0, B.s>>>(r.d_ops);
    }
    color_kernel<<<dim3((unsigned)((cw + 511) / 512), (unsigned)ch, (unsigned)cops.size()), 128, 0, B.s>>>(d_cops);
    JD_CUDA(cudaGetLastError());
    if (kernel_ms) JD_CUDA(cudaEventRecord(e1, B.s));
    std::vector<unsigned> errs((size_t)n, 0);
    JD_CUDA(cudaMemcpyAsync(errs.data(), d_err, (size_t)n * 4, cudaMemcpyDeviceToHost, B.s));
    if (!pixels_on_device)
        for (int i = 0; i < n; ++i)
            if (jobs[i].ok) JD_CUDA(cudaMemcpyAsync(outs[i].pixels, d_out + jobs[i].out_off, jobs[i].out_bytes, cudaMemcpyDeviceToHost, B.s));
    JD_CUDA(cudaStreamSynchronize(B.s));
    if (kernel_ms) { cudaEventElapsedTime(kernel_ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    int done = 0;
    for (int i = 0; i < n; ++i) {
        if (!jobs[i].ok) continue;
        if (errs[i]) { outs[i].status = JPEG_GPU_ERR_ARG; jg::set_error_text(result_text((int)errs[i])); }
        else ++done;
    }
    return done;
}

int decode(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity, int* width, int* height, int* ncomp, float* kernel_ms)
{
    if (!jpeg || !pixels) { jg::set_error_text("decode: null argument"); return 0; }
    jpeg_gpu_stream in = {jpeg, size};
    jpeg_gpu_decoded out = {pixels, capacity, 0, 0, 0, 0};
    const int ok = decode_batch(&in, 1, &out, 0, kernel_ms);
    if (width) *width = out.width;
    if (height) *height = out.height;
    if (ncomp) *ncomp = out.ncomp;
    return ok == 1 ? 1 : 0;
}

}  // namespace

extern "C" {

int jpeg_gpu_decode_info(const uint8_t* jpeg, size_t size, int* width, int* height, int* ncomp)
{
    if (!jpeg) return 0;
    jd::Info I;
    int rc = jd::parse(jpeg, size, &I, false, true);     // the headers only: walking a 600 KB scan for its markers is the decode call's business
    if (rc == jd::kOk) {
        // a file njDecode would reject (bad Huffman tables) is rejected here too; files of one encoder carry the same DHT
        // bytes, so the last verdict is remembered instead of building the 4 x 65536-entry tables for every file of a batch
        static std::mutex m;
        static std::vector<uint8_t> last_dht;
        static int last_rc = jd::kOk;
        std::lock_guard<std::mutex> lk(m);
        if (last_dht != I.dht) {
            std::vector<uint16_t> tables;
            last_rc = jd::build_vlc_tables(I.dht, &tables);
            last_dht = I.dht;
        }
        rc = last_rc;
    }
    if (rc != jd::kOk) { jg::set_error_text(result_text(rc)); return 0; }
    if (width) *width = I.width;
    if (height) *height = I.height;
    if (ncomp) *ncomp = I.ncomp;
    return 1;
}

int jpeg_gpu_decode(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity, int* width, int* height, int* ncomp)
{
    return decode(jpeg, size, pixels, capacity, width, height, ncomp, nullptr);
}

int jpeg_gpu_decode_batch(const jpeg_gpu_stream* streams, int n, jpeg_gpu_decoded* outs, int pixels_on_device, float* kernel_ms)
{
    // Host pixels out: the 3 bytes per pixel coming back are the longest stage of the call (256 x 1080p: 29 ms of download
    // against 7 - 12 ms of kernels), so a large batch is cut into chunks that three host threads take from a counter -- each
    // chunk on its own stream -- and the download of one chunk runs beside the parsing, upload and kernels of the next.
    // (With kernel_ms the batch stays in one piece: the timing is that of the kernels alone, back to back.)
    constexpr int kChunk = 32, kWorkers = 3;
    if (kernel_ms || pixels_on_device || !streams || !outs || n < 2 * kChunk) return decode_batch(streams, n, outs, pixels_on_device, kernel_ms);
    if (jpeg_gpu_device_count() == 0 && jpeg_gpu_init(nullptr, 0) <= 0) return 0;
    std::atomic<int> next{0}, done{0};
    std::mutex m;
    std::string err;
    auto work = [&] {
        for (int lo = next.fetch_add(kChunk); lo < n; lo = next.fetch_add(kChunk)) {
            const int k = std::min(kChunk, n - lo);
            const int ok = decode_batch(streams + lo, k, outs + lo, 0, nullptr);
            done += ok;
            if (ok != k) { std::lock_guard<std::mutex> lk(m); err = jpeg_gpu_last_error(); }
        }
    };
    std::vector<std::thread> pool;
    for (int k = 1; k < kWorkers; ++k) pool.emplace_back(work);
    work();
    for (std::thread& th : pool) th.join();
    if (!err.empty()) jg::set_error_text(err.c_str());
    return done.load();
}

int jpeg_gpu_decode_timed(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity, int* width, int* height, int* ncomp,
                          float* kernel_ms)
{
    return decode(jpeg, size, pixels, capacity, width, height, ncomp, kernel_ms);
}
}
