// jpeg_decode_api.cpp -- C entry points of the decoder (include/jpeg_gpu.h, "decode" section):
// parse on the host, everything else on the GPU, pixels back to the caller.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
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

struct DeviceBuffers {     // freed in order on every exit path
    cudaStream_t s = nullptr;
    std::vector<void*> ptrs;
    template <typename T>
    T* alloc(size_t n)
    {
        void* p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(n * sizeof(T), 16), s) != cudaSuccess) return nullptr;
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

int decode(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity, int* width, int* height, int* ncomp, float* kernel_ms)
{
    if (!jpeg || !pixels) { jg::set_error_text("decode: null argument"); return 0; }
    Info I;
    const int rc = parse(jpeg, size, &I);
    if (rc != kOk) { jg::set_error_text(result_text(rc)); return 0; }
    if (width) *width = I.width;
    if (height) *height = I.height;
    if (ncomp) *ncomp = I.ncomp;
    const size_t out_bytes = (size_t)I.width * I.height * I.ncomp;
    if (capacity < out_bytes) { jg::set_error_text("decode: output buffer too small"); return 0; }
    if (jpeg_gpu_device_count() == 0 && jpeg_gpu_init(nullptr, 0) <= 0) return 0;
    JD_CUDA(cudaSetDevice(jg::cuda_device_of(0)));

    DeviceBuffers B;
    JD_CUDA(cudaStreamCreateWithFlags(&B.s, cudaStreamNonBlocking));
    const int n_iv = (int)I.interval_off.size() - 1;
    uint8_t* d_data = B.alloc<uint8_t>(I.scan_end + 16);
    uint32_t* d_iv = B.alloc<uint32_t>(I.interval_off.size());
    uint16_t* d_vlc = B.alloc<uint16_t>(I.vlc.size());
    int16_t* d_coef = B.alloc<int16_t>(I.n_blocks * 64);
    uint8_t* d_planes = B.alloc<uint8_t>(I.plane_bytes);
    uint8_t* d_out = B.alloc<uint8_t>(out_bytes);
    unsigned* d_err = B.alloc<unsigned>(1);
    if (!d_data || !d_iv || !d_vlc || !d_coef || !d_planes || !d_out || !d_err) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
    JD_CUDA(cudaMemcpyAsync(d_data, jpeg, I.scan_end, cudaMemcpyHostToDevice, B.s));
    JD_CUDA(cudaMemcpyAsync(d_iv, I.interval_off.data(), I.interval_off.size() * 4, cudaMemcpyHostToDevice, B.s));
    JD_CUDA(cudaMemcpyAsync(d_vlc, I.vlc.data(), I.vlc.size() * 2, cudaMemcpyHostToDevice, B.s));
    JD_CUDA(cudaMemsetAsync(d_coef, 0, I.n_blocks * 64 * 2, B.s));
    JD_CUDA(cudaMemsetAsync(d_err, 0, 4, B.s));

    DevParams P;
    memset(&P, 0, sizeof P);
    P.data = d_data; P.interval_off = d_iv; P.n_intervals = n_iv; P.rstinterval = I.rstinterval; P.n_mcus = I.n_mcus;
    P.mbwidth = I.mbwidth; P.ncomp = I.ncomp; P.vlc = d_vlc; P.coef = d_coef; P.planes = d_planes; P.error = d_err;
    for (int c = 0; c < I.ncomp; ++c) {
        const Component& k = I.comp[c];
        DevComponent& d = P.comp[c];
        d.ssx = k.ssx; d.ssy = k.ssy; d.bw = k.bw; d.dctab = k.dctabsel; d.actab = k.actabsel; d.stride = k.stride;
        d.coef_off = k.coef_off; d.plane_off = k.plane_off;
        static const unsigned char zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                                             28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54,
                                             47, 55, 62, 63};
        for (int i = 0; i < 64; ++i) d.dq[zz[i]] = I.qtab[k.qtsel][i];
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) { JD_CUDA(cudaEventCreate(&e0)); JD_CUDA(cudaEventCreate(&e1)); JD_CUDA(cudaEventRecord(e0, B.s)); }

    decode_intervals_kernel<<<(n_iv + 63) / 64, 64, 0, B.s>>>(P);
    for (int c = 0; c < I.ncomp; ++c) {
        const unsigned long long nb = (unsigned long long)I.comp[c].bw * I.comp[c].bh;
        idct_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, B.s>>>(P, c, nb);
    }
    // njConvert (:817-836): every component is brought to the image size, H before V
    const uint8_t* plane[3]; int pw[3], ph[3], ps[3];
    for (int c = 0; c < I.ncomp; ++c) {
        plane[c] = d_planes + I.comp[c].plane_off; pw[c] = I.comp[c].width; ph[c] = I.comp[c].height; ps[c] = I.comp[c].stride;
        while (pw[c] < I.width || ph[c] < I.height) {
            if (pw[c] < I.width) {
                uint8_t* o = B.alloc<uint8_t>((size_t)pw[c] * ph[c] * 2);
                if (!o) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
                upsample_h_kernel<<<dim3((2 * pw[c] + 127) / 128, ph[c]), 128, 0, B.s>>>(plane[c], o, pw[c], ph[c], ps[c]);
                plane[c] = o; pw[c] <<= 1; ps[c] = pw[c];
            }
            if (ph[c] < I.height) {
                uint8_t* o = B.alloc<uint8_t>((size_t)pw[c] * ph[c] * 2);
                if (!o) { jg::set_error_text(result_text(kOutOfMem)); return 0; }
                upsample_v_kernel<<<dim3((pw[c] + 127) / 128, 2 * ph[c]), 128, 0, B.s>>>(plane[c], o, pw[c], ph[c], ps[c]);
                plane[c] = o; ph[c] <<= 1; ps[c] = pw[c];
            }
        }
    }
    const dim3 grid((I.width + 127) / 128, I.height);
    if (I.ncomp == 3) to_rgb_kernel<<<grid, 128, 0, B.s>>>(plane[0], ps[0], plane[1], ps[1], plane[2], ps[2], d_out, I.width, I.height);
    else to_gray_kernel<<<grid, 128, 0, B.s>>>(plane[0], ps[0], d_out, I.width, I.height);
    JD_CUDA(cudaGetLastError());
    if (kernel_ms) JD_CUDA(cudaEventRecord(e1, B.s));
    unsigned err = 0;
    JD_CUDA(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, B.s));
    JD_CUDA(cudaMemcpyAsync(pixels, d_out, out_bytes, cudaMemcpyDeviceToHost, B.s));
    JD_CUDA(cudaStreamSynchronize(B.s));
    if (kernel_ms) { cudaEventElapsedTime(kernel_ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    if (err) { jg::set_error_text(result_text((int)err)); return 0; }
    return 1;
}

}  // namespace

extern "C" {

int jpeg_gpu_decode_info(const uint8_t* jpeg, size_t size, int* width, int* height, int* ncomp)
{
    if (!jpeg) return 0;
    jd::Info I;
    const int rc = jd::parse(jpeg, size, &I);
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

int jpeg_gpu_decode_timed(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity, int* width, int* height, int* ncomp,
                          float* kernel_ms)
{
    return decode(jpeg, size, pixels, capacity, width, height, ncomp, kernel_ms);
}
}
