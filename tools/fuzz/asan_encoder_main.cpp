// The encode kernel SOURCE run by the CPU thread emulator (tests/emu) under AddressSanitizer: every global
// buffer of emu_encode is an exact-size heap block, so reads / writes outside pixels, raw scan, output,
// descriptors or a CTA's shared-memory block are reported.  (compute-sanitizer is closed on the GPU pool.)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
extern "C" int emu_encode(const uint8_t* pixels, int n_images, int w, int h, int ncomp, int stride, int flags, int subsampling,
                          int quality_mode, int quality, int win_words, int n_ctas,
                          uint8_t* scan_out, size_t scan_cap, unsigned long long* scan_bytes, unsigned* img_status,
                          int16_t* dbg_coefs, uint32_t* dbg_bits);
int main()
{
    struct Case { int n, w, h, nc, sub, qm, q, flags, win, ctas, noise; };
    const Case cases[] = {
        {1, 17, 13, 3, 0, 0, 3, 0, 0, 1, 0},   {2, 200, 120, 3, 0, 0, 2, 0, 0, 3, 0},  {1, 131, 67, 3, 0, 0, 3, 0, 0, 2, 0},
        {1, 96, 96, 3, 0, 0, 3, 0, 0, 2, 1},   {1, 96, 64, 3, 0, 0, 3, 0, 216, 2, 1},  {2, 120, 72, 3, 1, 1, 75, 0, 0, 2, 0},
        {1, 33, 47, 4, 1, 1, 90, 0, 0, 2, 0},  {2, 200, 130, 1, 0, 1, 85, 0, 0, 2, 0}, {1, 8, 8, 1, 0, 1, 85, 0, 0, 1, 0},
        {1, 176, 240, 3, 1, 1, 90, 0, 0, 1, 2}, {1, 176, 240, 3, 0, 0, 3, 0, 0, 1, 2},  {3, 64, 48, 4, 0, 0, 2, 2, 0, 1, 0},
        {1, 200, 120, 3, 0, 0, 3, 2, 0, 2, 0},  {1, 96, 96, 3, 0, 0, 3, 2, 0, 2, 1},    {1, 45, 37, 3, 0, 0, 3, 1, 0, 1, 0},
        {1, 64, 32, 3, 1, 1, 75, 1, 0, 1, 0},   {1, 300, 9, 3, 1, 1, 50, 0, 0, 2, 0},   {1, 9, 300, 1, 0, 1, 95, 2, 0, 2, 0},
    };
    unsigned seed = 12345;
    for (const Case& c : cases) {
        const size_t px_bytes = (size_t)c.n * c.w * c.h * c.nc;
        uint8_t* px = (uint8_t*)malloc(px_bytes);
        for (size_t i = 0; i < px_bytes; ++i) {
            seed = seed * 1664525u + 1013904223u;
            const size_t row = (i / ((size_t)c.w * c.nc)) % c.h;
            const bool noisy = c.noise == 1 || (c.noise == 2 && row > (size_t)c.h / 3 && row < 2 * (size_t)c.h / 3);
            px[i] = noisy ? (uint8_t)(seed >> 24) : (uint8_t)((i * 3 + (seed >> 28)) & 255);
        }
        const int mcu = (c.sub && c.nc != 1) ? 16 : 8, bpm = c.nc == 1 ? 1 : (c.sub ? 6 : 3);
        const size_t nblk = (size_t)((c.w + mcu - 1) / mcu) * ((c.h + mcu - 1) / mcu) * bpm;
        const size_t cap = nblk * 420 + 64;
        uint8_t* out = (uint8_t*)malloc(cap * c.n);
        std::vector<unsigned long long> sizes(c.n);
        std::vector<unsigned> status(c.n);
        const int rc = emu_encode(px, c.n, c.w, c.h, c.nc, 0, c.flags, c.sub, c.qm, c.q, c.win, c.ctas, out, cap, sizes.data(), status.data(), nullptr, nullptr);
        printf("%dx%dx%d n=%d sub=%d q=%d/%d flags=%d win=%d: rc=%d bytes=%llu status=%u\n", c.w, c.h, c.nc, c.n, c.sub, c.qm, c.q, c.flags, c.win, rc,
               sizes[0], status[0]);
        free(px); free(out);
        if (rc != 0) return 1;
    }
    puts("no sanitizer report");
    return 0;
}
