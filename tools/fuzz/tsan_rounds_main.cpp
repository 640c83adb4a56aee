#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
extern "C" int emu_decode_sub(const uint8_t* jpeg, size_t size, uint8_t* out, size_t cap, int* w, int* h, int* ncomp, int sub_log2, int* rounds);
extern "C" void emu_set_round_order(int o);
int main(int argc, char** argv) {
    emu_set_round_order(3);
    for (int i = 1; i < argc; ++i) {
        FILE* f = fopen(argv[i], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<uint8_t> buf(n); if (fread(buf.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
        int w = 0, h = 0, nc = 0, rounds = 0; uint8_t d;
        emu_decode_sub(buf.data(), n, &d, 0, &w, &h, &nc, 5, nullptr);
        std::vector<uint8_t> out((size_t)w * h * nc);
        int rc = emu_decode_sub(buf.data(), n, out.data(), out.size(), &w, &h, &nc, 5, &rounds);
        printf("%s: rc %d %dx%dx%d rounds %d\n", argv[i], rc, w, h, nc, rounds);
    }
}
