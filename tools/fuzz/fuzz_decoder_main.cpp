#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
extern "C" int emu_decode(const uint8_t* jpeg, size_t size, uint8_t* out, size_t cap, int* w, int* h, int* ncomp);
int main(int argc, char** argv) {
    int ok = 0, bad = 0;
    for (int i = 1; i < argc; ++i) {
        FILE* f = fopen(argv[i], "rb"); if (!f) continue;
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        uint8_t* buf = (uint8_t*)malloc(n ? n : 1);             // exact-size heap block: ASan sees any overread
        if (fread(buf, 1, n, f) != (size_t)n) { fclose(f); free(buf); continue; }
        fclose(f);
        int w = 0, h = 0, nc = 0;
        uint8_t dummy;
        int rc = emu_decode(buf, (size_t)n, &dummy, 0, &w, &h, &nc);
        if (rc == -1 && (size_t)w * h * nc < (64u << 20)) {
            std::vector<uint8_t> out((size_t)w * h * nc);
            rc = emu_decode(buf, (size_t)n, out.data(), out.size(), &w, &h, &nc);
        }
        if (rc == 0) ++ok; else ++bad;
        free(buf);
    }
    printf("decoded %d rejected %d\n", ok, bad);
    return 0;
}
