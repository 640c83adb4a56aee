/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin memory-sink wrapper around the UNMODIFIED reference encoder/decoder.  The
 * reference sources are compiled from where they lie (-I/root/reference); nothing
 * is copied into this repository.  The result goes to oracle/_ref/libtje_ref.so
 * (git-ignored, travels to the GPU box as a built artefact).
 *
 *   jpeg_enc.h  -> TinyJPEG  (tje_encode_with_func, jpeg_enc.h:1215-1271)
 *   jpeg_dec.h  -> NanoJPEG  (njInit/njDecode/njGetImage, jpeg_dec.h:130-171)
 *
 * Build flags matter: -O2 -ffp-contract=off.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define TJE_IMPLEMENTATION
#include "jpeg_enc.h"
#include "jpeg_dec.h"

typedef struct {
    uint8_t* data;
    size_t   cap;
    size_t   size;     /* bytes the encoder produced (may exceed cap) */
} ref_sink;

static void ref_sink_write(void* ctx, void* data, int n)
{
    ref_sink* s = (ref_sink*)ctx;
    if (s->size + (size_t)n <= s->cap) {
        memcpy(s->data + s->size, data, (size_t)n);
    }
    s->size += (size_t)n;
}

/* Returns the tje return code (1 ok / 0 error); *out_size = bytes produced. */
int ref_tje_encode_mem(int quality, int w, int h, int ncomp, const uint8_t* src,
                       uint8_t* out, size_t cap, size_t* out_size)
{
    ref_sink s;
    s.data = out; s.cap = cap; s.size = 0;
    int rc = tje_encode_with_func(ref_sink_write, &s, quality, w, h, ncomp, src);
    if (out_size) *out_size = s.size;
    return rc;
}

/* The file-API variant, so tests can pin its odd return-code behaviour
 * (jpeg_enc.h:1194-1213). */
int ref_tje_encode_file(const char* path, int quality, int w, int h, int ncomp,
                        const uint8_t* src)
{
    return tje_encode_to_file_at_quality(path, quality, w, h, ncomp, src);
}

/* NanoJPEG decode into caller memory. Returns 0 on success (nj_result_t),
 * negative if the caller buffer is too small. Not thread-safe (jpeg_dec.h:332). */
int ref_nj_decode(const uint8_t* jpeg, int size, uint8_t* out, size_t cap,
                  int* w, int* h, int* is_color)
{
    njInit();
    int rc = (int)njDecode(jpeg, size);
    if (rc == 0) {
        size_t n = (size_t)njGetImageSize();
        if (w) *w = njGetWidth();
        if (h) *h = njGetHeight();
        if (is_color) *is_color = njIsColor();
        if (n <= cap) memcpy(out, njGetImage(), n);
        else rc = -1;
    }
    njDone();
    return rc;
}
