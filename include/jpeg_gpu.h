/*
 * jpeg_gpu.h -- C ABI of the B200 (sm_100a) baseline JPEG encoder.
 *
 * This is the drop-in boundary for the JPEG write path of jstrom2002/ImageCodecs:
 *
 *     Image::write(".jpg")            codecs.cpp:106-107
 *       -> Image::writeJpg            codecs.cpp:851-854   (decl codecs.h:56)
 *         -> tje_encode_to_file       jpeg_enc.h:114-118 / :1177-1185
 *           -> tje_encode_to_file_at_quality   jpeg_enc.h:137-142 / :1194-1213
 *             -> tje_encode_with_func          jpeg_enc.h:154-160 / :1215-1271
 *
 * Every entry point is plain C: pointers, sizes, ints.  No C++ or torch types.
 * The library is self-contained (static cudart); it FAILS (returns 0 and sets
 * jpeg_gpu_last_error) when no CUDA device is usable -- there is no CPU fallback.
 *
 * Output contract: for the reference's native modes (JPEG_GPU_QMODE_TJE quality
 * 1..3, 3 or 4 channels, 4:4:4) the produced file is byte-identical to what
 * jpeg_enc.h writes for the same pixels.  Extended modes (IJG quality 1..100,
 * 4:2:0, 1-channel grayscale) are defined in DESIGN.md; IJG 50 == TJE 1 and
 * IJG 100 == TJE 3 byte for byte.
 */
#ifndef JPEG_GPU_H
#define JPEG_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define JPEG_GPU_API __declspec(dllexport)
#else
#define JPEG_GPU_API __attribute__((visibility("default")))
#endif

/* quality_mode */
#define JPEG_GPU_QMODE_TJE 0 /* quality 1..3 exactly as jpeg_enc.h:1231-1256 (3 = what writeJpg uses) */
#define JPEG_GPU_QMODE_IJG 1 /* quality 1..100, IJG scaling of the reference's two base tables (extended) */

/* subsampling */
#define JPEG_GPU_SUB_444 0 /* the only format the reference emits (jpeg_enc.h:1038) */
#define JPEG_GPU_SUB_420 1 /* extended: 16x16 MCUs, Y00 Y01 Y10 Y11 Cb Cr */

/* jpeg_gpu_image.flags: load-time swizzles, applied by the encode kernel while it reads the pixels
 * (no host pass over the pixels for BGR / bottom-up sources such as BMP files) */
#define JPEG_GPU_FLAG_SWAP_RB 1 /* channels 0 and 2 of every pixel are exchanged: the bytes produced are those of
                                   Image::swapBR() (codecs.cpp:193-251) followed by writeJpg.  Flagged images are read
                                   with byte loads (the transform runs ~1.5x slower); unflagged ones pay nothing */
#define JPEG_GPU_FLAG_RESTART 2 /* EXTENDED, opt-in (not in the reference encoder, so the bytes differ from
                                   jpeg_enc.h's): restart intervals of one tile = 8 (4:4:4) / 4 (4:2:0) / 24 (gray) MCUs.
                                   A DRI segment precedes SOS; every interval is padded to a byte boundary with 1-bits,
                                   RSTm follows (none after the last), DC prediction restarts at 0.  The stream decodes
                                   to exactly the pixels of the restart-free stream (checked with jpeg_dec.h) and its
                                   intervals can be decoded in parallel.  Bytes are defined by oracle/jpeg_oracle.c */

/* jpeg_gpu_output.status */
#define JPEG_GPU_OK 0
#define JPEG_GPU_ERR_ARG 1      /* rejected like jpeg_enc.h:954-960 / :1223-1226 would */
#define JPEG_GPU_ERR_CAPACITY 2 /* output buffer too small; `size` holds the bytes needed */
#define JPEG_GPU_ERR_CUDA 3     /* CUDA failure; see jpeg_gpu_last_error() */

/* The callback type of tje_encode_with_func (jpeg_enc.h:152). */
typedef void jpeg_gpu_write_func(void* context, void* data, int size);

typedef struct jpeg_gpu_image {
    const uint8_t* pixels; /* interleaved 8-bit, row-major, top-down (jpeg_enc.h:1101) */
    int width;             /* 1..65535 (jpeg_enc.h:958) */
    int height;            /* 1..65535 */
    int ncomp;             /* 3 = RGB, 4 = RGBA (alpha skipped), 1 = gray (extended) */
    int stride;            /* bytes from one row to the next; 0 means width*ncomp (the reference's only layout).
                              Negative: the rows are stored bottom-up (BMP, DIB): `pixels` still points at the
                              image's TOP row, which then is the last one in memory -- the bytes produced are
                              those of Image::flip() (codecs.cpp:162-191) on the bottom-up buffer + writeJpg */
    int quality_mode;      /* JPEG_GPU_QMODE_* */
    int quality;           /* 1..3 or 1..100 by mode */
    int subsampling;       /* JPEG_GPU_SUB_* */
    int pixels_on_device;  /* non-zero: `pixels` is a device pointer on the GPU that encodes this image */
    int flags;             /* JPEG_GPU_FLAG_* */
} jpeg_gpu_image;

typedef struct jpeg_gpu_output {
    uint8_t* data;   /* caller-owned buffer (host, or device if opts->outputs_on_device) */
    size_t capacity; /* bytes available at `data` */
    size_t size;     /* OUT: bytes of the complete JPEG file (also set on ERR_CAPACITY) */
    int status;      /* OUT: JPEG_GPU_OK or an error above */
} jpeg_gpu_output;

typedef struct jpeg_gpu_batch_opts {
    int device;            /* >=0: index into the initialised device list; -1: shard by image index */
    int outputs_on_device; /* outs[i].data are device pointers on the encoding GPU */
    void* stream;          /* optional cudaStream_t to launch on (only with device >= 0) */
    int debug_window_words;/* 0 = default; otherwise force the words of bits a tile may hold before it goes down the
                              slow path (tests; clamped to 216..384) */
} jpeg_gpu_batch_opts;

/* ---- lifetime ---------------------------------------------------------- */

/* Bind the encoder to `n_devices` CUDA devices (NULL/0: every visible device).
 * Returns the number of devices in use, 0 on failure.  Idempotent. */
JPEG_GPU_API int jpeg_gpu_init(const int* device_ids, int n_devices);
JPEG_GPU_API void jpeg_gpu_shutdown(void);
JPEG_GPU_API int jpeg_gpu_device_count(void);
/* Thread-local description of the last failure ("" if none). */
JPEG_GPU_API const char* jpeg_gpu_last_error(void);

/* ---- sizing / host-side marker emission -------------------------------- */

/* Upper bound of the encoded size for any pixel content. */
JPEG_GPU_API size_t jpeg_gpu_max_encoded_size(int width, int height, int ncomp, int subsampling);

/* SOI, APP0, COM, DQT, SOF0, DHT, SOS exactly as jpeg_enc.h:989-1077 lays them out
 * (655 bytes for the native modes).  Returns bytes written, 0 if rejected/too small. */
JPEG_GPU_API size_t jpeg_gpu_emit_headers(int width, int height, int ncomp, int quality_mode,
                                          int quality, int subsampling, uint8_t* out, size_t capacity);
/* The same for an image with flags (JPEG_GPU_FLAG_RESTART adds the DRI segment). */
JPEG_GPU_API size_t jpeg_gpu_emit_headers_for(const jpeg_gpu_image* image, uint8_t* out, size_t capacity);

/* ---- batch encode ------------------------------------------------------ */

/* Encode n images.  With opts->device == -1 the batch is split into contiguous index
 * ranges, one per initialised GPU (image i goes to GPU i*G/n), each GPU working on its
 * own stream; no inter-GPU communication takes place.  Returns the number of images
 * whose status is JPEG_GPU_OK. */
JPEG_GPU_API int jpeg_gpu_encode_batch(const jpeg_gpu_image* images, int n, jpeg_gpu_output* outs,
                                       const jpeg_gpu_batch_opts* opts);

/* ---- plans: a prepared batch whose device work can be re-launched ------- */
/* A plan owns the per-image tables, tile maps, scratch and a device output arena for a
 * fixed list of images on ONE device.  jpeg_gpu_plan_run only enqueues kernels (no
 * host synchronisation), so it can be timed with CUDA events or captured in a graph. */
typedef struct jpeg_gpu_plan jpeg_gpu_plan;

JPEG_GPU_API jpeg_gpu_plan* jpeg_gpu_plan_create(const jpeg_gpu_image* images, int n, int device,
                                                 int debug_window_words);
/* Point image i at new pixels of the same geometry (device pointer). */
JPEG_GPU_API int jpeg_gpu_plan_set_pixels(jpeg_gpu_plan* plan, int i, const uint8_t* device_pixels);
/* Upload host pixels for image i into plan-owned device memory (async on `stream`). */
JPEG_GPU_API int jpeg_gpu_plan_upload(jpeg_gpu_plan* plan, int i, const uint8_t* host_pixels, void* stream);
/* Enqueue the encode of every image of the plan on `stream` (cudaStream_t, may be NULL). */
JPEG_GPU_API int jpeg_gpu_plan_run(jpeg_gpu_plan* plan, void* stream);
/* With timing enabled every run records CUDA events around its kernels; after the run has
 * completed, kernel_times returns the device time of the encode kernels (pass 1) and of the
 * plan+stuff kernels (pass 2) of the LAST run, in milliseconds.  Returns 0 if unavailable. */
JPEG_GPU_API int jpeg_gpu_plan_enable_timing(jpeg_gpu_plan* plan, int enable);
JPEG_GPU_API int jpeg_gpu_plan_kernel_times(jpeg_gpu_plan* plan, float* encode_ms, float* stuff_ms);
/* The same per pass of the split pipeline: pass A (pixels -> coefficients, jpeg_transform.cuh), pass B (coefficients ->
 * unstuffed bits, jpeg_entropy.cuh), pass 2 (plan + stuff).  With the fused round-1 kernel (JPEG_GPU_PIPELINE=fused in
 * the environment when the plan is created; same bytes) entropy_ms is 0 and transform_ms is the fused kernel. */
JPEG_GPU_API int jpeg_gpu_plan_pass_times(jpeg_gpu_plan* plan, float* transform_ms, float* entropy_ms, float* stuff_ms);
JPEG_GPU_API int jpeg_gpu_plan_is_fused(const jpeg_gpu_plan* plan);
/* Number of kernel launches one jpeg_gpu_plan_run enqueues. */
JPEG_GPU_API int jpeg_gpu_plan_launches(const jpeg_gpu_plan* plan);
/* Wait for `stream`, then deliver headers + scans into outs[] (host or device buffers). */
JPEG_GPU_API int jpeg_gpu_plan_fetch(jpeg_gpu_plan* plan, jpeg_gpu_output* outs, int outputs_on_device,
                                     void* stream);
/* After a run has completed: encoded size of image i (whole file), 0 on error. */
JPEG_GPU_API size_t jpeg_gpu_plan_encoded_size(jpeg_gpu_plan* plan, int i);
/* Total blocks (8x8 data units) of the plan, in plan order. */
JPEG_GPU_API size_t jpeg_gpu_plan_num_blocks(const jpeg_gpu_plan* plan);
/* Stage dumps for parity tests: attach BEFORE plan_run; the kernel then also writes
 * coefs[nblocks*64] (int16, zigzag order, stream order) and block_bits[nblocks].
 * Both are device pointers or NULL. */
JPEG_GPU_API int jpeg_gpu_plan_attach_debug(jpeg_gpu_plan* plan, int16_t* dev_coefs, uint32_t* dev_block_bits);
JPEG_GPU_API void jpeg_gpu_plan_destroy(jpeg_gpu_plan* plan);

/* ---- drop-in twins of the reference's three entry points ---------------- */
/* Same argument meaning, same 1/0 return convention (jpeg_enc.h:111-112), same
 * callback contract: `func` is called on the calling thread, synchronously, with
 * chunks the encoder owns (jpeg_enc.h:147-150, :487-490).  Unlike the reference's file
 * variants, a rejected encode returns 0 instead of 1 (jpeg_enc.h:1210 is a bug we do
 * not reproduce); the bytes of every successful encode are identical. */

/* replaces tje_encode_to_file, jpeg_enc.h:114-118 (quality fixed at 3, :1183) */
JPEG_GPU_API int jpeg_gpu_encode_to_file(const char* dest_path, const int width, const int height,
                                         const int num_components, const unsigned char* src_data);
/* replaces tje_encode_to_file_at_quality, jpeg_enc.h:137-142 */
JPEG_GPU_API int jpeg_gpu_encode_to_file_at_quality(const char* dest_path, const int quality, const int width,
                                                    const int height, const int num_components,
                                                    const unsigned char* src_data);
/* replaces tje_encode_with_func, jpeg_enc.h:154-160 */
JPEG_GPU_API int jpeg_gpu_encode_with_func(jpeg_gpu_write_func* func, void* context, const int quality,
                                           const int width, const int height, const int num_components,
                                           const unsigned char* src_data);

/* ---- decode: the step on the other side of the format --------------- */
/* What Image::readJpg does with NanoJPEG (codecs.cpp:821-849: njInit, njDecode, njGetWidth / njGetHeight,
 * njGetImage; jpeg_dec.h:880-908), on the GPU: baseline JPEG, 1 or 3 components, power-of-two sampling
 * factors.  The pixels (RGB interleaved, or 8-bit gray) are bit-identical to NanoJPEG's, including its
 * chroma upsampling filter.  The entropy decode is parallel either way: one thread per restart interval
 * where the file has them (JPEG_GPU_FLAG_RESTART on the encode side), one thread per 32..128-byte subsequence
 * of the scan where it has none (everything jpeg_enc.h writes) -- speculative decodes that are repeated until
 * neighbouring subsequences agree (self-synchronisation of the Huffman code).  Only scans with irregular
 * marker bytes or more than 16 blocks per MCU are left to a single thread (correct, slow).
 * Return 1 on success, 0 on failure (jpeg_gpu_last_error names the nj_result_t). */
JPEG_GPU_API int jpeg_gpu_decode_info(const uint8_t* jpeg, size_t size, int* width, int* height, int* ncomp);
JPEG_GPU_API int jpeg_gpu_decode(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity,
                                 int* width, int* height, int* ncomp);
/* Many files per call: every kernel runs over the whole batch (image = one grid dimension).  outs[i].pixels /
 * capacity are the caller's buffers (device pointers if pixels_on_device); width, height, ncomp, status are filled
 * in.  kernel_ms (optional) receives the device time of the batch's kernels.  Returns the number decoded. */
typedef struct jpeg_gpu_stream { const uint8_t* data; size_t size; } jpeg_gpu_stream;
typedef struct jpeg_gpu_decoded { uint8_t* pixels; size_t capacity; int width, height, ncomp, status; } jpeg_gpu_decoded;
JPEG_GPU_API int jpeg_gpu_decode_batch(const jpeg_gpu_stream* streams, int n, jpeg_gpu_decoded* outs,
                                       int pixels_on_device, float* kernel_ms);
/* jpeg_gpu_decode; *kernel_ms receives the device time of the kernels (CUDA events), copies excluded.
 * All decode entry points run on the first initialised GPU. */
JPEG_GPU_API int jpeg_gpu_decode_timed(const uint8_t* jpeg, size_t size, uint8_t* pixels, size_t capacity,
                                       int* width, int* height, int* ncomp, float* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* JPEG_GPU_H */
