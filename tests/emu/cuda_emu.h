// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE.  A minimal "one OS thread per CUDA thread"
// emulation of the device primitives in imagecodecs_b200/csrc/jpeg_device.h, so that the
// real kernel source (jpeg_kernel.cuh) can be compiled with g++ and its control logic
// (tile bookkeeping, scans, bit packing, byte stuffing, decoupled look-back between
// concurrently running CTAs) checked against the oracle on a machine without a GPU.
// It is never linked into the product library.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#define JG_DEV inline
#define JG_DEV_NOINLINE static
#define JG_KERNEL(threads, min_ctas)
#define JG_GRID_CONSTANT
#define JG_TID (::jg::emu::tls.tid)
#define JG_CTA_ID (::jg::emu::tls.cta->id)
#define JG_GRID_DIM (::jg::emu::tls.cta->grid)
#define JG_DYNAMIC_SMEM(name) unsigned char* name = ::jg::emu::tls.cta->smem
#define JG_CONST_TABLE static const
#define JG_WARP_ANY(x) (x)
#define JG_RECONVERGE() ((void)0)

struct uint2 { unsigned x, y; };
struct float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };

namespace jg {
namespace emu {

struct Cta {
    pthread_barrier_t bar;            // all threads of the CTA
    pthread_barrier_t wbar[32];       // one per warp
    unsigned long long xch[32][32];   // warp exchange slots
    unsigned char* smem;
    int nthreads;
    int id;                           // blockIdx.x
    int grid;                         // gridDim.x
};
struct Tls { int tid; Cta* cta; };
extern thread_local Tls tls;

inline void warp_barrier() { pthread_barrier_wait(&tls.cta->wbar[tls.tid >> 5]); }

}  // namespace emu

JG_DEV float f_add(float a, float b) { return a + b; }   // TU is built with -ffp-contract=off
JG_DEV float f_sub(float a, float b) { return a - b; }
JG_DEV float f_mul(float a, float b) { return a * b; }
JG_DEV int f_floor_i(float a) { return (int)__builtin_floorf(a); }
typedef float2 f32x2;
JG_DEV f32x2 f2(float x, float y) { f32x2 r; r.x = x; r.y = y; return r; }
JG_DEV f32x2 f2_add(f32x2 a, f32x2 b) { return f2(a.x + b.x, a.y + b.y); }
JG_DEV f32x2 f2_sub(f32x2 a, f32x2 b) { return f2(a.x - b.x, a.y - b.y); }
JG_DEV f32x2 f2_mul(f32x2 a, f32x2 b) { return f2(a.x * b.x, a.y * b.y); }
JG_DEV float u8_to_f(unsigned v) { return (float)v; }
JG_DEV float u8_biased(unsigned w, int k) { const unsigned b = 0x4B000000u | ((w >> (8 * k)) & 0xffu); float f; memcpy(&f, &b, 4); return f; }
constexpr float kU8Bias = 8388608.0f;
JG_DEV float f_add_rd(float a, float b)      // binary32 sum rounded toward minus infinity (the double sum is exact for the magnitudes used)
{
    const double s = (double)a + (double)b;
    float f = (float)s;
    if ((double)f > s) f = __builtin_nextafterf(f, -__builtin_inff());
    return f;
}
JG_DEV f32x2 f2_add_rd(f32x2 a, f32x2 b) { return f2(f_add_rd(a.x, b.x), f_add_rd(a.y, b.y)); }
JG_DEV unsigned f_bits(float a) { unsigned b; memcpy(&b, &a, 4); return b; }

JG_DEV int i_clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
JG_DEV int i_ffs(unsigned v) { return __builtin_ffs((int)v); }
JG_DEV int i_popc(unsigned v) { return __builtin_popcount(v); }
JG_DEV unsigned bswap32(unsigned v) { return __builtin_bswap32(v); }
JG_DEV unsigned funnel_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (s & 31u)); }
JG_DEV unsigned byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long ab = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((ab >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
JG_DEV unsigned bit_reverse(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
JG_DEV unsigned add_min_u32(unsigned a, unsigned b, unsigned c) { return a + b < c ? a + b : c; }
JG_DEV unsigned funnel_l(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((((unsigned long long)hi) << 32) | lo) << (s & 31u)) >> 32); }
// "shared-memory addresses": offsets into the CTA's block
JG_DEV unsigned smem_addr(const void* p) { return (unsigned)((const unsigned char*)p - emu::tls.cta->smem); }
JG_DEV unsigned funnel_lc(unsigned lo, unsigned hi, unsigned s) { if (s > 32u) s = 32u; return (unsigned)((((((unsigned long long)hi) << 32) | lo) << s) >> 32); }
JG_DEV unsigned funnel_rc(unsigned lo, unsigned hi, unsigned s) { if (s > 32u) s = 32u; return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> s); }
JG_DEV void fence_proxy_async() {}
JG_DEV unsigned lds_u16(unsigned a) { unsigned short v; memcpy(&v, emu::tls.cta->smem + a, 2); return v; }
JG_DEV int lds_s16(unsigned a) { short v; memcpy(&v, emu::tls.cta->smem + a, 2); return (int)v; }
JG_DEV uint2 lds_u64(unsigned a) { uint2 v; memcpy(&v, emu::tls.cta->smem + a, 8); return v; }
JG_DEV void sts_u32(unsigned a, unsigned v) { memcpy(emu::tls.cta->smem + a, &v, 4); }
JG_DEV unsigned pinned(unsigned v) { return v; }
JG_DEV void sts_u16(unsigned a, unsigned v) { const unsigned short h = (unsigned short)v; memcpy(emu::tls.cta->smem + a, &h, 2); }
JG_DEV void sts_f32(unsigned a, float v) { memcpy(emu::tls.cta->smem + a, &v, 4); }
JG_DEV float lds_f32(unsigned a) { float v; memcpy(&v, emu::tls.cta->smem + a, 4); return v; }
JG_DEV void sts_v2f(unsigned a, f32x2 v) { memcpy(emu::tls.cta->smem + a, &v, 8); }
JG_DEV f32x2 lds_v2f(unsigned a) { f32x2 v; memcpy(&v, emu::tls.cta->smem + a, 8); return v; }
JG_DEV uint4 lds_v4(unsigned a) { uint4 v; memcpy(&v, emu::tls.cta->smem + a, 16); return v; }
JG_DEV unsigned v_minu2(unsigned a, unsigned b)
{
    const unsigned lo = (a & 0xffffu) < (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu);
    const unsigned hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return (hi << 16) | lo;
}
JG_DEV unsigned v_cmpne2(unsigned a, unsigned b)
{
    return (((a ^ b) & 0xffffu) ? 0xffffu : 0u) | (((a ^ b) >> 16) ? 0xffff0000u : 0u);
}
JG_DEV unsigned v_cmpeq4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; ++i)
        if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}

JG_DEV void cta_sync() { pthread_barrier_wait(&emu::tls.cta->bar); }

JG_DEV unsigned long long warp_exchange(unsigned long long v, int src_lane)
{
    emu::Cta* c = emu::tls.cta;
    const int w = emu::tls.tid >> 5, lane = emu::tls.tid & 31;
    c->xch[w][lane] = v;
    emu::warp_barrier();
    const unsigned long long r = (src_lane >= 0 && src_lane < 32) ? c->xch[w][src_lane] : v;
    emu::warp_barrier();
    return r;
}
JG_DEV unsigned warp_ballot(int pred)
{
    emu::Cta* c = emu::tls.cta;
    const int w = emu::tls.tid >> 5, lane = emu::tls.tid & 31;
    c->xch[w][lane] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(c->xch[w][i] & 1) << i;
    emu::warp_barrier();
    return m;
}
JG_DEV unsigned warp_shfl_u32(unsigned v, int lane) { return (unsigned)warp_exchange(v, lane); }
JG_DEV unsigned warp_shfl_up_u32(unsigned v, int d)
{
    const int lane = emu::tls.tid & 31;
    return (unsigned)warp_exchange(v, lane - d >= 0 ? lane - d : lane);
}
JG_DEV void warp_sync() { emu::warp_barrier(); }
JG_DEV unsigned warp_shfl_up_u32(unsigned v, int d);
JG_DEV unsigned warp_scan_incl_u32(unsigned v)
{
    const int lane = emu::tls.tid & 31;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned n = warp_shfl_up_u32(v, d);
        if (lane >= d) v += n;
    }
    return v;
}
JG_DEV unsigned warp_max_u32(unsigned v)
{
    for (int m = 16; m > 0; m >>= 1) {
        const unsigned o = (unsigned)warp_exchange(v, (emu::tls.tid & 31) ^ m);
        if (o > v) v = o;
    }
    return v;
}
JG_DEV unsigned warp_sum_u32(unsigned v)
{
    for (int m = 16; m > 0; m >>= 1) v += (unsigned)warp_exchange(v, (emu::tls.tid & 31) ^ m);
    return v;
}
JG_DEV float warp_shfl_xor_f32(float v, int m)
{
    unsigned bits; memcpy(&bits, &v, 4);
    bits = (unsigned)warp_exchange(bits, (emu::tls.tid & 31) ^ m);
    float r; memcpy(&r, &bits, 4); return r;
}
JG_DEV unsigned long long warp_shfl_u64(unsigned long long v, int lane) { return warp_exchange(v, lane); }
JG_DEV unsigned long long warp_shfl_xor_u64(unsigned long long v, int m)
{
    return warp_exchange(v, (emu::tls.tid & 31) ^ m);
}

JG_DEV uint32_t ldg_u32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
JG_DEV uint2 ldg_u64(const void* p) { uint2 v; memcpy(&v, p, 8); return v; }
JG_DEV uint4 ldg_u128(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
JG_DEV uint32_t ldg_u8(const void* p) { return *(const unsigned char*)p; }
JG_DEV int ldg_s16(const void* p) { short v; memcpy(&v, p, 2); return (int)v; }
// bulk async copies: done on the spot by the issuing thread; "waiting for the barrier" = every lane of the warp has issued
JG_DEV void mbar_init(unsigned long long*, unsigned) {}
JG_DEV void mbar_fence_init() {}
JG_DEV void mbar_expect_tx(unsigned long long*, unsigned) {}
JG_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
// the coefficient plane's TMA box: 32 rows of 64 int16 from row c1; rows past the end are zeros
JG_DEV void tma_load_2d(void* dst, const void* tmap, int c0, int c1, unsigned long long*)
{
    const unsigned long long* q = (const unsigned long long*)tmap;
    const short* base = (const short*)q[0];
    const long long rows = (long long)q[1];
    short* d = (short*)dst;
    for (int r = 0; r < 32; ++r)
        for (int e = 0; e < 72; ++e)
            d[r * 72 + e] = (c0 + e < 64 && (long long)c1 + r < rows) ? base[((long long)c1 + r) * 64 + c0 + e] : (short)0;
}
JG_DEV void mbar_wait(unsigned long long*, unsigned) { emu::warp_barrier(); }
JG_DEV void smem_atomic_or(unsigned* p, unsigned v) { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
JG_DEV unsigned gmem_atomic_add(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
JG_DEV void gmem_atomic_or(unsigned* p, unsigned v) { __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
JG_DEV unsigned long long ld_flag64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
JG_DEV void st_flag64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
JG_DEV void st_flag32(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
JG_DEV unsigned ld_flag32(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
JG_DEV void backoff() { sched_yield(); }

}  // namespace jg
