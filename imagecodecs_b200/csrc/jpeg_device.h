// jpeg_device.h -- the handful of device primitives the encode kernel is written against.
//
// Under nvcc these are the real sm_100a intrinsics.  When JG_EMULATE is defined (tests only,
// see tests/emu/) the same names are provided by a tiny thread-per-CUDA-thread emulator so
// that the control logic of the kernel (tile bookkeeping, scans, bit packing, stuffing,
// look-back) can be exercised on a machine without a GPU.  The emulation is test
// infrastructure; nothing in the product links against it.
#pragma once
#include <stdint.h>

#if defined(JG_EMULATE)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>

#define JG_DEV __device__ __forceinline__
#define JG_DEV_NOINLINE __device__ __noinline__
#define JG_TID ((int)threadIdx.x)
#define JG_CTA_ID ((int)blockIdx.x)
#define JG_GRID_DIM ((int)gridDim.x)
#define JG_KERNEL(threads, min_ctas) __global__ __launch_bounds__(threads, min_ctas)
#define JG_GRID_CONSTANT __grid_constant__
#define JG_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define JG_CONST_TABLE static __device__ __constant__
#define JG_WARP_ANY(x) __any_sync(0xffffffffu, (x))   // every lane of the warp calls it
#define JG_RECONVERGE() __syncwarp()      // all 32 lanes arrive: brings lanes that drifted apart in data-dependent loops back together

namespace jg {

// ---- exactly-rounded binary32 arithmetic: never contracted into FMA ----------------------
JG_DEV float f_add(float a, float b) { return __fadd_rn(a, b); }
JG_DEV float f_sub(float a, float b) { return __fsub_rn(a, b); }
JG_DEV float f_mul(float a, float b) { return __fmul_rn(a, b); }
JG_DEV int f_floor_i(float a) { return __float2int_rd(a); }   // (int)floorf(a)
// Two binary32 values per instruction (sm_100 FADD2).  Only additions are ever packed: ptxas fuses
// a packed multiply with a following packed add into FFMA2 even when both carry .rn (checked on
// CUDA 12.9, also with __fmul2_rn / __fadd2_rn), which would change the rounding.  x - y is
// x + (-y) (bit-identical; the negation folds into the operand).  build.py greps the SASS for FFMA.
typedef float2 f32x2;
JG_DEV f32x2 f2(float x, float y) { return make_float2(x, y); }
JG_DEV f32x2 f2_add(f32x2 a, f32x2 b) { return __fadd2_rn(a, b); }
JG_DEV f32x2 f2_sub(f32x2 a, f32x2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
JG_DEV f32x2 f2_mul(f32x2 a, f32x2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }   // two scalar FMULs, on purpose
JG_DEV float u8_to_f(unsigned v) { return (float)v; }          // exact (I2F: the XU pipe, a quarter of the FP32 rate)
// Off the XU pipe: byte k of w under the exponent of 2^23 IS the float 2^23 + byte (one PRMT); subtracting 2^23 (exact, and
// it packs) gives the byte as a float.
JG_DEV float u8_biased(unsigned w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (unsigned)k)); }
constexpr float kU8Bias = 8388608.0f;
// packed add, rounded toward minus infinity: x + (1.5 * 2^23 + c) has an ulp of 1, so the result is 1.5 * 2^23 + floor(x + c)
// and its low mantissa bits are that integer -- the floor without the XU pipe's F2I
JG_DEV f32x2 f2_add_rd(f32x2 a, f32x2 b) { return __fadd2_rd(a, b); }
JG_DEV float f_add_rd(float a, float b) { return __fadd_rd(a, b); }
JG_DEV unsigned f_bits(float a) { return __float_as_uint(a); }

// ---- integer helpers ---------------------------------------------------------------------
JG_DEV int i_clz(unsigned v) { return __clz((int)v); }
JG_DEV int i_ffs(unsigned v) { return __ffs((int)v); }
JG_DEV int i_popc(unsigned v) { return __popc(v); }
JG_DEV unsigned bswap32(unsigned v) { return __byte_perm(v, 0u, 0x0123u); }
// byte i of the result = byte (sel >> 4i) & 7 of the 8 bytes {a: 0-3, b: 4-7}
JG_DEV unsigned byte_perm(unsigned a, unsigned b, unsigned sel) { return __byte_perm(a, b, sel); }
JG_DEV unsigned funnel_l(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_l(lo, hi, s); }
// the same with the shift count CLAMPED to 32 instead of taken mod 32 (count 32: everything moves by a whole word)
JG_DEV unsigned funnel_lc(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_lc(lo, hi, s); }
JG_DEV unsigned funnel_rc(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_rc(lo, hi, s); }
// low 32 bits of (hi:lo) >> s, s in 0..31
JG_DEV unsigned funnel_r(unsigned lo, unsigned hi, unsigned s) { return __funnelshift_r(lo, hi, s); }
// per-byte compare: 0xff in every byte lane where a == b
JG_DEV unsigned v_cmpeq4(unsigned a, unsigned b) { return __vcmpeq4(a, b); }
JG_DEV unsigned bit_reverse(unsigned v) { return __brev(v); }
JG_DEV unsigned add_min_u32(unsigned a, unsigned b, unsigned c) { return __viaddmin_u32(a, b, c); }   // min(a + b, c), one VIADDMNMX
// per-halfword unsigned minimum (one VIMNMX.U16x2)
JG_DEV unsigned v_minu2(unsigned a, unsigned b) { return __vminu2(a, b); }
// per-halfword compare: 0xffff in every halfword lane where a != b
JG_DEV unsigned v_cmpne2(unsigned a, unsigned b) { return __vcmpne2(a, b); }

// ---- CTA / warp collectives --------------------------------------------------------------
JG_DEV void cta_sync() { __syncthreads(); }
JG_DEV unsigned warp_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
JG_DEV unsigned warp_shfl_u32(unsigned v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
JG_DEV unsigned warp_shfl_up_u32(unsigned v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
JG_DEV void warp_sync() { __syncwarp(); }
// inclusive prefix sum over the warp; the shuffle's own predicate replaces the lane compare (CUB idiom)
JG_DEV unsigned warp_scan_incl_u32(unsigned v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        asm volatile("{ .reg .pred p; .reg .u32 r; shfl.sync.up.b32 r|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, r; }"
                     : "+r"(v) : "r"(d));
    }
    return v;
}
JG_DEV unsigned warp_max_u32(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }   // one REDUX
JG_DEV unsigned warp_sum_u32(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }
JG_DEV float warp_shfl_xor_f32(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
JG_DEV unsigned long long warp_shfl_u64(unsigned long long v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
JG_DEV unsigned long long warp_shfl_xor_u64(unsigned long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ---- memory ------------------------------------------------------------------------------
JG_DEV uint32_t ldg_u32(const void* p) { return __ldg(reinterpret_cast<const unsigned*>(p)); }
JG_DEV uint2 ldg_u64(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
JG_DEV uint4 ldg_u128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
JG_DEV uint32_t ldg_u8(const void* p) { return __ldg(reinterpret_cast<const unsigned char*>(p)); }
JG_DEV int ldg_s16(const void* p) { return (int)__ldg(reinterpret_cast<const short*>(p)); }
JG_DEV void smem_atomic_or(unsigned* p, unsigned v) { atomicOr(p, v); }
JG_DEV unsigned gmem_atomic_add(unsigned* p, unsigned v) { return atomicAdd(p, v); }
JG_DEV void gmem_atomic_or(unsigned* p, unsigned v) { atomicOr(p, v); }
JG_DEV unsigned long long ld_flag64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
JG_DEV void st_flag64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
JG_DEV void st_flag32(unsigned* p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
JG_DEV unsigned ld_flag32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
JG_DEV void backoff() { __nanosleep(64); }

// ---- bulk async copy global -> shared (the TMA engine, UBLKCP in SASS) with mbarrier completion ---------
JG_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared memory by 32-bit address (the entropy coder's inner loop)
JG_DEV unsigned lds_u16(unsigned a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
JG_DEV int lds_s16(unsigned a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
JG_DEV uint2 lds_u64(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
JG_DEV void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
JG_DEV void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
JG_DEV void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
JG_DEV float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
JG_DEV void sts_v2f(unsigned a, f32x2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }
JG_DEV f32x2 lds_v2f(unsigned a) { f32x2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
JG_DEV uint4 lds_v4(unsigned a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// a value the compiler must keep in a register instead of recomputing it where it is used (shared-memory addresses in hot loops)
JG_DEV unsigned pinned(unsigned v) { unsigned r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
JG_DEV void mbar_init(unsigned long long* bar, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy (executed by the initialising threads, before the CTA barrier)
JG_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival + the number of bytes the copies of this phase will deliver
JG_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; dst and src 16-byte aligned.  Completes (complete_tx) on `bar`.
JG_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// generic-proxy stores to shared memory -> later async-proxy (TMA) writes of the same bytes
JG_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 2-D TMA tile load: the box of tensor map `tmap` (a __grid_constant__ kernel parameter) at element column c0, row c1
JG_DEV void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_addr(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_addr(bar)) : "memory");
}
// every lane that will read the data waits for the phase with this parity
JG_DEV void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

}  // namespace jg
#endif
