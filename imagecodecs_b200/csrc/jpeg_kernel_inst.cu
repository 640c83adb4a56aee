// jpeg_kernel_inst.cu -- one translation unit per kernel specialisation.
// Compiled five times with -DJG_LAYOUT=<0|1|2> -DJG_NC=<1|3|4> (see imagecodecs_b200/build.py)
// so the specialisations build in parallel.  Always with --fmad=false: the float math of
// jpeg_kernel.cuh additionally uses __fadd_rn/__fmul_rn so that no flag can re-fuse it.
#include "jpeg_kernel.cuh"
#include "jpeg_launch.h"

#ifndef JG_LAYOUT
#error "JG_LAYOUT / JG_NC must be defined"
#endif

#define JG_CAT2(a, b, c, d) a##b##c##d
#define JG_CAT(a, b, c, d) JG_CAT2(a, b, c, d)
#define JG_FN(prefix) JG_CAT(prefix, JG_LAYOUT, _, JG_NC)

namespace jg {

size_t JG_FN(smem_bytes_)() { return sizeof(Smem<JG_LAYOUT, JG_NC>); }

cudaError_t JG_FN(prepare_)(int* ctas_per_sm)
{
    auto kern = encode_tiles_kernel<JG_LAYOUT, JG_NC, false>;
    auto kern_deep = encode_tiles_kernel<JG_LAYOUT, JG_NC, true>;
    const int smem = (int)sizeof(Smem<JG_LAYOUT, JG_NC>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern_deep, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int a = 0, b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern_deep, kThreads, smem);
    *ctas_per_sm = a < b ? a : b;
    return e;
}

// deep: the two-iteration pipeline, for launches with few images (jpeg_kernel.cuh)
cudaError_t JG_FN(launch_)(int grid, cudaStream_t stream, const LaunchParams& P, const QuantSet& Q, bool deep)
{
    if (deep) encode_tiles_kernel<JG_LAYOUT, JG_NC, true><<<grid, kThreads, sizeof(Smem<JG_LAYOUT, JG_NC>), stream>>>(P, Q);
    else encode_tiles_kernel<JG_LAYOUT, JG_NC, false><<<grid, kThreads, sizeof(Smem<JG_LAYOUT, JG_NC>), stream>>>(P, Q);
    return cudaGetLastError();
}

}  // namespace jg
