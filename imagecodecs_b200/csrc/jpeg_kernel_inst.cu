// jpeg_kernel_inst.cu -- one translation unit per kernel specialisation.
// Compiled five times with -DJG_LAYOUT=<0|1|2> -DJG_NC=<1|3|4> (see imagecodecs_b200/build.py)
// so the specialisations build in parallel.  Always with --fmad=false: the float math of
// jpeg_kernel.cuh additionally uses __fadd_rn/__fmul_rn so that no flag can re-fuse it.
#include "jpeg_kernel.cuh"
#include "jpeg_transform.cuh"
#include "jpeg_launch.h"

#ifndef JG_LAYOUT
#error "JG_LAYOUT / JG_NC must be defined"
#endif

#define JG_CAT2(a, b, c, d) a##b##c##d
#define JG_CAT(a, b, c, d) JG_CAT2(a, b, c, d)
#define JG_FN(prefix) JG_CAT(prefix, JG_LAYOUT, _, JG_NC)

namespace jg {

size_t JG_FN(smem_bytes_)() { return sizeof(Smem<JG_LAYOUT, JG_NC>); }

template <int MODE>
static cudaError_t prepare_mode(int* ctas)
{
    auto kern = encode_tiles_kernel<JG_LAYOUT, JG_NC, MODE>;
    const int smem = (int)sizeof(Smem<JG_LAYOUT, JG_NC>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kThreads, smem);
    if (n < *ctas) *ctas = n;
    return e;
}

cudaError_t JG_FN(prepare_)(int* ctas_per_sm)
{
    int ctas = 1 << 20;
    cudaError_t e = prepare_mode<kModePlain>(&ctas);
    if (e == cudaSuccess) e = prepare_mode<kModeDeep>(&ctas);
    if (e == cudaSuccess) e = prepare_mode<kModeRestart>(&ctas);
    *ctas_per_sm = ctas;
    return e;
}

// mode: kModePlain / kModeDeep (launches with few images) / kModeRestart (jpeg_kernel.cuh)
cudaError_t JG_FN(launch_)(int grid, cudaStream_t stream, const LaunchParams& P, const QuantSet& Q, int mode)
{
    const size_t smem = sizeof(Smem<JG_LAYOUT, JG_NC>);
    if (mode == kModeDeep) encode_tiles_kernel<JG_LAYOUT, JG_NC, kModeDeep><<<grid, kThreads, smem, stream>>>(P, Q);
    else if (mode == kModeRestart) encode_tiles_kernel<JG_LAYOUT, JG_NC, kModeRestart><<<grid, kThreads, smem, stream>>>(P, Q);
    else encode_tiles_kernel<JG_LAYOUT, JG_NC, kModePlain><<<grid, kThreads, smem, stream>>>(P, Q);
    return cudaGetLastError();
}

// ---- pass A of the split pipeline ----
cudaError_t JG_FN(transform_prepare_)()
{
    return cudaFuncSetAttribute(transform_kernel<JG_LAYOUT, JG_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TSmem<JG_LAYOUT>));
}

cudaError_t JG_FN(transform_launch_)(cudaStream_t stream, const TransformParams& P, const QuantSet& Q)
{
    const int grid = (P.n_items + kWarps - 1) / kWarps;
    if (grid > 0) transform_kernel<JG_LAYOUT, JG_NC><<<grid, kThreads, sizeof(TSmem<JG_LAYOUT>), stream>>>(P, Q);
    return cudaGetLastError();
}

}  // namespace jg
