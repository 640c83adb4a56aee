// jpeg_entropy.cu -- pass B of the split pipeline (layout independent): see jpeg_entropy.cuh.
#include "jpeg_entropy.cuh"
#include "jpeg_launch.h"

namespace jg {

template <int MODE, bool DEFER>
static cudaError_t entropy_prepare_mode(int* ctas)
{
    auto kern = entropy_kernel<MODE, DEFER>;
    const int smem = (int)sizeof(EntSmem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kEntThreads, smem);
    if (n < *ctas) *ctas = n;
    return e;
}

cudaError_t entropy_prepare(int* ctas_per_sm)
{
    int ctas = 1 << 20;
    cudaError_t e = entropy_prepare_mode<kEntModePlain, false>(&ctas);
    if (e == cudaSuccess) e = entropy_prepare_mode<kEntModePlain, true>(&ctas);
    if (e == cudaSuccess) e = entropy_prepare_mode<kEntModeRestart, false>(&ctas);
    if (e == cudaSuccess) e = entropy_prepare_mode<kEntModeRestart, true>(&ctas);
    *ctas_per_sm = ctas;
    return e;
}

cudaError_t entropy_launch(int grid, cudaStream_t stream, const LaunchParams& P, const CoefMap& cmap, int mode)
{
    const size_t smem = sizeof(EntSmem);
    // mode: bit 1 = restart intervals, bit 0 = deferred write-out (launches with few images)
    if ((mode & 3) == 3) entropy_kernel<kEntModeRestart, true><<<grid, kEntThreads, smem, stream>>>(P, cmap);
    else if (mode & 2) entropy_kernel<kEntModeRestart, false><<<grid, kEntThreads, smem, stream>>>(P, cmap);
    else if (mode & 1) entropy_kernel<kEntModePlain, true><<<grid, kEntThreads, smem, stream>>>(P, cmap);
    else entropy_kernel<kEntModePlain, false><<<grid, kEntThreads, smem, stream>>>(P, cmap);
    return cudaGetLastError();
}

}  // namespace jg
