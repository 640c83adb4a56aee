// jpeg_entropy.cu -- pass B of the split pipeline (layout independent): see jpeg_entropy.cuh.
#include "jpeg_entropy.cuh"
#include "jpeg_launch.h"

namespace jg {

template <int MODE>
static cudaError_t entropy_prepare_mode(int* ctas)
{
    auto kern = entropy_kernel<MODE>;
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
    cudaError_t e = entropy_prepare_mode<kEntModePlain>(&ctas);
    if (e == cudaSuccess) e = entropy_prepare_mode<kEntModeRestart>(&ctas);
    *ctas_per_sm = ctas;
    return e;
}

cudaError_t entropy_launch(int grid, cudaStream_t stream, const LaunchParams& P, const CoefMap& cmap, bool restart)
{
    const size_t smem = sizeof(EntSmem);
    if (restart) entropy_kernel<kEntModeRestart><<<grid, kEntThreads, smem, stream>>>(P, cmap);
    else entropy_kernel<kEntModePlain><<<grid, kEntThreads, smem, stream>>>(P, cmap);
    return cudaGetLastError();
}

}  // namespace jg
