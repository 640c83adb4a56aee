// jpeg_stuff.cu -- the layout-independent second pass (plan + stuff kernels), one TU.
#include "jpeg_stuff.cuh"

namespace jg {

size_t stuff_smem_bytes() { return sizeof(StuffSmem); }

cudaError_t stuff_prepare(int* ctas_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, stuff_kernel, kThreads, sizeof(StuffSmem));
}

cudaError_t stuff_launch(int grid, cudaStream_t stream, const LaunchParams& P)
{
    plan_chunks_kernel<<<1, kThreads, sizeof(ScanSmem), stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    count_ff_kernel<<<grid, kCountThreads, sizeof(CountSmem), stream>>>(P);     // grid-stride: any grid covers every group
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    scan_groups_kernel<<<1, kThreads, sizeof(ScanSmem), stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    stuff_kernel<<<grid, kThreads, sizeof(StuffSmem), stream>>>(P);
    return cudaGetLastError();
}

}  // namespace jg
