// One instantiation of the general Tucker kernel per translation unit: VT_TUCKER_NM = 16, 32 or 64 (grids of up to that
// many nodes per axis).  build.py compiles this file three times; in one unit the three take eight minutes of cicc.
#include "tucker_internal.h"

#include <algorithm>
#include <cstring>

#ifndef VT_TUCKER_NM
#define VT_TUCKER_NM 64   // build.py passes 16, 32 and 64 in turn; a bare `nvcc -c` of this file gives the largest instance
#endif

namespace vt {

#include "tucker_kernel.inl"

#if VT_TUCKER_NM == 16
void launch_k_tucker_16(int grid, size_t smem, cudaStream_t stream, const TuckerParams& P)
{
    k_tucker<kThreadsSmall, 16><<<grid, kThreadsSmall, smem, stream>>>(P);
}
#elif VT_TUCKER_NM == 32
void launch_k_tucker_32(int grid, size_t smem, cudaStream_t stream, const TuckerParams& P)
{
    if (smem > 32 * 1024)
        VT_CUDA(cudaFuncSetAttribute(k_tucker<kThreadsSmall, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tucker<kThreadsSmall, 32><<<grid, kThreadsSmall, smem, stream>>>(P);
}
#else
void launch_k_tucker_64(int grid, size_t smem, cudaStream_t stream, const TuckerParams& P)
{
    if (smem > 32 * 1024)
        VT_CUDA(cudaFuncSetAttribute(k_tucker<kThreads, kMaxN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((3 * (size_t)(kMaxN | 1) * kMaxN + 2 * (size_t)tile_cap(kMaxN)) * sizeof(double))));
    k_tucker<kThreads, kMaxN><<<grid, kThreads, smem, stream>>>(P);
}
#endif

}  // namespace vt
