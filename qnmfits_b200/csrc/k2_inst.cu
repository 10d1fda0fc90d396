// k2_inst.cu — K2 (fit_general.cuh): the streamed dense Householder fallback.
#include "kernels.h"
#include "fit_general.cuh"

size_t k2_smem_bytes(int N, int L, int TR, int TK) { return GeneralSmem::bytes(N, L, TR, TK); }
const void *k2_kernel_ptr() { return (const void *)fit_general_kernel; }

cudaError_t k2_launch(int grid, size_t smem, cudaStream_t st, const FitParams &p, int TR, int TK)
{
    fit_general_kernel<<<grid, K2_THREADS, smem, st>>>(p, TR, TK);
    return cudaGetLastError();
}
