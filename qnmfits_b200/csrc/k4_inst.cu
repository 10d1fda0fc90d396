// k4_inst.cu — K4 (fit_panel.cuh): blocked Householder QR, trailing update on DMMA.
#include "kernels.h"
#include "fit_panel.cuh"

size_t k4_smem_bytes(int N, int L) { return PanelSmem::bytes(N, L); }
int k4_threads() { return K4_THREADS; }
const void *k4_kernel_ptr() { return (const void *)fit_panel_kernel; }

cudaError_t k4_launch(int grid, size_t smem, cudaStream_t st, const FitParams &p)
{
    fit_panel_kernel<<<grid, K4_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}
