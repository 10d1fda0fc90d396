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

#ifdef K4_TRACE
// developer build only: copy the clock stamps of CTA 0 (fit_panel.cuh, K4_STAMP) to the host
extern "C" int qnmfit_debug_trace(long long *out, int max_events)
{
    int n = 0;
    cudaMemcpyFromSymbol(&n, k4_trace_n, sizeof(int));
    if (n > 2048) n = 2048;
    if (n > max_events) n = max_events;
    cudaMemcpyFromSymbol(out, k4_trace_buf, sizeof(long long) * 4 * n);
    const int zero = 0;
    cudaMemcpyToSymbol(k4_trace_n, &zero, sizeof(int));
    return n;
}
#endif
