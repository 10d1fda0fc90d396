// k3_inst.cu — K3 (fit_struct.cuh): structured two-phase QR, forms (G, RPT).
#include "kernels.h"
#include "fit_struct.cuh"

typedef void (*struct_kernel_t)(const FitParams);

const int k3_forms[K3_FORMS][2] = {{1, 32}, {2, 16}, {2, 32}, {4, 8}, {4, 16}, {8, 8}};

static struct_kernel_t struct3_kernel(int G, int RPT)
{
    if (G == 1 && RPT == 32) return fit_struct3_kernel<1, 32>;
    if (G == 2 && RPT == 16) return fit_struct3_kernel<2, 16>;
    if (G == 2 && RPT == 32) return fit_struct3_kernel<2, 32>;
    if (G == 4 && RPT == 8) return fit_struct3_kernel<4, 8>;
    if (G == 4 && RPT == 16) return fit_struct3_kernel<4, 16>;
    if (G == 8 && RPT == 8) return fit_struct3_kernel<8, 8>;
    return nullptr;
}

size_t k3_smem_bytes(int N, int L) { return Struct3Smem::bytes(N, L); }
const void *k3_kernel_ptr(int G, int RPT) { return (const void *)struct3_kernel(G, RPT); }

cudaError_t k3_launch(int G, int RPT, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p)
{
    struct_kernel_t k = struct3_kernel(G, RPT);
    if (!k) return cudaErrorInvalidDeviceFunction;
    k<<<grid, block, smem, st>>>(p);
    return cudaGetLastError();
}
