// k1_dispatch.cu — routes a column count to the translation unit holding its K1 instance.
#include "kernels.h"
#include "fit_small.cuh"

#define K1_DECL(P)                                                                                   \
    size_t k1_smem_bytes_part##P(int N, int fpc, int stage_rows);                                    \
    const void *k1_kernel_ptr_part##P(int N, bool staged);                                           \
    cudaError_t k1_launch_part##P(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, \
                                  const FitParams &p);
K1_DECL(0) K1_DECL(1) K1_DECL(2) K1_DECL(3)

int k1_block_rows(int N) { return N <= 8 ? 4 : N <= QNMFIT_MB3_MAX_N ? 3 : 2; }   // SmallLayout<N>::MB
static_assert(SmallLayout<8>::MB == 4 && SmallLayout<9>::MB == 3 && SmallLayout<12>::MB == (12 <= QNMFIT_MB3_MAX_N ? 3 : 2),
              "k1_block_rows must mirror SmallLayout<N>::MB");

size_t k1_smem_bytes(int N, int fpc, int stage_rows)
{
    switch (k1_part_of(N)) {
    case 0: return k1_smem_bytes_part0(N, fpc, stage_rows);
    case 1: return k1_smem_bytes_part1(N, fpc, stage_rows);
    case 2: return k1_smem_bytes_part2(N, fpc, stage_rows);
    default: return k1_smem_bytes_part3(N, fpc, stage_rows);
    }
}

const void *k1_kernel_ptr(int N, bool staged)
{
    if (N < 1 || N > QNMFIT_MAX_MODES_SMALL) return nullptr;
    switch (k1_part_of(N)) {
    case 0: return k1_kernel_ptr_part0(N, staged);
    case 1: return k1_kernel_ptr_part1(N, staged);
    case 2: return k1_kernel_ptr_part2(N, staged);
    default: return k1_kernel_ptr_part3(N, staged);
    }
}

cudaError_t k1_launch(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p)
{
    switch (k1_part_of(N)) {
    case 0: return k1_launch_part0(N, staged, grid, block, smem, st, p);
    case 1: return k1_launch_part1(N, staged, grid, block, smem, st, p);
    case 2: return k1_launch_part2(N, staged, grid, block, smem, st, p);
    default: return k1_launch_part3(N, staged, grid, block, smem, st, p);
    }
}
