// k1_dispatch.cu — routes a column count to the translation unit holding its K1 instance.
#include "kernels.h"
#include "fit_small.cuh"

#define K1_DECL(P)                                                                                   \
    size_t k1_smem_bytes_part##P(int N, int fpc, int stage_rows);                                    \
    const void *k1_kernel_ptr_part##P(int N, bool staged);                                           \
    cudaError_t k1_launch_part##P(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, \
                                  const FitParams &p);
K1_DECL(0) K1_DECL(1) K1_DECL(2) K1_DECL(3) K1_DECL(4) K1_DECL(5)

int k1_block_rows(int N) { return N <= 8 ? 4 : N <= QNMFIT_MB3_MAX_N ? 3 : 2; }   // SmallLayout<N>::MB
static_assert(SmallLayout<8>::MB == 4 && SmallLayout<9>::MB == 3 && SmallLayout<12>::MB == (12 <= QNMFIT_MB3_MAX_N ? 3 : 2),
              "k1_block_rows must mirror SmallLayout<N>::MB");

// part holding the (N, threads) instance; -1: none
static int part_for(int N, int threads)
{
    if (N < 1 || N > QNMFIT_MAX_MODES_SMALL) return -1;
    if (threads == k1_threads(N)) return k1_part_of(N);
    if (threads == k1_alt_threads(N)) return k1_alt_part_of(N);
    return -1;
}

size_t k1_smem_bytes(int N, int threads, int fpc, int stage_rows)
{
    switch (part_for(N, threads)) {
    case 0: return k1_smem_bytes_part0(N, fpc, stage_rows);
    case 1: return k1_smem_bytes_part1(N, fpc, stage_rows);
    case 2: return k1_smem_bytes_part2(N, fpc, stage_rows);
    case 3: return k1_smem_bytes_part3(N, fpc, stage_rows);
    case 4: return k1_smem_bytes_part4(N, fpc, stage_rows);
    case 5: return k1_smem_bytes_part5(N, fpc, stage_rows);
    }
    return (size_t)-1;
}

const void *k1_kernel_ptr(int N, int threads, bool staged)
{
    switch (part_for(N, threads)) {
    case 0: return k1_kernel_ptr_part0(N, staged);
    case 1: return k1_kernel_ptr_part1(N, staged);
    case 2: return k1_kernel_ptr_part2(N, staged);
    case 3: return k1_kernel_ptr_part3(N, staged);
    case 4: return k1_kernel_ptr_part4(N, staged);
    case 5: return k1_kernel_ptr_part5(N, staged);
    }
    return nullptr;
}

cudaError_t k1_launch(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p)
{
    switch (part_for(N, block)) {
    case 0: return k1_launch_part0(N, staged, grid, block, smem, st, p);
    case 1: return k1_launch_part1(N, staged, grid, block, smem, st, p);
    case 2: return k1_launch_part2(N, staged, grid, block, smem, st, p);
    case 3: return k1_launch_part3(N, staged, grid, block, smem, st, p);
    case 4: return k1_launch_part4(N, staged, grid, block, smem, st, p);
    case 5: return k1_launch_part5(N, staged, grid, block, smem, st, p);
    }
    return cudaErrorInvalidDeviceFunction;
}

// ---- K1p (k1p_inst.cu, K1P_PARTS parts) -------------------------------------------------------
#define K1P_DECL(P)                                                                                  \
    size_t k1p_smem_bytes_part##P(int N, int fpc, int stage_rows);                                   \
    const void *k1p_kernel_ptr_part##P(int N, bool staged);                                          \
    cudaError_t k1p_launch_part##P(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, \
                                   const FitParams &p);
K1P_DECL(0) K1P_DECL(1) K1P_DECL(2) K1P_DECL(3) K1P_DECL(4) K1P_DECL(5) K1P_DECL(6) K1P_DECL(7) K1P_DECL(8) K1P_DECL(9)
static_assert(K1P_PARTS == 10 && k1p_part_of(K1P_MAX_N) == 9, "one K1P_DECL / table entry per part");

typedef size_t (*k1p_smem_fn)(int, int, int);
typedef const void *(*k1p_ptr_fn)(int, bool);
typedef cudaError_t (*k1p_launch_fn)(int, bool, int, int, size_t, cudaStream_t, const FitParams &);
#define K1P_TABLE(fn) { fn##_part0, fn##_part1, fn##_part2, fn##_part3, fn##_part4, fn##_part5, fn##_part6, fn##_part7, fn##_part8, fn##_part9 }
static const k1p_smem_fn k1p_smem_table[K1P_PARTS] = K1P_TABLE(k1p_smem_bytes);
static const k1p_ptr_fn k1p_ptr_table[K1P_PARTS] = K1P_TABLE(k1p_kernel_ptr);
static const k1p_launch_fn k1p_launch_table[K1P_PARTS] = K1P_TABLE(k1p_launch);

size_t k1p_smem_bytes(int N, int fpc, int stage_rows)
{
    if (N < K1P_MIN_N || N > K1P_MAX_N) return (size_t)-1;
    return k1p_smem_table[k1p_part_of(N)](N, fpc, stage_rows);
}

const void *k1p_kernel_ptr(int N, bool staged)
{
    if (N < K1P_MIN_N || N > K1P_MAX_N) return nullptr;
    return k1p_ptr_table[k1p_part_of(N)](N, staged);
}

cudaError_t k1p_launch(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p)
{
    if (N < K1P_MIN_N || N > K1P_MAX_N) return cudaErrorInvalidDeviceFunction;
    return k1p_launch_table[k1p_part_of(N)](N, staged, grid, block, smem, st, p);
}
