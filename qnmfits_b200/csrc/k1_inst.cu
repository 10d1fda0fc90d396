// k1_inst.cu — instances of K1 (fit_small.cuh).  Compiled K1_PARTS times with
// -DK1_PART=0..5 (N = 1-6, 7-8, 9-10, 11-12; 4, 5: N = 1-6, 7-8 with the second block size,
// kernels.h) so that the heavy unrolled kernels build in parallel; k1_dispatch.cu routes a
// column count and block size to its part.
#include "kernels.h"
#include "fit_small.cuh"

#ifndef K1_PART
#error "compile with -DK1_PART=0..5"
#endif

#if K1_PART >= 4
#define K1_RANGE (K1_PART - 4)
static constexpr int part_threads(int N) { return k1_alt_threads_ct(N); }
#else
#define K1_RANGE K1_PART
static constexpr int part_threads(int N) { return k1_threads_ct(N); }
#endif

typedef void (*small_kernel_t)(const FitParams);   // kernels take it as __grid_constant__

template <int N>
static small_kernel_t small_kernel_for(bool staged)
{
    return staged ? (small_kernel_t)fit_small_kernel<N, part_threads(N), true>
                  : (small_kernel_t)fit_small_kernel<N, part_threads(N), false>;
}

static small_kernel_t small_kernel(int N, bool staged)
{
    switch (N) {
#if K1_RANGE == 0
    case 1: return small_kernel_for<1>(staged);
    case 2: return small_kernel_for<2>(staged);
    case 3: return small_kernel_for<3>(staged);
    case 4: return small_kernel_for<4>(staged);
    case 5: return small_kernel_for<5>(staged);
    case 6: return small_kernel_for<6>(staged);
#elif K1_RANGE == 1
    case 7: return small_kernel_for<7>(staged);
    case 8: return small_kernel_for<8>(staged);
#elif K1_RANGE == 2
    case 9: return small_kernel_for<9>(staged);
    case 10: return small_kernel_for<10>(staged);
#else
    case 11: return small_kernel_for<11>(staged);
    case 12: return small_kernel_for<12>(staged);
#endif
    }
    return nullptr;
}

template <int N>
static size_t small_smem_bytes_for(int fpc, int stage_rows)
{
    return SmallSmem<N, part_threads(N)>::bytes(fpc, stage_rows);
}

#define K1_CAT2(a, b) a##b
#define K1_CAT(a, b) K1_CAT2(a, b)

size_t K1_CAT(k1_smem_bytes_part, K1_PART)(int N, int fpc, int stage_rows)
{
    switch (N) {
#if K1_RANGE == 0
    case 1: return small_smem_bytes_for<1>(fpc, stage_rows);
    case 2: return small_smem_bytes_for<2>(fpc, stage_rows);
    case 3: return small_smem_bytes_for<3>(fpc, stage_rows);
    case 4: return small_smem_bytes_for<4>(fpc, stage_rows);
    case 5: return small_smem_bytes_for<5>(fpc, stage_rows);
    case 6: return small_smem_bytes_for<6>(fpc, stage_rows);
#elif K1_RANGE == 1
    case 7: return small_smem_bytes_for<7>(fpc, stage_rows);
    case 8: return small_smem_bytes_for<8>(fpc, stage_rows);
#elif K1_RANGE == 2
    case 9: return small_smem_bytes_for<9>(fpc, stage_rows);
    case 10: return small_smem_bytes_for<10>(fpc, stage_rows);
#else
    case 11: return small_smem_bytes_for<11>(fpc, stage_rows);
    case 12: return small_smem_bytes_for<12>(fpc, stage_rows);
#endif
    }
    return 0;
}

const void *K1_CAT(k1_kernel_ptr_part, K1_PART)(int N, bool staged) { return (const void *)small_kernel(N, staged); }

cudaError_t K1_CAT(k1_launch_part, K1_PART)(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st,
                                            const FitParams &p)
{
    small_kernel_t k = small_kernel(N, staged);
    if (!k) return cudaErrorInvalidDeviceFunction;
    k<<<grid, block, smem, st>>>(p);
    return cudaGetLastError();
}
