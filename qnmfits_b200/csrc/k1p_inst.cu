// k1p_inst.cu — instances of K1p (fit_pair.cuh), N = 9 .. 24.  Compiled ten times with
// -DK1P_PART=0..9 (N = 9-12, 13-16, then one column count per part: 17 .. 24) so that the
// unrolled kernels build in parallel (kernels.h: k1p_part_of).
#include "kernels.h"
#include "fit_pair.cuh"

#ifndef K1P_PART
#error "compile with -DK1P_PART=0..9"
#endif

typedef void (*pair_kernel_t)(const FitParams);

template <int N>
static pair_kernel_t pair_kernel_for(bool staged)
{
    return staged ? (pair_kernel_t)fit_pair_kernel<N, k1p_cs_ct(N), k1p_mb_ct(N), k1p_threads_ct(N), true>
                  : (pair_kernel_t)fit_pair_kernel<N, k1p_cs_ct(N), k1p_mb_ct(N), k1p_threads_ct(N), false>;
}

static pair_kernel_t pair_kernel(int N, bool staged)
{
    switch (N) {
#if K1P_PART == 0
    case 9: return pair_kernel_for<9>(staged);
    case 10: return pair_kernel_for<10>(staged);
    case 11: return pair_kernel_for<11>(staged);
    case 12: return pair_kernel_for<12>(staged);
#elif K1P_PART == 1
    case 13: return pair_kernel_for<13>(staged);
    case 14: return pair_kernel_for<14>(staged);
    case 15: return pair_kernel_for<15>(staged);
    case 16: return pair_kernel_for<16>(staged);
#else
    case K1P_PART + 15: return pair_kernel_for<K1P_PART + 15>(staged);
#endif
    }
    return nullptr;
}

template <int N>
static size_t pair_smem_bytes_for(int fpc, int stage_rows)
{
    static_assert(PairLayout<N, k1p_cs_ct(N)>::E == k1p_entries_ct(N, k1p_cs_ct(N)), "kernels.h mirrors PairLayout::E");
    return PairSmem<N, k1p_cs_ct(N), k1p_threads_ct(N)>::bytes(fpc, stage_rows);
}

#define K1P_CAT2(a, b) a##b
#define K1P_CAT(a, b) K1P_CAT2(a, b)

size_t K1P_CAT(k1p_smem_bytes_part, K1P_PART)(int N, int fpc, int stage_rows)
{
    switch (N) {
#if K1P_PART == 0
    case 9: return pair_smem_bytes_for<9>(fpc, stage_rows);
    case 10: return pair_smem_bytes_for<10>(fpc, stage_rows);
    case 11: return pair_smem_bytes_for<11>(fpc, stage_rows);
    case 12: return pair_smem_bytes_for<12>(fpc, stage_rows);
#elif K1P_PART == 1
    case 13: return pair_smem_bytes_for<13>(fpc, stage_rows);
    case 14: return pair_smem_bytes_for<14>(fpc, stage_rows);
    case 15: return pair_smem_bytes_for<15>(fpc, stage_rows);
    case 16: return pair_smem_bytes_for<16>(fpc, stage_rows);
#else
    case K1P_PART + 15: return pair_smem_bytes_for<K1P_PART + 15>(fpc, stage_rows);
#endif
    }
    return (size_t)-1;
}

const void *K1P_CAT(k1p_kernel_ptr_part, K1P_PART)(int N, bool staged) { return (const void *)pair_kernel(N, staged); }

cudaError_t K1P_CAT(k1p_launch_part, K1P_PART)(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st,
                                               const FitParams &p)
{
    pair_kernel_t k = pair_kernel(N, staged);
    if (!k) return cudaErrorInvalidDeviceFunction;
    k<<<grid, block, smem, st>>>(p);
    return cudaGetLastError();
}
