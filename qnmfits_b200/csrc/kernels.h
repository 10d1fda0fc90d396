// kernels.h — internal C++ interface between the translation units of libqnmfit.so.
//
// The library is built from several .cu files so that the heavy template instances of the
// fit kernels compile in parallel (__graft_entry__.build): the host API (qnmfit_api.cu)
// sees the kernels only through the small launch / attribute functions declared here.
// Nothing in this header crosses the C ABI (include/qnmfit.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "qnmfit_common.cuh"

#ifndef K1_THREADS
#define K1_THREADS 256
#endif

// ---- K1 (fit_small.cuh): instances N = 1..12, with / without the staged window ----------
#define K1_PARTS 6
// threads per CTA: one 256-thread CTA per SM while the per-lane factor (N (N+1)/2 complex +
// N real in shared memory) allows it, fewer lanes for the wider factors
static constexpr int k1_threads_ct(int N) { return N <= 9 ? K1_THREADS : N == 10 ? 192 : 160; }
static inline int k1_threads(int N) { return k1_threads_ct(N); }
// Second block size for N <= 8 (parts 4, 5): seven warps.  A slab whose lanes fill 6.9 warps
// per SM (8192 fits x 4 lanes on 148 SMs: one rank's share of the 256 x 256 grid on 8 GPUs)
// runs as 147 CTAs of 224 threads on 147 SMs instead of 128 CTAs of 256 on 128.  The block size
// never enters a fit's arithmetic (a fit is lanes_per_fit lanes of ONE warp), so the choice is
// made per slab (qnmfit_api.cu, make_plan) and the bits stay those of any other split.
#ifndef K1_ALT_THREADS
#define K1_ALT_THREADS 224
#endif
static constexpr int k1_alt_threads_ct(int N) { return N <= 8 ? K1_ALT_THREADS : 0; }
static inline int k1_alt_threads(int N) { return k1_alt_threads_ct(N); }              // 0: none
static constexpr int k1_part_of(int N) { return N <= 6 ? 0 : N <= 8 ? 1 : N <= 10 ? 2 : 3; }
static constexpr int k1_alt_part_of(int N) { return N <= 6 ? 4 : 5; }
int k1_block_rows(int N);                                   // SmallLayout<N>::MB
// `threads` = k1_threads(N) or k1_alt_threads(N)
size_t k1_smem_bytes(int N, int threads, int fpc, int stage_rows);
const void *k1_kernel_ptr(int N, int threads, bool staged); // for cudaFuncSetAttribute / GetAttributes
cudaError_t k1_launch(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p);

// ---- K1p (fit_pair.cuh): N = 9..24, the columns of a row slice split over CS lanes ------------
#include "k1p_config.h"
static inline int k1p_cs(int N) { return k1p_cs_ct(N); }
static inline int k1p_mb(int N) { return k1p_mb_ct(N); }
static inline int k1p_threads(int N) { return k1p_threads_ct(N); }
size_t k1p_smem_bytes(int N, int fpc, int stage_rows);
const void *k1p_kernel_ptr(int N, bool staged);
cudaError_t k1p_launch(int N, bool staged, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p);

// ---- K2 (fit_general.cuh) -------------------------------------------------------------
#define K2_THREADS 256
size_t k2_smem_bytes(int N, int L, int TR, int TK);
const void *k2_kernel_ptr();
cudaError_t k2_launch(int grid, size_t smem, cudaStream_t st, const FitParams &p, int TR, int TK);

// ---- K3 (fit_struct.cuh): forms (G lanes per column, RPT rows per thread) ---------------
#define K3C_DEFAULT_G 4
#define K3C_DEFAULT_RPT 16
#define K3_FORMS 6
extern const int k3_forms[K3_FORMS][2];
size_t k3_smem_bytes(int N, int L);
const void *k3_kernel_ptr(int G, int RPT);                  // NULL: form not compiled
cudaError_t k3_launch(int G, int RPT, int grid, int block, size_t smem, cudaStream_t st, const FitParams &p);

// ---- K4 (fit_panel.cuh): blocked Householder, trailing update on the FP64 tensor cores ------
size_t k4_smem_bytes(int N, int L);
int k4_threads();
const void *k4_kernel_ptr();
cudaError_t k4_launch(int grid, size_t smem, cudaStream_t st, const FitParams &p);

// ---- multi-GPU epoch barrier (qnmfit_common.cuh) and the FP64 peak micro-benchmarks ------
cudaError_t peer_barrier_launch(cudaStream_t st, const FitParams &p);
cudaError_t fp64_peak_launch(int kind, int grid, int block, double *out, int iters);
cudaError_t fp64_latency_launch(int kind, double *out, int iters);   // kinds 20+n / 30+n: cycles per DMMA / DFMA, n chains, one warp
