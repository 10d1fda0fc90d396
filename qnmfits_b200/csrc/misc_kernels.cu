// misc_kernels.cu — the one-warp epoch barrier of the multi-GPU exchange and the FP64 peak
// micro-benchmarks (the roofline denominator; MEASURED_PEAKS.json has no FP64 entry).
#define QNMFIT_DEFINE_PEER_BARRIER
#include "kernels.h"

cudaError_t peer_barrier_launch(cudaStream_t st, const FitParams &p)
{
    peer_barrier_kernel<<<1, 32, 0, st>>>(p);
    return cudaGetLastError();
}

// Eight independent dependent-FMA chains per thread / four MMA accumulators per warp;
// results are stored so nothing is optimised away.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double b, double c)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// DFMA whose three source operands are all distinct registers and never repeat in the
// same operand slot of consecutive instructions (no operand-reuse-cache hits): the rate
// the register file can feed, which is what bounds register-blocked FP64 code like the
// Householder updates of K1/K2.
__global__ void __launch_bounds__(256) dfma_3op_kernel(double *out, int iters, double b, double c)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    double b0 = b, b1 = b + 1e-9, b2 = b + 2e-9, b3 = b + 3e-9, b4 = b + 4e-9, b5 = b + 5e-9, b6 = b + 6e-9,
           b7 = b + 7e-9;
    double c0 = c, c1 = c * 2, c2 = c * 3, c3 = c * 4, c4 = c * 5, c5 = c * 6, c6 = c * 7, c7 = c * 8;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a0) : "d"(b0), "d"(c0));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a1) : "d"(b1), "d"(c1));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a2) : "d"(b2), "d"(c2));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a3) : "d"(b3), "d"(c3));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a4) : "d"(b4), "d"(c4));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a5) : "d"(b5), "d"(c5));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a6) : "d"(b6), "d"(c6));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a7) : "d"(b7), "d"(c7));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Same, but the multiplier is shared by consecutive instructions (one operand served by
// the reuse cache, two register reads per DFMA) — the pattern of a rank-1 update.
__global__ void __launch_bounds__(256) dfma_2op_kernel(double *out, int iters, double b, double c)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    double c0 = c, c1 = c * 2, c2 = c * 3, c3 = c * 4, c4 = c * 5, c5 = c * 6, c6 = c * 7, c7 = c * 8;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a0) : "d"(b), "d"(c0));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a1) : "d"(b), "d"(c1));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a2) : "d"(b), "d"(c2));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a3) : "d"(b), "d"(c3));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a4) : "d"(b), "d"(c4));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a5) : "d"(b), "d"(c5));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a6) : "d"(b), "d"(c6));
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a7) : "d"(b), "d"(c7));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double b, double c)
{
    double a = threadIdx.x * 1e-3 + 0.5, bb = b;
    double c00 = c, c01 = c, c10 = c + 1, c11 = c + 1, c20 = c + 2, c21 = c + 2, c30 = c + 3, c31 = c + 3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c00), "+d"(c01) : "d"(a), "d"(bb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c10), "+d"(c11) : "d"(a), "d"(bb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c20), "+d"(c21) : "d"(a), "d"(bb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c30), "+d"(c31) : "d"(a), "d"(bb));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((c00 + c01) + (c10 + c11)) + ((c20 + c21) + (c30 + c31));
}

// kind 4: DFMA and DMMA interleaved in the same warp (are the vector and the tensor FP64
// datapaths one unit or two?); counts both kinds of flops.
__global__ void __launch_bounds__(256) dfma_dmma_mix_kernel(double *out, int iters, double b, double c)
{
    double a = threadIdx.x * 1e-3 + 0.5, bb = b;
    double c00 = c, c01 = c, c10 = c + 1, c11 = c + 1;
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c00), "+d"(c01) : "d"(a), "d"(bb));
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c10), "+d"(c11) : "d"(a), "d"(bb));
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] =
        ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) + ((c00 + c01) + (c10 + c11));
}

// kinds 20 + n (n = 1, 2, 4, 8): latency probe — ONE warp issuing DMMAs on n independent accumulators;
// out[0] = cycles per DMMA.  kinds 30 + n: the same for DFMA.
template <int CHAINS, bool MMA>
__global__ void __launch_bounds__(32) fp64_latency_kernel(double *out, int iters, double b, double c)
{
    double a = threadIdx.x * 1e-3 + 0.5, bb = b;
    double acc[CHAINS][2];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { acc[k][0] = c + k; acc[k][1] = c - k; }
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) {
                if (MMA)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[k][0]), "+d"(acc[k][1]) : "d"(a), "d"(bb));
                else
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k][0]) : "d"(a), "d"(bb));
            }
        }
    }
    const long long t1 = clock64();
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) sum += acc[k][0] + acc[k][1];
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / ((double)iters * 8.0 * CHAINS); out[1] = sum; }
}

// kind 40: dependent chain  x <- rcp(rsqrt(x) + c)  (the reflector scalars of the fit kernels);
// kind 41: dependent chain  x <- x + shfl_xor(x, 1)  (one level of a warp reduction);
// kind 42: dependent chain through shared memory  x <- lds(sts(x))  with __syncwarp.
__global__ void __launch_bounds__(32) misc_latency_kernel(double *out, int iters, int kind, double c)
{
    __shared__ double buf[32];
    double x = 1.0 + threadIdx.x * 1e-3;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (kind == 40) x = qf_rcp(qf_rsqrt(x) + c);
            else if (kind == 41) x = x + __shfl_xor_sync(0xffffffffu, x, 1);
            else { buf[threadIdx.x] = x; __syncwarp(); x = buf[threadIdx.x ^ 1] * c; __syncwarp(); }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / ((double)iters * 8.0); out[1] = x; }
}

cudaError_t fp64_latency_launch(int kind, double *out, int iters)
{
    if (kind >= 40) {
        misc_latency_kernel<<<1, 32>>>(out, iters, kind, 1.0 + 1e-9);
        return cudaGetLastError();
    }
    const bool mma = kind < 30;
    const int n = mma ? kind - 20 : kind - 30;
#define QF_LAT(N) if (n == N) { if (mma) fp64_latency_kernel<N, true><<<1, 32>>>(out, iters, 0.999999, 1e-9); \
                                else fp64_latency_kernel<N, false><<<1, 32>>>(out, iters, 0.999999, 1e-9); }
    QF_LAT(1) QF_LAT(2) QF_LAT(4) QF_LAT(8)
#undef QF_LAT
    return cudaGetLastError();
}

cudaError_t fp64_peak_launch(int kind, int grid, int block, double *out, int iters)
{
    if (kind == 0) dfma_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    else if (kind == 2) dfma_3op_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    else if (kind == 3) dfma_2op_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    else if (kind == 4) dfma_dmma_mix_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    else dmma_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    return cudaGetLastError();
}
