// Developer micro-benchmarks (not product): how the FP64 pipe of a B200 SM sub-partition
// behaves when DFMA streams are mixed with other instructions / limited ILP / few warps.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_variants/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define FMA(a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a) : "d"(b), "d"(c))
#define IMAD(x, y) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(x) : "r"(y))
#define LDS(v, addr) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr))

// CHAINS independent accumulators; per unrolled group of CHAINS DFMAs insert NI integer ops and NL LDS.
template <int CHAINS, int NI, int NL>
__global__ void __launch_bounds__(256) mix_kernel(double *out, int iters, double b, double c, int y)
{
    __shared__ double sh[256];
    sh[threadIdx.x] = c;
    __syncthreads();
    unsigned saddr = (unsigned)__cvta_generic_to_shared(&sh[threadIdx.x]);
    double a[CHAINS];
    double cc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x * 1e-3 + i; cc[i] = c * (i + 1); }
    int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double l0 = 0, l1 = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                FMA(a[i], b, cc[i]);
                if (NI > 0 && (i % (CHAINS / (NI > CHAINS ? CHAINS : NI))) == 0) {
                    if ((i & 3) == 0) IMAD(x0, y); else if ((i & 3) == 1) IMAD(x1, y);
                    else if ((i & 3) == 2) IMAD(x2, y); else IMAD(x3, y);
                }
                if (NL > 0 && (i % (CHAINS / NL)) == 0) { if (i & 1) LDS(l0, saddr); else LDS(l1, saddr); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(x0 + x1 + x2 + x3) + l0 + l1;
}

template <int CHAINS, int NI, int NL>
static void run(const char *name, int threads, int ctas_per_sm, int sms, double *out, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        mix_kernel<CHAINS, NI, NL><<<sms * ctas_per_sm, threads>>>(out, iters, 0.999999, 1e-9, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 1 && ms < best) best = ms;
    }
    double fmas = (double)iters * 8 * CHAINS * threads * ctas_per_sm * sms;
    double tf = 2 * fmas / (best * 1e-3) * 1e-12;
    printf("%-28s threads=%3d ctas/sm=%d chains=%2d int/grp=%d lds/grp=%d : %7.2f TFLOP/s (%5.1f%% of 37.2)\n", name, threads,
           ctas_per_sm, CHAINS, NI, NL, tf, 100 * tf / 37.2);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int it = 4000;
    run<8, 0, 0>("pure 8 chains", 256, 1, sms, out, it);
    run<8, 0, 0>("pure 8 chains", 128, 1, sms, out, it);
    run<4, 0, 0>("pure 4 chains", 128, 1, sms, out, it);
    run<4, 0, 0>("pure 4 chains", 256, 1, sms, out, it);
    run<2, 0, 0>("pure 2 chains", 128, 1, sms, out, it);
    run<2, 0, 0>("pure 2 chains", 256, 1, sms, out, it);
    run<2, 0, 0>("pure 2 chains", 512, 1, sms, out, it);
    run<3, 0, 0>("pure 3 chains", 128, 1, sms, out, it);
    run<3, 0, 0>("pure 3 chains", 256, 1, sms, out, it);
    run<1, 0, 0>("pure 1 chain", 128, 1, sms, out, it);
    run<1, 0, 0>("pure 1 chain", 256, 1, sms, out, it);
    run<1, 0, 0>("pure 1 chain", 512, 1, sms, out, it);
    run<8, 2, 0>("8 chains + 2 imad (20%)", 256, 1, sms, out, it);
    run<8, 2, 0>("8 chains + 2 imad (20%)", 128, 1, sms, out, it);
    run<8, 4, 0>("8 chains + 4 imad (33%)", 256, 1, sms, out, it);
    run<8, 8, 0>("8 chains + 8 imad (50%)", 256, 1, sms, out, it);
    run<8, 8, 0>("8 chains + 8 imad (50%)", 128, 1, sms, out, it);
    run<8, 0, 2>("8 chains + 2 lds", 256, 1, sms, out, it);
    run<8, 2, 2>("8 chains + 2 imad + 2 lds", 256, 1, sms, out, it);
    run<4, 1, 0>("4 chains + 1 imad", 256, 1, sms, out, it);
    run<4, 2, 0>("4 chains + 2 imad", 256, 1, sms, out, it);
    run<4, 4, 0>("4 chains + 4 imad", 256, 1, sms, out, it);
    run<4, 4, 0>("4 chains + 4 imad", 512, 1, sms, out, it);
    cudaFree(out);
    return 0;
}
