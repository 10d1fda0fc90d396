// Developer micro-benchmark 2 (not product): which instruction classes can issue in the
// shadow of a 2-cycle DFMA on a B200 SM sub-partition.  8 independent DFMA chains per
// thread, K "other" instructions per 8 DFMAs.
#include <cuda_runtime.h>
#include <stdio.h>

#define FMA(a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a) : "d"(b), "d"(c))

enum { O_NONE, O_IMAD, O_LOP, O_IADD, O_SHFL, O_SEL, O_LDS, O_MUFU, O_DSETP, O_FFMA, O_PRMT, O_STS };

template <int KIND>
__device__ __forceinline__ void other(int &x, int y, float &f, unsigned saddr, double &d, int i)
{
    if (KIND == O_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));
    if (KIND == O_LOP) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y));
    if (KIND == O_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (KIND == O_SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x));
    if (KIND == O_SEL) asm volatile("{.reg .pred p; setp.gt.s32 p, %1, 0; selp.b32 %0, %0, %1, p;}" : "+r"(x) : "r"(y));
    if (KIND == O_LDS) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(saddr));
    if (KIND == O_STS) asm volatile("st.shared.b32 [%1], %0;" :: "r"(x), "r"(saddr));
    if (KIND == O_MUFU) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d));
    if (KIND == O_FFMA) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f));
    if (KIND == O_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(x) : "r"(y));
}

template <int KIND, int K>
__global__ void __launch_bounds__(256) mix_kernel(double *out, int iters, double b, double c, int y)
{
    __shared__ int sh[256];
    sh[threadIdx.x] = y;
    __syncthreads();
    unsigned saddr = (unsigned)__cvta_generic_to_shared(&sh[threadIdx.x]);
    double a[8], cc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; cc[i] = c * (i + 1); }
    int x[4] = {(int)threadIdx.x, 1, 2, 3};
    float f[4] = {0.5f, 0.25f, 0.125f, 0.3f};
    double d[4] = {1.5, 2.5, 3.5, 4.5};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                FMA(a[i], b, cc[i]);
                if (K > 0 && (i % (8 / K)) == 0) other<KIND>(x[i & 3], y, f[i & 3], saddr, d[i & 3], i);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(x[0] + x[1] + x[2] + x[3]) + f[0] + f[1] + f[2] + f[3]
                                                 + d[0] + d[1] + d[2] + d[3];
}

template <int KIND, int K>
static void run(const char *name, int threads, int sms, double *out, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        mix_kernel<KIND, K><<<sms, threads>>>(out, iters, 0.999999, 1e-9, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 1 && ms < best) best = ms;
    }
    double fmas = (double)iters * 64 * threads * sms;
    double tf = 2 * fmas / (best * 1e-3) * 1e-12;
    double cyc = best * 1e-3 * 1.965e9 / iters / 8 / (threads / 128);   // SMSP cycles per group of 8 DFMA (+K other) per warp
    printf("%-8s K=%d threads=%3d : %6.2f TFLOP/s (%5.1f%%)  cycles per 8 DFMA + %d other = %.2f\n", name, K, threads, tf,
           100 * tf / 37.2, K, cyc);
}

#define ALL(KIND, name) \
    run<KIND, 2>(name, 256, sms, out, it); run<KIND, 4>(name, 256, sms, out, it); run<KIND, 8>(name, 256, sms, out, it); \
    run<KIND, 2>(name, 128, sms, out, it);

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    const int it = 3000;
    run<O_NONE, 0>("none", 256, sms, out, it);
    run<O_NONE, 0>("none", 128, sms, out, it);
    ALL(O_IMAD, "imad") ALL(O_LOP, "lop3") ALL(O_IADD, "iadd") ALL(O_SHFL, "shfl") ALL(O_SEL, "sel") ALL(O_LDS, "lds")
    ALL(O_STS, "sts") ALL(O_MUFU, "mufu64") ALL(O_FFMA, "ffma") ALL(O_PRMT, "prmt")
    cudaFree(out);
    return 0;
}
