// k1p_config.h — configuration of the K1p instances (fit_pair.cuh): lanes per row slice and rows per
// register block for every column count.  No CUDA dependency: included by kernels.h (the library)
// and by tests/hostsim/hostsim.cpp (the CPU lane emulation runs the same configuration).
#pragma once
#define K1P_MIN_N 9
#define K1P_MAX_N 24
// Lanes per row slice and rows per register block, measured on the 128 x 128 grid
// (profiles/k1p_variants_r02.txt: A = 2 lanes / 4 rows, C = 4 / 8, D = 2 / 6 (4 beyond N = 12),
// E = 4 / 6, F = 2 / 8 up to N = 10 and 2 / 5 beyond, G = 4 / 10): two lanes while their halves of
// the factor leave >= 224 lanes per CTA (N <= 12), four up to 16, eight beyond; taller blocks where the registers
// allow (eight-row blocks at N = 13, 14 were 2 - 5 % faster until the last register went: they spill now).
#ifndef K1P_CS
#define K1P_CS(N) ((N) <= 12 ? 2 : (N) <= 16 ? 4 : 8)
#endif
#ifndef K1P_MB
#define K1P_MB(N) ((N) <= 10 ? 6 : (N) == 11 ? 5 : (N) == 12 ? 4 : (N) <= 16 ? 6 : 8)
#endif
// AUTO prefers K1p from this column count on
#define K1P_AUTO_MIN_N 9
static constexpr int k1p_cs_ct(int N) { return K1P_CS(N); }
static constexpr int k1p_mb_ct(int N) { return K1P_MB(N); }
// entries of the factor per lane (PairLayout<N, CS>::E, checked in k1p_inst.cu)
static constexpr int k1p_entries_ct(int N, int CS)
{
    int e = N;
    for (int s = 1; CS * s < N + 1; ++s) e += N + 1 - CS * s;
    return e;
}
// threads per CTA: the most (multiple of 32, <= 256) whose factors leave room for a staged
// window of ~1000 rows and the frequency tables of fits of eight lanes (sixteen with CS = 8)
static constexpr int k1p_threads_ct(int N)
{
    for (int t = 256; t > 128; t -= 32)
        if (16 * k1p_entries_ct(N, K1P_CS(N)) * t + 48 * N * (t / (K1P_CS(N) == 8 ? 16 : 8)) + 24 * 1024 + 1024 <= 227 * 1024) return t;
    return 128;
}
// translation unit (k1p_inst.cu, -DK1P_PART) holding the instance of N columns
static constexpr int k1p_part_of(int N) { return N <= 12 ? 0 : N <= 16 ? 1 : N - 15; }
#define K1P_PARTS 10
