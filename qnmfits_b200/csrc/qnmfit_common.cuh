// qnmfit_common.cuh — arithmetic shared by the fit kernels.
//
// Everything here compiles both as sm_100a device code (nvcc) and, with
// -DQNMFIT_HOSTSIM, as plain host C++ (g++, -ffp-contract=off) for the lane-by-lane
// emulation used by the CPU tests (tests/hostsim).  The host build is a test harness
// only; the product library is always the CUDA build.
#pragma once

#include <math.h>
#include <stdint.h>
#include "qnmfit.h"

#ifdef QNMFIT_HOSTSIM
#include "hostsim_warp.h"    // tests/hostsim: warp collectives and __syncthreads, one fiber per thread
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define QF_DYN_SMEM(name) unsigned char *name = hswarp::dyn_smem()
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#define QF_HD static inline
#define QF_BOTH static inline
#define QF_MEM inline
#define QF_MEMBOTH inline
#define QF_DEV static inline
static inline double qf_mul_rn(double a, double b) { return a * b; }   // no contraction: -ffp-contract=off
static inline double qf_add_rn(double a, double b) { return a + b; }
static inline double qf_sub_rn(double a, double b) { return a - b; }
static inline double qf_rsqrt(double x) { return 1.0 / sqrt(x); }
static inline double qf_rcp(double x) { return 1.0 / x; }
static inline void qf_sincos(double y, double *s, double *c) { *s = sin(y); *c = cos(y); }
#else
#include <cuda_runtime.h>
#define QF_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define QF_HD __device__ __forceinline__
#define QF_BOTH __host__ __device__ __forceinline__
#define QF_MEM __device__ __forceinline__
#define QF_MEMBOTH __host__ __device__ __forceinline__
#define QF_DEV __device__ __forceinline__
// Un-fused, round-to-nearest products / sums: numpy forms frequencies and exponent
// arguments with separately rounded operations; FMA contraction would change bits.
QF_DEV double qf_mul_rn(double a, double b) { return __dmul_rn(a, b); }
QF_DEV double qf_add_rn(double a, double b) { return __dadd_rn(a, b); }
QF_DEV double qf_sub_rn(double a, double b) { return __dsub_rn(a, b); }

// 1/sqrt(x), x > 0 and normal: MUFU.RSQ64H seed (PTX: max rel. err 2^-22) + one
// cubically convergent step  y <- y (1 + e/2 + 3 e^2/8),  e = 1 - x y^2.  The error
// after the step is ~ (5/16) e^3 < 2^-65, i.e. the result is correct to ~1 ulp of
// rounding; the reflector only needs consistency to O(eps).  Five dependent FP64 ops.
QF_DEV double qf_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double xy = x * y;
    const double e = fma(-xy, y, 1.0);
    double p = fma(0.375, e, 0.5);
    p = p * e;
    return fma(y, p, y);
}

// 1/x, x normal: MUFU.RCP64H seed (2^-23) + two Newton steps (2^-46, 2^-92).
QF_DEV double qf_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
QF_DEV void qf_sincos(double y, double *s, double *c) { sincos(y, s, c); }
#endif

// ---- complex helpers (double2 = re, im) ------------------------------------

QF_HD double2 c_mul(double2 a, double2 b)
{
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

// exp(-i w tau) with numpy's argument rounding (reference qnmfits/qnmfits.py:281):
// numpy evaluates (-1j*w) exactly as (Im w, -Re w), multiplies each component by the
// real tau with one rounding, then exp(x) * (cos y, sin y).
QF_HD double2 design_entry(double2 w, double tau)
{
    double x = qf_mul_rn(w.y, tau);
    double y = qf_mul_rn(-w.x, tau);
    double s, c;
    qf_sincos(y, &s, &c);
    double e = exp(x);
    return make_double2(e * c, e * s);
}

// Frequency of mode j for one fit from the factored table, with the reference's
// rounding order: each constituent times 1/Mf (numpy divides complex by real as a
// multiplication with the correctly rounded reciprocal; reference qnm.py:235), the
// Python `sum` over constituents left to right (qnm.py:272-280), then the real
// factor (1 + delta) (qnmfits.py:274).
QF_HD double2 form_omega(const double *omega_tilde_row, const int32_t *mode_ptr, int j,
                         double inv_mf, double delta_factor)
{
    int p0 = mode_ptr[j], p1 = mode_ptr[j + 1];
    double re = 0.0, im = 0.0;
    for (int p = p0; p < p1; ++p) {
        double a = qf_mul_rn(omega_tilde_row[2 * p], inv_mf);
        double b = qf_mul_rn(omega_tilde_row[2 * p + 1], inv_mf);
        re = qf_add_rn(re, a);
        im = qf_add_rn(im, b);
    }
    return make_double2(qf_mul_rn(delta_factor, re), qf_mul_rn(delta_factor, im));
}

// Parameters shared by both kernels: a by-value copy of the fields of qnmfit_batch
// the device needs (see include/qnmfit.h for their meaning).
struct FitParams {
    int n_fits, n_modes, n_series, n_times;
    long long series_stride;
    long long first_fit;
    const double *times;
    const double2 *data;
    const int *row_begin;
    const int *row_end;
    const double *t0;
    int row_begin_all, row_end_all;
    double t0_all;
    const double2 *omega;
    const double *omega_tilde;
    const int *mode_ptr;
    const double *inv_Mf;
    const double *delta_factor;
    const int *chi_index;
    const int *mf_index;
    int n_chi, n_mf, n_constituents;
    const double2 *coef;
    const int *coef_index;
    const int *series_index;   // K1: per-fit data series (row of data[][]) or NULL
    const double2 *omega_rows; // [N][n_times] per-row frequencies (dynamic fits) or NULL
    const double2 *coef_rows;  // [L][N][n_times] per-row mixing coefficients (K2) or NULL
    int n_coef;
    int anchor_rows;
    double dt_nominal;
    double2 *C;
    double *mismatch;
    double *residual;
    double2 *R;
    int *status;
    double2 *model;
    long long model_stride;
    double *flagged_count;
    int *flag_list;            // [2 * flag_capacity]: (fit, status) of the flagged fits, or NULL
    int flag_capacity;
    const int *fit_index;      // [n_fits] or NULL: launch fit b reads the inputs of sweep fit fit_index[b]
    int omega_shared;
    // launch-time choices
    int lanes_per_fit;   // K1: power of two, 1..32
    int eval_only;       // 1: skip the solve, read C
    int fast_mismatch;   // K1: mismatch from QR by-products (uniform grid, no model output)
    int stage_begin;     // K1 staged variant: first staged row
    int stage_rows;      //                    number of staged rows
    // multi-GPU exchange fused into the kernels (qnmfit_fit_batch_peers); n_peers <= 1: off
    int n_peers, peer_rank;
    unsigned long long peer_epoch;
    long long peer_timeout_ns;
    double *peer_mismatch[QNMFIT_MAX_PEERS];
    double *peer_flagged[QNMFIT_MAX_PEERS];
    unsigned long long *peer_flags[QNMFIT_MAX_PEERS];
};

// Index under which launch fit `fit` finds its per-fit INPUTS (windows, t0, frequencies,
// table indices, data series): itself, or — in a launch over a subset of a sweep's fits (the
// repair of flagged fits) — the sweep's index of that fit.  Outputs are always stored at `fit`.
QF_HD int input_fit(const FitParams &p, int fit)
{
    return p.fit_index ? p.fit_index[fit] : fit;
}

QF_HD int fit_chi_index(const FitParams &p, int fit)
{
    if (p.chi_index) return p.chi_index[fit];
    return (int)((p.first_fit + fit) % (long long)p.n_chi);
}
QF_HD int fit_mf_index(const FitParams &p, int fit)
{
    if (p.mf_index) return p.mf_index[fit];
    return (int)((p.first_fit + fit) / (long long)p.n_chi);
}
QF_HD double2 fit_omega(const FitParams &p, int fit, int j)
{
    if (p.omega_rows) return p.omega_rows[(long long)j * p.n_times];   // placeholder; rows use row_omega()
    if (p.omega) return p.omega[(p.omega_shared ? 0ll : (long long)fit * p.n_modes) + j];
    int c = fit_chi_index(p, fit);
    double inv = p.inv_Mf[fit_mf_index(p, fit)];
    double df = p.delta_factor ? p.delta_factor[j] : 1.0;
    return form_omega(p.omega_tilde + 2ll * c * p.n_constituents, p.mode_ptr, j, inv, df);
}

// frequency of mode j at sample k: the per-row table of a dynamic fit, else the fit's own
QF_HD double2 row_omega(const FitParams &p, const double2 *om, int j, int k)
{
    return p.omega_rows ? p.omega_rows[(long long)j * p.n_times + k] : om[j];
}

// count a fit whose status word is non-zero and append (fit, status) to the launch's list of
// flagged fits (the counter's old value is the cursor), so that the host can repair exactly
// those fits without reading a status word per fit
QF_HD void note_status(const FitParams &p, int fit, int status)
{
    if (status != 0 && p.flagged_count) {
#ifdef QNMFIT_HOSTSIM
        const double old = *p.flagged_count;
        *p.flagged_count += 1.0;
#else
        const double old = atomicAdd(p.flagged_count, 1.0);
#endif
        if (p.flag_list && old < (double)p.flag_capacity) {
            const int cur = (int)old;
            p.flag_list[2 * cur] = fit;
            p.flag_list[2 * cur + 1] = status;
        }
    }
}

// numpy.linalg.lstsq truncates singular values below eps * max(M, N) * s_max
// (numpy/linalg/_linalg.py:2553).  The kernels decide "numpy MAY truncate here" in two steps:
//   1. prefilter on the diagonal of R: min |R_jj| <= prefilter(N) * eps * max(M, N) * max |R_jj|.
//      min/max of the diagonal of an unpivoted QR sits above s_min / s_max — by up to 684x on
//      overtone ladders of up to 12 columns, but the gap grows with the column count (1.7e5 at
//      19 columns, tests/test_gpu_parity.py::test_pair_kernel_many_fits_windows_series_and_eval;
//      2.8e4 on random label sets of 12 .. 15 columns, 1.5e5 at 20 .. 23; the bound is 2^(N-1))
//      — hence 1e5 up to eight columns and a factor 4 per column beyond, which from about
//      fourteen columns on sends every fit to step 2;
//   2. for the fits that pass it, an estimate of s_min by three steps of inverse iteration on
//      R^H R (two triangular solves each, start vector of ones; the estimate converges to
//      s_min from above) against MARGIN * eps * max(M, N) * ||R||_F  (||R||_F >= s_max).
// A flagged fit keeps the basic QR solution on the device; the host decides with the singular
// values of the exported factor and completes numpy's minimum-norm answer (qnmfits.py,
// _minimum_norm_from_factor).  Fits that are not flagged are full rank by numpy's criterion
// with a margin of >= MARGIN in s_min.
#define QNMFIT_RANK_PREFILTER 1.0e5
QF_HD double rank_prefilter(int n_modes)
{
    return n_modes <= 8 ? QNMFIT_RANK_PREFILTER : ldexp(QNMFIT_RANK_PREFILTER, 2 * (n_modes - 8));
}
#define QNMFIT_RANK_MARGIN 8.0
#define QNMFIT_EPS 2.220446049250313e-16

// Serial estimator (K1: lane 0 of the fit).  R(j, k), j < k < N: strictly upper entries;
// D(j): the real diagonal.  frob2 = ||R||_F^2.  Returns true when the fit must be flagged.
template <int N, class RF, class DF>
QF_HD bool rank_suspect_serial(const RF &R, const DF &D, double dim)
{
    double frob2 = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double d = D(j);
        if (d == 0.0) return true;
        frob2 = fma(d, d, frob2);
#pragma unroll
        for (int k = j + 1; k < N; ++k) {
            const double2 r = R(j, k);
            frob2 = fma(r.x, r.x, frob2);
            frob2 = fma(r.y, r.y, frob2);
        }
    }
    double2 x[N];
    const double x0 = 1.0 / sqrt((double)N);
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = make_double2(x0, 0.0);
    double nz = 1.0;
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
        // R^H y = x (forward): y_j = (x_j - sum_{i<j} conj(R_ij) y_i) / R_jj
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double ax = x[j].x, ay = x[j].y;
#pragma unroll
            for (int i = 0; i < j; ++i) {
                const double2 r = R(i, j);
                ax = fma(-r.x, x[i].x, ax); ax = fma(-r.y, x[i].y, ax);
                ay = fma(-r.x, x[i].y, ay); ay = fma(r.y, x[i].x, ay);
            }
            const double inv = 1.0 / D(j);
            x[j] = make_double2(ax * inv, ay * inv);
        }
        // R z = y (backward)
        double n2 = 0.0;
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            double ax = x[j].x, ay = x[j].y;
#pragma unroll
            for (int k = j + 1; k < N; ++k) {
                const double2 r = R(j, k);
                ax = fma(-r.x, x[k].x, ax); ax = fma(r.y, x[k].y, ax);
                ay = fma(-r.x, x[k].y, ay); ay = fma(-r.y, x[k].x, ay);
            }
            const double inv = 1.0 / D(j);
            x[j] = make_double2(ax * inv, ay * inv);
            n2 = fma(x[j].x, x[j].x, n2);
            n2 = fma(x[j].y, x[j].y, n2);
        }
        nz = sqrt(n2);
        if (!(nz < 1e300)) return true;              // overflow / NaN: as singular as it gets
        const double s = 1.0 / nz;
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = make_double2(x[j].x * s, x[j].y * s);
    }
    // ||(R^H R)^-1 x|| = nz with ||x|| = 1  =>  s_min^2 <= 1 / nz
    const double cut = QNMFIT_RANK_MARGIN * QNMFIT_EPS * dim;
    return !(1.0 / nz > cut * cut * frob2);
}

#define QNMFIT_ST_RANK_DEFICIENT_ 1
#define QNMFIT_ST_NONFINITE_ 2
#define QNMFIT_ST_UNDERDETERMINED_ 4

// ---- multi-GPU exchange (see include/qnmfit.h, qnmfit_fit_batch_peers) --------------
//
// peer_publish: called by the one thread that finished fit `fit`.  It stores the mismatch
// into every peer's result array — plain posted stores over NVLink that travel while the
// other fits are still being computed; nothing waits for them here (a system-scope fence
// per fit or per CTA was measured to cost ~6 us per wave of CTAs: the CTA cannot retire
// until the fence returns).  Their completion is guaranteed by the end of the kernel.
//
// peer_barrier_kernel (one warp, same stream, right after the fit kernel): lane r
// publishes the launch's flagged count and then its epoch into slot `rank` of peer r
// (release, system scope), then waits until peer r's epoch has arrived in the local flag
// array (acquire).  A peer that does not arrive within the timeout leaves NaN in its slot
// of the local flagged[] array and the host raises.
#ifndef QNMFIT_HOSTSIM
QF_DEV unsigned long long qf_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// fixed-order butterfly sum over the 32 lanes of a warp
QF_DEV double warp_sum(double v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// The same estimate computed by one warp (K2 / K3 / K4: N up to 64, R in shared memory).  All
// 32 lanes call it; the result is uniform.  Lane l holds entries l and l + 32 of the iteration
// vector in registers and both triangular solves run column by column: the owner of entry j
// finishes it (one multiplication by the reciprocal diagonal), broadcasts it with a shuffle,
// and every lane updates its own entries with one complex FMA — no reduction per step (the
// first form summed a row per step over the warp: two five-level butterflies and a division
// for each of the 6 N steps, 75 us at N = 40).  `x` (scratch in shared memory) is not used.
template <class RF, class DF>
QF_DEV bool rank_suspect_warp(const RF &R, const DF &D, int N, double dim, double2 *x, int lane)
{
    (void)x;
    const unsigned full = 0xffffffffu;
    double frob2 = 0.0;
    bool zero_diag = false;
    for (int j = 0; j < N; ++j) {
        const double d = D(j);
        zero_diag |= (d == 0.0);
        if (lane == 0) frob2 = fma(d, d, frob2);
        for (int k = j + 1 + lane; k < N; k += 32) {
            const double2 r = R(j, k);
            frob2 = fma(r.x, r.x, frob2);
            frob2 = fma(r.y, r.y, frob2);
        }
    }
    if (zero_diag) return true;
    frob2 = warp_sum(frob2);
    const int k0 = lane, k1 = lane + 32;
    const bool has0 = k0 < N, has1 = k1 < N;
    const double inv0 = has0 ? 1.0 / D(k0) : 0.0, inv1 = has1 ? 1.0 / D(k1) : 0.0;
    const double x0 = 1.0 / sqrt((double)N);
    double2 v0 = make_double2(has0 ? x0 : 0.0, 0.0), v1 = make_double2(has1 ? x0 : 0.0, 0.0);
    double nz = 1.0;
    for (int it = 0; it < 3; ++it) {
        for (int j = 0; j < N; ++j) {                // R^H y = x, column by column
            if (j == k0) { v0.x *= inv0; v0.y *= inv0; }
            if (j == k1) { v1.x *= inv1; v1.y *= inv1; }
            const double sx = j < 32 ? v0.x : v1.x, sy = j < 32 ? v0.y : v1.y;
            const double yx = __shfl_sync(full, sx, j & 31), yy = __shfl_sync(full, sy, j & 31);
            if (has0 && k0 > j) {                    // x_k -= conj(R_jk) y_j
                const double2 r = R(j, k0);
                v0.x = fma(-r.x, yx, v0.x); v0.x = fma(-r.y, yy, v0.x);
                v0.y = fma(-r.x, yy, v0.y); v0.y = fma(r.y, yx, v0.y);
            }
            if (has1 && k1 > j) {
                const double2 r = R(j, k1);
                v1.x = fma(-r.x, yx, v1.x); v1.x = fma(-r.y, yy, v1.x);
                v1.y = fma(-r.x, yy, v1.y); v1.y = fma(r.y, yx, v1.y);
            }
        }
        double n2 = 0.0;
        for (int k = N - 1; k >= 0; --k) {           // R z = y, column by column
            if (k == k0) { v0.x *= inv0; v0.y *= inv0; }
            if (k == k1) { v1.x *= inv1; v1.y *= inv1; }
            const double sx = k < 32 ? v0.x : v1.x, sy = k < 32 ? v0.y : v1.y;
            const double zx = __shfl_sync(full, sx, k & 31), zy = __shfl_sync(full, sy, k & 31);
            n2 = fma(zx, zx, n2);
            n2 = fma(zy, zy, n2);
            if (k0 < k) {                            // y_j -= R_jk z_k
                const double2 r = R(k0, k);
                v0.x = fma(-r.x, zx, v0.x); v0.x = fma(r.y, zy, v0.x);
                v0.y = fma(-r.x, zy, v0.y); v0.y = fma(-r.y, zx, v0.y);
            }
            if (k1 < k) {
                const double2 r = R(k1, k);
                v1.x = fma(-r.x, zx, v1.x); v1.x = fma(r.y, zy, v1.x);
                v1.y = fma(-r.x, zy, v1.y); v1.y = fma(-r.y, zx, v1.y);
            }
        }
        nz = sqrt(n2);
        if (!(nz < 1e300)) return true;
        const double sc = 1.0 / nz;
        v0.x *= sc; v0.y *= sc; v1.x *= sc; v1.y *= sc;
    }
    const double cut = QNMFIT_RANK_MARGIN * QNMFIT_EPS * dim;
    return !(1.0 / nz > cut * cut * frob2);
}

#ifndef QNMFIT_HOSTSIM
QF_DEV void peer_publish(const FitParams &p, int fit, double mm)
{
    if (p.n_peers <= 1) return;
    const long long gi = p.first_fit + fit;
    for (int r = 0; r < p.n_peers; ++r)
        if (r != p.peer_rank) p.peer_mismatch[r][gi] = mm;
}

#ifdef QNMFIT_DEFINE_PEER_BARRIER   /* defined once, in misc_kernels.cu */
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ FitParams p)
{
    const int W = p.n_peers, me = p.peer_rank, r = threadIdx.x;
    // the fit kernel has completed: the count of flagged fits is final; reset it for the next launch
    double flagged = 0.0;
    if (r == 0 && p.flagged_count)
        flagged = __longlong_as_double((long long)atomicExch((unsigned long long *)p.flagged_count, 0ull));
    flagged = __shfl_sync(0xffffffffu, flagged, 0);
    if (r >= W) return;
    *(volatile double *)(p.peer_flagged[r] + me) = flagged;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_flags[r] + me), "l"(p.peer_epoch) : "memory");
    const unsigned long long start = qf_globaltimer();
    const unsigned long long limit = p.peer_timeout_ns > 0 ? (unsigned long long)p.peer_timeout_ns : 30000000000ull;
    const unsigned long long *f = p.peer_flags[me] + r;
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= p.peer_epoch) break;
        if (qf_globaltimer() - start > limit) {
            *(volatile double *)(p.peer_flagged[me] + r) = __longlong_as_double(0x7ff8000000000000ll);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}
#endif  // QNMFIT_DEFINE_PEER_BARRIER
#else
static inline void peer_publish(const FitParams &, int, double) {}
#endif
