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

// count a fit whose status word is non-zero
QF_HD void note_status(const FitParams &p, int status)
{
    if (status != 0 && p.flagged_count) {
#ifdef QNMFIT_HOSTSIM
        *p.flagged_count += 1.0;
#else
        atomicAdd(p.flagged_count, 1.0);
#endif
    }
}

// numpy.linalg.lstsq truncates singular values below eps * max(M, N) * s_max.  The kernels
// only see the diagonal of R, and min |R_jj| / max |R_jj| can sit several hundred times above
// s_min / s_max (measured up to 684x on overtone ladders), so the flag is raised with that
// margin: it means "numpy MAY truncate here"; the host decides with the singular values of
// the exported factor (qnmfits.py, _minimum_norm_from_factor).
#define QNMFIT_RANK_FLAG_MARGIN 1024.0
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

// fixed-order butterfly sum over the 32 lanes of a warp
QF_DEV double warp_sum(double v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

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
