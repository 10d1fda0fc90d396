// fit_struct.cuh — K3: one CTA per fit, N pivot columns + L right-hand sides <= 64,
// structured two-phase Householder QR with register-resident tiles.
//
// Replaces, for one start time / grid point, the body of the reference's
// multimode_ringdown_fit (qnmfits/qnmfits.py:606-652) and of ringdown_fit
// (qnmfits.py:274-293) when N > 8.  The reference stacks L copies of the K x N matrix of
// exponentials, each scaled column-wise by the mixing coefficients of one spherical mode,
//
//   a[(i,k), j] = coef[i][j] * E[k][j],   E[k][j] = exp(-i w_j (t_k - t0))   (qnmfits.py:628-631)
//
// and hands the (L K) x N matrix to LAPACK.  Here the structure is used:
//
//   phase 1   QR of [E | d_1 ... d_L]  (K rows, N + L columns, only the first N are
//             pivots):  E = Q_E R_E,  Y_i = Q_E^H d_i.  Then A_i = Q_E (R_E D_i) with
//             D_i = diag(coef[i][:]), so the big problem equals the small one
//   phase 2   QR of the stacked triangles [R_E D_i | Y_i], i = 1..L  (L N rows, N + 1
//             columns), rows ordered by their first non-zero column so that the
//             reflections of a tile start there;
//   then      back-substitution, residual = |annihilated rhs entries|^2 of both phases
//             (no cancellation), and the mismatch from the by-products on uniform grids
//             (as in K1) or a second streaming pass otherwise.
//
// Flops per fit: 8KN^2 + 16KNL (phase 1) + ~(8/3) L N^3 (phase 2) instead of 8 (LK) N^2
// — 4.4e7 instead of 2.9e8 at L=21, K=1000, N=40.
//
// Mapping.  The CTA (256 threads) owns one fit.  A tile is TR = 8 G rows; thread
// (column pair p, row group g) keeps rows 8g..8g+7 of columns 2p and 2p+1 in registers,
// G lanes of a warp share a column pair (G = 8, 16, 32 chosen so that ceil((N+L)/2) G
// <= 256).  Reflection j: the owner lanes reduce the column norm (xor-shuffles inside
// the group), form the reflector scalars and publish v through shared memory (double
// buffered: one __syncthreads per reflection); every thread loads its 8 rows of v ONCE
// (128 B — shared-memory bandwidth, not FP64, bounded the first version, which re-read v
// for a single column), forms the dot products of both its columns, reduces them over
// the G lanes, updates its entries of row j of R (shared memory) and its rows.  Warps
// whose columns are all retired only pass the barrier.
#pragma once
#include "qnmfit_common.cuh"
#include "fit_general.cuh"

#ifndef QNMFIT_HOSTSIM

#define K3_THREADS 256
#define K3_WARPS (K3_THREADS / 32)
#define K3_RPT 8                  // rows per thread
#define K3_VST (K3_RPT + 1)       // row-group stride of the v buffer (double2): conflict-free LDS.128
#define K3_TK 16                  // time chunk of the second pass

struct StructSmem {
    double2 *R1;      // [N][NC]    R_E (strictly upper) | Y
    double2 *R2;      // [N][N+1]   second-phase factor | Q^H d   (aliases R1 when phase 2 is skipped)
    double2 *vbuf;    // [2][G * K3_VST]
    double2 *om, *qq, *qw;   // [N]
    double2 *coef;    // [L][N]
    double2 *Cv;      // [N]
    double2 *cc;      // [L][N]     coef * C
    double2 *E;       // [K3_TK][N] second pass / end rows
    double *diag1, *diag2;   // [N]
    double *scal;     // [2][2]     v0, beta
    double *red;      // [K3_WARPS][8]
    double *ends;     // [64][3]     end-point terms per series
    static size_t bytes(int N, int L, int G)
    {
        const size_t NC = (size_t)N + L;
        return sizeof(double2) * ((size_t)N * NC + (size_t)N * (N + 1) + 2 * (size_t)G * K3_VST + 3 * (size_t)N
                                  + 2 * (size_t)L * N + N + (size_t)K3_TK * N)
             + sizeof(double) * (2 * (size_t)N + 4 + K3_WARPS * 8 + 64 * 3);
    }
    __device__ void carve(void *base, int N, int L, int G)
    {
        const int NC = N + L;
        double2 *p = (double2 *)base;
        R1 = p; p += N * NC;
        R2 = p; p += N * (N + 1);
        vbuf = p; p += 2 * G * K3_VST;
        om = p; p += N; qq = p; p += N; qw = p; p += N;
        coef = p; p += L * N;
        Cv = p; p += N;
        cc = p; p += L * N;
        E = p; p += K3_TK * N;
        double *d = (double *)p;
        diag1 = d; d += N;
        diag2 = d; d += N;
        scal = d; d += 4;
        red = d; d += K3_WARPS * 8;
        ends = d;
    }
};

// sum of |x|^2 over the thread's rows of one column, four chains
__device__ __forceinline__ double k3_norm2(const double2 (&X)[K3_RPT])
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int r = 0; r < K3_RPT; r += 4) {
        a0 = fma(X[r].x, X[r].x, a0); a1 = fma(X[r + 1].x, X[r + 1].x, a1);
        a2 = fma(X[r + 2].x, X[r + 2].x, a2); a3 = fma(X[r + 3].x, X[r + 3].x, a3);
        a0 = fma(X[r].y, X[r].y, a0); a1 = fma(X[r + 1].y, X[r + 1].y, a1);
        a2 = fma(X[r + 2].y, X[r + 2].y, a2); a3 = fma(X[r + 3].y, X[r + 3].y, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// conj(v) . x over the thread's rows, two complex chains
__device__ __forceinline__ void k3_dot(const double2 (&v)[K3_RPT], const double2 (&X)[K3_RPT], double &sr, double &si)
{
    double sr0 = 0.0, si0 = 0.0, sr1 = 0.0, si1 = 0.0;
#pragma unroll
    for (int r = 0; r < K3_RPT; r += 2) {
        sr0 = fma(v[r].x, X[r].x, sr0); si0 = fma(v[r].x, X[r].y, si0);
        sr1 = fma(v[r + 1].x, X[r + 1].x, sr1); si1 = fma(v[r + 1].x, X[r + 1].y, si1);
        sr0 = fma(v[r].y, X[r].y, sr0); si0 = fma(-v[r].y, X[r].x, si0);
        sr1 = fma(v[r + 1].y, X[r + 1].y, sr1); si1 = fma(-v[r + 1].y, X[r + 1].x, si1);
    }
    sr = sr0 + sr1;
    si = si0 + si1;
}

// x -= (pr + i pi) v
__device__ __forceinline__ void k3_axpy(const double2 (&v)[K3_RPT], double2 (&X)[K3_RPT], double pr, double pi)
{
#pragma unroll
    for (int r = 0; r < K3_RPT; ++r) {
        double bx = X[r].x, by = X[r].y;
        bx = fma(-pr, v[r].x, bx);
        by = fma(-pr, v[r].y, by);
        bx = fma(pi, v[r].y, bx);
        by = fma(-pi, v[r].x, by);
        X[r].x = bx;
        X[r].y = by;
    }
}

// Reflections jstart..N-1 of [R; tile].  X0 / X1: the thread's 8 rows of columns c0 = 2p
// and c0 + 1 (row group g).  nrm2 must hold the thread's partial |column jstart|^2 when
// it owns that column.  Columns >= ncols do not exist; columns < jstart of the tile
// must be zero.
template <int G>
__device__ __forceinline__ void k3_reflect(double2 (&X0)[K3_RPT], double2 (&X1)[K3_RPT], double &nrm2,
                                           const int c0, const int g, const int warp, const int ncols,
                                           const int N, const int jstart, double2 *Rm, const int ldr,
                                           double *diag, double2 *vbuf, double *scal, int &buf)
{
    constexpr int CPW = 2 * (32 / G);    // columns per warp
    const int wfirst = warp * CPW, wlast = wfirst + CPW - 1;
#pragma unroll 1
    for (int j = jstart; j < N; ++j) {
        if (wfirst <= j && j <= wlast) {             // the owner warp
            double sig = nrm2;
#pragma unroll
            for (int s = 1; s < G; s <<= 1) sig += __shfl_xor_sync(0xffffffffu, sig, s);
            if ((j >> 1) == (c0 >> 1)) {
                const double r = diag[j];
                const double t = fma(r, r, sig) + 1e-300;   // see fit_small.cuh: an all-zero column needs no branch
                const double y = qf_rsqrt(t);
                const double nrm = t * y;
                const double ar = fabs(r);
                const double v0 = copysign(ar + nrm, r);
                const double beta = qf_rcp(nrm * (ar + nrm));
                if (g == 0) {
                    diag[j] = -copysign(nrm, r);
                    scal[buf * 2] = v0;
                    scal[buf * 2 + 1] = beta;
                }
                double2 *vw = vbuf + buf * (G * K3_VST) + K3_VST * g;
                if (j & 1) {
#pragma unroll
                    for (int r2 = 0; r2 < K3_RPT; ++r2) vw[r2] = X1[r2];
                } else {
#pragma unroll
                    for (int r2 = 0; r2 < K3_RPT; ++r2) vw[r2] = X0[r2];
                }
            }
        }
        __syncthreads();
        if (wlast > j && wfirst < ncols) {           // warp still has trailing columns
            const bool act0 = c0 > j && c0 < ncols, act1 = c0 + 1 > j && c0 + 1 < ncols;
            const double2 *vs = vbuf + buf * (G * K3_VST) + K3_VST * g;
            double2 v[K3_RPT];
#pragma unroll
            for (int r = 0; r < K3_RPT; ++r) v[r] = vs[r];
            double sr0, si0, sr1, si1;
            k3_dot(v, X0, sr0, si0);
            k3_dot(v, X1, sr1, si1);
#pragma unroll
            for (int s = 1; s < G; s <<= 1) {
                sr0 += __shfl_xor_sync(0xffffffffu, sr0, s);
                si0 += __shfl_xor_sync(0xffffffffu, si0, s);
                sr1 += __shfl_xor_sync(0xffffffffu, sr1, s);
                si1 += __shfl_xor_sync(0xffffffffu, si1, s);
            }
            const double v0 = scal[buf * 2], beta = scal[buf * 2 + 1];
            if (act0) {
                double2 Rjc = Rm[j * ldr + c0];
                const double pr = fma(v0, Rjc.x, sr0) * beta;
                const double pi = fma(v0, Rjc.y, si0) * beta;
                if (g == 0) Rm[j * ldr + c0] = make_double2(fma(-v0, pr, Rjc.x), fma(-v0, pi, Rjc.y));
                k3_axpy(v, X0, pr, pi);
            }
            if (act1) {
                double2 Rjc = Rm[j * ldr + c0 + 1];
                const double pr = fma(v0, Rjc.x, sr1) * beta;
                const double pi = fma(v0, Rjc.y, si1) * beta;
                if (g == 0) Rm[j * ldr + c0 + 1] = make_double2(fma(-v0, pr, Rjc.x), fma(-v0, pi, Rjc.y));
                k3_axpy(v, X1, pr, pi);
            }
            if (wfirst <= j + 1 && j + 1 <= wlast)   // look-ahead for the next pivot
                nrm2 = ((j + 1) & 1) ? k3_norm2(X1) : k3_norm2(X0);
        }
        buf ^= 1;
    }
}

// Rows first..first+7 of column c of the phase-1 matrix [E | d_1..d_L] (zero outside the window).
__device__ __forceinline__ void k3_load1(double2 (&X)[K3_RPT], const FitParams &p, const StructSmem &sm, const int c,
                                         const int N, const int NC, const int first, const int re, const double t0,
                                         double &sdd)
{
    const double2 zero = make_double2(0.0, 0.0);
    if (c < N) {
        if (p.dt_nominal > 0.0 && !p.omega_rows) {
            if (first < re) {
                const double dt = p.dt_nominal;
                double tau = qf_sub_rn(p.times[first], t0);
                double2 z = design_entry(sm.om[c], tau);
                const double2 q = sm.qq[c], w = sm.qw[c];
#pragma unroll
                for (int r = 0; r < K3_RPT; ++r) {
                    X[r] = first + r < re ? z : zero;
                    const int kn = first + r + 1 < re ? first + r + 1 : re - 1;
                    const double tau_n = qf_sub_rn(p.times[kn], t0);
                    const double de = qf_sub_rn(qf_sub_rn(tau_n, tau), dt);
                    tau = tau_n;
                    z = c_mul(z, make_double2(fma(w.x, de, q.x), fma(w.y, de, q.y)));
                }
            } else {
#pragma unroll
                for (int r = 0; r < K3_RPT; ++r) X[r] = zero;
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < K3_RPT; ++r) {
                double2 e = zero;
                if (first + r < re) e = design_entry(row_omega(p, sm.om, c, first + r), qf_sub_rn(p.times[first + r], t0));
#pragma unroll
                for (int q = 0; q < K3_RPT; ++q) if (q == r) X[q] = e;   // static register index
            }
        }
    } else if (c < NC) {
        const double2 *dsrc = p.data + (long long)(c - N) * p.series_stride;
#pragma unroll
        for (int r = 0; r < K3_RPT; ++r) X[r] = first + r < re ? dsrc[first + r] : zero;
        sdd += k3_norm2(X);
    } else {
#pragma unroll
        for (int r = 0; r < K3_RPT; ++r) X[r] = zero;
    }
}

// Rows q0..q0+7 of column c of the phase-2 matrix: row q = rr * L + i is row rr of
// [R_E D_i | Y_i]  (rr-major, so that a tile's rows share their leading zeros).
__device__ __forceinline__ void k3_load2(double2 (&X)[K3_RPT], const StructSmem &sm, const int c, const int N,
                                         const int L, const int NC, const int q0, const int rows2)
{
#pragma unroll
    for (int r = 0; r < K3_RPT; ++r) {
        const int q = q0 + r;
        double2 v = make_double2(0.0, 0.0);
        if (q < rows2 && c <= N) {
            const int rr = q / L, i = q - rr * L;
            if (c == N) v = sm.R1[rr * NC + N + i];
            else if (c == rr) { const double2 cf = sm.coef[i * N + c]; const double d = sm.diag1[rr]; v = make_double2(cf.x * d, cf.y * d); }
            else if (c > rr) v = c_mul(sm.coef[i * N + c], sm.R1[rr * NC + c]);
        }
        X[r] = v;
    }
}

template <int G>
__global__ void __launch_bounds__(K3_THREADS, 2) fit_struct_kernel(const FitParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PPW = 32 / G;            // column pairs per warp
    constexpr int TR = K3_RPT * G;
    const int N = p.n_modes, L = p.n_series, NC = N + L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = 2 * (warp * PPW + lane / G), g = lane % G;
    const int fit = blockIdx.x;
    StructSmem sm;
    sm.carve(smem_raw, N, L, G);

    int rb = p.row_begin ? p.row_begin[fit] : p.row_begin_all;
    int re = p.row_end ? p.row_end[fit] : p.row_end_all;
    const double t0 = p.t0 ? p.t0[fit] : p.t0_all;
    if (rb < 0) rb = 0;
    if (re > p.n_times) re = p.n_times;
    if (re < rb) re = rb;
    const int K = re - rb;
    const long long Mrows = (long long)K * L;
    const double2 *coef_g = nullptr;
    if (p.coef) {
        const int ci = p.coef_index ? p.coef_index[fit] : fit_chi_index(p, fit);
        coef_g = p.coef + (long long)ci * L * N;
    }
    const bool two_phase = (coef_g != nullptr) || L > 1;

    for (int j = tid; j < N; j += K3_THREADS) {
        const double2 w = fit_omega(p, fit, j);
        sm.om[j] = w;
        if (p.dt_nominal > 0.0) {
            const double2 q = design_entry(w, p.dt_nominal);
            sm.qq[j] = q;
            sm.qw[j] = c_mul(q, make_double2(w.y, -w.x));
        }
        sm.diag1[j] = 0.0;
        sm.diag2[j] = 0.0;
    }
    for (int e = tid; e < N * NC; e += K3_THREADS) sm.R1[e] = make_double2(0.0, 0.0);
    for (int e = tid; e < N * (N + 1); e += K3_THREADS) sm.R2[e] = make_double2(0.0, 0.0);
    for (int e = tid; e < L * N; e += K3_THREADS) sm.coef[e] = coef_g ? coef_g[e] : make_double2(1.0, 0.0);
    __syncthreads();

    int status = 0;
    double sdd = 0.0, res2 = 0.0;
    int buf = 0;
    double2 X0[K3_RPT], X1[K3_RPT];

    if (!p.eval_only) {
        // ---------------- phase 1: [E | d_1..d_L] ----------------
        const int ntiles = (K + TR - 1) / TR;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            const int first = rb + tile * TR + K3_RPT * g;
            k3_load1(X0, p, sm, c0, N, NC, first, re, t0, sdd);
            k3_load1(X1, p, sm, c0 + 1, N, NC, first, re, t0, sdd);
            double nrm2 = (c0 == 0) ? k3_norm2(X0) : 0.0;
            k3_reflect<G>(X0, X1, nrm2, c0, g, warp, NC, N, 0, sm.R1, NC, sm.diag1, sm.vbuf, sm.scal, buf);
            if (c0 >= N && c0 < NC) res2 += k3_norm2(X0);
            if (c0 + 1 >= N && c0 + 1 < NC) res2 += k3_norm2(X1);
        }
        __syncthreads();

        // ---------------- phase 2: stacked [R_E D_i | Y_i] ----------------
        double2 *Rf = sm.R1;
        double *dgf = sm.diag1;
        int ldf = NC;
        if (two_phase) {
            Rf = sm.R2; dgf = sm.diag2; ldf = N + 1;
            const int rows2 = N * L;
            const int ntiles2 = (rows2 + TR - 1) / TR;
#pragma unroll 1
            for (int tile = 0; tile < ntiles2; ++tile) {
                const int q0 = tile * TR + K3_RPT * g;
                const int jstart = (tile * TR) / L;        // first non-zero column of the tile
                k3_load2(X0, sm, c0, N, L, NC, q0, rows2);
                k3_load2(X1, sm, c0 + 1, N, L, NC, q0, rows2);
                double nrm2 = 0.0;
                if (c0 == (jstart & ~1)) nrm2 = (jstart & 1) ? k3_norm2(X1) : k3_norm2(X0);
                k3_reflect<G>(X0, X1, nrm2, c0, g, warp, N + 1, N, jstart, sm.R2, N + 1, sm.diag2, sm.vbuf, sm.scal,
                              buf);
                if (c0 == N) res2 += k3_norm2(X0);
                if (c0 + 1 == N) res2 += k3_norm2(X1);
            }
            __syncthreads();
        }

        // ---------------- back-substitution (warp 0) ----------------
        if (warp == 0) {
            double dmax = 0.0, dmin = 1e300;
            for (int j = lane; j < N; j += 32) {
                const double a = fabs(dgf[j]);
                dmax = fmax(dmax, a);
                dmin = fmin(dmin, a);
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, s));
                dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, s));
            }
            const double dim = (double)(Mrows > N ? Mrows : N);
            if (!(dmin > 2.220446049250313e-16 * dim * dmax)) status |= QNMFIT_ST_RANK_DEFICIENT_;
            if (Mrows <= N) status |= QNMFIT_ST_UNDERDETERMINED_;
            if (p.R) {
                double2 *Rout = p.R + (long long)fit * N * (N + 1);
                for (int e = lane; e < N * (N + 1); e += 32) {
                    const int j = e / (N + 1), k = e - j * (N + 1);
                    double2 v = make_double2(0.0, 0.0);
                    if (k == j) v = make_double2(dgf[j], 0.0);
                    else if (k > j) v = Rf[j * ldf + k];
                    Rout[e] = v;
                }
            }
            for (int j = N - 1; j >= 0; --j) {
                double ax = 0.0, ay = 0.0;
                for (int k = j + 1 + lane; k < N; k += 32) {
                    const double2 Rjk = Rf[j * ldf + k];
                    const double2 cv = sm.Cv[k];
                    ax = fma(Rjk.x, cv.x, ax);
                    ax = fma(-Rjk.y, cv.y, ax);
                    ay = fma(Rjk.x, cv.y, ay);
                    ay = fma(Rjk.y, cv.x, ay);
                }
                ax = warp_sum(ax);
                ay = warp_sum(ay);
                if (lane == 0) {
                    const double2 b = Rf[j * ldf + N];
                    const double d = dgf[j];
                    double2 cv = make_double2(0.0, 0.0);
                    if (d != 0.0) cv = make_double2((b.x - ax) / d, (b.y - ay) / d);
                    sm.Cv[j] = cv;
                }
                __syncwarp();
            }
            for (int j = lane; j < N; j += 32) {
                const double2 cv = sm.Cv[j];
                if (p.C) p.C[(long long)fit * N + j] = cv;
                if (!(fabs(cv.x) < 1e300) || !(fabs(cv.y) < 1e300)) status |= QNMFIT_ST_NONFINITE_;
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) status |= __shfl_xor_sync(0xffffffffu, status, s);
        }
        // block sums of sdd / res2 in a fixed order
        sdd = warp_sum(sdd);
        res2 = warp_sum(res2);
        if (lane == 0) { sm.red[warp * 8] = sdd; sm.red[warp * 8 + 1] = res2; }
        __syncthreads();
        sdd = 0.0; res2 = 0.0;
        for (int w = 0; w < K3_WARPS; ++w) { sdd += sm.red[w * 8]; res2 += sm.red[w * 8 + 1]; }
        __syncthreads();

        if (p.fast_mismatch && K > 0) {
            // mismatch from the by-products (see small_fast_finalize in fit_small.cuh)
            double2 *Rfin = two_phase ? sm.R2 : sm.R1;
            const int ldfin = two_phase ? N + 1 : NC;
            double cn2 = 0.0;
            for (int j = 0; j < N; ++j) {   // every thread: same fixed order
                const double2 b = Rfin[j * ldfin + N];
                cn2 = fma(b.x, b.x, cn2);
                cn2 = fma(b.y, b.y, cn2);
            }
            for (int e = tid; e < 2 * N; e += K3_THREADS) {
                const int j = e % N;
                const int row = e < N ? rb : re - 1;
                sm.E[e] = design_entry(row_omega(p, sm.om, j, row), qf_sub_rn(p.times[row], t0));
            }
            __syncthreads();
            double e0 = 0.0, e1 = 0.0, e2 = 0.0;      // end-point terms of the three sums
            for (int i = tid; i < L; i += K3_THREADS) {
#pragma unroll 1
                for (int end = 0; end < 2; ++end) {
                    const double2 *Er = sm.E + end * N;
                    double mx = 0.0, my = 0.0;
                    for (int j = 0; j < N; ++j) {
                        const double2 a = Er[j];
                        const double2 cj = c_mul(sm.coef[i * N + j], sm.Cv[j]);
                        mx = fma(a.x, cj.x, mx);
                        my = fma(a.x, cj.y, my);
                        mx = fma(-a.y, cj.y, mx);
                        my = fma(a.y, cj.x, my);
                    }
                    const double2 d = p.data[(long long)i * p.series_stride + (end ? re - 1 : rb)];
                    e0 += fma(mx, d.x, my * d.y);
                    e1 += fma(mx, mx, my * my);
                    e2 += fma(d.x, d.x, d.y * d.y);
                }
            }
            // L < 64 <= K3_THREADS: at most one series per thread; sum in series order
            double *ends = sm.ends;
            if (tid < L) { ends[tid * 3] = e0; ends[tid * 3 + 1] = e1; ends[tid * 3 + 2] = e2; }
            __syncthreads();
            if (tid == 0) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int i = 0; i < L; ++i) { a0 += ends[i * 3]; a1 += ends[i * 3 + 1]; a2 += ends[i * 3 + 2]; }
                const double num = cn2 - 0.5 * a0, n1 = cn2 - 0.5 * a1, n2 = sdd - 0.5 * a2;
                p.mismatch[fit] = 1.0 - num / sqrt(n1 * n2);
                if (p.residual) p.residual[fit] = res2;
                if (p.status) p.status[fit] = status;
                note_status(p, status);
            }
            return;
        }
    } else {
        for (int j = tid; j < N; j += K3_THREADS) sm.Cv[j] = p.C[(long long)fit * N + j];
    }
    __syncthreads();

    // ---------------- second pass: model and trapezoid-weighted sums ----------------
    for (int e = tid; e < L * N; e += K3_THREADS) sm.cc[e] = c_mul(sm.coef[e], sm.Cv[e % N]);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int nchunks = (K + K3_TK - 1) / K3_TK;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int k0 = rb + ch * K3_TK;
        const int kn = min(K3_TK, re - k0);
        const int rows = kn * L;
        __syncthreads();
        for (int e = tid; e < kn * N; e += K3_THREADS) {
            const int kk = e / N, j = e - kk * N;
            sm.E[kk * N + j] = design_entry(row_omega(p, sm.om, j, k0 + kk), qf_sub_rn(p.times[k0 + kk], t0));
        }
        __syncthreads();
        for (int r = tid; r < rows; r += K3_THREADS) {
            const int i = r / kn, kk = r - i * kn;
            const double2 *Er = sm.E + kk * N;
            const double2 *ci = sm.cc + i * N;
            double mx = 0.0, my = 0.0;
            for (int j = 0; j < N; ++j) {
                const double2 a = Er[j], cj = ci[j];
                mx = fma(a.x, cj.x, mx);
                my = fma(a.x, cj.y, my);
                mx = fma(-a.y, cj.y, mx);
                my = fma(a.y, cj.x, my);
            }
            const double2 d = p.data[(long long)i * p.series_stride + k0 + kk];
            const int row = k0 + kk;
            if (p.model) p.model[(long long)fit * p.model_stride + (long long)i * K + (row - rb)] = make_double2(mx, my);
            const int rm = row - 1 < rb ? rb : row - 1;
            const int rp = row + 1 > re - 1 ? re - 1 : row + 1;
            const double w = 0.5 * (p.times[rp] - p.times[rm]);
            s0 = fma(w, fma(mx, d.x, my * d.y), s0);
            s1 = fma(w, fma(mx, mx, my * my), s1);
            s2 = fma(w, fma(d.x, d.x, d.y * d.y), s2);
            const double ex = mx - d.x, ey = my - d.y;
            s3 += fma(ex, ex, ey * ey);
        }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    __syncthreads();
    if (lane == 0) {
        sm.red[warp * 8 + 0] = s0; sm.red[warp * 8 + 1] = s1;
        sm.red[warp * 8 + 2] = s2; sm.red[warp * 8 + 3] = s3;
    }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int w = 0; w < K3_WARPS; ++w) {
            a0 += sm.red[w * 8 + 0]; a1 += sm.red[w * 8 + 1];
            a2 += sm.red[w * 8 + 2]; a3 += sm.red[w * 8 + 3];
        }
        p.mismatch[fit] = 1.0 - a0 / sqrt(a1 * a2);
        // the second pass recomputes |model - data|^2 directly; on the solve path the
        // annihilated-entry sum of the factorisation is the better conditioned value
        if (p.residual) p.residual[fit] = p.eval_only ? a3 : res2;
        if (p.status) p.status[fit] = status;
        note_status(p, status);
    }
}
#endif  // !QNMFIT_HOSTSIM
