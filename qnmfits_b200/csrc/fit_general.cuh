// fit_general.cuh — K2: one CTA per fit, any N <= 64 columns, L >= 1 stacked series.
//
// Replaces, for one grid point / start time, the body of the reference's
// multimode_ringdown_fit (qnmfits/qnmfits.py:606-652) — and of ringdown_fit
// (qnmfits.py:274-293) when N exceeds what K1 keeps in registers:
//
//   a[(i,k), j] = coef[i][j] * exp(-i w_j (t_k - t0))       (qnmfits.py:628-631)
//
// The (L*K) x N matrix (13 MB at L=21, K=1000, N=40) is never formed.  The CTA
// streams tiles of TR rows: the exponentials E[k][j] of a chunk of time samples are
// evaluated once (direct exp/sincos, numpy's argument rounding) and reused by all L
// series of the chunk; the tile [coef (.) E | d] is written column-major to shared
// memory and folded into the resident upper-triangular factor [R | Q^H d] by N
// Householder reflections of the stacked matrix [R; tile] (sequential TSQR).  Each
// warp recomputes the reflector (norm by warp shuffles) and applies it to its share
// of the trailing columns, so one __syncthreads per reflection suffices.  Then
// back-substitution, and a second streaming pass for the model and the
// trapezoid-weighted sums of the sky-averaged mismatch (qnmfits.py:123-139).
#pragma once
#include "qnmfit_common.cuh"

#define K2_THREADS 256
#define K2_WARPS (K2_THREADS / 32)

struct GeneralSmem {
    double2 *R;       // [N][N+1] row-major; strictly-upper part + rhs column used
    double2 *T;       // [N+1][TR] column-major tile
    double2 *E;       // [TK][N]   exponentials of the time chunk
    double2 *om;      // [N]
    double2 *cc;      // [L][N]    coef * C (second pass)
    double2 *Cv;      // [N]
    double *diag;     // [2][N]    real diagonal, double-buffered by tile parity
    double *red;      // [K2_WARPS][4]
    static size_t bytes(int N, int L, int TR, int TK)
    {
        return sizeof(double2) * ((size_t)N * (N + 1) + (size_t)(N + 1) * TR + (size_t)TK * N + N
                                  + (size_t)L * N + N)
             + sizeof(double) * (2 * (size_t)N + K2_WARPS * 4);
    }
    __device__ void carve(void *base, int N, int L, int TR, int TK)
    {
        double2 *p = (double2 *)base;
        R = p; p += N * (N + 1);
        T = p; p += (N + 1) * TR;
        E = p; p += TK * N;
        om = p; p += N;
        cc = p; p += L * N;
        Cv = p; p += N;
        double *d = (double *)p;
        diag = d; d += 2 * N;
        red = d;
    }
};

// tile_rows / time_chunk are chosen by the host: TK = max(1, TR / L), rows used = TK * L.
__global__ void __launch_bounds__(K2_THREADS, 1)
fit_general_kernel(const __grid_constant__ FitParams p, const int TR, const int TK)
{
    QF_DYN_SMEM(smem_raw);
    const int N = p.n_modes, L = p.n_series;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fit = blockIdx.x;
    GeneralSmem sm;
    sm.carve(smem_raw, N, L, TR, TK);

    const int fi = input_fit(p, fit);             // index of the fit's inputs (see input_fit)
    int rb = p.row_begin ? p.row_begin[fi] : p.row_begin_all;
    int re = p.row_end ? p.row_end[fi] : p.row_end_all;
    const double t0 = p.t0 ? p.t0[fi] : p.t0_all;
    if (rb < 0) rb = 0;
    if (re > p.n_times) re = p.n_times;
    if (re < rb) re = rb;
    const int K = re - rb;
    const long long Mrows = (long long)K * L;
    const double2 *coef = nullptr;
    if (p.coef) {
        const int ci = p.coef_index ? p.coef_index[fi] : fit_chi_index(p, fi);
        coef = p.coef + (long long)ci * L * N;
    }

    for (int j = tid; j < N; j += K2_THREADS) sm.om[j] = fit_omega(p, fi, j);
    for (int e = tid; e < N * (N + 1); e += K2_THREADS) sm.R[e] = make_double2(0.0, 0.0);
    for (int e = tid; e < 2 * N; e += K2_THREADS) sm.diag[e] = 0.0;
    __syncthreads();

    int status = 0;
    const int ntiles = (K + TK - 1) / TK;

    if (!p.eval_only) {
        int parity = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int k0 = rb + tile * TK;
            const int kn = min(TK, re - k0);          // time samples in this tile
            const int rows = kn * L;                  // valid tile rows
            // 1. exponentials of the chunk
            for (int e = tid; e < kn * N; e += K2_THREADS) {
                const int kk = e / N, j = e - kk * N;
                const double tau = qf_sub_rn(p.times[k0 + kk], t0);
                sm.E[kk * N + j] = design_entry(row_omega(p, sm.om, j, k0 + kk), tau);
            }
            __syncthreads();
            // 2. tile [coef (.) E | d], column-major, zero padded to TR rows
            for (int e = tid; e < (N + 1) * TR; e += K2_THREADS) {
                const int k = e / TR, r = e - k * TR;
                double2 v = make_double2(0.0, 0.0);
                if (r < rows) {
                    const int i = r / kn, kk = r - i * kn;
                    if (k < N) {
                        v = sm.E[kk * N + k];
                        if (p.coef_rows) v = c_mul(p.coef_rows[((long long)i * N + k) * p.n_times + k0 + kk], v);
                        else if (coef) v = c_mul(coef[i * N + k], v);
                    } else {
                        v = p.data[(long long)i * p.series_stride + k0 + kk];
                    }
                }
                sm.T[e] = v;
            }
            __syncthreads();
            // 3. N reflections of [R; T]
            for (int j = 0; j < N; ++j) {
                const double2 *Tj = sm.T + (size_t)j * TR;
                double sig = 0.0;
                for (int r = lane; r < TR; r += 32) {
                    const double2 b = Tj[r];
                    sig = fma(b.x, b.x, sig);
                    sig = fma(b.y, b.y, sig);
                }
                sig = warp_sum(sig);
                const double rjj = sm.diag[parity * N + j];
                const double t = fma(rjj, rjj, sig);
                const bool ok = t > 1e-280;
                const double nrm = ok ? sqrt(t) : 0.0;
                const double ar = fabs(rjj);
                const double v0 = copysign(ar + nrm, rjj);
                const double den = nrm * (ar + nrm);
                const double beta = ok ? 1.0 / den : 0.0;
                if (tid == 0) sm.diag[(parity ^ 1) * N + j] = ok ? -copysign(nrm, rjj) : rjj;
                for (int k = j + 1 + warp; k <= N; k += K2_WARPS) {
                    double2 *Tk = sm.T + (size_t)k * TR;
                    // read R_jk before the shuffles: they order every lane's read
                    // ahead of lane 0's write below
                    double2 Rjk = sm.R[j * (N + 1) + k];
                    double sr = 0.0, si = 0.0;
                    for (int r = lane; r < TR; r += 32) {
                        const double2 b = Tj[r], c = Tk[r];
                        sr = fma(b.x, c.x, sr);
                        si = fma(b.x, c.y, si);
                        sr = fma(b.y, c.y, sr);
                        si = fma(-b.y, c.x, si);
                    }
                    sr = warp_sum(sr);
                    si = warp_sum(si);
                    sr = fma(v0, Rjk.x, sr) * beta;
                    si = fma(v0, Rjk.y, si) * beta;
                    if (lane == 0) {
                        Rjk.x = fma(-v0, sr, Rjk.x);
                        Rjk.y = fma(-v0, si, Rjk.y);
                        sm.R[j * (N + 1) + k] = Rjk;
                    }
                    for (int r = lane; r < TR; r += 32) {
                        const double2 b = Tj[r];
                        double2 c = Tk[r];
                        c.x = fma(-sr, b.x, c.x);
                        c.y = fma(-sr, b.y, c.y);
                        c.x = fma(si, b.y, c.x);
                        c.y = fma(-si, b.x, c.y);
                        Tk[r] = c;
                    }
                }
                __syncthreads();
            }
            parity ^= 1;
        }
        // 4. back-substitution (warp 0)
        if (warp == 0) {
            const double *dg = sm.diag + parity * N;
            double dmax = 0.0, dmin = 1e300;
            for (int j = lane; j < N; j += 32) {
                const double a = fabs(dg[j]);
                dmax = fmax(dmax, a);
                dmin = fmin(dmin, a);
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, s));
                dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, s));
            }
            const double dim = (double)(Mrows > N ? Mrows : N);
            if (!(dmin > rank_prefilter(N) * QNMFIT_EPS * dim * dmax)) {   // rare: confirm (qnmfit_common.cuh)
                if (rank_suspect_warp([&](int j, int k) { return sm.R[j * (N + 1) + k]; }, [&](int j) { return dg[j]; },
                                      N, dim, sm.Cv, lane))
                    status |= QNMFIT_ST_RANK_DEFICIENT_;
            }
            if (Mrows <= N) status |= QNMFIT_ST_UNDERDETERMINED_;
            if (p.R) {
                double2 *Rout = p.R + (long long)fit * N * (N + 1);
                for (int e = lane; e < N * (N + 1); e += 32) {
                    const int j = e / (N + 1), k = e - j * (N + 1);
                    double2 v = make_double2(0.0, 0.0);
                    if (k == j) v = make_double2(dg[j], 0.0);
                    else if (k > j) v = sm.R[e];
                    Rout[e] = v;
                }
            }
            for (int j = N - 1; j >= 0; --j) {
                double ax = 0.0, ay = 0.0;
                for (int k = j + 1 + lane; k < N; k += 32) {
                    const double2 Rjk = sm.R[j * (N + 1) + k];
                    const double2 c = sm.Cv[k];
                    ax = fma(Rjk.x, c.x, ax);
                    ax = fma(-Rjk.y, c.y, ax);
                    ay = fma(Rjk.x, c.y, ay);
                    ay = fma(Rjk.y, c.x, ay);
                }
                ax = warp_sum(ax);
                ay = warp_sum(ay);
                if (lane == 0) {
                    const double2 b = sm.R[j * (N + 1) + N];
                    const double d = dg[j];
                    double2 c = make_double2(0.0, 0.0);
                    if (d != 0.0) c = make_double2((b.x - ax) / d, (b.y - ay) / d);
                    sm.Cv[j] = c;
                }
                __syncwarp();
            }
            for (int j = lane; j < N; j += 32) {
                const double2 c = sm.Cv[j];
                if (p.C) p.C[(long long)fit * N + j] = c;
                if (!(fabs(c.x) < 1e300) || !(fabs(c.y) < 1e300)) status |= QNMFIT_ST_NONFINITE_;
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) status |= __shfl_xor_sync(0xffffffffu, status, s);
        }
    } else {
        for (int j = tid; j < N; j += K2_THREADS) sm.Cv[j] = p.C[(long long)fit * N + j];
    }
    __syncthreads();

    // 5. coef * C, then the second streaming pass
    for (int e = tid; e < L * N; e += K2_THREADS) {
        const int j = e % N;
        sm.cc[e] = coef ? c_mul(coef[e], sm.Cv[j]) : sm.Cv[j];
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = rb + tile * TK;
        const int kn = min(TK, re - k0);
        const int rows = kn * L;
        __syncthreads();
        for (int e = tid; e < kn * N; e += K2_THREADS) {
            const int kk = e / N, j = e - kk * N;
            const double tau = qf_sub_rn(p.times[k0 + kk], t0);
            sm.E[kk * N + j] = design_entry(row_omega(p, sm.om, j, k0 + kk), tau);
        }
        __syncthreads();
        for (int r = tid; r < rows; r += K2_THREADS) {
            const int i = r / kn, kk = r - i * kn;
            const double2 *Er = sm.E + kk * N;
            const double2 *ci = sm.cc + i * N;
            double mx = 0.0, my = 0.0;
            for (int j = 0; j < N; ++j) {
                const double2 a = Er[j];
                const double2 c = p.coef_rows
                    ? c_mul(p.coef_rows[((long long)i * N + j) * p.n_times + k0 + kk], sm.Cv[j]) : ci[j];
                mx = fma(a.x, c.x, mx);
                my = fma(a.x, c.y, my);
                mx = fma(-a.y, c.y, mx);
                my = fma(a.y, c.x, my);
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
    if (lane == 0) {
        sm.red[warp * 4 + 0] = s0; sm.red[warp * 4 + 1] = s1;
        sm.red[warp * 4 + 2] = s2; sm.red[warp * 4 + 3] = s3;
    }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int w = 0; w < K2_WARPS; ++w) {
            a0 += sm.red[w * 4 + 0]; a1 += sm.red[w * 4 + 1];
            a2 += sm.red[w * 4 + 2]; a3 += sm.red[w * 4 + 3];
        }
        const double mm = 1.0 - a0 / sqrt(a1 * a2);
        p.mismatch[fit] = mm;
        if (p.residual) p.residual[fit] = a3;
        if (p.status) p.status[fit] = status;
        note_status(p, fit, status);
        peer_publish(p, fit, mm);
    }
}
