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
// — 3.0e7 instead of 2.9e8 at L=21, K=1000, N=40.
//
// Mapping.  G lanes (1, 2, 4 or 8, template parameter) own ONE column of the tile, RPT rows
// each, in registers (tile height G * RPT).  The dot product b^H x_c needs log2(G) shuffle
// levels of two doubles (none for G = 1), the column norm of the next pivot is accumulated
// in the update pass, the reflector b is broadcast through a padded, double-buffered
// shared-memory buffer (one __syncthreads per reflection), and one CTA = one fit =
// G * ceil((N+L)/32) warps, so several fits share an SM and fill each other's latency.
// R_E and the second factor are kept in packed upper-triangular form (~46 KB per fit at
// N=40, L=21).  Earlier forms (a column PAIR per thread with 8-lane shuffle reductions and
// the reflector held in registers; reflector re-read per column with 4-way bank conflicts)
// and the measurements that retired them are in profiles/README.md; tools/k3_time.py
// sweeps (G, RPT) through QNMFIT_K3G.
#pragma once
#include "qnmfit_common.cuh"

#define K3C_MAXROWS 64             // tallest tile (rows) any variant uses: sizes the v buffer
#define K3C_TK 16                 // time chunk of the second pass

// Packed factor with W+1 columns (W = columns right of column 0): row j holds columns
// j+1..W; the real diagonal lives in its own array.
struct PackedR {
    double2 *a;
    int W;
    __device__ __forceinline__ double2 &at(int j, int k) const { return a[j * W - (j * (j - 1)) / 2 + (k - j - 1)]; }
    __host__ __device__ static size_t entries(int N, int W) { return (size_t)N * W - (size_t)N * (N - 1) / 2; }
};

struct Struct3Smem {
    PackedR R1;       // N rows, N+L columns: R_E | Y
    PackedR R2;       // N rows, N+1 columns: second factor | Q^H d
    double2 *vbuf;    // [2][K3C_MAXROWS + 8]   (row groups padded by one entry)
    double2 *om, *qq, *qw, *Cv;   // [N]
    double2 *scratch; // second pass / end rows: cc [L][N] then E [K3C_TK][N]
    double *diag1, *diag2;        // [N]
    double *scal;     // [2][2]
    double *red;      // [16][8]
    double *ends;     // [64][3]
    static size_t bytes(int N, int L)
    {
        const size_t r1 = PackedR::entries(N, N + L - 1), r2 = PackedR::entries(N, N);
        const size_t scr = (size_t)L * N + (size_t)K3C_TK * N;
        // the scratch area overlays R1 when phase 2 runs (R1 is dead by then)
        const bool two_phase_possible = L > 1;
        const size_t r1_alloc = two_phase_possible ? (r1 > scr ? r1 : scr) : r1;
        const size_t scr_alloc = two_phase_possible ? 0 : scr;
        return sizeof(double2) * (r1_alloc + r2 + scr_alloc + 2 * (size_t)(K3C_MAXROWS + 8) + 4 * (size_t)N)
             + sizeof(double) * (2 * (size_t)N + 4 + 128 + 64 * 3);
    }
    __device__ void carve(void *base, int N, int L)
    {
        const size_t r1 = PackedR::entries(N, N + L - 1), r2 = PackedR::entries(N, N);
        const size_t scr = (size_t)L * N + (size_t)K3C_TK * N;
        const bool overlay = L > 1;
        double2 *p = (double2 *)base;
        R1.a = p; R1.W = N + L - 1;
        p += overlay ? (r1 > scr ? r1 : scr) : r1;
        R2.a = p; R2.W = N; p += r2;
        if (overlay) scratch = R1.a; else { scratch = p; p += scr; }
        vbuf = p; p += 2 * (K3C_MAXROWS + 8);
        om = p; p += N; qq = p; p += N; qw = p; p += N; Cv = p; p += N;
        double *d = (double *)p;
        diag1 = d; d += N;
        diag2 = d; d += N;
        scal = d; d += 4;
        red = d; d += 128;
        ends = d;
    }
};

template <int RPT>
__device__ __forceinline__ double k3c_norm2(const double2 (&X)[RPT])
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int r = 0; r < RPT; r += 4) {
        a0 = fma(X[r].x, X[r].x, a0); a1 = fma(X[r + 1].x, X[r + 1].x, a1);
        a2 = fma(X[r + 2].x, X[r + 2].x, a2); a3 = fma(X[r + 3].x, X[r + 3].x, a3);
        a0 = fma(X[r].y, X[r].y, a0); a1 = fma(X[r + 1].y, X[r + 1].y, a1);
        a2 = fma(X[r + 2].y, X[r + 2].y, a2); a3 = fma(X[r + 3].y, X[r + 3].y, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

template <int G>
__device__ __forceinline__ double k3c_group_sum(double v)
{
#pragma unroll
    for (int s = 1; s < G; s <<= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// Reflections jstart..N-1 of [R; tile]; X = the thread's RPT rows (row group g) of column
// c; part = the thread's share of |column|^2, kept current for every live column.
template <int G, int RPT>
__device__ __forceinline__ void k3c_reflect(double2 (&X)[RPT], double &part, const int c, const int g,
                                            const int ncols, const int N, const int jstart, const PackedR &R,
                                            double *diag, double2 *vbuf, double *scal, int &buf)
{
    constexpr int CPW = 32 / G;                              // columns per warp
    constexpr int VST = RPT + 1;                             // padded row-group stride of vbuf
    constexpr int VBUF = K3C_MAXROWS + 8;
    const int wfirst = (c / CPW) * CPW, wlast = wfirst + CPW - 1;
    double tot = k3c_group_sum<G>(part);                     // |column|^2 over the G lanes
#pragma unroll 1
    for (int j = jstart; j < N; ++j) {
        if (c == j) {                    // owner lanes: scalars and broadcast of the reflector
            const double r = diag[j];
            const double t = fma(r, r, tot) + 1e-300;        // see fit_small.cuh: an all-zero column needs no branch
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
            double2 *vw = vbuf + buf * VBUF + g * VST;
#pragma unroll
            for (int r2 = 0; r2 < RPT; ++r2) vw[r2] = X[r2];
        }
        __syncthreads();
        if (wlast > j && wfirst < ncols) {                   // warp-uniform: trailing columns left
            const bool act = c > j && c < ncols;
            const double2 *v = vbuf + buf * VBUF + g * VST;
            double sr0 = 0.0, si0 = 0.0, sr1 = 0.0, si1 = 0.0, sr2 = 0.0, si2 = 0.0, sr3 = 0.0, si3 = 0.0;
#pragma unroll
            for (int r = 0; r < RPT; r += 4) {
                const double2 b0 = v[r], b1 = v[r + 1], b2 = v[r + 2], b3 = v[r + 3];
                sr0 = fma(b0.x, X[r].x, sr0); si0 = fma(b0.x, X[r].y, si0);
                sr1 = fma(b1.x, X[r + 1].x, sr1); si1 = fma(b1.x, X[r + 1].y, si1);
                sr2 = fma(b2.x, X[r + 2].x, sr2); si2 = fma(b2.x, X[r + 2].y, si2);
                sr3 = fma(b3.x, X[r + 3].x, sr3); si3 = fma(b3.x, X[r + 3].y, si3);
                sr0 = fma(b0.y, X[r].y, sr0); si0 = fma(-b0.y, X[r].x, si0);
                sr1 = fma(b1.y, X[r + 1].y, sr1); si1 = fma(-b1.y, X[r + 1].x, si1);
                sr2 = fma(b2.y, X[r + 2].y, sr2); si2 = fma(-b2.y, X[r + 2].x, si2);
                sr3 = fma(b3.y, X[r + 3].y, sr3); si3 = fma(-b3.y, X[r + 3].x, si3);
            }
            // R_jc is read by all G lanes of the column and rewritten by lane 0 of the group: the
            // load sits BEFORE the group sums, whose shuffles order it ahead of that store (a
            // converged warp executes the loads first anyway; the lock-step host emulation, which
            // runs a thread alone from one collective to the next, does not)
            const double2 old = act ? R.at(j, c) : make_double2(0.0, 0.0);
            const double sr = k3c_group_sum<G>((sr0 + sr1) + (sr2 + sr3));
            const double si = k3c_group_sum<G>((si0 + si1) + (si2 + si3));
            if (act) {
                const double v0 = scal[buf * 2], beta = scal[buf * 2 + 1];
                double2 &Rjc = R.at(j, c);
                const double pr = fma(v0, old.x, sr) * beta;
                const double pi = fma(v0, old.y, si) * beta;
                if (g == 0) Rjc = make_double2(fma(-v0, pr, old.x), fma(-v0, pi, old.y));
                double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; r += 2) {
                    const double2 b0 = v[r], b1 = v[r + 1];
                    double x0 = X[r].x, y0 = X[r].y, x1 = X[r + 1].x, y1 = X[r + 1].y;
                    x0 = fma(-pr, b0.x, x0); y0 = fma(-pr, b0.y, y0);
                    x1 = fma(-pr, b1.x, x1); y1 = fma(-pr, b1.y, y1);
                    x0 = fma(pi, b0.y, x0); y0 = fma(-pi, b0.x, y0);
                    x1 = fma(pi, b1.y, x1); y1 = fma(-pi, b1.x, y1);
                    X[r] = make_double2(x0, y0);
                    X[r + 1] = make_double2(x1, y1);
                    n0 = fma(x0, x0, n0); n1 = fma(y0, y0, n1);
                    n2 = fma(x1, x1, n2); n3 = fma(y1, y1, n3);
                }
                part = (n0 + n1) + (n2 + n3);
            }
            if (G > 1) {
                if (wfirst <= j + 1 && j + 1 <= wlast) tot = k3c_group_sum<G>(part);   // the next pivot's warp
            } else {
                tot = part;
            }
        }
        buf ^= 1;
    }
}

// Rows first..first+RPT-1 of column c of the phase-1 matrix [E | d_1..d_L].
template <int RPT>
__device__ __forceinline__ void k3c_load1(double2 (&X)[RPT], const FitParams &p, const Struct3Smem &sm,
                                          const int c, const int N, const int NC, const int first, const int re,
                                          const double t0, double &sdd)
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
                for (int r = 0; r < RPT; ++r) {
                    X[r] = first + r < re ? z : zero;
                    const int kn = first + r + 1 < re ? first + r + 1 : re - 1;
                    const double tau_n = qf_sub_rn(p.times[kn], t0);
                    const double de = qf_sub_rn(qf_sub_rn(tau_n, tau), dt);
                    tau = tau_n;
                    z = c_mul(z, make_double2(fma(w.x, de, q.x), fma(w.y, de, q.y)));
                }
            } else {
#pragma unroll
                for (int r = 0; r < RPT; ++r) X[r] = zero;
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < RPT; ++r) {
                double2 e = zero;
                if (first + r < re) e = design_entry(row_omega(p, sm.om, c, first + r), qf_sub_rn(p.times[first + r], t0));
#pragma unroll
                for (int q = 0; q < RPT; ++q) if (q == r) X[q] = e;   // static register index
            }
        }
    } else if (c < NC) {
        const double2 *dsrc = p.data + (long long)(c - N) * p.series_stride;
#pragma unroll
        for (int r = 0; r < RPT; ++r) X[r] = first + r < re ? dsrc[first + r] : zero;
        sdd += k3c_norm2<RPT>(X);
    } else {
#pragma unroll
        for (int r = 0; r < RPT; ++r) X[r] = zero;
    }
}

// Rows q0..q0+RPT-1 of column c of the phase-2 matrix: row q = rr * L + i is row rr of
// [R_E D_i | Y_i] (rr-major, so that a tile's rows share their leading zeros).
template <int RPT>
__device__ __forceinline__ void k3c_load2(double2 (&X)[RPT], const Struct3Smem &sm, const double2 *coef,
                                          const int c, const int N, const int L, const int q0, const int rows2)
{
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int q = q0 + r;
        double2 v = make_double2(0.0, 0.0);
        if (q < rows2 && c <= N) {
            const int rr = q / L, i = q - rr * L;
            if (c == N) v = sm.R1.at(rr, N + i);
            else if (c >= rr) {
                const double2 cf = coef ? coef[i * N + c] : make_double2(1.0, 0.0);
                v = c == rr ? make_double2(cf.x * sm.diag1[rr], cf.y * sm.diag1[rr]) : c_mul(cf, sm.R1.at(rr, c));
            }
        }
        X[r] = v;
    }
}

// Everything after the factorisation, shared by K3 and K4 (fit_panel.cuh): rank flag,
// export of the factor, back-substitution, then either the mismatch from the by-products of
// the factorisation (uniform grids) or the second streaming pass (model, trapezoid sums).
// SM provides R1, R2 (PackedR), diag1, diag2, Cv, om, scratch, red, ends as Struct3Smem does.
// sdd = this thread's share of sum |d|^2, res2 = its share of the annihilated rhs entries.
template <class SM>
__device__ __forceinline__ void struct_finish(const FitParams &p, const SM &sm, const int fit, const int N, const int L,
                                              const int rb, const int re, const double t0, const double2 *coef,
                                              const bool two_phase, double sdd, double res2)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = nthr >> 5;
    const int K = re - rb;
    const long long Mrows = (long long)K * L;
    int status = 0;
    if (!p.eval_only) {
        const PackedR &Rf = two_phase ? sm.R2 : sm.R1;
        double *dgf = two_phase ? sm.diag2 : sm.diag1;
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
            if (!(dmin > rank_prefilter(N) * QNMFIT_EPS * dim * dmax)) {   // rare: confirm (qnmfit_common.cuh)
                if (rank_suspect_warp([&](int j, int k) { return Rf.at(j, k); }, [&](int j) { return dgf[j]; }, N, dim,
                                      sm.Cv, lane))
                    status |= QNMFIT_ST_RANK_DEFICIENT_;
            }
            if (Mrows <= N) status |= QNMFIT_ST_UNDERDETERMINED_;
            if (p.R) {
                double2 *Rout = p.R + (long long)fit * N * (N + 1);
                for (int e = lane; e < N * (N + 1); e += 32) {
                    const int j = e / (N + 1), k = e - j * (N + 1);
                    double2 v = make_double2(0.0, 0.0);
                    if (k == j) v = make_double2(dgf[j], 0.0);
                    else if (k > j) v = Rf.at(j, k);
                    Rout[e] = v;
                }
            }
            for (int j = N - 1; j >= 0; --j) {
                double ax = 0.0, ay = 0.0;
                for (int k = j + 1 + lane; k < N; k += 32) {
                    const double2 Rjk = Rf.at(j, k);
                    const double2 cv = sm.Cv[k];
                    ax = fma(Rjk.x, cv.x, ax);
                    ax = fma(-Rjk.y, cv.y, ax);
                    ay = fma(Rjk.x, cv.y, ay);
                    ay = fma(Rjk.y, cv.x, ay);
                }
                ax = warp_sum(ax);
                ay = warp_sum(ay);
                if (lane == 0) {
                    const double2 b = Rf.at(j, N);
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
        for (int w = 0; w < nwarps; ++w) { sdd += sm.red[w * 8]; res2 += sm.red[w * 8 + 1]; }
        double cn2 = 0.0;                 // ||Q^H d||^2, every thread in the same fixed order
        for (int j = 0; j < N; ++j) {
            const double2 b = Rf.at(j, N);
            cn2 = fma(b.x, b.x, cn2);
            cn2 = fma(b.y, b.y, cn2);
        }
        __syncthreads();                  // R1 may be overlaid by the scratch area from here on

        if (p.fast_mismatch && K > 0) {
            // mismatch from the by-products (see small_fast_finalize in fit_small.cuh)
            double2 *E = sm.scratch;
            for (int e = tid; e < 2 * N; e += nthr) {
                const int j = e % N;
                const int row = e < N ? rb : re - 1;
                E[e] = design_entry(row_omega(p, sm.om, j, row), qf_sub_rn(p.times[row], t0));
            }
            __syncthreads();
            for (int i = tid; i < L; i += nthr) {
                double e0 = 0.0, e1 = 0.0, e2 = 0.0;  // end-point terms of the three sums
#pragma unroll 1
                for (int end = 0; end < 2; ++end) {
                    const double2 *Er = E + end * N;
                    double mx = 0.0, my = 0.0;
                    for (int j = 0; j < N; ++j) {
                        const double2 a = Er[j];
                        const double2 cj = coef ? c_mul(coef[i * N + j], sm.Cv[j]) : sm.Cv[j];
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
                sm.ends[i * 3] = e0; sm.ends[i * 3 + 1] = e1; sm.ends[i * 3 + 2] = e2;
            }
            __syncthreads();
            if (tid == 0) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int i = 0; i < L; ++i) { a0 += sm.ends[i * 3]; a1 += sm.ends[i * 3 + 1]; a2 += sm.ends[i * 3 + 2]; }
                const double num = cn2 - 0.5 * a0, n1 = cn2 - 0.5 * a1, n2 = sdd - 0.5 * a2;
                const double mm = 1.0 - num / sqrt(n1 * n2);
                p.mismatch[fit] = mm;
                if (p.residual) p.residual[fit] = res2;
                if (p.status) p.status[fit] = status;
                note_status(p, fit, status);
                peer_publish(p, fit, mm);
            }
            return;
        }
    } else {
        for (int j = tid; j < N; j += nthr) sm.Cv[j] = p.C[(long long)fit * N + j];
    }
    __syncthreads();

    // ---------------- second pass: model and trapezoid-weighted sums ----------------
    double2 *cc = sm.scratch, *E = sm.scratch + L * N;
    for (int e = tid; e < L * N; e += nthr) cc[e] = coef ? c_mul(coef[e], sm.Cv[e % N]) : sm.Cv[e % N];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int nchunks = (K + K3C_TK - 1) / K3C_TK;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int k0 = rb + ch * K3C_TK;
        const int kn = min(K3C_TK, re - k0);
        const int rows = kn * L;
        __syncthreads();
        for (int e = tid; e < kn * N; e += nthr) {
            const int kk = e / N, j = e - kk * N;
            E[kk * N + j] = design_entry(row_omega(p, sm.om, j, k0 + kk), qf_sub_rn(p.times[k0 + kk], t0));
        }
        __syncthreads();
        for (int r = tid; r < rows; r += nthr) {
            const int i = r / kn, kk = r - i * kn;
            const double2 *Er = E + kk * N;
            const double2 *ci = cc + i * N;
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
        for (int w = 0; w < nwarps; ++w) {
            a0 += sm.red[w * 8 + 0]; a1 += sm.red[w * 8 + 1];
            a2 += sm.red[w * 8 + 2]; a3 += sm.red[w * 8 + 3];
        }
        const double mm = 1.0 - a0 / sqrt(a1 * a2);
        p.mismatch[fit] = mm;
        if (p.residual) p.residual[fit] = p.eval_only ? a3 : res2;
        if (p.status) p.status[fit] = status;
        note_status(p, fit, status);
        peer_publish(p, fit, mm);
    }
}

template <int G, int RPT>
#define K3C_MINB(G, RPT) (RPT >= 32 ? (G == 1 ? 4 : 2) : G <= 2 ? 4 : G == 4 ? 2 : 1)
__global__ void __launch_bounds__(64 * G, K3C_MINB(G, RPT)) fit_struct3_kernel(const __grid_constant__ FitParams p)
{
    QF_DYN_SMEM(smem_raw);
    constexpr int TR = G * RPT;
    static_assert(TR <= K3C_MAXROWS, "tile taller than the v buffer");
    const int N = p.n_modes, L = p.n_series, NC = N + L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = nthr >> 5;
    const int c = tid / G, g = tid % G;
    const int fit = blockIdx.x;
    Struct3Smem sm;
    sm.carve(smem_raw, N, L);

    const int fi = input_fit(p, fit);             // index of the fit's inputs (see input_fit)
    int rb = p.row_begin ? p.row_begin[fi] : p.row_begin_all;
    int re = p.row_end ? p.row_end[fi] : p.row_end_all;
    const double t0 = p.t0 ? p.t0[fi] : p.t0_all;
    if (rb < 0) rb = 0;
    if (re > p.n_times) re = p.n_times;
    if (re < rb) re = rb;
    const int K = re - rb;
    const double2 *coef = nullptr;
    if (p.coef) {
        const int ci = p.coef_index ? p.coef_index[fi] : fit_chi_index(p, fi);
        coef = p.coef + (long long)ci * L * N;
    }
    const bool two_phase = (coef != nullptr) || L > 1;

    for (int j = tid; j < N; j += nthr) {
        const double2 w = fit_omega(p, fi, j);
        sm.om[j] = w;
        if (p.dt_nominal > 0.0) {
            const double2 q = design_entry(w, p.dt_nominal);
            sm.qq[j] = q;
            sm.qw[j] = c_mul(q, make_double2(w.y, -w.x));
        }
        sm.diag1[j] = 0.0;
        sm.diag2[j] = 0.0;
    }
    {
        const int n1 = (int)PackedR::entries(N, NC - 1), n2 = (int)PackedR::entries(N, N);
        for (int e = tid; e < n1; e += nthr) sm.R1.a[e] = make_double2(0.0, 0.0);
        for (int e = tid; e < n2; e += nthr) sm.R2.a[e] = make_double2(0.0, 0.0);
    }
    __syncthreads();

    double sdd = 0.0, res2 = 0.0;
    int buf = 0;
    double2 X[RPT];

    if (!p.eval_only) {
        // ---------------- phase 1: [E | d_1..d_L] ----------------
        const int ntiles = (K + TR - 1) / TR;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            k3c_load1<RPT>(X, p, sm, c, N, NC, rb + tile * TR + g * RPT, re, t0, sdd);
            double part = k3c_norm2<RPT>(X);
            k3c_reflect<G, RPT>(X, part, c, g, NC, N, 0, sm.R1, sm.diag1, sm.vbuf, sm.scal, buf);
            if (c >= N && c < NC) res2 += part;
        }
        __syncthreads();

        // ---------------- phase 2: stacked [R_E D_i | Y_i] ----------------
        if (two_phase) {
            const int rows2 = N * L;
            const int ntiles2 = (rows2 + TR - 1) / TR;
#pragma unroll 1
            for (int tile = 0; tile < ntiles2; ++tile) {
                const int jstart = (tile * TR) / L;             // first non-zero column of the tile
                k3c_load2<RPT>(X, sm, coef, c, N, L, tile * TR + g * RPT, rows2);
                double part = k3c_norm2<RPT>(X);
                k3c_reflect<G, RPT>(X, part, c, g, N + 1, N, jstart, sm.R2, sm.diag2, sm.vbuf, sm.scal, buf);
                if (c == N) res2 += part;
            }
            __syncthreads();
        }
    }
    struct_finish(p, sm, fit, N, L, rb, re, t0, coef, two_phase, sdd, res2);
}
