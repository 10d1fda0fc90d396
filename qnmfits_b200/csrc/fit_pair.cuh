// fit_pair.cuh — K1p: single-series fits with 9 .. 24 columns, columns split over lanes.
//
// Same algorithm and the same reference lines as K1 (fit_small.cuh: sequential TSQR of the
// lane's rows in register blocks, R-combine tree, back-substitution, mismatch from the
// factorisation's by-products; reference qnmfits/qnmfits.py:274-293).  What changes is who
// holds what.  K1 gives every lane ALL N + 1 columns of its rows, and beyond eight columns
// that no longer fits: the register block shrinks to three and two rows, the per-lane factor
// (N (N+1) / 2 complex in shared memory) cuts the CTA to 160 lanes, and the N = 12 instance
// spills.  Here a row slice is owned by a GROUP of CS adjacent lanes (CS = 2, 4 or 8) and lane h
// of the group holds only the columns k with (N - k) mod CS == h (k = N: the right-hand side),
// in "slots" s = (N - k) / CS counted from the right-hand side backwards:
//
//   * registers per lane: MB x ceil((N+1)/CS) complex for the block — the footprint of K1 at
//     N = 8 or less — so the blocks stay MB = 4 (or 8) rows tall and nothing spills;
//   * shared memory per lane: its columns of the factor only (about 1/CS of the triangle), so
//     the CTA keeps 224 - 256 lanes;
//   * a reflection needs column j everywhere: its owner broadcasts the MB block entries and
//     with width-CS shuffles (2 MB doubles; R_jj is read from the owner's shared memory); the norm and the reflector scalars are
//     then computed redundantly by the CS lanes (no lane idles), and every lane updates its own
//     trailing columns.  Counting slots from the right makes the lanes' trailing-column
//     counts differ by at most one for every j (ceil vs floor of (N - j) / CS);
//   * control flow is warp-uniform (all shuffles use the full mask): every lane of a warp
//     runs the warp's longest block count and the tree levels of all groups; what a lane
//     has no rows / no partner for is computed on zero rows and its stores are predicated
//     off, so a fit's bits do not depend on its neighbours in the warp.
//
// Lane 0 of the fit back-substitutes, reading the other lanes' columns of the factor
// straight from shared memory.
#pragma once
#include "qnmfit_common.cuh"
#include "fit_small.cuh"     // SmallLane, SmallAcc, small_fast_finalize, small_finalize, SMALL_STAGE_PAD
#ifdef QNMFIT_HOSTSIM
#include "hostsim_warp.h"    // tests/hostsim: the warp collectives, emulated with one fiber per lane
#endif

// Once-per-fit stages are real calls: their register needs must not perturb the allocation
// of the block loop (with them inlined, the N = 13, 14 instances spilled ~100 B inside it).
#ifndef QNMFIT_HOSTSIM
#define PAIR_COLD static __device__ __noinline__
#else
#define PAIR_COLD static inline
#endif

template <int N, int CS>
struct PairLayout {
    static_assert(CS == 2 || CS == 4 || CS == 8, "2, 4 or 8 lanes per row slice");
    static constexpr int S = (N + 1 + CS - 1) / CS;            // slots (columns) per lane
    // frequency tables: one row of N entries per fit behind CS - 1 leading entries (and one
    // trailing), so that slot s of lane h — column N - CS s - h, which may be the data column N
    // or lie left of column 0 — reads entry tab(N - h) - CS s of its fit's row without a clamp.
    // What it finds there for those two cases (a neighbouring fit's entry, or padding) is
    // never used: the data slot is overwritten with the data, and a slot left of column 0 is
    // neither a reflector nor a trailing column of any reflection.
    static constexpr int NT = N;
    QF_MEMBOTH static constexpr int tab(int k) { return CS - 1 + k; }
    // rows of the factor stored for slot s (lane 0's column is the longest of the slot; the
    // right-hand side, slot 0 of lane 0, has N rows)
    QF_MEMBOTH static constexpr int rows(int s) { return s == 0 ? N : N + 1 - CS * s; }
    QF_MEMBOTH static constexpr int base(int s) { return s == 0 ? 0 : N + (s - 1) * (N + 1) - CS * ((s - 1) * s / 2); }
    static constexpr int E = base(S - 1) + rows(S - 1);        // entries per lane
    QF_MEMBOTH static constexpr int lane_of(int k) { return (N - k) % CS; }
    QF_MEMBOTH static constexpr int slot_of(int k) { return (N - k) / CS; }
    // R[j][k], j <= k <= N (k == j: the real diagonal in .x), in lane lane_of(k) of the group
    QF_MEMBOTH static constexpr int entry(int j, int k) { return base(slot_of(k)) + j; }
};

template <int N, int CS, int THREADS>
struct PairSmem {
    double2 *R;         // [E][THREADS]    the lanes' columns of their factors
    double2 *om;        // [fpc][NT]       frequencies of the CTA's fits (PairLayout::tab)
    double2 *qq;        // [fpc][NT]       exp(-i w dt)
    double2 *qw;        // [fpc][NT]       exp(-i w dt) * (-i w)
    const double2 *ds;  // [stage_rows]    staged data window (or global data)
    const double *ts;   // [stage_rows]    staged times (or global times)
    int fpc, t_off;

    QF_MEMBOTH static size_t bytes(int fpc, int stage_rows)
    {
        if (stage_rows > 0) stage_rows += SMALL_STAGE_PAD;
        return sizeof(double2) * (size_t)(PairLayout<N, CS>::E * THREADS + 3 * (PairLayout<N, CS>::NT * fpc + CS) + stage_rows)
             + sizeof(double) * (size_t)stage_rows;
    }
    QF_MEM void carve(void *basep, int fpc_, int stage_rows)
    {
        if (stage_rows > 0) stage_rows += SMALL_STAGE_PAD;
        fpc = fpc_;
        double2 *p2 = (double2 *)basep;
        R = p2; p2 += PairLayout<N, CS>::E * THREADS;
        om = p2; p2 += PairLayout<N, CS>::NT * fpc + CS;
        qq = p2; p2 += PairLayout<N, CS>::NT * fpc + CS;
        qw = p2; p2 += PairLayout<N, CS>::NT * fpc + CS;
        ds = p2; p2 += stage_rows;
        ts = (const double *)p2;
    }
};

// A lane's share of one fit: the rows of its group (lf / CS of lanes_per_fit / CS groups).
// Lanes without a fit and empty windows get an empty share at `safe_row` (a row that may be read).
template <int CS>
QF_HD SmallLane pair_lane_setup(const FitParams &p, int cta, int tid, int threads, bool per_fit_data, int mb,
                                int safe_row)
{
    SmallLane L;
    const int lpf = p.lanes_per_fit;
    const int groups = lpf / CS;
    const int fpc = threads / lpf;
    L.slot = tid / lpf;
    L.lf = tid % lpf;
    const int fit = cta * fpc + L.slot;
    L.fit = fit < p.n_fits ? fit : -1;
    L.rb = L.re = L.lo = L.hi = safe_row;
    L.nblk = 0;
    L.t0 = 0.0;
    L.d_off = 0;
    if (L.fit >= 0) {
        const int fi = input_fit(p, fit);
        if (per_fit_data && p.series_index) L.d_off = (long long)p.series_index[fi] * p.series_stride;
        L.rb = p.row_begin ? p.row_begin[fi] : p.row_begin_all;
        L.re = p.row_end ? p.row_end[fi] : p.row_end_all;
        L.t0 = p.t0 ? p.t0[fi] : p.t0_all;
        if (L.rb < 0) L.rb = 0;
        if (L.re > p.n_times) L.re = p.n_times;
        if (L.re <= L.rb) { L.rb = L.re = L.lo = L.hi = safe_row; return L; }
        const int M = L.re - L.rb;
        int rpl = (M + groups - 1) / groups;
        rpl = (rpl + mb - 1) / mb * mb;
        L.nblk = rpl / mb;
        L.lo = L.rb + (L.lf / CS) * rpl;
        if (L.lo > L.re) L.lo = L.re;
        L.hi = L.lo + rpl;
        if (L.hi > L.re) L.hi = L.re;
    }
    return L;
}

// Fold the MB x (columns of this lane) block into the group's factor: reflections
// JSTART .. N-1 of [R; B].  Rt = the lane's entries (sm.R + tid); `live`: the group has rows
// in this block (otherwise B is zero and nothing is stored).
template <int N, int CS, int MB, int THREADS, int JSTART>
QF_HD void pair_absorb(double2 (&B)[MB][PairLayout<N, CS>::S], double2 *Rt, const int h, const bool live)
{
    typedef PairLayout<N, CS> LY;
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int j = JSTART; j < N; ++j) {
        const int d = N - j, own = d % CS, sj = d / CS;       // column j: lane `own`, slot sj
        const int ST = (d + CS - 1) / CS;                     // slots that can hold a trailing column (k > j)
        // ---- R_jj and column j of the block, from their owner.  R_jj is read in place, from the
        // owner's shared memory, BEFORE the shuffles: the owner overwrites it further down, and
        // only the shuffles order that store behind the other lanes' loads (the lock-step host
        // emulation, which runs a lane from one collective to the next, caught the load sitting
        // behind them).
        const double r = Rt[(LY::base(sj) + j) * THREADS + (own - h)].x;
        double2 v[MB];
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            v[i].x = __shfl_sync(full, B[i][sj].x, own, CS);
            v[i].y = __shfl_sync(full, B[i][sj].y, own, CS);
        }
        // ---- (A) column norm and the raw dot products v^H B_s of the trailing slots
        double sr[LY::S], si[LY::S];
        double sig0 = 1e-300, sig1 = 0.0;     // seed: an exactly zero column needs no branch (fit_small.cuh)
#pragma unroll
        for (int s = 0; s < ST; ++s) { sr[s] = v[0].x * B[0][s].x; si[s] = v[0].x * B[0][s].y; }
#pragma unroll
        for (int s = 0; s < ST; ++s) { sr[s] = fma(v[0].y, B[0][s].y, sr[s]); si[s] = fma(-v[0].y, B[0][s].x, si[s]); }
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            if (i & 1) { sig1 = fma(v[i].x, v[i].x, sig1); sig1 = fma(v[i].y, v[i].y, sig1); }
            else { sig0 = fma(v[i].x, v[i].x, sig0); sig0 = fma(v[i].y, v[i].y, sig0); }
        }
        const double t = fma(r, r, sig0 + sig1);
#pragma unroll
        for (int i = 1; i < MB; ++i) {
#pragma unroll
            for (int s = 0; s < ST; ++s) { sr[s] = fma(v[i].x, B[i][s].x, sr[s]); si[s] = fma(v[i].x, B[i][s].y, si[s]); }
#pragma unroll
            for (int s = 0; s < ST; ++s) { sr[s] = fma(v[i].y, B[i][s].y, sr[s]); si[s] = fma(-v[i].y, B[i][s].x, si[s]); }
        }
        // ---- (B) reflector scalars (redundantly in the CS lanes)
        const double y = qf_rsqrt(t);
        const double nrm = t * y;
        const double ar = fabs(r);
        const double v0 = copysign(ar + nrm, r);          // v = [v0; b]
        const double beta = qf_rcp(nrm * (ar + nrm));     // 2 / v^H v
        if (live && h == own) Rt[(LY::base(sj) + j) * THREADS] = make_double2(-copysign(nrm, r), 0.0);
        // ---- (C) row j of R and the rank-1 update, own trailing columns
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            // slot s of lane h holds column k = N - CS s - h; trailing iff k > j
            const bool act = (CS * s + CS - 1 < d) || (CS * s + h < d);
            double2 Rjk = make_double2(0.0, 0.0);
            if (act) Rjk = Rt[(LY::base(s) + j) * THREADS];
            const double pr = fma(v0, Rjk.x, sr[s]) * beta;
            const double pi = fma(v0, Rjk.y, si[s]) * beta;
            Rjk.x = fma(-v0, pr, Rjk.x);
            Rjk.y = fma(-v0, pi, Rjk.y);
            if (live && act) Rt[(LY::base(s) + j) * THREADS] = Rjk;
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                double bx = B[i][s].x, by = B[i][s].y;
                bx = fma(-pr, v[i].x, bx);
                by = fma(-pr, v[i].y, by);
                bx = fma(pi, v[i].y, bx);
                by = fma(-pi, v[i].x, by);
                B[i][s].x = bx;
                B[i][s].y = by;
            }
        }
    }
}

// |right-hand-side entries|^2 of the block (slot 0; meaningful in lane 0 of the group)
template <int MB, int S>
QF_HD void pair_acc_rhs(const double2 (&B)[MB][S], double &a0, double &a1)
{
#pragma unroll
    for (int i = 0; i < MB; ++i) {
        if (i & 1) { a1 = fma(B[i][0].x, B[i][0].x, a1); a1 = fma(B[i][0].y, B[i][0].y, a1); }
        else { a0 = fma(B[i][0].x, B[i][0].x, a0); a0 = fma(B[i][0].y, B[i][0].y, a0); }
    }
}

// Leaf stage: sequential TSQR over the group's rows.  Rows are generated by the recurrence
// z <- z (q + q(-i w) de), de = the step's deviation from the nominal one, re-anchored with a
// direct exp/sincos every anchor_rows rows (fit_small.cuh); dt_nominal == 0: every row direct.
template <int N, int CS, int MB, int THREADS>
QF_HD void pair_leaf(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, const SmallLane &L, const int h,
                     int tid, SmallAcc &acc)
{
    typedef PairLayout<N, CS> LY;
    constexpr int S = LY::S;
    const double *ts = sm.ts - sm.t_off;
    const double2 *ds = sm.ds - sm.t_off + L.d_off;
    const int fpc = sm.fpc;
    const double dt = p.dt_nominal, t0 = L.t0;
    const bool direct = !(dt > 0.0);
    int ablk = (p.anchor_rows > 0 ? p.anchor_rows : QNMFIT_DEFAULT_ANCHOR_ROWS) / MB;
    if (ablk < 1) ablk = 1;
    // this lane's columns: slot s holds column N - CS s - h (slot 0 of lane 0: the data); see
    // PairLayout::NT for what the table holds for the data slot and left of column 0
    const double2 *om = sm.om + L.slot * LY::NT + LY::tab(N - h);
    const double2 *qq = sm.qq + L.slot * LY::NT + LY::tab(N - h);
    const double2 *qw = sm.qw + L.slot * LY::NT + LY::tab(N - h);
    const unsigned full = 0xffffffffu;
    const int nblk_w = __reduce_max_sync(full, L.nblk);
    const int last = L.re > L.rb ? L.re - 1 : L.rb;
    double2 z[S];
    double2 B[MB][S];
#pragma unroll
    for (int s = 0; s < S; ++s) z[s] = make_double2(0.0, 0.0);
    double tau = 0.0;
    double sdd1 = 0.0, res1 = 0.0;
    const double2 zero = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int blk = 0; blk < nblk_w; ++blk) {
        const int row0 = L.lo + blk * MB;
        const bool live = row0 < L.hi;
        const int ra = live ? row0 : L.rb;
        if (direct || blk % ablk == 0) {
            tau = qf_sub_rn(ts[ra], t0);
#pragma unroll
            for (int s = 0; s < S; ++s) z[s] = design_entry(om[-CS * s], tau);
        }
        if (!direct && __all_sync(full, row0 + MB <= L.hi)) {
            // every lane of the warp has a full block: no per-row selects
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                const int r = row0 + i;
                const int rn = r + 1 < last ? r + 1 : last;
                const double2 dval = ds[r];
                const double tau_n = qf_sub_rn(ts[rn], t0);
                const double de = qf_sub_rn(qf_sub_rn(tau_n, tau), dt);
                tau = tau_n;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    B[i][s] = z[s];
                    const double2 q = qq[-CS * s], w = qw[-CS * s];
                    z[s] = c_mul(z[s], make_double2(fma(w.x, de, q.x), fma(w.y, de, q.y)));
                }
                if (h == 0) B[i][0] = dval;
            }
        } else {
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                const int r = ra + i;
                const bool valid = live && r < L.hi;
                const int rc = r < last ? r : last;
                const int rn = r + 1 < last ? r + 1 : last;
                const double2 dval = ds[rc];
#pragma unroll
                for (int s = 0; s < S; ++s) B[i][s] = valid ? z[s] : zero;
                if (h == 0) B[i][0] = valid ? dval : zero;
                const double tau_n = qf_sub_rn(ts[rn], t0);
                if (direct) {
                    if (i < MB - 1) {
#pragma unroll
                        for (int s = 0; s < S; ++s) z[s] = design_entry(om[-CS * s], tau_n);
                    }
                } else {
                    const double de = qf_sub_rn(qf_sub_rn(tau_n, tau), dt);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const double2 q = qq[-CS * s], w = qw[-CS * s];
                        z[s] = c_mul(z[s], make_double2(fma(w.x, de, q.x), fma(w.y, de, q.y)));
                    }
                }
                tau = tau_n;
            }
        }
        pair_acc_rhs<MB, S>(B, acc.sdd, sdd1);
        pair_absorb<N, CS, MB, THREADS, 0>(B, sm.R + tid, h, live);
        pair_acc_rhs<MB, S>(B, acc.res2, res1);
    }
    acc.sdd += sdd1;
    acc.res2 += res1;
}

// Absorb rows B0 .. B0+MB-1 of the partner group's triangle (thread pt = the partner lane
// with the same h).  Row r of a triangle is zero left of column r: reflections start at B0.
template <int N, int CS, int MB, int THREADS, int B0>
QF_HD void pair_tree_block(const PairSmem<N, CS, THREADS> &sm, int tid, int pt, const int h, const bool live,
                           SmallAcc &acc)
{
    typedef PairLayout<N, CS> LY;
    if constexpr (B0 < N) {
        // The previous block ended with reflection N - 1, whose owner stored R_jj behind that
        // reflection's shuffles; when this block starts at the same column (B0 = N - 1) the other
        // lanes read that entry with no collective in between.  A converged warp executes the
        // store first, but only a barrier makes that order a guarantee (found by the lock-step
        // host emulation, which runs each lane from one collective to the next).
        __syncwarp();
        double2 B[MB][LY::S];
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            const int row = B0 + i;
#pragma unroll
            for (int s = 0; s < LY::S; ++s) {
                const int k = N - CS * s - h;
                const int nrows = k == N ? N : k + 1;       // rows stored for this column (<= 0: no column)
                double2 val = make_double2(0.0, 0.0);
                if (live && row < N && row < nrows) val = sm.R[(LY::base(s) + row) * THREADS + pt];
                B[i][s] = val;
            }
        }
        pair_absorb<N, CS, MB, THREADS, B0>(B, sm.R + tid, h, live);
        double dummy = 0.0;
        pair_acc_rhs<MB, LY::S>(B, acc.res2, dummy);
        acc.res2 += dummy;
        pair_tree_block<N, CS, MB, THREADS, B0 + MB>(sm, tid, pt, h, live, acc);
    }
}

// "numpy MAY truncate a singular value here" (qnmfit_common.cuh), decided by the lanes of the fit
// together: the estimate of s_min by inverse iteration on R^H R, as rank_suspect_warp, with the
// iteration vector spread over the fit's lanes_per_fit lanes (lane lf holds entries lf,
// lf + lpf, ...) and both triangular solves column by column — the owner of an entry finishes
// it, a width-lpf shuffle broadcasts it, every lane updates its own entries.  Called by ALL
// lanes of the warp (uniform control flow); the result is uniform within a fit.  The serial
// form (lane 0 alone, rank_suspect_serial) cost 24 % of the N = 16 kernel once the prefilter
// sent every fit of that size here.
template <int N, int CS, int THREADS>
PAIR_COLD bool pair_rank_suspect(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, const SmallLane &L, int tid)
{
    typedef PairLayout<N, CS> LY;
    constexpr int ME = (N + CS - 1) / CS;              // entries per lane at the smallest lanes_per_fit (= CS)
    const unsigned full = 0xffffffffu;
    const int lpf = p.lanes_per_fit;
    const double2 *R0 = sm.R + (tid - L.lf);           // lane 0 of the fit: its group holds the final factor
    auto Rat = [&](int j, int k) {                     // R[j][k], j <= k < N
        const int dd = N - k;
        return R0[(LY::base(dd / CS) + j) * THREADS + dd % CS];
    };
    const int M = L.re - L.rb;
    const double dim = (double)(M > N ? M : N);
    double dmax = 0.0, dmin = 1e300, frob2 = 0.0;
    bool zero_diag = false;
    double inv[ME];
    double2 v[ME];
    int col_off[ME], row_off[ME];                      // entry k = lf + m lpf of this lane: R[j][k] = R0[col_off + j THREADS],
    const double x0 = 1.0 / sqrt((double)N);           //                                    R[k][c] = R0[off(c) + row_off]
#pragma unroll
    for (int m = 0; m < ME; ++m) {
        inv[m] = 0.0; v[m] = make_double2(0.0, 0.0);
        const int k = L.lf + m * lpf, dd = N - (k < N ? k : N - 1);
        col_off[m] = LY::base(dd / CS) * THREADS + dd % CS;
        row_off[m] = (k < N ? k : 0) * THREADS;
    }
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        const double dj = Rat(j, j).x;
        const double a = fabs(dj);
        dmax = a > dmax ? a : dmax;
        dmin = a < dmin ? a : dmin;
        zero_diag |= (dj == 0.0);
#pragma unroll
        for (int m = 0; m < ME; ++m)
            if (j == L.lf + m * lpf) { inv[m] = 1.0 / dj; v[m] = make_double2(x0, 0.0); }
    }
    const bool need = L.fit >= 0 && !(dmin > rank_prefilter(N) * QNMFIT_EPS * dim * dmax);
    if (!__any_sync(full, need)) return false;
    // ||R||_F^2: each lane its own columns
#pragma unroll 1
    for (int k = L.lf; k < N; k += lpf)
        for (int j = 0; j <= k; ++j) {
            const double2 r = Rat(j, k);
            frob2 = fma(r.x, r.x, frob2);
            frob2 = fma(r.y, r.y, frob2);
        }
    for (int s = 1; s < lpf; s <<= 1) frob2 += __shfl_xor_sync(full, frob2, s);
    double nz = 1.0;
    bool bad = zero_diag;
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
#pragma unroll 1
        for (int j = 0; j < N; ++j) {                  // R^H y = x, column by column
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int m = 0; m < ME; ++m)
                if (j == L.lf + m * lpf) { v[m].x *= inv[m]; v[m].y *= inv[m]; sx = v[m].x; sy = v[m].y; }
            const double yx = __shfl_sync(full, sx, j % lpf, lpf), yy = __shfl_sync(full, sy, j % lpf, lpf);
#pragma unroll
            for (int m = 0; m < ME; ++m) {
                const int k = L.lf + m * lpf;
                if (k > j && k < N) {                  // x_k -= conj(R_jk) y_j
                    const double2 r = R0[col_off[m] + j * THREADS];
                    v[m].x = fma(-r.x, yx, v[m].x); v[m].x = fma(-r.y, yy, v[m].x);
                    v[m].y = fma(-r.x, yy, v[m].y); v[m].y = fma(r.y, yx, v[m].y);
                }
            }
        }
        double n2 = 0.0;
#pragma unroll 1
        for (int k = N - 1; k >= 0; --k) {             // R z = y, column by column
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int m = 0; m < ME; ++m)
                if (k == L.lf + m * lpf) { v[m].x *= inv[m]; v[m].y *= inv[m]; sx = v[m].x; sy = v[m].y; }
            const double zx = __shfl_sync(full, sx, k % lpf, lpf), zy = __shfl_sync(full, sy, k % lpf, lpf);
            n2 = fma(zx, zx, n2);
            n2 = fma(zy, zy, n2);
            const int koff = LY::base((N - k) / CS) * THREADS + (N - k) % CS;      // column k (uniform)
#pragma unroll
            for (int m = 0; m < ME; ++m) {
                const int j = L.lf + m * lpf;
                if (j < k) {                           // y_j -= R_jk z_k
                    const double2 r = R0[koff + row_off[m]];
                    v[m].x = fma(-r.x, zx, v[m].x); v[m].x = fma(r.y, zy, v[m].x);
                    v[m].y = fma(-r.x, zy, v[m].y); v[m].y = fma(-r.y, zx, v[m].y);
                }
            }
        }
        nz = sqrt(n2);
        if (!(nz < 1e300)) bad = true;                 // overflow / NaN: as singular as it gets
        const double sc = bad ? 0.0 : 1.0 / nz;
#pragma unroll
        for (int m = 0; m < ME; ++m) { v[m].x *= sc; v[m].y *= sc; }
    }
    const double cut = QNMFIT_RANK_MARGIN * QNMFIT_EPS * dim;
    return need && (bad || !(1.0 / nz > cut * cut * frob2));
}

// Back-substitution by lane 0 of the fit (it reads all CS lanes' columns of the factor);
// leaves C in the right-hand-side entries of lane 0.
template <int N, int CS, int THREADS>
PAIR_COLD void pair_backsub(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, const SmallLane &L, int tid,
                        const bool suspect, int &status, SmallAcc &acc)
{
    typedef PairLayout<N, CS> LY;
    if (L.fit < 0 || L.lf != 0) return;
    const int M = L.re - L.rb;
    const double2 *R0 = sm.R + tid;
    auto Roff = [&](int j, int k) { return R0[LY::entry(j, k) * THREADS + LY::lane_of(k)]; };
    auto Rdiag = [&](int j) { return R0[LY::entry(j, j) * THREADS + LY::lane_of(j)].x; };
    if (suspect) status |= QNMFIT_ST_RANK_DEFICIENT_;
    if (M <= N) status |= QNMFIT_ST_UNDERDETERMINED_;
    if (p.R) {
        double2 *Rout = p.R + (long long)L.fit * N * (N + 1);
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k <= N; ++k) {
                double2 val = make_double2(0.0, 0.0);
                if (k == j) val = make_double2(Rdiag(j), 0.0);
                else if (k > j) val = Roff(j, k);
                Rout[j * (N + 1) + k] = val;
            }
    }
    double2 C[N];
    double cn2 = 0.0;
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        double2 a = Roff(j, N);
        cn2 = fma(a.x, a.x, cn2);
        cn2 = fma(a.y, a.y, cn2);
#pragma unroll
        for (int k = j + 1; k < N; ++k) {
            const double2 Rjk = Roff(j, k);
            a.x = fma(-Rjk.x, C[k].x, a.x);
            a.x = fma(Rjk.y, C[k].y, a.x);
            a.y = fma(-Rjk.x, C[k].y, a.y);
            a.y = fma(-Rjk.y, C[k].x, a.y);
        }
        const double dd = Rdiag(j);
        if (dd != 0.0) { C[j].x = a.x / dd; C[j].y = a.y / dd; }
        else C[j] = make_double2(0.0, 0.0);
    }
    acc.cn2 = cn2;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        sm.R[LY::entry(j, N) * THREADS + tid] = C[j];
        if (p.C) p.C[(long long)L.fit * N + j] = C[j];
        if (!(fabs(C[j].x) < 1e300) || !(fabs(C[j].y) < 1e300)) status |= QNMFIT_ST_NONFINITE_;
    }
}

// Fast path (uniform grids), as small_fast_partials: the lanes of the fit share the columns
// of the first and last model row.
template <int N, int CS, int THREADS>
QF_HD void pair_fast_partials(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, const SmallLane &L,
                              int tid, const SmallAcc &acc, double (&part)[6])
{
    typedef PairLayout<N, CS> LY;
    part[0] = acc.sdd; part[1] = acc.res2;
    part[2] = part[3] = part[4] = part[5] = 0.0;
    if (L.fit < 0 || L.re <= L.rb) return;
    const int t0lane = tid - L.lf;
    const double tau_f = qf_sub_rn(sm.ts[L.rb - sm.t_off], L.t0);
    const double tau_l = qf_sub_rn(sm.ts[L.re - 1 - sm.t_off], L.t0);
    for (int j = L.lf; j < N; j += p.lanes_per_fit) {
        const double2 w = sm.om[L.slot * LY::NT + LY::tab(j)];
        const double2 C = sm.R[(LY::base(0) + j) * THREADS + t0lane];
        const double2 af = design_entry(w, tau_f), al = design_entry(w, tau_l);
        part[2] = fma(af.x, C.x, part[2]); part[2] = fma(-af.y, C.y, part[2]);
        part[3] = fma(af.x, C.y, part[3]); part[3] = fma(af.y, C.x, part[3]);
        part[4] = fma(al.x, C.x, part[4]); part[4] = fma(-al.y, C.y, part[4]);
        part[5] = fma(al.x, C.y, part[5]); part[5] = fma(al.y, C.x, part[5]);
    }
}

// Second pass (general path): model rows and the trapezoid-weighted inner products, one row
// at a time; the CS lanes of a group take contiguous parts of the group's rows.
template <int N, int CS, int THREADS>
PAIR_COLD void pair_eval(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, const SmallLane &L, const int h,
                     int tid, double (&sums)[4])
{
    typedef PairLayout<N, CS> LY;
    sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
    if (L.fit < 0 || L.re <= L.rb) return;
    double2 C[N];
    if (p.eval_only) {
#pragma unroll
        for (int j = 0; j < N; ++j) C[j] = p.C[(long long)L.fit * N + j];
    } else {
        const int t0lane = tid - L.lf;
#pragma unroll
        for (int j = 0; j < N; ++j) C[j] = sm.R[(LY::base(0) + j) * THREADS + t0lane];
    }
    const double *ts = sm.ts - sm.t_off;
    const double2 *ds = sm.ds - sm.t_off + L.d_off;
    const int per = (L.hi - L.lo + CS - 1) / CS;
    const int a = L.lo + h * per < L.hi ? L.lo + h * per : L.hi;
    const int b = a + per < L.hi ? a + per : L.hi;
    const bool direct = !(p.dt_nominal > 0.0);
    const int arows = p.anchor_rows > 0 ? p.anchor_rows : QNMFIT_DEFAULT_ANCHOR_ROWS;
    double2 z[N];
    double tau = 0.0;
#pragma unroll 1
    for (int r = a; r < b; ++r) {
        const double tau_r = qf_sub_rn(ts[r], L.t0);
        if (direct || (r - a) % arows == 0) {
#pragma unroll
            for (int j = 0; j < N; ++j) z[j] = design_entry(sm.om[L.slot * LY::NT + LY::tab(j)], tau_r);
        } else {
            const double de = qf_sub_rn(qf_sub_rn(tau_r, tau), p.dt_nominal);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 q = sm.qq[L.slot * LY::NT + LY::tab(j)], w = sm.qw[L.slot * LY::NT + LY::tab(j)];
                z[j] = c_mul(z[j], make_double2(fma(w.x, de, q.x), fma(w.y, de, q.y)));
            }
        }
        tau = tau_r;
        double mx = 0.0, my = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            mx = fma(z[j].x, C[j].x, mx);
            my = fma(z[j].x, C[j].y, my);
            mx = fma(-z[j].y, C[j].y, mx);
            my = fma(z[j].y, C[j].x, my);
        }
        const double2 dv = ds[r];
        if (p.model) p.model[(long long)L.fit * p.model_stride + (r - L.rb)] = make_double2(mx, my);
        const int rm = r - 1 < L.rb ? L.rb : r - 1;
        const int rp = r + 1 > L.re - 1 ? L.re - 1 : r + 1;
        const double wgt = 0.5 * (ts[rp] - ts[rm]);
        sums[0] = fma(wgt, fma(mx, dv.x, my * dv.y), sums[0]);
        sums[1] = fma(wgt, fma(mx, mx, my * my), sums[1]);
        sums[2] = fma(wgt, fma(dv.x, dv.x, dv.y * dv.y), sums[2]);
        const double ex = mx - dv.x, ey = my - dv.y;
        sums[3] += fma(ex, ex, ey * ey);
    }
}

// Entry idx of the CTA's padded frequency tables (PairLayout::NT): the CTA fills idx = 0 .. fpc N + CS - 1.
template <int N, int CS, int THREADS>
QF_HD void pair_fill_tables(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, int cta_first, int fpc, int idx)
{
    const int e = idx - (CS - 1);                 // entry of the [fpc][N] table, or padding
    const int slot = e >= 0 ? e / N : 0, j = e - slot * N;
    const int fit = cta_first + slot;
    double2 w = make_double2(0.0, 0.0), q = w, qw = w;
    if (e >= 0 && slot < fpc && fit < p.n_fits) {
        w = fit_omega(p, input_fit(p, fit), j);
        if (p.dt_nominal > 0.0) {
            q = design_entry(w, p.dt_nominal);
            qw = c_mul(q, make_double2(w.y, -w.x));
        }
    }
    sm.om[idx] = w; sm.qq[idx] = q; sm.qw[idx] = qw;
}

// Everything a lane does once the CTA's shared memory is set up (tables filled, factors zeroed,
// window staged).  Called by all lanes of a warp together: the kernel below, and the lock-step
// host emulation of tests/hostsim.
template <int N, int CS, int MB, int THREADS>
QF_HD void pair_lane_body(const FitParams &p, const PairSmem<N, CS, THREADS> &sm, const SmallLane &L, int tid)
{
    const int lpf = p.lanes_per_fit;
    const int h = L.lf % CS;
    int status = 0;
    SmallAcc acc;
    acc.sdd = acc.res2 = acc.cn2 = 0.0;
    if (!p.eval_only) {
        pair_leaf<N, CS, MB, THREADS>(p, sm, L, h, tid, acc);
        if (h != 0) acc.sdd = acc.res2 = 0.0;               // only lane 0 of a group holds the right-hand side
        const int grp = L.lf / CS;
        for (int s = 1; s < lpf / CS; s <<= 1) {            // R-combine over the fit's row groups
            __syncwarp();
            const bool live = L.fit >= 0 && (grp % (2 * s)) == 0;
            SmallAcc tacc; tacc.res2 = 0.0;
            pair_tree_block<N, CS, MB, THREADS, 0>(sm, tid, live ? tid + s * CS : tid, h, live, tacc);
            if (live && h == 0) acc.res2 += tacc.res2;
        }
        __syncwarp();
        const bool suspect = pair_rank_suspect<N, CS, THREADS>(p, sm, L, tid);
        pair_backsub<N, CS, THREADS>(p, sm, L, tid, suspect, status, acc);
        __syncwarp();
    }
    if (p.fast_mismatch) {
        double part[6];
        pair_fast_partials<N, CS, THREADS>(p, sm, L, tid, acc, part);
        for (int s = 1; s < lpf; s <<= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) part[q] += __shfl_xor_sync(0xffffffffu, part[q], s);
        }
        if (L.fit >= 0 && L.lf == 0) {
            if (L.re > L.rb) {
                small_fast_finalize(p, L, sm.ds[L.rb - sm.t_off + L.d_off], sm.ds[L.re - 1 - sm.t_off + L.d_off], part,
                                    acc.cn2, status);
            } else {
                const double2 z = make_double2(0.0, 0.0);
                small_fast_finalize(p, L, z, z, part, acc.cn2, status);
            }
        }
        return;
    }
    double sums[4];
    pair_eval<N, CS, THREADS>(p, sm, L, h, tid, sums);
    for (int s = 1; s < lpf; s <<= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) sums[q] += __shfl_xor_sync(0xffffffffu, sums[q], s);
    }
    small_finalize(p, L, sums, status);
}

#ifndef QNMFIT_HOSTSIM
template <int N, int CS, int MB, int THREADS, bool STAGED>
__global__ void __launch_bounds__(THREADS, 1) fit_pair_kernel(const __grid_constant__ FitParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef PairLayout<N, CS> LY;
    PairSmem<N, CS, THREADS> sm;
    const int lpf = p.lanes_per_fit;
    const int fpc = THREADS / lpf;
    const int tid = threadIdx.x;
    sm.carve(smem_raw, fpc, STAGED ? p.stage_rows : 0);
    if (STAGED) {
        double *ts_w = const_cast<double *>(sm.ts);
        double2 *ds_w = const_cast<double2 *>(sm.ds);
        for (int r = tid; r < p.stage_rows + SMALL_STAGE_PAD; r += THREADS) {
            const int src = p.stage_begin + (r < p.stage_rows ? r : p.stage_rows - 1);
            ts_w[r] = p.times[src];
            ds_w[r] = p.data[src];
        }
        sm.t_off = p.stage_begin;
    } else {
        sm.ts = p.times;
        sm.ds = p.data;
        sm.t_off = 0;
    }
    for (int idx = tid; idx < fpc * N + CS; idx += THREADS)
        pair_fill_tables<N, CS, THREADS>(p, sm, blockIdx.x * fpc, fpc, idx);
    const SmallLane L = pair_lane_setup<CS>(p, blockIdx.x, tid, THREADS, !STAGED, MB, STAGED ? p.stage_begin : 0);
#pragma unroll 1
    for (int e = 0; e < LY::E; ++e) sm.R[e * THREADS + tid] = make_double2(0.0, 0.0);
    __syncthreads();
    pair_lane_body<N, CS, MB, THREADS>(p, sm, L, tid);
}
#endif  // !QNMFIT_HOSTSIM
