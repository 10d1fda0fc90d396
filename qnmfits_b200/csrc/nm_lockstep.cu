// nm_lockstep.cu — host code only: the lock-step Nelder-Mead that drives the batched
// free-frequency search (qnmfit_nm_* of include/qnmfit.h).
//
// The reference calls scipy.optimize.minimize(method='Nelder-Mead', bounds=...) once per
// waveform (qnmfits/qnmfits.py:2031-2038 free_frequency_fit, :1520-1560 calculate_epsilon);
// every objective call is one least-squares fit.  Here B searches advance together and the
// objective is ONE batched device launch per optimiser step, so what is left on the host is
// the simplex bookkeeping of B problems between two launches.  In numpy that bookkeeping
// (qnmfits_b200/_neldermead.py: the specification of this file, kept and tested against it)
// cost more than the launches: ~2 ms per step at B = 4096 against ~0.2 ms for the fits.  This
// is the same state machine in C++, written as a coroutine by hand: qnmfit_nm_step() takes the
// objective values of the points it handed out last time and returns the next set of points.
//
// The arithmetic follows scipy/optimize/_optimize.py::_minimize_neldermead (scipy 1.18,
// non-adaptive: rho = 1, chi = 2, psi = 0.5, sigma = 0.5) operation for operation — plain
// multiplications and additions in numpy's order, no contraction (this unit is compiled with
// -ffp-contract=off) — so that a problem follows scipy's trajectory exactly when the objective
// returns the same floats.  One point needs care: scipy orders the simplex with np.argsort,
// whose order of EQUAL values depends on the CPU (numpy dispatches to an unstable SIMD network
// sort where AVX2 / AVX-512 exist).  A row with distinct values has one sorted order whatever
// the algorithm; the rows with ties or NaNs are handed, all rows of a step in one call, to the
// caller's `order` callback (the Python wrapper passes np.argsort itself), or sorted stably when
// there is none.  (They are common: a simplex pressed against a bound has coinciding vertices.)
#include <cmath>
#include <cstdint>
#include <limits>
#include <new>
#include <vector>

#include "qnmfit.h"

namespace {

constexpr double RHO = 1.0, CHI = 2.0, PSI = 0.5, SIGMA = 0.5;
constexpr double NONZDELT = 0.05, ZDELT = 0.00025;

enum Phase { P_START, P_INIT, P_REFLECT, P_SECOND, P_SHRINK, P_DONE };
enum Kind : uint8_t { K_EXPAND, K_ACCEPT, K_OUTSIDE, K_INSIDE };

inline double clip(double v, double lo, double hi)
{
    // numpy.clip = minimum(maximum(v, lo), hi); a NaN stays a NaN in both forms
    v = (v < lo) ? lo : v;
    return (v > hi) ? hi : v;
}

// the order numpy sorts in: NaNs last
inline bool sorts_before(double a, double b) { return a < b || (b != b && a == a); }

}  // namespace

struct qnmfit_nm {
    int64_t B;
    int N;
    std::vector<double> sim;      // [B][N + 1][N]
    std::vector<double> fsim;     // [B][N + 1]
    std::vector<double> lower, upper;
    std::vector<int64_t> fcalls, iters, status;
    std::vector<uint8_t> active;
    double xatol, fatol, maxiter, maxfun;
    int64_t n_calls;
    qnmfit_nm_order_fn order;
    void *order_user;

    Phase phase;
    int k;                        // vertex being evaluated (P_INIT, P_SHRINK)
    std::vector<int64_t> req;     // problems of the request that is out, in request order
    // the iteration in flight: problems whose reflection point was evaluated, and per problem
    std::vector<int64_t> a;
    std::vector<double> xbar, xr, second;     // [len(a)][N]
    std::vector<double> fxr, f2;
    std::vector<uint8_t> kind, ok2;
    std::vector<int64_t> req_pos;             // P_SECOND: position in `a` of each requested point
    std::vector<int64_t> s;                   // problems that shrink
    std::vector<double> tmp_x, tmp_f;
    std::vector<int64_t> perm, everyone, amb_rows, amb_order;
    std::vector<double> amb_values;

    double *vertex(int64_t b, int j) { return &sim[(b * (N + 1) + j) * N]; }
    double *values(int64_t b) { return &fsim[b * (N + 1)]; }
    bool has_budget(int64_t b) const { return (double)fcalls[b] < maxfun; }
    void stop(int64_t b) { active[b] = 0; status[b] = 1; }

    void apply_order(int64_t b, const int64_t *pm)
    {
        const int n = N + 1;
        bool sorted = true;
        for (int i = 0; i < n; ++i) sorted = sorted && pm[i] == i;
        if (sorted) return;
        double *f = values(b), *x = vertex(b, 0);
        for (int i = 0; i < n; ++i) {
            tmp_f[i] = f[pm[i]];
            for (int c = 0; c < N; ++c) tmp_x[i * N + c] = x[pm[i] * N + c];
        }
        for (int i = 0; i < n; ++i) f[i] = tmp_f[i];
        for (int i = 0; i < n * N; ++i) x[i] = tmp_x[i];
    }

    // Order the simplices of `rows` by value.  Rows whose order is ambiguous (equal values,
    // NaNs) are collected and handed to the caller's `order` in ONE call.
    void sort_rows(const int64_t *rows, int64_t n_rows)
    {
        const int n = N + 1;
        amb_rows.clear(); amb_values.clear();
        for (int64_t r = 0; r < n_rows; ++r) {
            const int64_t b = rows[r];
            const double *f = values(b);
            for (int i = 0; i < n; ++i) perm[i] = i;
            for (int i = 1; i < n; ++i) {         // insertion sort: stable
                const int64_t p = perm[i];
                int j = i;
                while (j > 0 && sorts_before(f[p], f[perm[j - 1]])) { perm[j] = perm[j - 1]; --j; }
                perm[j] = p;
            }
            bool ambiguous = f[perm[n - 1]] != f[perm[n - 1]];
            for (int i = 1; i < n && !ambiguous; ++i) ambiguous = f[perm[i]] == f[perm[i - 1]];
            if (ambiguous && order) {
                amb_rows.push_back(b);
                amb_values.insert(amb_values.end(), f, f + n);
            } else {
                apply_order(b, perm.data());
            }
        }
        if (amb_rows.empty()) return;
        amb_order.assign(amb_rows.size() * n, 0);
        order(amb_values.data(), (int64_t)amb_rows.size(), n, amb_order.data(), order_user);
        for (size_t r = 0; r < amb_rows.size(); ++r) {
            // a permutation, or the callback is broken: then the row is left as it is
            const int64_t *pm = &amb_order[r * n];
            uint64_t seen_lo = 0, seen_hi = 0;
            bool ok = true;
            for (int i = 0; i < n && ok; ++i) {
                if (pm[i] < 0 || pm[i] >= n) { ok = false; break; }
                uint64_t &word = pm[i] < 64 ? seen_lo : seen_hi;
                const uint64_t bit = uint64_t(1) << (pm[i] & 63);
                ok = !(word & bit);
                word |= bit;
            }
            if (ok) apply_order(amb_rows[r], pm);
        }
    }

    void put(double *x_out, int64_t *idx_out, int64_t b, const double *x)
    {
        const int64_t at = (int64_t)req.size();
        for (int c = 0; c < N; ++c) x_out[at * N + c] = x[c];
        idx_out[at] = b;
        req.push_back(b);
    }

    // scipy's while-condition and convergence test, then the reflection points of every
    // problem that goes on; false when nothing is left to do
    bool begin_iteration(double *x_out, int64_t *idx_out)
    {
        req.clear();
        a.clear();
        xbar.resize((size_t)B * N); xr.resize((size_t)B * N);    // filled by position in `a`
        const double n_vars = (double)N;
        for (int64_t b = 0; b < B; ++b) {
            if (!active[b]) continue;
            if ((double)fcalls[b] >= maxfun) { stop(b); continue; }
            if ((double)iters[b] >= maxiter) { active[b] = 0; status[b] = 2; continue; }
            // np.max propagates a NaN; a comparison with it is false
            double size = 0.0, spread = 0.0;
            bool nan = false;
            const double *x0 = vertex(b, 0), *f = values(b);
            for (int j = 1; j <= N; ++j) {
                const double *xj = vertex(b, j);
                for (int c = 0; c < N; ++c) {
                    const double d = std::fabs(xj[c] - x0[c]);
                    if (d != d) nan = true; else if (d > size) size = d;
                }
                const double d = std::fabs(f[0] - f[j]);
                if (d != d) nan = true; else if (d > spread) spread = d;
            }
            if (!nan && size <= xatol && spread <= fatol) { active[b] = 0; continue; }
            if (!has_budget(b)) { stop(b); continue; }      // (cannot happen after the test above)
            const double *worst = vertex(b, N);
            const size_t at = a.size() * N;
            for (int c = 0; c < N; ++c) {
                double sum = x0[c];
                for (int j = 1; j < N; ++j) sum = sum + x0[j * N + c];
                const double xb = sum / n_vars;
                xbar[at + c] = xb;
                xr[at + c] = clip((1 + RHO) * xb - RHO * worst[c], lower[c], upper[c]);
            }
            a.push_back(b);
            put(x_out, idx_out, b, &xr[at]);
        }
        return !a.empty();
    }

    // after the reflection values: classify and hand out the expansion / contraction points
    void begin_second(double *x_out, int64_t *idx_out)
    {
        const size_t n = a.size();
        kind.assign(n, K_ACCEPT); ok2.assign(n, 0);
        f2.assign(n, std::numeric_limits<double>::infinity());
        second.assign(n * N, 0.0);
        req.clear(); req_pos.clear();
        for (size_t p = 0; p < n; ++p) {
            const int64_t b = a[p];
            const double *f = values(b), *worst = vertex(b, N), *xb = &xbar[p * N];
            double *x2 = &second[p * N];
            if (fxr[p] < f[0]) {
                kind[p] = K_EXPAND;
                for (int c = 0; c < N; ++c) x2[c] = (1 + RHO * CHI) * xb[c] - RHO * CHI * worst[c];
            } else if (fxr[p] < f[N - 1]) {
                kind[p] = K_ACCEPT;
                continue;
            } else if (fxr[p] < f[N]) {
                kind[p] = K_OUTSIDE;
                for (int c = 0; c < N; ++c) x2[c] = (1 + PSI * RHO) * xb[c] - PSI * RHO * worst[c];
            } else {
                kind[p] = K_INSIDE;
                for (int c = 0; c < N; ++c) x2[c] = (1 - PSI) * xb[c] + PSI * worst[c];
            }
            for (int c = 0; c < N; ++c) x2[c] = clip(x2[c], lower[c], upper[c]);
            if (has_budget(b)) {
                ok2[p] = 1;
                req_pos.push_back((int64_t)p);
                put(x_out, idx_out, b, x2);
            } else {
                stop(b);
            }
        }
    }

    // with both values known: replace the worst vertex, or mark the problem for a shrink
    void update()
    {
        s.clear();
        for (size_t p = 0; p < a.size(); ++p) {
            const int64_t b = a[p];
            double *f = values(b);
            const double *x = &xr[p * N];
            double fx = fxr[p];
            bool replace = kind[p] == K_ACCEPT, shrink = false;
            if (ok2[p]) {
                const bool take = (kind[p] == K_EXPAND && f2[p] < fxr[p]) ||
                                  (kind[p] == K_OUTSIDE && f2[p] <= fxr[p]) ||
                                  (kind[p] == K_INSIDE && f2[p] < f[N]);
                if (take) { x = &second[p * N]; fx = f2[p]; }
                if (kind[p] == K_EXPAND) replace = true;
                else if (kind[p] == K_OUTSIDE || kind[p] == K_INSIDE) { replace = take; shrink = !take; }
            }
            if (replace) {
                double *w = vertex(b, N);
                for (int c = 0; c < N; ++c) w[c] = x[c];
                f[N] = fx;
            }
            if (shrink) s.push_back(b);
        }
    }

    // vertex k of every problem that shrinks (and is still running) moves towards the best one
    bool begin_shrink(double *x_out, int64_t *idx_out)
    {
        req.clear();
        bool any = false;
        for (int64_t b : s) {
            if (!active[b]) continue;
            any = true;
            double *xk = vertex(b, k);
            const double *x0 = vertex(b, 0);
            for (int c = 0; c < N; ++c) xk[c] = clip(x0[c] + SIGMA * (xk[c] - x0[c]), lower[c], upper[c]);
            if (has_budget(b)) put(x_out, idx_out, b, xk);
            else stop(b);
        }
        return any;
    }

    void end_iteration()
    {
        for (int64_t b : a)
            if (active[b]) iters[b] += 1;
        sort_rows(a.data(), (int64_t)a.size());
    }

    int64_t step(const double *f_prev, double *x_out, int64_t *idx_out)
    {
        // consume the values of the request that is out
        if (!req.empty()) {
            if (!f_prev) return QNMFIT_E_NULL;
            n_calls += 1;
            for (size_t i = 0; i < req.size(); ++i) fcalls[req[i]] += 1;
            switch (phase) {
            case P_INIT:
                for (size_t i = 0; i < req.size(); ++i) values(req[i])[k] = f_prev[i];
                break;
            case P_REFLECT:
                fxr.assign(f_prev, f_prev + req.size());     // req == a here
                break;
            case P_SECOND:
                for (size_t i = 0; i < req.size(); ++i) f2[req_pos[i]] = f_prev[i];
                break;
            case P_SHRINK:
                for (size_t i = 0; i < req.size(); ++i) values(req[i])[k] = f_prev[i];
                break;
            default:
                break;
            }
            req.clear();
        } else if (phase == P_REFLECT) {
            fxr.clear();
        }
        // advance until there is something to evaluate
        for (;;) {
            switch (phase) {
            case P_START:
                phase = P_INIT;
                k = -1;
                // fall through
            case P_INIT:
                ++k;
                if (k <= N) {
                    for (int64_t b = 0; b < B; ++b)
                        if (has_budget(b)) put(x_out, idx_out, b, vertex(b, k));
                    if (!req.empty()) return (int64_t)req.size();
                    continue;
                }
                everyone.resize(B);
                for (int64_t b = 0; b < B; ++b) everyone[b] = b;
                sort_rows(everyone.data(), B);
                if (!begin_iteration(x_out, idx_out)) { phase = P_DONE; return 0; }
                phase = P_REFLECT;
                return (int64_t)req.size();
            case P_REFLECT:
                begin_second(x_out, idx_out);
                phase = P_SECOND;
                if (!req.empty()) return (int64_t)req.size();
                continue;
            case P_SECOND:
                update();
                phase = P_SHRINK;
                k = 0;
                continue;
            case P_SHRINK:
                ++k;
                if (k <= N && !s.empty() && begin_shrink(x_out, idx_out)) {
                    if (!req.empty()) return (int64_t)req.size();
                    continue;
                }
                end_iteration();
                if (!begin_iteration(x_out, idx_out)) { phase = P_DONE; return 0; }
                phase = P_REFLECT;
                return (int64_t)req.size();
            case P_DONE:
                return 0;
            }
        }
    }
};

extern "C" {

int qnmfit_nm_create(int64_t n_problems, int n_vars, const double *x0, const double *lower,
                     const double *upper, double xatol, double fatol, double maxiter, double maxfun,
                     qnmfit_nm_order_fn order, void *order_user, qnmfit_nm **out)
{
    if (!out) return QNMFIT_E_NULL;
    *out = nullptr;
    if (!x0 || !lower || !upper) return QNMFIT_E_NULL;
    if (n_problems < 0 || n_vars < 1 || n_vars > QNMFIT_NM_MAX_VARS) return QNMFIT_E_SHAPE;
    qnmfit_nm *nm = new (std::nothrow) qnmfit_nm();
    if (!nm) return QNMFIT_E_SHAPE;
    try {
        const int64_t B = nm->B = n_problems;
        const int N = nm->N = n_vars;
        nm->lower.assign(lower, lower + N);
        nm->upper.assign(upper, upper + N);
        nm->xatol = xatol; nm->fatol = fatol; nm->maxiter = maxiter; nm->maxfun = maxfun;
        nm->n_calls = 0;
        nm->order = order; nm->order_user = order_user;
        nm->phase = P_START; nm->k = 0;
        nm->sim.assign((size_t)B * (N + 1) * N, 0.0);
        nm->fsim.assign((size_t)B * (N + 1), std::numeric_limits<double>::infinity());
        nm->fcalls.assign(B, 0); nm->iters.assign(B, 1); nm->status.assign(B, 0);
        nm->active.assign(B, 1);
        nm->tmp_x.assign((size_t)(N + 1) * N, 0.0); nm->tmp_f.assign(N + 1, 0.0);
        nm->perm.assign(N + 1, 0);
        nm->req.reserve(B); nm->a.reserve(B);
        // the start simplex as scipy builds it: x0 clipped, one vertex per coordinate moved by
        // 5 % (0.00025 from zero), reflected at the upper bound, clipped
        for (int64_t b = 0; b < B; ++b) {
            double *v0 = nm->vertex(b, 0);
            for (int c = 0; c < N; ++c) v0[c] = clip(x0[b * N + c], lower[c], upper[c]);
            for (int j = 1; j <= N; ++j) {
                double *v = nm->vertex(b, j);
                for (int c = 0; c < N; ++c) v[c] = v0[c];
                v[j - 1] = (v0[j - 1] != 0.0) ? (1 + NONZDELT) * v0[j - 1] : ZDELT;
            }
            for (int j = 0; j <= N; ++j) {
                double *v = nm->vertex(b, j);
                for (int c = 0; c < N; ++c) {
                    if (v[c] > upper[c]) v[c] = 2 * upper[c] - v[c];
                    v[c] = clip(v[c], lower[c], upper[c]);
                }
            }
        }
    } catch (...) {
        delete nm;
        return QNMFIT_E_SHAPE;
    }
    *out = nm;
    return 0;
}

int64_t qnmfit_nm_step(qnmfit_nm *nm, const double *f_prev, double *x_out, int64_t *idx_out)
{
    if (!nm || !x_out || !idx_out) return QNMFIT_E_NULL;
    try {
        return nm->step(f_prev, x_out, idx_out);
    } catch (...) {
        return QNMFIT_E_SHAPE;
    }
}

int qnmfit_nm_result(const qnmfit_nm *nm, double *x, double *fun, int64_t *nit, int64_t *nfev,
                     int64_t *status, int64_t *n_calls)
{
    if (!nm) return QNMFIT_E_NULL;
    const int N = nm->N;
    for (int64_t b = 0; b < nm->B; ++b) {
        if (x)
            for (int c = 0; c < N; ++c) x[b * N + c] = nm->sim[(size_t)b * (N + 1) * N + c];
        if (fun) {
            // np.min: a NaN wins
            const double *f = &nm->fsim[(size_t)b * (N + 1)];
            double m = f[0];
            for (int j = 1; j <= N; ++j)
                if (m == m && (f[j] != f[j] || f[j] < m)) m = f[j];
            fun[b] = m;
        }
        if (nit) nit[b] = nm->iters[b];
        if (nfev) nfev[b] = nm->fcalls[b];
        if (status) status[b] = nm->status[b];
    }
    if (n_calls) *n_calls = nm->n_calls;
    return 0;
}

int qnmfit_nm_destroy(qnmfit_nm *nm)
{
    delete nm;
    return 0;
}

}  // extern "C"
