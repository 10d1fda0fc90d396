"""
Lock-step Nelder-Mead for many independent minimisations.

The reference drives ``scipy.optimize.minimize(method='Nelder-Mead', bounds=...)`` once
per waveform (qnmfits/qnmfits.py:1992-2041 free_frequency_fit, :1520-1560
calculate_epsilon), each objective call being one least-squares fit.  scipy's optimiser
cannot be batched, so this module restates its bounded Nelder-Mead
(scipy/optimize/_optimize.py::_minimize_neldermead, scipy 1.18, non-adaptive
coefficients rho=1, chi=2, psi=0.5, sigma=0.5, initial simplex 5 % / 0.00025, clip to
the bounds, xatol & fatol test, maxiter = maxfun = 200 N) for B problems advancing
together: per iteration ONE batched objective call for the reflection points, one for
the expansion / contraction points of the problems that need them, and N for shrinks.

With an objective that returns the same floats as the scalar one, every problem follows
exactly scipy's trajectory (tests/test_neldermead.py checks x, fun, nit and nfev for
equality).  The only deliberate difference: when ``maxfun`` is hit in the middle of an
iteration scipy abandons that iteration through an exception; here the problem simply
stops before the call.

Two implementations of the same state machine.  ``minimize_lockstep_numpy`` below is the
specification.  Its bookkeeping costs ~2 ms per optimiser step at B = 4096 (fancy indexing
over a dozen small arrays), ten times the batched fit it waits for, so ``minimize_lockstep``
runs the C++ form in libqnmfit.so (``qnmfit_nm_*``, csrc/nm_lockstep.cu: host code, one call
per optimiser step) whenever the problem has at most ``QNMFIT_NM_MAX_VARS`` variables; the
tests hold the two to identical results, ties in the simplex order included (the C++ form
asks numpy's own argsort for the order of equal values, which is not stable).
"""
import ctypes as C

import numpy as np

RHO, CHI, PSI, SIGMA = 1.0, 2.0, 0.5, 0.5
NONZDELT, ZDELT = 0.05, 0.00025


class LockstepResult:
    def __init__(self, x, fun, nit, nfev, status, n_calls):
        self.x, self.fun, self.nit, self.nfev, self.status = x, fun, nit, nfev, status
        self.success = status == 0
        self.n_calls = n_calls        # batched objective calls issued

    def __repr__(self):
        return (f"LockstepResult(B={len(self.fun)}, nit<= {int(self.nit.max())}, "
                f"nfev<= {int(self.nfev.max())}, batched calls={self.n_calls})")


def initial_simplex(x0, lower, upper):
    """(B, N+1, N) start simplices from (B, N) start points, as scipy builds them."""
    x0 = np.clip(np.asarray(x0, dtype=float), lower, upper)
    B, N = x0.shape
    sim = np.empty((B, N + 1, N), dtype=float)
    sim[:, 0] = x0
    for k in range(N):
        y = x0.copy()
        nz = y[:, k] != 0
        y[nz, k] = (1 + NONZDELT) * y[nz, k]
        y[~nz, k] = ZDELT
        sim[:, k + 1] = y
    msk = sim > upper
    sim = np.where(msk, 2 * upper - sim, sim)
    return np.clip(sim, lower, upper)


def _limits(N, maxiter, maxfun):
    """scipy's defaults for the two budgets."""
    if maxiter is None and maxfun is None:
        maxiter = maxfun = N * 200
    elif maxiter is None:
        maxiter = N * 200 if maxfun == np.inf else np.inf
    elif maxfun is None:
        maxfun = N * 200 if maxiter == np.inf else np.inf
    return maxiter, maxfun


@C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int64, C.c_int, C.POINTER(C.c_int64), C.c_void_p)
def _numpy_order(values, n_rows, n, order, _user):
    """Order of the simplices whose values hold ties or NaNs: what scipy's ``np.argsort(fsim)``
    gives for each of them (numpy sorts the rows of a 2-D array with the same routine)."""
    got = np.argsort(np.ctypeslib.as_array(values, (n_rows, n)), axis=1)
    np.ctypeslib.as_array(order, (n_rows, n))[...] = got


def minimize_lockstep(fun, x0, bounds, xatol=1e-4, fatol=1e-4, maxiter=None, maxfun=None):
    """Minimise B independent functions of N variables.

    fun(X, idx) -> f : X float64 (n, N) trial points, idx int64 (n,) the problems they
    belong to (ascending, at most one point per problem per call); returns float (n,).
    x0: (B, N) or (N,) start point(s) (then B must be given by idx range = 1).
    bounds: sequence of N (low, high) pairs.

    Runs the state machine of libqnmfit.so (``qnmfit_nm_step``: one C call between two
    objective calls); more than ``QNMFIT_NM_MAX_VARS`` variables go to the numpy form.
    """
    from . import _cabi
    x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=float)))
    B, N = x0.shape
    if N > _cabi.NM_MAX_VARS:
        return minimize_lockstep_numpy(fun, x0, bounds, xatol, fatol, maxiter, maxfun)
    lib = _cabi.load_library()
    lower = np.array([b[0] for b in bounds], dtype=float)
    upper = np.array([b[1] for b in bounds], dtype=float)
    if len(lower) != N:
        raise ValueError("bounds must hold one (low, high) pair per variable")
    maxiter, maxfun = _limits(N, maxiter, maxfun)
    handle = C.c_void_p()
    rc = lib.qnmfit_nm_create(B, N, x0.ctypes.data, lower.ctypes.data, upper.ctypes.data, float(xatol),
                              float(fatol), float(maxiter), float(maxfun), _numpy_order, None, C.byref(handle))
    if rc:
        raise _cabi.QnmfitError(rc, "qnmfit_nm_create: bad arguments")
    try:
        X = np.empty((max(B, 1), N), dtype=np.float64)
        idx = np.empty(max(B, 1), dtype=np.int64)
        f = None
        while True:
            n = lib.qnmfit_nm_step(handle, None if f is None else f.ctypes.data, X.ctypes.data, idx.ctypes.data)
            if n <= 0:
                break
            # the objective gets its own copies: the next step overwrites X and idx
            f = np.ascontiguousarray(fun(X[:n].copy(), idx[:n].copy()), dtype=np.float64).reshape(-1)
            if len(f) != n:
                raise ValueError(f"the objective returned {len(f)} values for {n} points")
        if n < 0:
            raise _cabi.QnmfitError(int(n), "qnmfit_nm_step: bad arguments")
        x = np.empty((B, N), dtype=np.float64)
        val = np.empty(B, dtype=np.float64)
        nit, nfev, status = (np.empty(B, dtype=np.int64) for _ in range(3))
        n_calls = C.c_int64()
        lib.qnmfit_nm_result(handle, x.ctypes.data, val.ctypes.data, nit.ctypes.data, nfev.ctypes.data,
                             status.ctypes.data, C.addressof(n_calls))
    finally:
        lib.qnmfit_nm_destroy(handle)
    return LockstepResult(x, val, nit, nfev, status, int(n_calls.value))


def minimize_lockstep_numpy(fun, x0, bounds, xatol=1e-4, fatol=1e-4, maxiter=None, maxfun=None):
    """The same search written with numpy: the specification of ``qnmfit_nm_*`` (and the form
    used beyond ``QNMFIT_NM_MAX_VARS`` variables).  Same arguments as ``minimize_lockstep``."""
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    B, N = x0.shape
    lower = np.array([b[0] for b in bounds], dtype=float)
    upper = np.array([b[1] for b in bounds], dtype=float)
    maxiter, maxfun = _limits(N, maxiter, maxfun)

    sim = initial_simplex(x0, lower, upper)
    fsim = np.full((B, N + 1), np.inf)
    fcalls = np.zeros(B, dtype=np.int64)
    everyone = np.arange(B)
    n_calls = 0
    for k in range(N + 1):
        ok = fcalls < maxfun
        idx = everyone[ok]
        if len(idx):
            fsim[idx, k] = fun(sim[idx, k], idx)
            fcalls[idx] += 1
            n_calls += 1

    def sort_rows(rows):
        order = np.argsort(fsim[rows], axis=1)
        fsim[rows] = np.take_along_axis(fsim[rows], order, axis=1)
        sim[rows] = np.take_along_axis(sim[rows], order[:, :, None], axis=1)

    sort_rows(everyone)
    iterations = np.ones(B, dtype=np.int64)
    active = np.ones(B, dtype=bool)
    status = np.zeros(B, dtype=np.int64)

    def evaluate(points, idx):
        """Objective for problems idx that still have budget; others get +inf and stop."""
        nonlocal n_calls
        ok = fcalls[idx] < maxfun
        out = np.full(len(idx), np.inf)
        if ok.any():
            out[ok] = fun(points[ok], idx[ok])
            fcalls[idx[ok]] += 1
            n_calls += 1
        stopped = idx[~ok]
        active[stopped] = False
        status[stopped] = 1
        return out, ok

    while True:
        # loop condition of scipy's while
        over_f = active & (fcalls >= maxfun)
        over_i = active & ~over_f & (iterations >= maxiter)
        status[over_f] = 1
        status[over_i] = 2
        active &= ~(over_f | over_i)
        a = everyone[active]
        if len(a) == 0:
            break
        # convergence test
        size = np.max(np.abs(sim[a, 1:] - sim[a, :1]).reshape(len(a), -1), axis=1)
        spread = np.max(np.abs(fsim[a, :1] - fsim[a, 1:]), axis=1)
        done = (size <= xatol) & (spread <= fatol)
        active[a[done]] = False
        a = a[~done]
        if len(a) == 0:
            break

        xbar = np.add.reduce(sim[a, :-1], 1) / N
        worst = sim[a, -1]
        xr = np.clip((1 + RHO) * xbar - RHO * worst, lower, upper)
        fxr, ok = evaluate(xr, a)
        # problems that ran out of budget leave the iteration untouched
        a, xbar, worst, xr, fxr = a[ok], xbar[ok], worst[ok], xr[ok], fxr[ok]

        f0, fm2, fm1 = fsim[a, 0], fsim[a, -2], fsim[a, -1]
        expand = fxr < f0
        accept = ~expand & (fxr < fm2)
        outside = ~expand & ~accept & (fxr < fm1)
        inside = ~expand & ~accept & ~outside

        second = np.empty_like(xr)
        second[expand] = (1 + RHO * CHI) * xbar[expand] - RHO * CHI * worst[expand]
        second[outside] = (1 + PSI * RHO) * xbar[outside] - PSI * RHO * worst[outside]
        second[inside] = (1 - PSI) * xbar[inside] + PSI * worst[inside]
        need = expand | outside | inside
        second = np.clip(second, lower, upper)
        f2 = np.full(len(a), np.inf)
        ok2 = np.zeros(len(a), dtype=bool)
        if need.any():
            f2[need], ok2[need] = evaluate(second[need], a[need])

        new_x = xr.copy()
        new_f = fxr.copy()
        replace = accept.copy()
        shrink = np.zeros(len(a), dtype=bool)
        # expansion
        e_good = expand & ok2 & (f2 < fxr)
        new_x[e_good], new_f[e_good] = second[e_good], f2[e_good]
        replace |= expand & ok2
        # outside contraction
        o_good = outside & ok2 & (f2 <= fxr)
        new_x[o_good], new_f[o_good] = second[o_good], f2[o_good]
        replace |= o_good
        shrink |= outside & ok2 & ~o_good
        # inside contraction
        i_good = inside & ok2 & (f2 < fm1)
        new_x[i_good], new_f[i_good] = second[i_good], f2[i_good]
        replace |= i_good
        shrink |= inside & ok2 & ~i_good

        sim[a[replace], -1] = new_x[replace]
        fsim[a[replace], -1] = new_f[replace]

        s = a[shrink]
        if len(s):
            for j in range(1, N + 1):
                live = active[s]
                sj = s[live]
                if len(sj) == 0:
                    break
                sim[sj, j] = np.clip(sim[sj, 0] + SIGMA * (sim[sj, j] - sim[sj, 0]), lower, upper)
                fj, okj = evaluate(sim[sj, j], sj)
                fsim[sj[okj], j] = fj[okj]

        finished = a[active[a]]
        iterations[finished] += 1
        sort_rows(a)

    return LockstepResult(sim[:, 0].copy(), np.min(fsim, axis=1), iterations, fcalls, status, n_calls)
