"""Host logic of the batched free-frequency search: the lock-step Nelder-Mead must
follow scipy's bounded Nelder-Mead (what the reference calls, qnmfits.py:2031-2038)
step for step when the objective returns the same floats."""
import numpy as np
import pytest
from scipy.optimize import minimize

from qnmfits_b200 import _neldermead as nm


def _objectives():
    def rosen(x, a):
        return (a - x[..., 0]) ** 2 + 30.0 * (x[..., 1] - x[..., 0] ** 2) ** 2

    def bowl(x, a):          # minimum on / outside the boundary for some a
        return (x[..., 0] - 2.5 * a) ** 2 + (x[..., 1] + 0.4 * a) ** 2 + 0.1 * np.sin(5 * x[..., 0])

    def flat(x, a):          # many ties: exercises the sort order
        return np.round(np.abs(x[..., 0] - a) + np.abs(x[..., 1] + 0.5), 3)
    return {"rosen": rosen, "bowl": bowl, "flat": flat}


@pytest.mark.parametrize("name", ["rosen", "bowl", "flat"])
@pytest.mark.parametrize("xatol,maxfun", [(1e-8, None), (1e-6, 37)])
def test_lockstep_follows_scipy(name, xatol, maxfun):
    f = _objectives()[name]
    rng = np.random.default_rng(11)
    B = 40
    a = rng.uniform(0.1, 1.0, B)
    x0 = np.column_stack([rng.uniform(-0.2, 2.2, B), rng.uniform(-1.1, 0.1, B)])
    x0[0] = [1.0, -0.5]
    x0[1] = [2.0, 0.0]          # on the upper bounds: the reflected initial simplex
    x0[2] = [0.0, -1.0]         # zero coordinate: zdelt
    bounds = [(0, 2), (-1, 0)]
    calls = []

    def fun(X, idx):
        calls.append(len(idx))
        assert np.all(np.diff(idx) > 0)
        return f(X, a[idx])

    kw = {} if maxfun is None else {"maxfun": maxfun}
    got = nm.minimize_lockstep(fun, x0, bounds, xatol=xatol, **kw)
    import warnings
    for b in range(B):
        opts = {"xatol": xatol, "disp": False}
        if maxfun is not None:
            opts["maxfev"] = maxfun
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = minimize(lambda x: float(f(x, a[b])), x0[b], method="Nelder-Mead", bounds=bounds, options=opts)
        if maxfun is None:
            assert np.array_equal(got.x[b], ref.x), (b, got.x[b], ref.x)
            assert got.fun[b] == ref.fun and got.nit[b] == ref.nit and got.nfev[b] == ref.nfev, b
            assert got.status[b] == ref.status
        else:
            # budget exhausted mid-iteration: scipy abandons the iteration by exception
            assert got.nfev[b] <= maxfun and got.status[b] == ref.status
            assert got.fun[b] <= ref.fun + 1e-12 or abs(got.fun[b] - ref.fun) < 1e-3
    assert got.n_calls == len(calls) and max(calls) <= B


def test_initial_simplex_matches_scipy():
    lower, upper = np.array([0.0, -1.0]), np.array([2.0, 0.0])
    x0 = np.array([[1.0, -0.5], [1.99, -0.001], [0.0, 0.0], [3.0, -2.0]])
    sim = nm.initial_simplex(x0, lower, upper)
    for b in range(len(x0)):
        seen = []
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            minimize(lambda x: seen.append(np.array(x)) or 0.0, x0[b], method="Nelder-Mead",
                     bounds=list(zip(lower, upper)), options={"maxfev": 3})
        np.testing.assert_array_equal(sim[b], np.array(seen[:3]))


def _same(got, want):
    for key in ("x", "fun", "nit", "nfev", "status"):
        assert np.array_equal(getattr(got, key), getattr(want, key), equal_nan=True), key
    assert got.n_calls == want.n_calls


@pytest.mark.parametrize("N", [1, 2, 3, 5])
@pytest.mark.parametrize("limits", [{}, {"maxfun": 23}, {"maxiter": 9}, {"maxfun": 3}, {"maxiter": 40, "maxfun": 61}])
def test_native_state_machine_equals_the_numpy_form(N, limits):
    """``minimize_lockstep`` (qnmfit_nm_* of libqnmfit.so, host code) against its specification
    ``minimize_lockstep_numpy``: same requests in the same order, same results to the bit —
    smooth objectives, plateaus (ties in the simplex order: numpy's argsort decides), NaNs and
    infinities, budgets that end a search before, inside and after an iteration."""
    rng = np.random.default_rng(100 + N)
    B = 64
    centre = rng.uniform(-0.5, 1.5, (B, N))
    bounds = [(-1.0 + 0.1 * k, 1.0 + 0.2 * k) for k in range(N)]
    x0 = rng.uniform(-1.2, 1.6, (B, N))
    x0[0] = 0.0                                   # zero coordinates
    x0[1] = [b[1] for b in bounds]                # on the upper bounds

    def smooth(X, idx):
        d = X - centre[idx]
        return np.sum(d * d * (1.0 + np.arange(N)), axis=1) + 0.05 * np.sin(7 * X[:, 0])

    def plateau(X, idx):
        return np.round(np.sum(np.abs(X - centre[idx]), axis=1), 2)

    def rough(X, idx):
        f = smooth(X, idx)
        f[(idx % 7 == 3) & (X[:, 0] > 0.5)] = np.nan
        f[(idx % 5 == 1) & (X[:, 0] < -0.2)] = np.inf
        return f

    for objective in (smooth, plateau, rough):
        logs = []

        def make(tag):
            def fun(X, idx):
                logs.append((tag, X.copy(), idx.copy()))
                assert X.dtype == np.float64 and idx.dtype == np.int64 and np.all(np.diff(idx) > 0)
                return objective(X, idx)
            return fun
        got = nm.minimize_lockstep(make("c"), x0, bounds, xatol=1e-6, fatol=1e-5, **limits)
        want = nm.minimize_lockstep_numpy(make("np"), x0, bounds, xatol=1e-6, fatol=1e-5, **limits)
        _same(got, want)
        mine = [entry for entry in logs if entry[0] == "c"]
        spec = [entry for entry in logs if entry[0] == "np"]
        assert len(mine) == len(spec) == got.n_calls
        for (_, Xc, ic), (_, Xn, i_n) in zip(mine, spec):
            assert np.array_equal(ic, i_n) and np.array_equal(Xc, Xn)


def test_native_state_machine_edge_cases():
    lib_err = pytest.raises(Exception)
    f = lambda X, idx: np.sum(X * X, axis=1)  # noqa: E731
    # no problems at all; a single problem given as a flat start point
    empty = nm.minimize_lockstep(f, np.zeros((0, 2)), [(0, 1), (0, 1)])
    assert empty.x.shape == (0, 2) and empty.n_calls == 0
    one = nm.minimize_lockstep(f, [0.5, 0.25], [(-1, 1), (-1, 1)], xatol=1e-9, fatol=1e-12)
    ref = nm.minimize_lockstep_numpy(f, [0.5, 0.25], [(-1, 1), (-1, 1)], xatol=1e-9, fatol=1e-12)
    _same(one, ref)
    assert np.all(np.abs(one.x) < 1e-8) and one.status[0] == 0
    # an objective that returns the wrong number of values is an error, not a crash
    with lib_err:
        nm.minimize_lockstep(lambda X, idx: np.zeros(len(idx) + 1), np.zeros((3, 2)), [(0, 1), (0, 1)])
    with pytest.raises(ValueError):
        nm.minimize_lockstep(f, np.zeros((3, 2)), [(0, 1)])
    # beyond QNMFIT_NM_MAX_VARS variables the numpy form takes over
    from qnmfits_b200 import _cabi
    N = _cabi.NM_MAX_VARS + 1
    big = nm.minimize_lockstep(f, np.full((2, N), 0.3), [(-1, 1)] * N, maxfun=N + 5)
    assert big.x.shape == (2, N) and np.all(big.nfev <= N + 5)


def test_native_state_machine_random_configurations():
    """Thirty random problems (1 - 6 variables, 1 - 40 searches, random bounds, tolerances and
    budgets, objectives quantised to a random grid so that ties are frequent): the C++ form and
    the numpy form agree to the bit."""
    rng = np.random.default_rng(2024)
    for trial in range(30):
        N = int(rng.integers(1, 7))
        B = int(rng.integers(1, 41))
        lo = rng.uniform(-2, 0, N)
        hi = lo + rng.uniform(0.5, 3, N)
        bounds = list(zip(lo, hi))
        x0 = rng.uniform(lo - 0.3, hi + 0.3, (B, N))
        centre = rng.uniform(lo - 0.5, hi + 0.5, (B, N))
        scale = rng.uniform(0.2, 5.0, N)
        quantum = float(rng.choice([0.0, 1e-3, 1e-2, 0.1]))
        limits = [{}, {"maxfun": int(rng.integers(1, 60))}, {"maxiter": int(rng.integers(1, 30))}][trial % 3]

        def fun(X, idx):
            f = np.sum(scale * (X - centre[idx]) ** 2, axis=1)
            return np.round(f / quantum) * quantum if quantum else f
        kw = dict(xatol=float(rng.choice([1e-8, 1e-4, 1e-2])), fatol=float(rng.choice([1e-10, 1e-4, 1e-1])), **limits)
        _same(nm.minimize_lockstep(fun, x0, bounds, **kw), nm.minimize_lockstep_numpy(fun, x0, bounds, **kw))


def test_native_state_machine_survives_a_broken_order_callback():
    """The `order` callback is only trusted when it returns permutations: garbage leaves the
    simplices of that step unsorted (a worse search, not a crash or an out-of-range access);
    without a callback ties are resolved stably."""
    import ctypes as C
    from qnmfits_b200 import _cabi
    lib = _cabi.load_library()

    @_cabi.NM_ORDER_FN
    def garbage(values, n_rows, n, order, _user):
        for i in range(n_rows * n):
            order[i] = 7 if i % 2 else -3

    for callback in (garbage, _cabi.NM_ORDER_FN()):
        B, N = 5, 2
        x0 = np.tile([0.3, 0.6], (B, 1))
        lower, upper = np.zeros(N), np.ones(N)
        handle = C.c_void_p()
        assert lib.qnmfit_nm_create(B, N, x0.ctypes.data, lower.ctypes.data, upper.ctypes.data, 1e-4, 1e-4,
                                    400.0, 400.0, callback, None, C.byref(handle)) == 0
        X, idx = np.empty((B, N)), np.empty(B, dtype=np.int64)
        f, steps = None, 0
        while True:
            n = lib.qnmfit_nm_step(handle, None if f is None else f.ctypes.data, X.ctypes.data, idx.ctypes.data)
            assert 0 <= n <= B
            if n == 0:
                break
            assert np.all((X[:n] >= 0) & (X[:n] <= 1)) and np.all(np.diff(idx[:n]) > 0)
            f = np.round(np.sum((X[:n] - 0.5) ** 2, axis=1), 2)      # plateaus: ties at every step
            steps += 1
            assert steps < 2000
        x, nfev = np.empty((B, N)), np.empty(B, dtype=np.int64)
        assert lib.qnmfit_nm_result(handle, x.ctypes.data, None, None, nfev.ctypes.data, None, None) == 0
        assert np.all((x >= 0) & (x <= 1)) and np.all(nfev <= 400)
        assert lib.qnmfit_nm_destroy(handle) == 0
    # argument errors are reported, not dereferenced
    assert lib.qnmfit_nm_create(1, 0, x0.ctypes.data, lower.ctypes.data, upper.ctypes.data, 1e-4, 1e-4, 1.0, 1.0,
                                _cabi.NM_ORDER_FN(), None, C.byref(handle)) == -2
    assert lib.qnmfit_nm_step(None, None, X.ctypes.data, idx.ctypes.data) == -1
