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
