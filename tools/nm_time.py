#!/usr/bin/env python
"""Developer tool (CPU only): host cost per optimiser step of the lock-step Nelder-Mead, C++ form
(qnmfit_nm_* in libqnmfit.so) against its numpy specification, with a cheap numpy objective in
place of the device fits; checks that the two return identical results."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qnmfits_b200 import _neldermead as nm  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rng = np.random.default_rng(0)
    centre = np.stack([rng.uniform(0.3, 0.8, B), rng.uniform(-0.3, -0.05, B)], 1)
    spent = [0.0, 0]

    def fun(X, idx):
        t = time.perf_counter()
        d = X - centre[idx]
        f = d[:, 0] ** 2 * 30 + d[:, 1] ** 2 * 80 + 0.5 * d[:, 0] * d[:, 1]
        spent[0] += time.perf_counter() - t
        spent[1] += 1
        return f

    x0 = np.tile([1.0, -0.5], (B, 1))       # the reference's start point and bounds (qnmfits.py:2031-2038)
    results = {}
    for form in (nm.minimize_lockstep, nm.minimize_lockstep_numpy):
        best = None
        for _ in range(3):
            spent[:] = [0.0, 0]
            t = time.perf_counter()
            res = form(fun, x0, [(0, 2), (-1, 0)], xatol=1e-8)
            total = time.perf_counter() - t
            host = (total - spent[0]) / spent[1] * 1e6
            best = host if best is None else min(best, host)
        results[form.__name__] = res
        print(f"{form.__name__:26s} B = {B}: {res.n_calls} steps, nit <= {int(res.nit.max())}, "
              f"host bookkeeping {best:8.1f} us per step")
    a, b = results.values()
    for key in ("x", "fun", "nit", "nfev", "status"):
        assert np.array_equal(getattr(a, key), getattr(b, key), equal_nan=True), key
    print("identical results")


if __name__ == "__main__":
    main()
