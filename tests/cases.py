"""Inputs of the golden cases (must match tests/golden/make_golden.py)."""
import numpy as np

from qnmfits_b200 import workloads


def cfg1_cases():
    wl = workloads.config1()
    m = wl.modes
    return wl, {
        "base": dict(modes=m, Mf=0.95, chif=0.69, t0=0.0),
        "offgrid_geq": dict(modes=m, Mf=0.93, chif=0.66, t0=3.37, T=77.7),
        "closest": dict(modes=m, Mf=0.95, chif=0.69, t0=3.37, T=77.7, t0_method="closest"),
        "delta_float": dict(modes=m[:4], Mf=0.95, chif=0.69, t0=10.0, delta=0.01),
        "delta_list": dict(modes=m[:3], Mf=0.95, chif=0.69, t0=10.0, delta=[0.0, 0.01, -0.02]),
        "quadratic": dict(modes=[(2, 2, 0, 1), (2, 2, 1, 1), (2, 2, 0, 1, 2, 2, 0, 1)], Mf=0.95,
                          chif=0.69, t0=15.0),
        "mirror": dict(modes=[(2, 2, 0, 1), (2, 2, 0, -1), (2, 2, 1, 1), (2, 2, 1, -1)], Mf=0.95,
                       chif=0.69, t0=5.0),
        "one_mode": dict(modes=[(2, 2, 0, 1)], Mf=0.95, chif=0.69, t0=20.0),
        "twelve": dict(modes=[(2, 2, n, 1) for n in range(12)], Mf=0.95, chif=0.69, t0=0.0),
        "duplicate_label": dict(modes=[(2, 2, n, 1) for n in (0, 1, 9, 10)], Mf=0.95, chif=0.69,
                                t0=5.0),
    }


#: tolerance on amplitudes per case: max(1e-8, 100 * cond * eps) relative to max|C|
#:; cond measured from the golden singular values.
def amp_tol(s):
    cond = float(s[0] / s[-1]) if s[-1] > 0 else np.inf
    return max(1e-8, 100.0 * cond * np.finfo(float).eps)


MM_SPH = [(2, 2), (3, 2), (4, 2), (2, 0), (2, -2)]
MM_MODES = [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1), (2, 2, 0, -1), (2, -2, 0, 1), (2, 0, 0, 1),
            (2, 0, 1, 1), (4, 2, 0, 1)]


def cfg4_small():
    return workloads.config4(n_t0=6, spherical=MM_SPH, modes=MM_MODES)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.abs(b)))


# dynamic (time-dependent spectrum) cases of tests/golden/make_golden_dynamic.py
DYN_SPH = [(2, 2), (3, 2), (4, 2)]
DYN_MODES = [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1), (2, 2, 0, -1), (4, 2, 0, 1)]


def drift(times):
    """Mf(t), chif(t) relaxing exponentially to (0.95, 0.69)."""
    x = np.exp(-np.clip(times, 0, None) / 15.0)
    return 0.95 - 0.03 * x, 0.69 - 0.05 * x
