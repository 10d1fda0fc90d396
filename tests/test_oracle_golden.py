"""The oracle (oracle/qnmfits_oracle.py) against the golden fixtures produced by the
unmodified reference, and against the live reference when it is present."""
import warnings

import numpy as np
import pytest

from oracle import qnmfits_oracle as orc
import cases
from cases import rel_err


def test_oracle_cfg1_cases_match_reference_fixtures(qf, oracle_tables, golden):
    g = golden("cfg1")
    wl, cs = cases.cfg1_cases()
    assert np.array_equal(wl.times, g["times"]) and np.array_equal(wl.data, g["data"])
    for name, kw in cs.items():
        fit = orc.ringdown_fit(oracle_tables, wl.times, wl.data, **kw)
        assert np.array_equal(fit["frequencies"], g[name + "__frequencies"]), name
        assert np.array_equal(fit["model_times"], g[name + "__model_times"]), name
        assert int(fit["rank"]) == int(g[name + "__rank"]), name
        # same numpy, same LAPACK, same inputs: identical to the last bits
        np.testing.assert_allclose(fit["C"], g[name + "__C"], rtol=1e-13, atol=0, err_msg=name)
        np.testing.assert_allclose(fit["s"], g[name + "__s"], rtol=1e-13, err_msg=name)
        assert abs(fit["mismatch"] - float(g[name + "__mismatch"])) < 1e-15, name
        assert fit["residual"].shape == g[name + "__residual"].shape, name


def test_oracle_sweeps_match_reference_fixtures(qf, oracle_tables, golden):
    from qnmfits_b200 import workloads
    g2, g3 = golden("cfg2"), golden("cfg3")
    wl = workloads.config2(n_t0=40)
    mm = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, wl.modes, 0.95, 0.69, wl.t0_array)
    np.testing.assert_allclose(mm, g2["mismatch"], rtol=0, atol=1e-15)
    mmc = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, wl.modes[:4], 0.95, 0.69,
                                wl.t0_array[:10], t0_method="closest",
                                T_array=np.linspace(50, 80, 10))
    np.testing.assert_allclose(mmc, g2["mismatch_closest"], rtol=0, atol=1e-15)
    grid = orc.mismatch_M_chi_grid(oracle_tables, wl.times, wl.data, wl.modes, (0.85, 1.05),
                                   (0.59, 0.79), 0.0, T=100, res=12)
    np.testing.assert_allclose(grid, g3["grid"], rtol=0, atol=1e-15)


def test_oracle_multimode_matches_reference_fixtures(qf, oracle_tables, golden):
    g = golden("cfg4")
    wl = cases.cfg4_small()
    for lm in cases.MM_SPH:
        assert np.array_equal(wl.data[lm], g[f"data_{lm[0]}_{lm[1]}"])
    fit = orc.multimode_ringdown_fit(oracle_tables, wl.times, wl.data, cases.MM_MODES, 0.95, 0.69,
                                     5.0, T=80)
    np.testing.assert_allclose(fit["C"], g["C"], rtol=1e-12)
    assert abs(fit["mismatch"] - float(g["mismatch"])) < 1e-15
    for lm in cases.MM_SPH:
        np.testing.assert_allclose(fit["weighted_C"][lm], g[f"weighted_C_{lm[0]}_{lm[1]}"],
                                   rtol=1e-12, atol=1e-300)
    sweep = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, cases.MM_MODES, 0.95, 0.69,
                                  wl.t0_array, T_array=70)
    np.testing.assert_allclose(sweep, g["t0_sweep"], rtol=0, atol=1e-15)


def test_g1_injection_recovery(qf, oracle_tables, golden):
    """examples/correcting_measured_amplitude.ipynb: C = 1-1j recovered at t0 = 0 with zero
    mismatch, and C(t0=10) = C(0) exp(-i w 10): amplitudes are referenced to t0."""
    g = golden("g1")
    fit0 = orc.ringdown_fit(oracle_tables, g["times"], g["data"], [(2, 2, 0, 1)], 1, 0.7, 0)
    fit10 = orc.ringdown_fit(oracle_tables, g["times"], g["data"], [(2, 2, 0, 1)], 1, 0.7, 10)
    assert abs(fit0["C"][0] - (1 - 1j)) < 1e-13 and abs(fit0["mismatch"]) < 1e-14
    assert abs(fit10["C"][0] - (1 - 1j) * np.exp(-1j * g["omega"][0] * 10)) < 1e-13
    np.testing.assert_allclose(fit10["C"], g["C10"], rtol=1e-13)


def test_oracle_against_live_reference(qf, oracle_tables, reference):
    if reference is None:
        pytest.skip("/root/reference not present")
    from qnmfits_b200 import workloads
    wl = workloads.config1()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t0, T, method in ((0.0, 100, "geq"), (7.77, 55.5, "geq"), (7.77, 55.5, "closest")):
            a = reference.ringdown_fit(wl.times, wl.data, wl.modes, 0.9, 0.7, t0, method, T)
            b = orc.ringdown_fit(oracle_tables, wl.times, wl.data, wl.modes, 0.9, 0.7, t0, method, T)
            assert np.array_equal(a["frequencies"], b["frequencies"])
            assert np.array_equal(a["model_times"], b["model_times"])
            assert rel_err(b["C"], a["C"]) < 1e-13
            assert abs(a["mismatch"] - b["mismatch"]) < 1e-15


def test_oracle_omega_grid_and_epsilon_vs_reference_golden(golden, oracle_tables):
    """mismatch_omega_grid (incl. the 'closest' re-slicing quirk, reference
    qnmfits.py:1759-1768) and calculate_epsilon restated in the oracle vs the unmodified
    reference's outputs (tests/golden/make_golden_next.py)."""
    from qnmfits_b200 import workloads
    g = golden("next")
    wl = workloads.config1()
    m2 = wl.modes[:2]
    got = orc.mismatch_omega_grid(oracle_tables, wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1),
                                  5.0, T=80, res=7)
    np.testing.assert_allclose(got, g["omega_grid_geq"], rtol=0, atol=1e-15)
    got = orc.mismatch_omega_grid(oracle_tables, wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1),
                                  3.37, t0_method='closest', T=60, res=5)
    np.testing.assert_allclose(got, g["omega_grid_closest"], rtol=0, atol=1e-15)
    eps = orc.calculate_epsilon(oracle_tables, wl.times, wl.data, wl.modes[:4], 0.95, 0.69, 10.0)
    np.testing.assert_allclose(eps, g["eps_single"], rtol=0, atol=1e-12)
    wl4 = cases.cfg4_small()
    eps = orc.calculate_epsilon(oracle_tables, wl4.times, wl4.data, cases.MM_MODES, 0.95, 0.69, 5.0, T=80,
                                x0=[0.97, 0.65])
    np.testing.assert_allclose(eps, g["eps_multimode_x0"], rtol=0, atol=1e-12)


def test_oracle_dynamic_fits_vs_reference_golden(golden, oracle_tables):
    """Time-dependent Kerr spectrum (reference qnmfits.py:318-475, 676-911) restated in the
    oracle vs the unmodified reference's outputs (tests/golden/make_golden_dynamic.py)."""
    from qnmfits_b200 import workloads
    g = golden("dynamic")
    wl = workloads.config1()
    Mf_t, chi_t = cases.drift(wl.times)
    fit = orc.dynamic_ringdown_fit(oracle_tables, wl.times, wl.data, wl.modes[:5], Mf_t, chi_t, 2.0, T=70)
    assert np.array_equal(fit["frequencies"], g["single_frequencies"])
    assert rel_err(fit["C"], g["single_C"]) < 1e-12
    assert abs(fit["mismatch"] - float(g["single_mismatch"])) < 1e-15
    wl4 = cases.cfg4_small()
    Mf4, chi4 = cases.drift(wl4.times)
    fit = orc.dynamic_multimode_ringdown_fit(oracle_tables, wl4.times, wl4.data, cases.DYN_MODES, Mf4, chi4, 5.0,
                                             T=80, spherical_modes=cases.DYN_SPH)
    assert rel_err(fit["C"], g["multi_C"]) < 1e-12
    assert abs(fit["mismatch"] - float(g["multi_mismatch"])) < 1e-15
    np.testing.assert_allclose(fit["weighted_C"][cases.DYN_SPH[1]], g["multi_weighted_1"], rtol=1e-12)
