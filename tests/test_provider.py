"""Host side of the boundary: label -> value resolution is bit-exact with the reference."""
import ast
import ctypes as C

import numpy as np
import pytest

import cases


def test_omega_and_mu_match_reference_fixture_bitwise(qf, golden):
    g = golden("provider")
    labels = [ast.literal_eval(s) for s in g["labels"]]
    for i, chi in enumerate(g["spins"]):
        got = np.array(qf.qnm.omega_list(labels, chi, float(g["Mf"])))
        assert np.array_equal(got, g["omega"][i]), (chi, got - g["omega"][i])
        mu = np.array([complex(v) for v in qf.qnm.mu_list([tuple(x) for x in g["mu_indices"]], chi)])
        assert np.array_equal(mu, g["mu"][i]), chi


def test_mu_is_int_zero_for_m_mismatch(qf):
    assert qf.qnm.mu(2, 2, 2, 1, 0, 1, 0.7) == 0 and type(qf.qnm.mu(2, 2, 2, 1, 0, 1, 0.7)) is int


def test_mirror_symmetry(qf):
    a = qf.qnm.omega(2, 2, 0, -1, 0.7)
    b = qf.qnm.omega(2, -2, 0, 1, 0.7)
    assert a == -np.conjugate(b)


def test_multiplet_reindexing_without_cook_data(qf):
    """Without the Cook-Zalutskiy files (2,2,9) and (2,2,10) resolve to the same
    sequence (reference qnm.py:128-134) -> exactly duplicated columns."""
    assert qf.qnm.omega(2, 2, 9, 1, 0.5) == qf.qnm.omega(2, 2, 10, 1, 0.5)
    assert qf.qnm.omega(2, 1, 11, 1, 0.5) != qf.qnm.omega(2, 1, 10, 1, 0.5)


def test_vectorised_tabulation_equals_scalar_calls(qf):
    chis = np.linspace(0.59, 0.79, 33)
    modes = [(2, 2, 0, 1), (2, 2, 3, -1), (2, 2, 0, 1, 3, 3, 0, 1)]
    table, ptr = qf.qnm.constituent_table(modes, chis)
    assert list(ptr) == [0, 1, 2, 4]
    for c, chi in enumerate(chis):
        parts = [qf.qnm.omega(2, 2, 0, 1, chi), qf.qnm.omega(2, 2, 3, -1, chi),
                 qf.qnm.omega(2, 2, 0, 1, chi), qf.qnm.omega(3, 3, 0, 1, chi)]
        assert np.array_equal(table[c], np.array(parts))


def _form_omega_like_device(table_row, ptr, j, inv_mf, delta_factor):
    """Python restatement of form_omega() in csrc/qnmfit_common.cuh (separately rounded
    products and sums)."""
    re, im = np.float64(0.0), np.float64(0.0)
    for p in range(ptr[j], ptr[j + 1]):
        re = re + np.float64(table_row[p].real) * inv_mf
        im = im + np.float64(table_row[p].imag) * inv_mf
    return complex(np.float64(delta_factor) * re, np.float64(delta_factor) * im)


def test_factored_frequency_formation_is_bit_exact(qf, golden):
    """omega_j = delta_j * sum_p(table[c,p] * (1/Mf)) reproduces the reference's
    delta_factor*np.array(qnm.omega_list(modes, chif, Mf)) to the last bit."""
    rng = np.random.default_rng(3)
    modes = [(2, 2, 0, 1), (2, 2, 5, 1), (3, 2, 1, -1), (2, 2, 0, 1, 2, 2, 0, 1),
             (2, 2, 0, 1, 3, 3, 0, 1, 2, 0, 1, -1)]
    deltas = np.array([0.0, 0.01, -0.02, 0.0, 0.3])
    chis = np.linspace(0.2, 0.9, 9)
    table, ptr = qf.qnm.constituent_table(modes, chis)
    for Mf in rng.uniform(0.5, 1.5, 40):
        inv = np.float64(1.0) / np.float64(Mf)
        for c, chi in enumerate(chis):
            want = (deltas + 1) * np.array(qf.qnm.omega_list(modes, chi, Mf))
            got = np.array([_form_omega_like_device(table[c], ptr, j, inv, deltas[j] + 1)
                            for j in range(len(modes))])
            assert np.array_equal(got, want), (Mf, chi)


def test_default_provider_falls_back_to_the_builtin_leaver_solver(qf):
    """Without the `qnm` PyPI package the provider serves Kerr tables from qnmfits_b200.kerr
    (with a notice) instead of failing: omega_220(0.7) is the value the reference's notebook
    prints."""
    from qnmfits_b200.qnm import set_table_provider, qnm as qnm_class
    from qnmfits_b200 import synthetic
    try:
        import qnm as _pypi  # noqa: F401
        pytest.skip("the qnm package is installed: the fallback is not reached")
    except ImportError:
        pass
    set_table_provider(None)
    try:
        import qnmfits_b200.qnm as mod
        mod._warned_fallback = False
        with pytest.warns(RuntimeWarning, match="built-in Leaver"):
            w = qnm_class().omega(2, 2, 0, 1, 0.7)
        assert abs(w - (0.53260024 - 0.08079287j)) < 1e-8
    finally:
        set_table_provider(synthetic.modes_cache)
