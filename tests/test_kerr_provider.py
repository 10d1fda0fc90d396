"""The built-in Kerr table provider (qnmfits_b200/kerr.py): Leaver's
continued fraction + the Cook-Zalutskiy spectral angular solver, pinned to literature values
and to the one Kerr number the reference itself prints — omega_220(chi = 0.7) in
examples/working_with_qnms.ipynb (golden G2)."""
import numpy as np
import pytest

from qnmfits_b200 import kerr

# Leaver (1985) table 1 / Berti, Cardoso & Starinets (2009) appendix, M = 1, s = -2
SCHWARZSCHILD = {
    2: [0.37367168 - 0.08896232j, 0.34671099 - 0.27391488j, 0.30105345 - 0.47827698j,
        0.25150496 - 0.70514820j, 0.20751458 - 0.94684489j, 0.16929940 - 1.19560805j],
    3: [0.59944329 - 0.09270305j, 0.58264380 - 0.28129811j, 0.55168490 - 0.47909275j],
    4: [0.80917838 - 0.09416396j, 0.79663153 - 0.28433435j],
}


def test_schwarzschild_overtones_match_the_literature():
    for ell, values in SCHWARZSCHILD.items():
        got = kerr.schwarzschild_omegas(-2, ell, len(values) - 1)
        for n, (a, b) in enumerate(zip(got, values)):
            assert abs(a - b) < 2e-8, (ell, n, a, b)
    # m does not matter without spin
    w0 = kerr.solve_mode(-2, 2, 1, 0, 0.0, 0.37 - 0.09j)[0]
    assert abs(w0 - SCHWARZSCHILD[2][0]) < 2e-8


def test_l2_overtones_beyond_the_algebraically_special_frequency():
    """l = 2: overtone 8 of Schwarzschild is the algebraically special frequency -2i (M = 1), which the
    continued fraction cannot reach; the ladder continues below it with the index of Leaver's table
    (1985, table 1, 2M omega: n = 10 (0.126527, -4.605289), 11 (0.153107, -5.121653),
    12 (0.165196, -5.630885) in his counting from 1) and of the `qnm` package.  Corotating Kerr
    sequences of those overtones follow to chi = 0.99; the multiplets of n = 8 are not attempted."""
    got = kerr.schwarzschild_omegas(-2, 2, 11)
    assert got[8] == -2j
    for n, ref in ((9, 0.126527 - 4.605289j), (10, 0.153107 - 5.121653j), (11, 0.165196 - 5.630885j)):
        assert abs(2 * got[n] - ref) < 2e-6, (n, got[n])
    seq = kerr.modes_cache(-2, 2, 2, 9)
    assert seq.a[0] == 0.0 and abs(seq.a[-1] - 0.99) < 1e-12
    assert abs(seq.omega[0] - got[9]) < 1e-10
    assert seq.omega[-1].real > 0.8 and np.all(seq.omega.real > 0.06)    # away from the imaginary axis, up to 0.99
    assert np.all(seq.omega.imag < 0) and np.max(np.abs(np.diff(seq.omega))) < 0.05
    # the root really is the n = 9 one: it solves the radial equation in ANOTHER inversion too
    w_check = kerr.solve_mode(-2, 2, 2, 7, float(seq.a[100]), seq.omega[100])[0]
    assert abs(w_check - seq.omega[100]) < 1e-9
    with pytest.raises(NotImplementedError):
        kerr.modes_cache(-2, 2, 2, 8)


def test_notebook_value_and_kerr_literature():
    seq = kerr.modes_cache(-2, 2, 2, 0)
    assert seq.a[0] == 0.0 and abs(seq.a[-1] - 0.99) < 1e-12 and np.all(np.diff(seq.a) > 0)
    at = lambda a: seq.omega[int(np.argmin(np.abs(seq.a - a)))]       # noqa: E731  (grid points)
    assert abs(at(0.7) - (0.53260024 - 0.08079287j)) < 1e-8            # the reference's notebook (G2)
    assert abs(at(0.98) - (0.8254 - 0.0386j)) < 1e-4                   # Berti et al. tables
    assert abs(at(0.99) - (0.8709 - 0.0294j)) < 1e-4
    assert abs(at(0.5) - (0.4641 - 0.0856j)) < 1e-4
    # counter-rotating partner and an overtone of another multipole
    assert abs(kerr.modes_cache(-2, 2, -2, 0).omega[70] - (0.3098 - 0.0887j)) < 1e-4
    # the l = m = 2, n = 5 sequence is the one that leaves the zero-damped family (Onozawa 1997)
    tails = [kerr.modes_cache(-2, 2, 2, n).omega[-1] for n in (4, 5, 6)]
    assert abs(tails[0].real - 0.8688) < 1e-3 and abs(tails[2].real - 0.8680) < 1e-3
    assert abs(tails[1] - (0.5064 - 0.7114j)) < 1e-3


def test_solution_satisfies_both_equations_and_is_truncation_independent():
    rng = np.random.default_rng(3)
    for key in [(-2, 2, 2, 0), (-2, 3, 2, 1), (-2, 4, -1, 2), (-2, 2, 0, 5)]:
        s, l, m, n = key
        seq = kerr.modes_cache(*key)
        for i in rng.integers(1, len(seq.a), size=3):
            a, w, A, C = float(seq.a[i]), seq.omega[i], seq.A[i], seq.C[i]
            # angular: M(c) C = A C, unit norm, real positive l' = l component
            l_min, c1, c2 = kerr._cos_matrices(s, m, kerr.L_MAX)
            ls = np.arange(l_min, kerr.L_MAX + 1)
            c = a * w
            mat = np.diag(ls * (ls + 1.0) - s * (s + 1.0)).astype(complex) - c * c * c2 + 2 * c * s * c1
            assert np.max(np.abs(mat @ C - A * C)) < 1e-10
            assert abs(np.linalg.norm(C) - 1) < 1e-13
            assert abs(C[l - l_min].imag) < 1e-13 and C[l - l_min].real > 0.5
            # radial: the same root with four times as many continued-fraction terms
            w4 = kerr.solve_mode(s, l, m, n, a, w, A, n_terms=400000)[0]
            assert abs(w4 - w) < 5e-11, (key, a)
    # no spin: no mixing, spherical separation constant
    seq = kerr.modes_cache(-2, 3, 2, 1)
    assert np.array_equal(np.abs(seq.C[0]) > 1e-14, np.arange(2, kerr.L_MAX + 1) == 3)
    assert abs(seq.A[0] - (3 * 4 - 2)) < 1e-12


def test_cos_matrix_elements_against_closed_forms():
    # <s l+1 m| cos |s l m> = sqrt(((l+1)^2 - m^2)((l+1)^2 - s^2) / ((2l+1)(2l+3))) / (l+1), and
    # <s l m| cos |s l m> = -m s / (l (l+1))
    s, m = -2, 1
    l_min, c1, c2 = kerr._cos_matrices(s, m, 10)
    for i, l in enumerate(range(l_min, 10)):
        up = np.sqrt(((l + 1) ** 2 - m * m) * ((l + 1) ** 2 - s * s) / ((2 * l + 1) * (2 * l + 3))) / (l + 1)
        assert abs(abs(c1[i + 1, i]) - up) < 1e-13 and abs(c1[i, i + 1] - c1[i + 1, i]) < 1e-13
        assert abs(c1[i, i] - (-m * s / (l * (l + 1.0)))) < 1e-13
    # cos^2 = cos . cos (exact away from the truncated edge of the basis)
    inner = slice(0, c1.shape[0] - 1)
    assert np.allclose(c2[inner, inner], (c1 @ c1)[inner, inner], atol=1e-12)


def test_spline_of_the_refined_grid_reproduces_direct_solutions():
    from scipy.interpolate import UnivariateSpline
    rng = np.random.default_rng(11)
    for key in [(-2, 2, 2, 0), (-2, 2, 2, 3), (-2, 3, 3, 0)]:
        seq = kerr.modes_cache(*key)
        re = UnivariateSpline(seq.a, seq.omega.real, s=0)
        im = UnivariateSpline(seq.a, seq.omega.imag, s=0)
        for a in list(rng.uniform(0.0, 0.99, 5)) + [0.9893]:
            guess = complex(re(a), im(a))
            assert abs(kerr.solve_mode(*key, float(a), guess)[0] - guess) < 3e-9, (key, a)


def test_fit_with_kerr_tables_recovers_an_injection(oracle_tables):
    """End to end on the host: the provider feeding the oracle (the reference's algorithm)."""
    from oracle import qnmfits_oracle as orc
    tables = orc.OracleTables(kerr.modes_cache)
    modes = [(2, 2, n, 1) for n in range(4)] + [(2, 2, 0, -1), (3, 2, 0, 1)]
    times = np.arange(-100, 1201) * 0.1
    rng = np.random.default_rng(0)
    C = rng.normal(size=len(modes)) + 1j * rng.normal(size=len(modes))
    omega = np.array(tables.omega_list(modes, 0.69, 0.95))
    assert np.all(omega.imag < 0) and omega[4].real < 0                 # mirror mode: -conj
    data = np.where(times >= 0, (C[None, :] * np.exp(-1j * omega[None, :] * times[:, None])).sum(axis=1), 0)
    fit = orc.ringdown_fit(tables, times, data, modes, 0.95, 0.69, 0.0, T=100)
    assert fit['mismatch'] < 1e-12
    assert np.max(np.abs(fit['C'] - C)) < 1e-6 * np.max(np.abs(C))
    with pytest.raises(NotImplementedError):
        kerr.modes_cache(-2, 2, 2, 8)


def test_against_the_qnm_package_when_it_is_installed():
    """Parity of the built-in solver with the ``qnm`` PyPI package (Stein 2019), the reference's
    actual table source (reference qnmfits/qnm.py:134): activates by itself wherever that package
    and its data are importable (it is absent from the build and GPU images, where parity with
    the PACKAGE stays unpinned — DESIGN.md section 7).  Frequencies to 1e-8, mixing coefficients
    to 1e-7 INCLUDING the phase convention (unit norm, ell' = ell component real positive)."""
    qnm_pkg = pytest.importorskip("qnm")
    try:
        probe = qnm_pkg.modes_cache(-2, 2, 2, 0)
    except Exception as exc:                       # the package needs a one-off qnm.download_data()
        pytest.skip(f"qnm is importable but its tables are not: {exc}")
    for l, m, n in ((2, 2, 0), (2, 2, 3), (2, -2, 1), (3, 2, 0), (4, 4, 1), (2, 0, 2)):
        theirs = qnm_pkg.modes_cache(-2, l, m, n)
        ours = kerr.modes_cache(-2, l, m, n)
        for a in (0.0, 0.3, 0.69, 0.9):
            w_t, A_t, C_t = theirs(a=a)
            k = int(np.argmin(np.abs(ours.a - a)))
            w_o, C_o = kerr.solve_mode(-2, l, m, n, float(a), ours.omega[k])[0], None
            assert abs(w_o - w_t) < 1e-8, (l, m, n, a, w_o, w_t)
            # mixing coefficients on the sequence's own grid point nearest to a
            w_g, A_g, C_g = theirs(a=float(ours.a[k]))
            ncol = min(len(C_g), ours.C.shape[1])
            np.testing.assert_allclose(ours.C[k, :ncol], np.asarray(C_g)[:ncol], rtol=0, atol=1e-7,
                                       err_msg=f"mixing coefficients of {(l, m, n)} at a = {ours.a[k]}")
    del probe
