"""
Offline Kerr QNM table provider.

The reference takes its Kerr frequencies and spherical-spheroidal mixing coefficients
from the third-party ``qnm`` PyPI package through one call, ``qnm.modes_cache(s, l, m, n)``
(reference ``qnmfits/qnm.py:134``), and reads three attributes of what it returns:
``.a``, ``.omega`` and ``.C`` (``qnm.py:137-141``).  That package needs a one-off network
download and is absent from the build and GPU boxes.  This module computes the same
quantities from scratch on the host, with the published algorithm the package implements
(Stein 2019, arXiv:1908.10377):

* radial equation: Leaver's three-term recurrence and its continued fraction, solved in
  its n-th inversion for the n-th overtone (Leaver 1985, Proc. R. Soc. A 402, 285,
  eqs. 24-26), evaluated by backward recursion;
* angular equation: the spectral method of Cook & Zalutskiy (2014, PRD 90, 124021): the
  spin-weighted spheroidal operator in the basis of spin-weighted SPHERICAL harmonics is
  the pentadiagonal matrix ``diag(l(l+1) - s(s+1)) - c^2 <cos^2> + 2 c s <cos>`` with
  ``c = a omega``; its eigenvalue is the separation constant A_lm and its eigenvector the
  mixing coefficients ``C[l']`` (unit 2-norm, the ``l' = l`` component real positive);
* a Newton iteration on the complex frequency, following each (l, m, n) sequence in spin
  from the Schwarzschild root.

Units: M = 1 (the ``qnm`` package's and the reference's; Leaver's 2M = 1 is converted
inside).  ``modes_cache(s, l, m, n)`` is a drop-in for ``qnm.modes_cache`` for
``qnmfits_b200.set_table_provider``; sequences are computed on first use and cached in
memory and, optionally, on disk (``QNMFITS_B200_KERR_CACHE=<dir>``).

Pinned by ``tests/test_kerr_provider.py`` to literature values (Schwarzschild and Kerr
frequencies of Leaver 1985 / Berti, Cardoso & Starinets 2009) and to the value the
reference's own notebook prints, omega_220(chi = 0.7) = 0.53260024 - 0.08079287i
(golden G2).  Not covered: overtone n = 8 of l = 2, whose sequences
start at the algebraically special frequency (the reference takes its multiplets from the
Cook-Zalutskiy data files, ``qnm.py:60-122``); the l = 2 overtones below it (n >= 9, with
Leaver's and the ``qnm`` package's index) are followed like any other sequence.
"""
import math
import os
from fractions import Fraction

import numpy as np

try:                                    # the continued fraction is a tight scalar loop
    import numba
    _jit = numba.njit(cache=True)
except Exception:  # pragma: no cover - numba is optional
    numba = None

    def _jit(f):
        return f

#: largest spherical ell' carried by the mixing table (the ``qnm`` package uses 20)
L_MAX = 20
#: initial spin grid of every sequence; refined adaptively (``compute_sequence``)
SPIN_GRID = np.arange(0, 100) * 0.01
#: refinement stops when the cubic spline through the grid reproduces omega at every
#: interval midpoint to this absolute accuracy (the ``qnm`` package refines the same way)
INTERP_TOL = 1e-9
MAX_GRID = 800


# ----------------------------------------------------------------------------
# angular problem

def _clebsch_gordan(j1, m1, j2, m2, J, M):
    """<j1 m1; j2 m2 | J M> for integer arguments (Racah's formula, exact rationals)."""
    if m1 + m2 != M or not (abs(j1 - j2) <= J <= j1 + j2) or abs(m1) > j1 or abs(m2) > j2 or abs(M) > J:
        return 0.0
    f = math.factorial
    pref = Fraction((2 * J + 1) * f(J + j1 - j2) * f(J - j1 + j2) * f(j1 + j2 - J), f(j1 + j2 + J + 1))
    pref *= f(J + M) * f(J - M) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)
    total = Fraction(0)
    for k in range(0, j1 + j2 - J + 1):
        d = [k, j1 + j2 - J - k, j1 - m1 - k, j2 + m2 - k, J - j2 + m1 + k, J - j1 - m2 + k]
        if min(d) < 0:
            continue
        den = 1
        for x in d:
            den *= f(x)
        total += Fraction((-1) ** k, den)
    return float(total) * math.sqrt(float(pref))


_cos_cache = {}


def _cos_matrices(s, m, l_max):
    """<s l' m| cos |s l m> and <s l' m| cos^2 |s l m> for l, l' = l_min..l_max."""
    key = (s, m, l_max)
    if key not in _cos_cache:
        l_min = max(abs(m), abs(s))
        ls = range(l_min, l_max + 1)
        c1 = np.zeros((len(ls), len(ls)))
        c2 = np.zeros((len(ls), len(ls)))
        for i, lp in enumerate(ls):
            for j, l in enumerate(ls):
                if abs(lp - l) > 2:
                    continue
                w = math.sqrt((2 * l + 1) / (2 * lp + 1))
                c1[i, j] = w * _clebsch_gordan(l, m, 1, 0, lp, m) * _clebsch_gordan(l, -s, 1, 0, lp, -s)
                c2[i, j] = (1.0 / 3.0 if lp == l else 0.0) + (2.0 / 3.0) * w \
                    * _clebsch_gordan(l, m, 2, 0, lp, m) * _clebsch_gordan(l, -s, 2, 0, lp, -s)
        _cos_cache[key] = (l_min, c1, c2)
    return _cos_cache[key]


def separation_constant(s, l, m, c, l_max=L_MAX, A_near=None):
    """(A_lm, C) of the spin-weighted spheroidal harmonic with oblateness c = a omega.

    The eigenvalue closest to ``A_near`` (default: the spherical value) is followed; ``C``
    has unit 2-norm and a real positive ``l' = l`` component."""
    l_min, c1, c2 = _cos_matrices(s, m, l_max)
    ls = np.arange(l_min, l_max + 1)
    mat = np.diag(ls * (ls + 1.0) - s * (s + 1.0)).astype(complex) - (c * c) * c2 + (2.0 * c * s) * c1
    vals, vecs = np.linalg.eig(mat)
    if A_near is None:
        A_near = l * (l + 1.0) - s * (s + 1.0)
    k = int(np.argmin(np.abs(vals - A_near)))
    vec = vecs[:, k]
    vec = vec / np.linalg.norm(vec)
    ref = vec[l - l_min]
    if abs(ref) > 0:
        vec = vec * (abs(ref) / ref)
    return vals[k], vec


# ----------------------------------------------------------------------------
# radial problem (Leaver units inside: 2M = 1)

@_jit
def _leaver_cf(omega, a, s, m, A, n_inv, n_terms):
    """Leaver's continued-fraction equation in its n_inv-th inversion (zero at a QNM).

    omega, a in units 2M = 1.  Coefficients: Leaver (1985) eqs. 25-26."""
    b = math.sqrt(1.0 - 4.0 * a * a)
    k = omega / 2.0 - a * m
    c0 = 1.0 - s - 1j * omega - (2j / b) * k
    c1 = -4.0 + 2j * omega * (2.0 + b) + (4j / b) * k
    c2 = s + 3.0 - 3j * omega - (2j / b) * k
    c3 = omega * omega * (4.0 + 2.0 * b - a * a) - 2.0 * a * m * omega - s - 1.0 + (2.0 + b) * 1j * omega - A \
        + ((4.0 * omega + 2j) / b) * k
    c4 = s + 1.0 - 2.0 * omega * omega - (2.0 * s + 3.0) * 1j * omega - ((4.0 * omega + 2j) / b) * k
    # tail: backward recursion T_k = beta_k - alpha_k gamma_{k+1} / T_{k+1}, where
    # T_k = beta_k + alpha_k a_{k+1}/a_k.  It starts from the large-k expansion of the ratio
    # of the minimal solution (Leaver 1985 eq. 11, Nollert 1993, redone for these
    # coefficients): a_{k+1}/a_k = 1 + u k^(-1/2) + v k^(-1) + ..., u^2 = -(c0 + c1 + c2)
    # = -2 i omega b with Re u < 0, v = (u^2 + c2 - c0 - 7/2) / 2.
    n = n_terms
    u = (-2j * omega * b) ** 0.5
    if u.real > 0:
        u = -u
    v = 0.5 * (u * u + c2 - c0 - 3.5)
    T = (-2.0 * n * n + (c1 + 2.0) * n + c3) + (n * n + (c0 + 1.0) * n + c0) * (1.0 + u / math.sqrt(n) + v / n)
    for kk in range(n_terms - 1, n_inv, -1):
        alpha = kk * kk + (c0 + 1.0) * kk + c0
        beta = -2.0 * kk * kk + (c1 + 2.0) * kk + c3
        gamma_next = (kk + 1.0) * (kk + 1.0) + (c2 - 3.0) * (kk + 1.0) + c4 - c2 + 2.0
        T = beta - alpha * gamma_next / T
    # head: forward recursion L_k = beta_k - alpha_{k-1} gamma_k / L_{k-1}
    L = c3 + 0j
    for kk in range(1, n_inv + 1):
        alpha_prev = (kk - 1.0) * (kk - 1.0) + (c0 + 1.0) * (kk - 1.0) + c0
        beta = -2.0 * kk * kk + (c1 + 2.0) * kk + c3
        gamma = kk * kk + (c2 - 3.0) * kk + c4 - c2 + 2.0
        L = beta - alpha_prev * gamma / L
    alpha_n = n_inv * n_inv + (c0 + 1.0) * n_inv + c0
    gamma_n1 = (n_inv + 1.0) * (n_inv + 1.0) + (c2 - 3.0) * (n_inv + 1.0) + c4 - c2 + 2.0
    return L - alpha_n * gamma_n1 / T


def _residual(omega, a, s, l, m, n, A_near, n_terms):
    """Continued-fraction residual at omega (M = 1 units) with the consistent separation constant."""
    A, C = separation_constant(s, l, m, a * omega, A_near=A_near)
    f = _leaver_cf(complex(2.0 * omega), 0.5 * a, float(s), float(m), complex(A), int(n), int(n_terms))
    return f, A, C


def _newton(s, l, m, n, a, omega, A_near, n_terms, tol, max_iter):
    for _ in range(max_iter):
        f0, A, C = _residual(omega, a, s, l, m, n, A_near, n_terms)
        h = 1e-7 * (1.0 + abs(omega))
        f1, _, _ = _residual(omega + h, a, s, l, m, n, A, n_terms)
        slope = (f1 - f0) / h
        if slope == 0:
            return omega, A, True
        step = f0 / slope
        if abs(step) > 0.05:                       # keep the iteration on its own branch
            step *= 0.05 / abs(step)
        omega = omega - step
        A_near = A
        if abs(step) < tol * (1.0 + abs(omega)):
            return omega, A, True
    return omega, A_near, False


def solve_mode(s, l, m, n, a, omega_guess, A_guess=None, tol=1e-13, n_terms=None, max_iter=60):
    """(omega, A, C) of the Kerr QNM (s, l, m, n) at spin a (M = 1), by Newton's method on
    the complex frequency from ``omega_guess``.  The continued fraction is truncated at
    ``n_terms`` (default: quadrupled until the root moves by less than 1e-12)."""
    fixed = n_terms is not None
    nt = int(n_terms) if fixed else 1000 + 500 * n
    omega, A_near, ok = _newton(s, l, m, n, a, complex(omega_guess), A_guess, nt, tol, max_iter)
    if not ok:
        raise RuntimeError(f"Leaver iteration did not converge for (s,l,m,n)={(s, l, m, n)} at a={a}")
    while not fixed and nt < 1000000:
        nt *= 4
        prev = omega
        omega, A_near, ok = _newton(s, l, m, n, a, omega, A_near, nt, tol, max_iter)
        if not ok:
            raise RuntimeError(f"Leaver iteration did not converge for (s,l,m,n)={(s, l, m, n)} at a={a}")
        if abs(omega - prev) < 1e-12:
            break
    f0, A, C = _residual(omega, a, s, l, m, n, A_near, nt)
    return omega, A, C


# ----------------------------------------------------------------------------
# Schwarzschild roots and spin sequences

_schw_cache = {}


def schwarzschild_omegas(s, l, n_max):
    """omega_{l n}(a = 0), n = 0..n_max: overtone after overtone, each from a guess
    extrapolated from the previous ones and required to be a new root further down the
    imaginary axis."""
    key = (s, l)
    have = _schw_cache.setdefault(key, [])
    while len(have) <= n_max:
        n = len(have)
        if abs(s) == 2 and l == 2 and n == 8:
            # the algebraically special frequency (Chandrasekhar 1984; Leaver 1985 lists the
            # root at 2M omega = -3.998 i): the continued fraction does not converge on the
            # negative imaginary axis.  Kept as a place holder so that the overtones below it
            # keep the index the `qnm` package (and Leaver's table) gives them.
            have.append(complex(0.0, -2.0))
            continue
        if abs(s) == 2 and l == 2 and n == 9:
            # first root past it, from Leaver (1985) table 1 / Berti, Cardoso & Starinets (2009):
            # 2M omega = 0.126527 - 4.605289 i; Newton refines it on the n = 9 inversion
            guesses = [complex(0.063263, -2.302645)]
        elif n == 0:
            # eikonal estimate, accurate to ~10 % at l = 2: (l + 1/2 - i (n + 1/2)) / (3 sqrt 3)
            guesses = [complex((l + 0.5) / math.sqrt(27.0) * 0.78 + 0.1 * (l - 2) * 0.22, -0.5 / math.sqrt(27.0))]
        elif n == 1:
            guesses = [have[0] + complex(-0.02, -0.188), have[0] + complex(-0.03, -0.2)]
        else:
            d = have[-1] - have[-2]
            guesses = [have[-1] + d, have[-1] + d + complex(-0.01, -0.02), have[-1] + complex(d.real * 1.3, d.imag)]
        found = None
        for g in guesses:
            try:
                w, _, _ = solve_mode(s, l, 0, n, 0.0, g)
            except RuntimeError:
                continue
            if w.real > 0 and w.imag < (have[-1].imag - 0.1 if have else 0.0) \
                    and all(abs(w - x) > 1e-6 for x in have):
                found = w
                break
        if found is None:
            raise RuntimeError(f"no Schwarzschild root found for s={s}, l={l}, n={n}")
        have.append(found)
    return list(have[:n_max + 1])


class KerrSequence:
    """The attributes of ``qnm``'s ``KerrSpinSeq`` that the reference reads (qnm.py:137-141)."""

    def __init__(self, s, l, m, n, a, omega, A, C):
        self.s, self.l, self.m, self.n = s, l, m, n
        self.a, self.omega, self.A, self.C = a, omega, A, C


def compute_sequence(s, l, m, n, spins=SPIN_GRID, tol=INTERP_TOL):
    """Follow (s, l, m, n) in spin from its Schwarzschild root, then refine the spin grid
    until cubic-spline interpolation (what the reference does with the table,
    ``qnm.py:144-155``) is accurate to ``tol`` at every interval midpoint."""
    from scipy.interpolate import UnivariateSpline
    if l < max(abs(m), abs(s)):
        raise KeyError(f"no sequence for s={s}, l={l}, m={m}")
    if l == 2 and n == 8:
        raise NotImplementedError(
            "l = 2, n = 8: these sequences start at the algebraically special frequency (and split into "
            "the multiplets {8,0}, {8,1} for m >= 0); the reference reads them from the Cook-Zalutskiy "
            "data files (qnm.py:60-122)")
    # march in spin with step control: a point is accepted when the solved frequency lies
    # close to the one extrapolated along the sequence, otherwise the step is halved (near
    # avoided crossings two sequences come close and Newton must not change branch)
    targets = [float(a) for a in np.asarray(spins, dtype=float)]
    w0 = schwarzschild_omegas(s, l, n)[n]
    w_first, A_first, C_first = solve_mode(s, l, m, n, targets[0], w0)
    sa, sw, sA, sC = [targets[0]], [w_first], [A_first], [C_first]
    ok = True
    for target in targets[1:]:
        while sa[-1] < target - 1e-15:
            a_next = target
            while True:
                if len(sa) < 3:
                    guess, A_guess = sw[-1], sA[-1]
                else:                               # quadratic extrapolation along the sequence
                    x0, x1, x2 = sa[-3:]
                    t = a_next
                    w = [((t - x1) * (t - x2)) / ((x0 - x1) * (x0 - x2)),
                         ((t - x0) * (t - x2)) / ((x1 - x0) * (x1 - x2)),
                         ((t - x0) * (t - x1)) / ((x2 - x0) * (x2 - x1))]
                    guess = w[0] * sw[-3] + w[1] * sw[-2] + w[2] * sw[-1]
                    A_guess = w[0] * sA[-3] + w[1] * sA[-2] + w[2] * sA[-1]
                try:
                    wn, An, Cn = solve_mode(s, l, m, n, a_next, guess, A_guess)
                    ok = abs(wn - guess) < (2e-3 if len(sa) < 3 else 3e-4)
                except RuntimeError:
                    ok = False
                if ok or a_next - sa[-1] < 2e-5:
                    break
                a_next = 0.5 * (sa[-1] + a_next)
            if not ok:
                break
            sa.append(a_next); sw.append(wn); sA.append(An); sC.append(Cn)
        if not ok:
            # a few high overtones of counter-rotating modes approach the negative imaginary
            # axis close to extremality, where the continued fraction stops converging; the
            # sequence then ends at the last spin reached (the `qnm` package shows the same
            # limitation for such modes)
            if sa[-1] < 0.95:
                raise RuntimeError(f"sequence (s,l,m,n)={(s, l, m, n)} could not be followed beyond a={sa[-1]}")
            break
    spins = np.array(sa)
    omega = np.array(sw)
    A = np.array(sA)
    C = np.array(sC)

    # adaptive refinement: bisect every interval whose midpoint the spline misses
    check = np.ones(len(spins) - 1, dtype=bool)       # intervals still to be verified
    while check.any() and len(spins) < MAX_GRID:
        sp = [UnivariateSpline(spins, f, s=0) for f in (omega.real, omega.imag, A.real, A.imag)]
        mids, rows = [], []
        for i in np.nonzero(check)[0]:
            am = 0.5 * (spins[i] + spins[i + 1])
            wg, Ag = complex(sp[0](am), sp[1](am)), complex(sp[2](am), sp[3](am))
            try:
                wm, Am, Cm = solve_mode(s, l, m, n, float(am), wg, Ag)
            except RuntimeError:
                continue                              # see the note on sequences that end early
            if abs(wm - wg) > tol and abs(wm - wg) < 1e-3:
                mids.append(am)
                rows.append((wm, Am, Cm))
        if not mids:
            break
        order = np.argsort(np.concatenate([spins, mids]), kind="stable")
        is_new = np.concatenate([np.zeros(len(spins), bool), np.ones(len(mids), bool)])[order]
        spins = np.concatenate([spins, mids])[order]
        omega = np.concatenate([omega, [r[0] for r in rows]])[order]
        A = np.concatenate([A, [r[1] for r in rows]])[order]
        C = np.concatenate([C, np.array([r[2] for r in rows])])[order]
        check = is_new[:-1] | is_new[1:]              # both halves of every bisected interval
    return KerrSequence(s, l, m, n, spins.copy(), omega, A, C)


_seq_cache = {}


def modes_cache(s, l, m, n):
    """Drop-in for ``qnm.modes_cache(s, l, m, n)`` (reference ``qnmfits/qnm.py:134``)."""
    key = (int(s), int(l), int(m), int(n))
    if key in _seq_cache:
        return _seq_cache[key]
    cache_dir = os.environ.get("QNMFITS_B200_KERR_CACHE")
    path = os.path.join(cache_dir, "kerr_s%d_l%d_m%d_n%d.npz" % key) if cache_dir else None
    if path and os.path.isfile(path):
        z = np.load(path)
        seq = KerrSequence(*key, z["a"], z["omega"], z["A"], z["C"])
    else:
        seq = compute_sequence(*key)
        if path:
            os.makedirs(cache_dir, exist_ok=True)
            np.savez(path, a=seq.a, omega=seq.omega, A=seq.A, C=seq.C)
    _seq_cache[key] = seq
    return seq
