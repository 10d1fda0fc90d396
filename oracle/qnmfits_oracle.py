"""
TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's least-squares
ringdown-fitting path.  Not product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker / the timed CPU arm.  Nothing under
``qnmfits_b200/`` imports it.

Parity status: PINNED against the reference itself.  The reference is pure Python
and imports in the build container (``oracle/ref_loader.py``); the script
``tests/golden/make_golden.py`` runs the unmodified reference functions and commits
their outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures on every run and against the live reference when
``/root/reference`` exists.  The reference has no tests of its own; the only
published known answers are the notebook values G1/G2, which
need the real Kerr tables (``qnm`` PyPI package, unpinned in the reference's
pyproject.toml:25) that are not installable offline.  G1's "mismatch 0, C = 1-1j"
is table-independent and is tested; G2's numeric value, omega_220(0.7) =
0.53260024-0.08079287i, is reproduced by the from-scratch Kerr provider
``qnmfits_b200/kerr.py`` (``tests/test_kerr_provider.py``), which can feed this oracle
like any other ``modes_cache``.

The least-squares solve itself is third-party in the reference too:
``numpy.linalg.lstsq(a, b, rcond=None)`` -> LAPACK zgelsd (numpy unpinned in the
reference's pyproject.toml:18; here numpy 2.3.5 / OpenBLAS 0.3.30).  The oracle
calls the same function, so "parity with the oracle" is parity with the
reference's numbers on this numpy.

Each function cites the reference lines it follows (paths relative to
/root/reference).  The per-fit structure (frequencies re-tabulated per call, one
``np.exp`` per column, ``lstsq``, ``einsum``, trapezoid mismatch) is kept on purpose:
this file is also the CPU baseline that bench.py times.
"""
import numpy as np
from scipy.interpolate import UnivariateSpline


class OracleTables:
    """Label -> value logic of the reference's provider (qnmfits/qnm.py:124-393)."""

    multiplets = ((2, 0, 8), (2, 1, 8), (2, 2, 8))  # qnm.py:67 (s = -2)

    def __init__(self, modes_cache):
        self.modes_cache = modes_cache
        self.splines = {}

    def _seq(self, ell, m, n, s):
        key = (ell, m, n, s)
        if key not in self.splines:
            # qnm.py:128-134 — overtone index shift above a multiplet
            n_load = n
            for el0, m0, n0 in self.multiplets:
                if ell == el0 and m == m0 and n > n0 + 1:
                    n_load -= 1
            seq = self.modes_cache(s, ell, m, n_load)
            a = seq.a
            # qnm.py:144-155 — interpolating splines of real and imaginary parts
            w = (UnivariateSpline(a, np.real(seq.omega), s=0),
                 UnivariateSpline(a, np.imag(seq.omega), s=0))
            mu = [(UnivariateSpline(a, np.real(c), s=0),
                   UnivariateSpline(a, np.imag(c), s=0)) for c in seq.C.T]
            self.splines[key] = (w, mu)
        return self.splines[key]

    def omega(self, ell, m, n, sign, chif, Mf=1, s=-2):
        # qnm.py:220-235
        w = self._seq(ell, m * sign, n, s)[0]
        val = w[0](chif) + 1j * w[1](chif)
        if sign == -1:
            val = -np.conjugate(val)
        return val / Mf

    def omega_list(self, modes, chif, Mf=1, s=-2):
        # qnm.py:272-280 — nonlinear modes are sums over 4-tuples
        res = []
        for mode in modes:
            total = 0
            for i in range(0, len(mode), 4):
                ell, m, n, sign = mode[i:i + 4]
                total = total + self.omega(ell, m, n, sign, chif, Mf, s)
            res.append(total)
        return res

    def mu(self, ell, m, ellp, mp, nprime, sign, chif, s=-2):
        # qnm.py:336-361
        if mp != m:
            return 0
        ms, mps = m * sign, mp * sign
        col = ell - (abs(ms) if abs(ms) > abs(s) else abs(s))
        f = self._seq(ellp, mps, nprime, s)[1][col]
        val = f[0](chif) + 1j * f[1](chif)
        if sign == -1:
            val = (-1) ** (ell + ellp) * np.conjugate(val)
        return val

    def mu_list(self, indices, chif, s=-2):
        # qnm.py:390-391 (six-index unpack: nonlinear labels raise ValueError)
        return [self.mu(ell, m, ellp, mp, nprime, sign, chif, s)
                for ell, m, ellp, mp, nprime, sign in indices]


def ringdown(time, start_time, complex_amplitudes, frequencies):
    """qnmfits/qnmfits.py:56-70 — sum of damped sinusoids, zero before start_time."""
    time = np.asarray(time)
    h = np.zeros(len(time), dtype=complex)
    keep = time >= start_time
    tau = (time - start_time)[keep]
    h[keep] = np.sum([complex_amplitudes[n] * np.exp(-1j * frequencies[n] * tau)
                      for n in range(len(frequencies))], axis=0)
    return h


def mismatch(times, wf_1, wf_2):
    """qnmfits/qnmfits.py:90-97."""
    num = np.real(np.trapezoid(wf_1 * np.conjugate(wf_2), x=times))
    den = np.sqrt(np.trapezoid(np.real(wf_1 * np.conjugate(wf_1)), x=times)
                  * np.trapezoid(np.real(wf_2 * np.conjugate(wf_2)), x=times))
    return 1 - (num / den)


def multimode_mismatch(times, wf_dict_1, wf_dict_2):
    """qnmfits/qnmfits.py:123-139 — sums over the keys of the first dict."""
    keys = list(wf_dict_1.keys())
    num = np.real(sum([np.trapezoid(wf_dict_1[k] * np.conjugate(wf_dict_2[k]), x=times)
                       for k in keys]))
    n1 = sum([np.trapezoid(np.real(wf_dict_1[k] * np.conjugate(wf_dict_1[k])), x=times)
              for k in keys])
    n2 = sum([np.trapezoid(np.real(wf_dict_2[k] * np.conjugate(wf_dict_2[k])), x=times)
              for k in keys])
    return 1 - (num / np.sqrt(n1 * n2))


def window(times, t0, T, t0_method):
    """Row range [start, stop) of the analysis window (qnmfits/qnmfits.py:231-244).

    For 'geq' the reference uses a boolean mask; for sorted times that is a
    contiguous range, which is what is returned (the mask itself is returned too).
    """
    if t0_method == 'geq':
        mask = (times >= t0) & (times < t0 + T)
        return mask
    if t0_method == 'closest':
        start = np.argmin((times - t0) ** 2)
        stop = np.argmin((times - t0 - T) ** 2)
        return slice(start, stop)
    raise ValueError("t0_method must be 'geq' or 'closest'")


def delta_factor(delta, n_modes):
    """qnmfits/qnmfits.py:256-271 (bad input -> ValueError instead of print+crash)."""
    if type(delta) is int:
        delta = float(delta)
    if type(delta) is list and len(delta) == n_modes:
        delta = np.array(delta)
    if (isinstance(delta, np.ndarray) and len(delta) == n_modes) or type(delta) is float:
        return delta + 1
    raise ValueError("delta must be a float or an array with length len(modes)")


def lstsq_fit(times_masked, data_masked, frequencies, t0, coef=None):
    """Design matrix + lstsq + model (qnmfits/qnmfits.py:280-290 and :628-639).

    coef is None for a single series, or an (L, N) table for L stacked series.
    """
    tau = times_masked - t0
    if coef is None:
        a = np.array([np.exp(-1j * frequencies[j] * tau)
                      for j in range(len(frequencies))]).T
    else:
        a = np.concatenate([
            np.array([coef[i][j] * np.exp(-1j * frequencies[j] * tau)
                      for j in range(len(frequencies))]).T
            for i in range(len(coef))])
    C, res, rank, s = np.linalg.lstsq(a, data_masked, rcond=None)
    model = np.einsum('ij,j->i', a, C)
    return a, C, res, rank, s, model


def ringdown_fit(tables, times, data, modes, Mf, chif, t0, t0_method='geq', T=100,
                 delta=0.0):
    """qnmfits/qnmfits.py:231-312."""
    sel = window(times, t0, T, t0_method)
    t_m, d_m = times[sel], data[sel]
    frequencies = delta_factor(delta, len(modes)) * np.array(
        tables.omega_list(modes, chif, Mf))
    a, C, res, rank, s, model = lstsq_fit(t_m, d_m, frequencies, t0)
    return {
        'residual': res, 'rank': rank, 's': s,
        'mismatch': mismatch(t_m, model, d_m), 'C': C,
        'data': d_m, 'model': model, 'model_times': t_m, 't0': t0,
        'modes': modes, 'mode_labels': [str(mode) for mode in modes],
        'frequencies': frequencies,
    }


def multimode_ringdown_fit(tables, times, data_dict, modes, Mf, chif, t0,
                           t0_method='geq', T=100, spherical_modes=None,
                           coef_override=None):
    """qnmfits/qnmfits.py:573-670.

    ``coef_override`` (L, N) replaces the mu table; it is how the documented
    superset (caller-supplied coefficient columns, e.g. for quadratic QNMs, which
    the reference cannot express — qnm.py:390) is checked.
    """
    if spherical_modes is None:
        spherical_modes = list(data_dict.keys())
    sel = window(times, t0, T, t0_method)
    t_m = times[sel]
    d_masked = {lm: data_dict[lm][sel] for lm in spherical_modes}
    stacked = np.concatenate([d_masked[lm] for lm in spherical_modes])
    frequencies = np.array(tables.omega_list(modes, chif, Mf))
    if coef_override is None:
        mu_lists = [tables.mu_list([lm + mode for mode in modes], chif)
                    for lm in spherical_modes]
    else:
        if callable(coef_override):                  # spin-dependent columns (grids)
            coef_override = coef_override(chif)
        mu_lists = [list(row) for row in coef_override]
    a, C, res, rank, s, model = lstsq_fit(t_m, stacked, frequencies, t0, mu_lists)
    K = len(t_m)
    model_dict, weighted_C = {}, {}
    for i, lm in enumerate(spherical_modes):
        model_dict[lm] = model[i * K:(i + 1) * K]
        weighted_C[lm] = np.array(mu_lists[i]) * C
    return {
        'residual': res, 'mismatch': multimode_mismatch(t_m, model_dict, d_masked),
        'C': C, 'weighted_C': weighted_C, 'data': d_masked, 'model': model_dict,
        'model_times': t_m, 't0': t0, 'modes': modes,
        'mode_labels': [str(mode) for mode in modes], 'frequencies': frequencies,
    }


def mismatch_t0_array(tables, times, data, modes, Mf, chif, t0_array, t0_method='geq',
                      T_array=100, spherical_modes=None, delta=0.0, coef_override=None):
    """qnmfits/qnmfits.py:1259-1301 (fixed-spectrum branch only)."""
    if type(T_array) != np.ndarray:
        T_array = T_array * np.ones(len(t0_array))
    out = []
    for t0, T in zip(t0_array, T_array):
        if type(data) == dict:
            fit = multimode_ringdown_fit(tables, times, data, modes, Mf, chif, t0,
                                         t0_method, T, spherical_modes, coef_override)
        else:
            fit = ringdown_fit(tables, times, data, modes, Mf, chif, t0, t0_method,
                               T, delta)
        out.append(fit['mismatch'])
    return out


def grid_axes(Mf_minmax, chif_minmax, res):
    """qnmfits/qnmfits.py:1382-1383."""
    return (np.linspace(Mf_minmax[0], Mf_minmax[1], res),
            np.linspace(chif_minmax[0], chif_minmax[1], res))


def mismatch_M_chi_grid(tables, times, data, modes, Mf_minmax, chif_minmax, t0,
                        t0_method='geq', T=100, res=50, spherical_modes=None,
                        delta=0.0, flat_indices=None, coef_override=None):
    """qnmfits/qnmfits.py:1382-1415.

    ``flat_indices`` restricts the loop to a subset of the res*res flat indices (the
    bounded sample that bench.py times); the return value is then the 1-D array of
    mismatches for those indices instead of the (res, res) grid.
    """
    Mf_array, chif_array = grid_axes(Mf_minmax, chif_minmax, res)
    idx = range(len(Mf_array) * len(chif_array)) if flat_indices is None \
        else flat_indices
    out = []
    for i in idx:
        Mf = Mf_array[int(i / len(Mf_array))]
        chif = chif_array[i % len(chif_array)]
        if type(data) is dict:
            fit = multimode_ringdown_fit(tables, times, data, modes, Mf, chif, t0,
                                         t0_method, T, spherical_modes, coef_override)
        else:
            fit = ringdown_fit(tables, times, data, modes, Mf, chif, t0, t0_method,
                               T, delta)
        out.append(fit['mismatch'])
    out = np.array(out)
    if flat_indices is None:
        out = np.reshape(out, (len(Mf_array), len(chif_array)))
    return out


def free_frequency_fit(tables, times, data, t0, modes=(), Mf=None, chif=None, t0_method='geq',
                       T=100, min_method='Nelder-Mead', return_result=False):
    """qnmfits/qnmfits.py:1972-2043: scipy minimize over (Re w, Im w) of the mismatch of a
    fit with the fixed modes plus one free frequency; x0 = [1, -0.5], bounds
    [(0, 2), (-1, 0)], xatol = 1e-8."""
    from scipy.optimize import minimize
    sel = window(times, t0, T, t0_method)
    t_m, d_m = times[sel], data[sel]
    fixed = np.array(tables.omega_list(list(modes), chif, Mf)) if len(modes) else np.zeros(0, complex)

    def mismatch_f_tau(x):
        frequencies = np.hstack([fixed, x[0] + 1j * x[1]])
        a, C, res, rank, s, model = lstsq_fit(t_m, d_m, frequencies, t0)
        return mismatch(t_m, model, d_m)

    res = minimize(mismatch_f_tau, [1, -0.5], method=min_method, bounds=[(0, 2), (-1, 0)],
                   options={'xatol': 1e-8, 'disp': False})
    omega = res.x[0] + 1j * res.x[1]
    return (omega, res) if return_result else omega


def mismatch_omega_grid(tables, times, data, modes, Mf, chif, re_minmax, im_minmax, t0,
                        t0_method='geq', T=100, res=50):
    """qnmfits/qnmfits.py:1745-1827, including its re-masking of the already masked arrays
    inside the loop (:1759-1768; harmless for 'geq', drops the last sample per iteration
    for 'closest') and the final transpose (:1825)."""
    re_array = np.linspace(re_minmax[0], re_minmax[1], res)
    im_array = np.linspace(im_minmax[0], im_minmax[1], res)
    fixed = list(tables.omega_list(list(modes), chif, Mf)) if len(modes) else []
    mm_list = []
    for i in range(len(re_array) * len(im_array)):
        re = re_array[int(i / len(re_array))]
        im = im_array[i % len(im_array)]
        sel = window(times, t0, T, t0_method)
        times, data = times[sel], data[sel]
        frequencies = np.array(fixed + [re + 1j * im])
        a, C, r, rank, s, model = lstsq_fit(times, data, frequencies, t0)
        mm_list.append(mismatch(times, model, data))
    return np.reshape(np.array(mm_list), (len(re_array), len(im_array))).T


def calculate_epsilon(tables, times, data, modes, Mf, chif, t0, t0_method='geq', T=100,
                      spherical_modes=None, min_method='Nelder-Mead', delta=0.0, x0=None):
    """qnmfits/qnmfits.py:1514-1594."""
    from scipy.optimize import minimize
    if x0 is None:
        x0 = [Mf, chif]

    def mismatch_M_chi(x):
        c = min(max(x[1], 0), 0.99)
        if type(data) == dict:
            return multimode_ringdown_fit(tables, times, data, modes, x[0], c, t0, t0_method, T,
                                          spherical_modes)['mismatch']
        return ringdown_fit(tables, times, data, modes, x[0], c, t0, t0_method, T, delta)['mismatch']

    res = minimize(mismatch_M_chi, x0, method=min_method, bounds=[(0, 2.0), (0, 0.99)],
                   options={'xatol': 1e-6, 'disp': False})
    dM, dc = res.x[0] - Mf, res.x[1] - chif
    return np.sqrt(dM**2 + dc**2), res.x[0], res.x[1]


def dynamic_ringdown_fit(tables, times, data, modes, Mf, chif, t0, t0_method='geq', T=100):
    """qnmfits/qnmfits.py:396-475: per-sample Mf / chif, a = exp(-1j*frequencies*(times-t0)).T
    with frequencies of shape (N, K)."""
    sel = window(times, t0, T, t0_method)
    t_m, d_m = times[sel], data[sel]
    Mf = np.full(len(t_m), Mf) if type(Mf) in [float, np.float64] else Mf[sel]
    chif = np.full(len(t_m), chif) if type(chif) in [float, np.float64] else chif[sel]
    frequencies = np.array(tables.omega_list(modes, chif, Mf))
    a = np.exp(-1j * frequencies * (t_m - t0)).T
    C, res, rank, s = np.linalg.lstsq(a, d_m, rcond=None)
    model = np.einsum('ij,j->i', a, C)
    return {'residual': res, 'mismatch': mismatch(t_m, model, d_m), 'C': C, 'data': d_m, 'model': model,
            'model_times': t_m, 't0': t0, 'modes': modes, 'mode_labels': [str(m) for m in modes],
            'frequencies': frequencies}


def dynamic_multimode_ringdown_fit(tables, times, data_dict, modes, Mf, chif, t0, t0_method='geq',
                                   T=100, spherical_modes=None):
    """qnmfits/qnmfits.py:773-911: a[(i,k), j] = mu_ij(chif_k) exp(-i w_j(Mf_k, chif_k) (t_k - t0))."""
    if spherical_modes is None:
        spherical_modes = list(data_dict.keys())
    sel = window(times, t0, T, t0_method)
    t_m = times[sel]
    d_masked = {lm: data_dict[lm][sel] for lm in spherical_modes}
    data = np.concatenate([d_masked[lm] for lm in spherical_modes])
    Mf = Mf[sel]
    chif = np.full(len(t_m), chif) if type(chif) in [float, np.float64] else chif[sel]
    frequencies = np.array(tables.omega_list(modes, chif, Mf)).T
    frequencies = np.vstack(len(spherical_modes) * [frequencies])
    mu = np.vstack([np.array([np.broadcast_to(v, t_m.shape) for v in
                              tables.mu_list([lm + mode for mode in modes], chif)]).T
                    for lm in spherical_modes])
    stacked_times = np.vstack(len(spherical_modes) * [t_m[:, None]])
    a = mu * np.exp(-1j * frequencies * (stacked_times - t0))
    C, res, rank, s = np.linalg.lstsq(a, data, rcond=None)
    model = np.einsum('ij,j->i', a, C)
    weighted = mu * C
    K = len(t_m)
    model_dict = {lm: model[i * K:(i + 1) * K] for i, lm in enumerate(spherical_modes)}
    return {'residual': res, 'mismatch': multimode_mismatch(t_m, model_dict, d_masked), 'C': C,
            'weighted_C': {lm: weighted[i * K:(i + 1) * K] for i, lm in enumerate(spherical_modes)},
            'data': d_masked, 'model': model_dict, 'model_times': t_m, 't0': t0, 'modes': modes,
            'mode_labels': [str(m) for m in modes], 'frequencies': frequencies}
