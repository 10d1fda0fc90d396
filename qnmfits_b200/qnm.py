"""
QNM table provider — host side of the drop-in boundary.

Mirrors the interface of the reference's ``qnmfits.qnm.qnm`` class (reference
``qnmfits/qnm.py:36-393``): ``omega``, ``omega_list``, ``mu``, ``mu_list`` with the
same argument order, label conventions and return types.  Tabulation stays on the
host (as BASELINE.json's north_star asks): cubic interpolating splines in spin built
by the same SciPy call the reference makes, so label -> value resolution is
bit-identical; what this module adds is the *factored, vectorised* form the device
kernels consume (``constituent_table``): the spline values for the few unique spins
of a sweep instead of one Python call per grid point.

Label semantics kept from the reference:

* ``(ell, m, n, sign)``; ``sign=-1`` selects the mirror mode: the sequence loaded is
  ``(ell, -m, n)`` and the value is ``-conj(omega)`` (reference qnm.py:220,232-233).
* a mode tuple of length 4p is a p-th order (quadratic, cubic, ...) QNM whose
  frequency is the Python ``sum`` of its constituents, each already divided by
  ``Mf`` (reference qnm.py:235,272-280).
* ``mu`` is the int ``0`` when ``m' != m`` — tested *before* the mirror sign flip
  (reference qnm.py:336-341); the ell column is ``ell - max(|m|, |s|)`` evaluated
  after the flip (reference qnm.py:345-348); mirror value
  ``(-1)**(ell+ell') * conj(mu)`` (reference qnm.py:358-359).
* overtone re-indexing around the Cook-Zalutskiy multiplets: for (2,0), (2,1),
  (2,2) a request ``n > 9`` loads sequence ``n - 1`` (reference qnm.py:128-134);
  the n = 8, 9 multiplet members themselves come from ``KerrQNM_08.h5`` /
  ``KerrQNM_09.h5`` when present (reference qnm.py:60-122).
"""
from pathlib import Path

import numpy as np
from scipy.interpolate import UnivariateSpline

_table_provider = None


def set_table_provider(modes_cache):
    """Install the callable used in place of ``qnm.modes_cache(s, l, m, n)``.

    ``modes_cache`` must return an object with ``.a``, ``.omega`` and ``.C`` (see
    reference qnmfits/qnm.py:134-141).  Passing ``None`` restores the default (the
    ``qnm`` PyPI package).  Existing ``qnm`` instances drop their spline caches.
    """
    global _table_provider
    _table_provider = modes_cache
    for inst in list(qnm._instances):
        inst._reset_sequences()


_warned_fallback = False


def _default_provider():
    """The installed table source: what ``set_table_provider`` installed, else the ``qnm``
    PyPI package (what the reference uses, qnm.py:134), else this package's own Leaver
    solver (``qnmfits_b200.kerr``) with a one-time notice."""
    global _warned_fallback
    if _table_provider is not None:
        return _table_provider
    try:
        import qnm as qnm_loader  # the PyPI package (Stein 2019)
    except ImportError:
        from . import kerr
        if not _warned_fallback:
            import warnings
            warnings.warn(
                "The 'qnm' package (Kerr QNM tables) is not installed: using the built-in Leaver "
                "solver qnmfits_b200.kerr (same algorithm; frequencies agree to ~1e-9, overtone "
                "n = 8 of l = 2 unavailable).  Call qnmfits_b200.set_table_provider(...) to "
                "install another source.", RuntimeWarning, stacklevel=3)
            _warned_fallback = True
        return kerr.modes_cache
    return qnm_loader.modes_cache


def _spline_pair(x, values):
    """Interpolating cubic splines for real and imaginary part (qnm.py:144-155)."""
    return (UnivariateSpline(x, np.real(values), s=0),
            UnivariateSpline(x, np.imag(values), s=0))


class qnm:
    """Frequencies and spherical-spheroidal mixing coefficients of Kerr QNMs."""

    _instances = []
    _epoch = 0

    #: multiplets that the Leaver solver of the ``qnm`` package cannot follow
    #: (reference qnm.py:67), as (ell, m, n, s)
    multiplet_list = [(2, 0, 8, -2), (2, 1, 8, -2), (2, 2, 8, -2)]

    def __init__(self, data_dir=None):
        self._qnm_funcs = {}
        self._interpolated_qnm_funcs = {}
        self._tabulated = {}
        self.download_check = {}
        self._data_dir = Path(data_dir) if data_dir is not None \
            else Path(__file__).parent / 'Data'
        self._load_cook_multiplets()
        qnm._instances.append(self)

    # ------------------------------------------------------------------ tables

    def _reset_sequences(self):
        qnm._epoch += 1                  # prepared sweeps keyed on the tables are stale now
        self._interpolated_qnm_funcs = {}
        self._tabulated = {}
        self._load_cook_multiplets()

    def _load_cook_multiplets(self):
        """Cook & Zalutskiy n=8,9 multiplet data, if the HDF5 files are present.

        Dataset path ``n08/m+02/{2,2,{8,0}}``; columns
        ``[chi, Re w, Im w, Re A, Im A, Re mu_0, Im mu_0, ...]`` (reference
        qnm.py:79-98).
        """
        for ell, m, n, s in self.multiplet_list:
            path = self._data_dir / f'KerrQNM_{n:02}.h5'
            self.download_check[n] = path.exists()
            if not self.download_check[n]:
                continue
            import h5py
            with h5py.File(path, 'r') as f:
                for i in (0, 1):
                    name = f'n{n:02}/m{m:+03}/{{{ell},{m},{{{n},{i}}}}}'
                    table = np.array(f[name])
                    spins = table[:, 0]
                    w = _spline_pair(spins, table[:, 1] + 1j * table[:, 2])
                    mus = [
                        _spline_pair(spins, re + 1j * im)
                        for re, im in zip(table[:, 5::2].T, table[:, 6::2].T)
                    ]
                    self._interpolated_qnm_funcs[(ell, m, n + i, s)] = [w, mus]

    def _interpolate(self, ell, m, n, s=-2):
        n_load = n
        for ellp, mp, nprime, sp in self.multiplet_list:
            if (ell == ellp) and (m == mp) and (n > nprime + 1):
                n_load -= 1
        seq = _default_provider()(s, ell, m, n_load)
        spins = seq.a
        w = _spline_pair(spins, seq.omega)
        mus = [_spline_pair(spins, col) for col in np.asarray(seq.C).T]
        self._interpolated_qnm_funcs[ell, m, n, s] = [w, mus]

    def _funcs(self, ell, m, n, s):
        key = (ell, m, n, s)
        if key not in self._interpolated_qnm_funcs:
            self._interpolate(ell, m, n, s)
        return self._interpolated_qnm_funcs[key]

    # ----------------------------------------------------------- reference API

    def omega(self, ell, m, n, sign, chif, Mf=1, s=-2):
        """Complex frequency omega_{ell m n}(Mf, chif) (reference qnm.py:162-235)."""
        m = m * sign
        re_f, im_f = self._funcs(ell, m, n, s)[0]
        omega = re_f(chif) + 1j * im_f(chif)
        if sign == -1:
            omega = -np.conjugate(omega)
        return omega / Mf

    def _scalar_key(self, kind, labels, *scalars):
        """Memo key of a list call with scalar spin / mass, or None (array arguments,
        unhashable labels): repeated calls on the same remnant skip the FITPACK
        evaluations — the values are those the first call computed."""
        try:
            if any(np.ndim(v) != 0 for v in scalars):
                return None
            return (kind, tuple(map(tuple, labels))) + tuple(
                v if isinstance(v, (int, float)) else float(v) for v in scalars)
        except TypeError:
            return None

    def omega_list(self, modes, chif, Mf=1, s=-2):
        """List of mode frequencies; 4p-tuples sum p constituents (qnm.py:237-280)."""
        key = self._scalar_key('wl', modes, chif, Mf, s)
        if key is not None:
            hit = self._tabulated.get(key)
            if hit is not None:
                return list(hit)
        out = []
        for mode in modes:
            parts = [
                self.omega(*mode[i:i + 4], chif, Mf, s)
                for i in range(0, len(mode), 4)
            ]
            out.append(sum(parts))
        if key is not None:
            self._memo(key, lambda: tuple(out))
        return out

    def mu(self, ell, m, ellp, mp, nprime, sign, chif, s=-2):
        """Spherical-spheroidal mixing coefficient (reference qnm.py:293-361)."""
        if mp != m:
            return 0
        m = m * sign
        mp = mp * sign
        index = ell - max(abs(m), abs(s))
        re_f, im_f = self._funcs(ellp, mp, nprime, s)[1][index]
        mu = re_f(chif) + 1j * im_f(chif)
        if sign == -1:
            mu = (-1) ** (ell + ellp) * np.conjugate(mu)
        return mu

    def mu_list(self, indices, chif, s=-2):
        """Mixing coefficients for (ell, m, ell', m', n', sign) tuples (qnm.py:363-393)."""
        key = self._scalar_key('ml', indices, chif, s)
        if key is not None:
            hit = self._tabulated.get(key)
            if hit is not None:
                return list(hit)
        out = [
            self.mu(ell, m, ellp, mp, nprime, sign, chif, s)
            for ell, m, ellp, mp, nprime, sign in indices
        ]
        if key is not None:
            self._memo(key, lambda: tuple(out))
        return out

    # ------------------------------------------------- factored device tables

    def _memo(self, key, compute):
        """Memoise a tabulated column per (label, spin array): the reference memoises its
        splines the same way (qnm.py:225-226); a sweep repeated on the same spin grid
        (other t0, other data) then skips the FITPACK evaluation."""
        hit = self._tabulated.get(key)
        if hit is None:
            if len(self._tabulated) > 4096:
                self._tabulated.clear()
            hit = self._tabulated[key] = compute()
        return hit

    def constituent_table(self, modes, chif_values, s=-2, with_max=False):
        """Mf-independent frequency table for a sweep over spins.

        Returns ``(table, mode_ptr)``: ``table`` is complex128 of shape
        ``(len(chif_values), P)`` holding, for every unique spin, the mirror-resolved
        dimensionless frequency of each of the ``P`` constituents of ``modes`` (a
        linear mode has one, a quadratic mode two, ...); ``mode_ptr`` is int32 of
        length ``len(modes)+1`` with the constituent range of each mode.  The device
        forms ``omega_j = delta_factor_j * sum_p(table[c, p] * (1/Mf))`` with the
        rounding order of the reference (qnm.py:235 then the Python ``sum`` of
        qnm.py:272-280 then qnmfits.py:274), so the per-point frequencies match it
        bit for bit.  ``with_max``: also return max |table| (the frequency bound of the
        device's row recurrence).  The assembled table is memoised per (modes, spins) as
        well — read-only arrays are returned.
        """
        chif_values = np.atleast_1d(np.asarray(chif_values, dtype=float))
        chi_key = chif_values.tobytes()
        try:
            whole_key = ('table', tuple(map(tuple, modes)), s, chi_key)
            hit = self._tabulated.get(whole_key)
        except TypeError:
            whole_key = hit = None
        if hit is not None:
            return hit if with_max else hit[:2]
        cols = []
        mode_ptr = [0]
        for mode in modes:
            if len(mode) == 0 or len(mode) % 4 != 0:
                raise ValueError(
                    f"mode label {mode!r} must have a multiple of 4 entries")
            for i in range(0, len(mode), 4):
                ell, m, n, sign = (int(v) for v in mode[i:i + 4])
                cols.append(self._memo(
                    ('w', ell, m, n, sign, s, chi_key),
                    lambda: np.asarray(self.omega(ell, m, n, sign, chif_values, 1.0, s),
                                       dtype=complex)))
            mode_ptr.append(len(cols))
        table = np.ascontiguousarray(np.stack(cols, axis=1)) if cols else \
            np.zeros((len(chif_values), 0), dtype=complex)
        mode_ptr = np.asarray(mode_ptr, dtype=np.int32)
        table.setflags(write=False)
        mode_ptr.setflags(write=False)
        hit = (table, mode_ptr, float(np.max(np.abs(table))) if table.size else 0.0)
        if whole_key is not None:
            self._tabulated[whole_key] = hit
        return hit if with_max else hit[:2]

    def mu_table(self, spherical_modes, modes, chif_values, s=-2):
        """Mixing coefficients mu[c, i, j] for spins c, spherical modes i, QNMs j.

        Entry rules follow ``mu`` exactly (int 0 -> 0+0j).  Only 4-tuples are valid
        here: the reference raises for nonlinear labels in the multimode fit
        (qnm.py:390 unpacks six indices); see ``multimode_ringdown_fit`` for the
        documented superset.
        """
        chif_values = np.atleast_1d(np.asarray(chif_values, dtype=float))
        chi_key = chif_values.tobytes()
        out = np.zeros((len(chif_values), len(spherical_modes), len(modes)),
                       dtype=complex)
        for i, (ell, m) in enumerate(spherical_modes):
            for j, mode in enumerate(modes):
                ellp, mp, nprime, sign = (int(v) for v in mode)
                if mp != m:
                    continue          # int 0 in the reference (qnm.py:336-337)
                out[:, i, j] = self._memo(
                    ('mu', int(ell), int(m), ellp, mp, nprime, sign, s, chi_key),
                    lambda: np.asarray(self.mu(ell, m, ellp, mp, nprime, sign, chif_values, s),
                                       dtype=complex))
        return out
