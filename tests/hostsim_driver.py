"""Drive the K1 lane-emulation harness (tests/hostsim) with numpy host arrays through
the same qnmfit_batch descriptor the CUDA library takes."""
import ctypes as C
import os

import numpy as np

from qnmfits_b200 import _cabi
from qnmfits_b200._engine import nominal_step

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_hs = None


def lib():
    global _hs
    if _hs is None:
        _hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libqnmfit_hostsim.so"))
        _hs.hostsim_fit_small.argtypes = [C.POINTER(_cabi.Batch), C.c_int, C.c_int]
        _hs.hostsim_fit_pair.argtypes = [C.POINTER(_cabi.Batch), C.c_int, C.c_int]
        _hs.hostsim_fit_small_cta.argtypes = [C.POINTER(_cabi.Batch), C.c_int, C.c_int]
        _hs.hostsim_fit_struct.argtypes = [C.POINTER(_cabi.Batch), C.c_int]
        _hs.hostsim_fit_panel.argtypes = [C.POINTER(_cabi.Batch), C.c_int]
        _hs.hostsim_fit_general.argtypes = [C.POINTER(_cabi.Batch), C.c_int]
    return _hs


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(times, data, *, n_fits, n_modes, window, t0, lpf=4, eval_only=False, dt=None,
        anchor_rows=0, omega=None, omega_shared=False, table=None, mode_ptr=None, inv_Mf=None,
        delta_factor=None, n_chi=0, n_mf=0, first_fit=0, C_in=None, want_model=False,
        uniform_weights=0, series_index=None, pair=False, order=0, cta=False, staged=False):
    """``pair=True``: the K1p code (csrc/fit_pair.cuh), warps in lock step; ``cta=True``: K1's kernel
    function itself on an emulated CTA (``staged``: with the window staged in shared memory); otherwise
    K1's stages called lane by lane.  ``order``: how the emulated threads are resumed between
    collectives — 0 ascending, 1 descending, >= 2 seed of an order shuffled anew for every pass."""
    times = np.ascontiguousarray(times, dtype=float)
    data = np.ascontiguousarray(data, dtype=complex)
    keep = [times, data]
    kw = {}
    if isinstance(window[0], np.ndarray):
        rb = np.ascontiguousarray(window[0], np.int32)
        re = np.ascontiguousarray(window[1], np.int32)
        keep += [rb, re]
        kw.update(row_begin=_p(rb), row_end=_p(re), row_begin_all=int(rb.min()),
                  row_end_all=int(re.max()))
    else:
        kw.update(row_begin_all=int(window[0]), row_end_all=int(window[1]))
    if np.ndim(t0) == 0:
        kw.update(t0_all=float(t0))
    else:
        t0a = np.ascontiguousarray(t0, dtype=float)
        keep.append(t0a)
        kw.update(t0=_p(t0a))
    if omega is not None:
        om = np.ascontiguousarray(omega, dtype=complex)
        keep.append(om)
        kw.update(omega=_p(om), omega_shared=1 if omega_shared else 0)
        wmax = float(np.max(np.abs(om)))
    else:
        tab = np.ascontiguousarray(table, dtype=complex)
        mp = np.ascontiguousarray(mode_ptr, np.int32)
        inv = np.ascontiguousarray(inv_Mf, dtype=float)
        keep += [tab, mp, inv]
        kw.update(omega_tilde=_p(tab), mode_ptr=_p(mp), inv_Mf=_p(inv), n_chi=n_chi, n_mf=n_mf,
                  n_constituents=tab.shape[1])
        wmax = float(np.max(np.abs(tab)) * np.max(inv)) * 3
        if delta_factor is not None:
            df = np.ascontiguousarray(delta_factor, dtype=float)
            keep.append(df)
            kw.update(delta_factor=_p(df))
    if series_index is not None:
        si = np.ascontiguousarray(series_index, np.int32)
        keep.append(si)
        kw.update(series_index=_p(si))
    if dt is None:
        dt = nominal_step(times[kw["row_begin_all"]:kw["row_end_all"]], wmax)
    Mmax = kw["row_end_all"] - kw["row_begin_all"]
    Cbuf = np.zeros((n_fits, n_modes), complex) if C_in is None else \
        np.ascontiguousarray(C_in, dtype=complex)
    mm = np.zeros(n_fits)
    res = np.zeros(n_fits)
    R = np.zeros((n_fits, n_modes, n_modes + 1), complex)
    st = np.zeros(n_fits, np.int32)
    model = np.zeros((n_fits, Mmax), complex) if want_model else None
    b = _cabi.Batch(n_fits=n_fits, n_modes=n_modes, n_series=1, n_times=len(times),
                    series_stride=data.shape[-1], first_fit=first_fit, times=_p(times), data=_p(data),
                    dt_nominal=float(dt), anchor_rows=anchor_rows, C=_p(Cbuf), mismatch=_p(mm),
                    residual=_p(res), R=_p(R), status=_p(st), model=_p(model),
                    model_stride=Mmax if want_model else 0, uniform_weights=int(uniform_weights), **kw)
    entry = lib().hostsim_fit_pair if pair else lib().hostsim_fit_small_cta if cta else lib().hostsim_fit_small
    rc = entry(C.byref(b), int(lpf), (1 if eval_only else 0) | (2 if cta and staged else 0)
               | (int(order) << 8 if (pair or cta) else 0))
    assert rc == 0, rc
    return dict(C=Cbuf, mismatch=mm, residual=res, R=R, status=st, model=model, dt=dt)


def run_struct(times, data, *, n_fits, n_modes, window, t0, omega, coef=None, eval_only=False, dt=None,
               uniform_weights=0, C_in=None, want_model=False, order=0, panel=False, general=False,
               omega_rows=None, coef_rows=None):
    """K3 (csrc/fit_struct.cuh) or, with ``panel=True``, K4 (csrc/fit_panel.cuh, its DMMA emulated with the
    PTX fragment layout) or, with ``general=True``, K2 (csrc/fit_general.cuh): the kernel function itself on an
    emulated CTA per fit.  ``data`` is
    (L, K_tot); ``omega`` (N,) shared by all fits; ``coef`` (L, N) or None (single series, no table)."""
    times = np.ascontiguousarray(times, dtype=float)
    data = np.ascontiguousarray(np.atleast_2d(data), dtype=complex)
    L, K_tot = data.shape
    om = np.ascontiguousarray(omega, dtype=complex).reshape(1, -1)
    keep = [times, data, om]
    kw = dict(omega=_p(om), omega_shared=1)
    if isinstance(window[0], np.ndarray):
        rb = np.ascontiguousarray(window[0], np.int32)
        re = np.ascontiguousarray(window[1], np.int32)
        keep += [rb, re]
        kw.update(row_begin=_p(rb), row_end=_p(re), row_begin_all=int(rb.min()), row_end_all=int(re.max()))
    else:
        kw.update(row_begin_all=int(window[0]), row_end_all=int(window[1]))
    if np.ndim(t0) == 0:
        kw.update(t0_all=float(t0))
    else:
        t0a = np.ascontiguousarray(t0, dtype=float)
        keep.append(t0a)
        kw.update(t0=_p(t0a))
    if coef is not None:
        cf = np.ascontiguousarray(coef, dtype=complex).reshape(1, L, n_modes)
        ci = np.zeros(n_fits, np.int32)
        keep += [cf, ci]
        kw.update(coef=_p(cf), coef_index=_p(ci), n_coef=1)
    if omega_rows is not None:                       # per-sample frequencies [N][K_tot] (dynamic fits)
        orows = np.ascontiguousarray(omega_rows, dtype=complex)
        keep.append(orows)
        kw.update(omega_rows=_p(orows))
    if coef_rows is not None:                        # per-sample mixing coefficients [L][N][K_tot] (K2 only)
        crows = np.ascontiguousarray(coef_rows, dtype=complex)
        keep.append(crows)
        kw.update(coef_rows=_p(crows))
    if dt is None:
        dt = nominal_step(times[kw["row_begin_all"]:kw["row_end_all"]], float(np.max(np.abs(om))))
    Mmax = kw["row_end_all"] - kw["row_begin_all"]
    Cbuf = np.zeros((n_fits, n_modes), complex) if C_in is None else np.ascontiguousarray(C_in, dtype=complex)
    mm, res = np.zeros(n_fits), np.zeros(n_fits)
    st = np.zeros(n_fits, np.int32)
    model = np.zeros((n_fits, L * Mmax), complex) if want_model else None
    b = _cabi.Batch(n_fits=n_fits, n_modes=n_modes, n_series=L, n_times=K_tot, series_stride=K_tot,
                    times=_p(times), data=_p(data), dt_nominal=float(dt), C=_p(Cbuf), mismatch=_p(mm),
                    residual=_p(res), status=_p(st), model=_p(model), model_stride=L * Mmax if want_model else 0,
                    uniform_weights=int(uniform_weights), **kw)
    entry = lib().hostsim_fit_general if general else lib().hostsim_fit_panel if panel else lib().hostsim_fit_struct
    rc = entry(C.byref(b), (1 if eval_only else 0) | (int(order) << 8))
    assert rc == 0, rc
    return dict(C=Cbuf, mismatch=mm, residual=res, status=st, model=model, dt=dt)
