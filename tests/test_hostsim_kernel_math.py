"""CPU check of the K1 kernel's arithmetic: tests/hostsim compiles the device code of
csrc/fit_small.cuh for the host and steps the lanes in the kernel's order.  Compared
with the oracle (numpy lstsq) on the golden cases.  (The CUDA build itself is checked
by the -m gpu tests.)"""
import numpy as np
import pytest

import cases
import hostsim_driver as hs
from oracle import qnmfits_oracle as orc
from qnmfits_b200 import qnmfits as api
from qnmfits_b200 import workloads


@pytest.fixture(autouse=True, scope="module")
def _one_blas_thread():
    """The numpy references of this module are thousands of tiny dense solves: with one BLAS thread per
    core they spend their time in thread hand-offs (20 tests took 86 s instead of 8).  Local to this
    module — the golden comparisons elsewhere are to 1e-13 and depend on the BLAS code path."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover - optional
        yield
        return
    with threadpool_limits(limits=1):
        yield


def _single(qf, wl, kw, lpf, **extra):
    kw = dict(kw)
    modes = kw.pop("modes")
    Mf, chif, t0 = kw.pop("Mf"), kw.pop("chif"), kw.pop("t0")
    T = kw.pop("T", 100)
    method = kw.pop("t0_method", "geq")
    delta = kw.pop("delta", 0.0)
    sel = api._window(wl.times, t0, T, method)
    tm, dm = wl.times[sel], wl.data[sel]
    freq = api._delta_factor(delta, len(modes)) * np.array(qf.qnm.omega_list(modes, chif, Mf))
    return hs.run(tm, dm, n_fits=1, n_modes=len(modes), window=(0, len(tm)), t0=t0, lpf=lpf,
                  omega=freq.reshape(1, -1), omega_shared=True, **extra), freq


@pytest.mark.parametrize("lpf", [1, 4, 32])
def test_single_fit_cases_vs_golden(qf, golden, lpf):
    g = golden("cfg1")
    wl, cs = cases.cfg1_cases()
    for name, kw in cs.items():
        if len(kw["modes"]) > 8 or name == "duplicate_label":
            continue
        out, freq = _single(qf, wl, kw, lpf)
        C_ref, s = g[name + "__C"], g[name + "__s"]
        assert np.array_equal(freq, g[name + "__frequencies"]), name
        err = np.max(np.abs(out["C"][0] - C_ref)) / np.max(np.abs(C_ref))
        assert err < cases.amp_tol(s), (name, err)
        assert abs(out["mismatch"][0] - float(g[name + "__mismatch"])) < 1e-10, name
        assert out["status"][0] == 0
        np.testing.assert_allclose(out["residual"][0], g[name + "__residual"][0], rtol=1e-9)
        sv = np.linalg.svd(out["R"][0][:, :len(s)], compute_uv=False)
        np.testing.assert_allclose(sv, s, rtol=1e-9)


def test_model_output_and_eval_only(qf, golden):
    g = golden("cfg1")
    wl, cs = cases.cfg1_cases()
    out, freq = _single(qf, wl, cs["offgrid_geq"], 8, want_model=True)
    M = len(g["offgrid_geq__model"])
    np.testing.assert_allclose(out["model"][0][:M], g["offgrid_geq__model"], rtol=0,
                               atol=1e-9 * np.max(np.abs(g["offgrid_geq__model"])))
    # eval-only with the reference's amplitudes reproduces its mismatch
    sel = api._window(wl.times, 3.37, 77.7, "geq")
    ev = hs.run(wl.times[sel], wl.data[sel], n_fits=1, n_modes=8, window=(0, M), t0=3.37, lpf=4,
                omega=freq.reshape(1, -1), omega_shared=True, eval_only=True,
                C_in=g["offgrid_geq__C"].reshape(1, -1))
    assert abs(ev["mismatch"][0] - float(g["offgrid_geq__mismatch"])) < 1e-13


def test_direct_mode_on_nonuniform_grid(qf, golden):
    g = golden("nonuniform")
    t, d = g["times"], g["data"]
    modes = workloads.overtone_modes(5)
    sel = api._window(t, 1.0, 60, "geq")
    freq = np.array(qf.qnm.omega_list(modes, 0.69, 0.95))
    out = hs.run(t[sel], d[sel], n_fits=1, n_modes=5, window=(0, int(sel.sum())), t0=1.0, lpf=8,
                 omega=freq.reshape(1, -1), omega_shared=True)
    assert out["dt"] == 0.0            # direct evaluation selected
    assert cases.rel_err(out["C"][0], g["fit__C"]) < 1e-8
    assert abs(out["mismatch"][0] - float(g["fit__mismatch"])) < 1e-10


def test_recurrence_needs_the_jitter_correction(qf, golden):
    """np.arange(n)*0.1 is not bit-uniform; anchors + first-order correction keep the
    generated rows within ~1e-15 of direct evaluation regardless of anchor spacing."""
    g = golden("cfg1")
    wl, cs = cases.cfg1_cases()
    ref_direct, _ = _single(qf, wl, cs["base"], 4, dt=0.0)
    for anchor in (8, 32, 128):
        out, _ = _single(qf, wl, cs["base"], 4, anchor_rows=anchor)
        err = np.max(np.abs(out["C"][0] - ref_direct["C"][0])) / np.max(np.abs(ref_direct["C"][0]))
        assert err < 2e-9, (anchor, err)


def test_t0_sweep_vs_golden(qf, golden):
    g = golden("cfg2")
    wl = workloads.config2(n_t0=40)
    begin = np.empty(40, np.int32)
    end = np.empty(40, np.int32)
    for i, t0 in enumerate(wl.t0_array):
        begin[i], end[i] = api._window_rows(wl.times, t0, 100.0, "geq")
    freq = np.array(qf.qnm.omega_list(wl.modes, 0.95, 0.69)).reshape(1, -1)
    freq = np.array(qf.qnm.omega_list(wl.modes, 0.69, 0.95)).reshape(1, -1)
    for lpf in (2, 16):
        out = hs.run(wl.times, wl.data, n_fits=40, n_modes=8, window=(begin, end), t0=wl.t0_array,
                     lpf=lpf, omega=freq, omega_shared=True)
        np.testing.assert_allclose(out["mismatch"], g["mismatch"], rtol=0, atol=1e-10)


def test_grid_vs_golden_and_sharding_invariance(qf, golden):
    g = golden("cfg3")
    wl = workloads.config3(res=12)
    Mf = np.linspace(0.85, 1.05, 12)
    chi = np.linspace(0.59, 0.79, 12)
    table, ptr = qf.qnm.constituent_table(wl.modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    common = dict(n_modes=8, window=win, t0=0.0, table=table, mode_ptr=ptr, inv_Mf=1.0 / Mf,
                  n_chi=12, n_mf=12)
    full = hs.run(wl.times, wl.data, n_fits=144, lpf=4, **common)
    np.testing.assert_allclose(full["mismatch"].reshape(12, 12), g["grid"], rtol=0, atol=1e-10)
    # two slabs (as two ranks would compute them) are bit-identical to the full launch
    a = hs.run(wl.times, wl.data, n_fits=72, first_fit=0, lpf=4, **common)
    b = hs.run(wl.times, wl.data, n_fits=72, first_fit=72, lpf=4, **common)
    assert np.array_equal(np.concatenate([a["mismatch"], b["mismatch"]]), full["mismatch"])


def test_grid_with_quadratic_mode_and_delta(qf, golden):
    g = golden("cfg3")
    wl = workloads.config3(res=5)
    modes = [(2, 2, 0, 1), (2, 2, 1, 1), (2, 2, 0, 1, 2, 2, 0, 1)]
    Mf = np.linspace(0.85, 1.05, 5)
    chi = np.linspace(0.59, 0.79, 5)
    table, ptr = qf.qnm.constituent_table(modes, chi)
    win = api._window_rows(wl.times, 15.0, 60, "geq")
    out = hs.run(wl.times, wl.data, n_fits=25, n_modes=3, window=win, t0=15.0, lpf=8, table=table,
                 mode_ptr=ptr, inv_Mf=1.0 / Mf, n_chi=5, n_mf=5,
                 delta_factor=np.array([1.0, 1.01, 1.0]))
    np.testing.assert_allclose(out["mismatch"].reshape(5, 5), g["grid_quadratic"], rtol=0,
                               atol=1e-10)


def test_ragged_and_tiny_windows(qf, oracle_tables):
    """Windows that are not multiples of the block size, shorter than the lanes, and
    barely overdetermined."""
    wl = workloads.config1()
    modes = wl.modes[:3]
    freq = np.array(qf.qnm.omega_list(modes, 0.69, 0.95)).reshape(1, -1)
    for M in (4, 5, 7, 33, 127):
        for lpf in (1, 8, 32):
            out = hs.run(wl.times, wl.data, n_fits=1, n_modes=3, window=(500, 500 + M), t0=0.0,
                         lpf=lpf, omega=freq, omega_shared=True)
            want = orc.ringdown_fit(oracle_tables, wl.times, wl.data, modes, 0.95, 0.69, 0.0,
                                    T=wl.times[500 + M - 1] + 0.05)
            assert len(want["model_times"]) == M
            tol = cases.amp_tol(want["s"])
            err = np.max(np.abs(out["C"][0] - want["C"])) / np.max(np.abs(want["C"]))
            assert err < tol, (M, lpf, err)
            assert abs(out["mismatch"][0] - want["mismatch"]) < 1e-10, (M, lpf)


def test_fast_mismatch_path_equals_general_path(qf, golden):
    """uniform_weights=1: mismatch and residual from the by-products of the
    factorisation (no second pass) agree with the weighted second pass and with the
    reference to 1e-10, for full grids, ragged t0 windows and every lanes-per-fit."""
    g2, g3 = golden("cfg2"), golden("cfg3")
    wl = workloads.config3(res=12)
    Mf = np.linspace(0.85, 1.05, 12)
    chi = np.linspace(0.59, 0.79, 12)
    table, ptr = qf.qnm.constituent_table(wl.modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    common = dict(n_fits=144, n_modes=8, window=win, t0=0.0, table=table, mode_ptr=ptr,
                  inv_Mf=1.0 / Mf, n_chi=12, n_mf=12)
    slow = hs.run(wl.times, wl.data, lpf=4, uniform_weights=0, **common)
    for lpf in (1, 4, 16):
        fast = hs.run(wl.times, wl.data, lpf=lpf, uniform_weights=1, **common)
        np.testing.assert_allclose(fast["mismatch"].reshape(12, 12), g3["grid"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(fast["mismatch"], slow["mismatch"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(fast["residual"], slow["residual"], rtol=1e-9)
        assert np.array_equal(fast["C"], hs.run(wl.times, wl.data, lpf=lpf, uniform_weights=0,
                                                **common)["C"])
    wl2 = workloads.config2(n_t0=40)
    begin = np.empty(40, np.int32)
    end = np.empty(40, np.int32)
    for i, t0 in enumerate(wl2.t0_array):
        begin[i], end[i] = api._window_rows(wl2.times, t0, 100.0, "geq")
    freq = np.array(qf.qnm.omega_list(wl2.modes, 0.69, 0.95)).reshape(1, -1)
    fast = hs.run(wl2.times, wl2.data, n_fits=40, n_modes=8, window=(begin, end), t0=wl2.t0_array,
                  lpf=8, omega=freq, omega_shared=True, uniform_weights=1)
    np.testing.assert_allclose(fast["mismatch"], g2["mismatch"], rtol=0, atol=1e-10)


def test_per_fit_data_series_for_the_free_frequency_search(qf, oracle_tables):
    """series_index: each fit of one launch reads its own waveform (the batched
    free-frequency objective), explicit per-fit frequencies."""
    from oracle import qnmfits_oracle as orc
    wl = workloads.config5(n_waveforms=7, n_fixed=2)
    fixed = np.array(oracle_tables.omega_list(wl.modes, wl.chif, wl.Mf))
    rng = np.random.default_rng(4)
    pick = np.array([6, 2, 2, 5, 0], dtype=np.int32)
    trial = rng.uniform(0.3, 1.7, len(pick)) - 1j * rng.uniform(0.05, 0.9, len(pick))
    omega = np.column_stack([np.tile(fixed, (len(pick), 1)), trial])
    win = (int(np.searchsorted(wl.times, 0.0)), int(np.searchsorted(wl.times, 100.0)))
    for uw in (0, 1):
        out = hs.run(wl.times, wl.data, n_fits=len(pick), n_modes=3, window=win, t0=0.0, lpf=8,
                     omega=omega, series_index=pick, uniform_weights=uw)
        for k, b in enumerate(pick):
            sel = orc.window(wl.times, 0.0, 100, 'geq')
            a, C, res, rank, s, model = orc.lstsq_fit(wl.times[sel], wl.data[b][sel], omega[k], 0.0)
            assert np.max(np.abs(out["C"][k] - C)) / np.max(np.abs(C)) < 1e-9
            assert abs(out["mismatch"][k] - orc.mismatch(wl.times[sel], model, wl.data[b][sel])) < 1e-11


def test_oracle_free_frequency_fit_vs_reference_golden(golden, oracle_tables):
    """The oracle's free_frequency_fit (scipy Nelder-Mead on the numpy mismatch) against
    what the unmodified reference returned (tests/golden/make_golden_cfg5.py)."""
    from oracle import qnmfits_oracle as orc
    g = golden("cfg5")
    wl = workloads.config5(n_waveforms=3, n_fixed=1)
    for b in range(3):
        w = orc.free_frequency_fit(oracle_tables, wl.times, wl.data[b], 0.0, modes=wl.modes, Mf=wl.Mf,
                                   chif=wl.chif)
        assert abs(w - g["fixed1_omega"][b]) < 1e-12


@pytest.mark.parametrize("N", [9, 10, 11, 12])
def test_wider_factors_use_shorter_blocks(qf, oracle_tables, N):
    """K1 beyond eight columns: 3-row blocks for N = 9, 10 and 2-row blocks for N = 11, 12
    (SmallLayout<N>::MB).  Full-rank label sets against the oracle for several lane splits,
    both mismatch paths, ragged windows and a start-time sweep."""
    from oracle import qnmfits_oracle as orc
    wl = workloads.config1()
    modes = [(2, 2, n, 1) for n in range(8)] + [(3, 2, n, 1) for n in range(N - 8)]
    w = np.array(qf.qnm.omega_list(modes, 0.69, 0.95))
    ref = orc.ringdown_fit(oracle_tables, wl.times, wl.data, modes, 0.95, 0.69, 0.0)
    assert int(ref["rank"]) == N
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    for lpf in (1, 4, 32):
        for uw in (0, 1):
            out = hs.run(wl.times, wl.data, n_fits=1, n_modes=N, window=win, t0=0.0, lpf=lpf,
                         omega=w.reshape(1, -1), omega_shared=True, uniform_weights=uw)
            err = np.max(np.abs(out["C"][0] - ref["C"])) / np.max(np.abs(ref["C"]))
            assert err < cases.amp_tol(ref["s"]), (lpf, uw, err)
            assert abs(out["mismatch"][0] - ref["mismatch"]) < 1e-10
            assert out["status"][0] == 0
    for M in (N + 1, 23, 64, 101):                       # windows that are no multiple of the block height
        rb, re = 500, 500 + M
        a, C, r, rank, s, model = orc.lstsq_fit(wl.times[rb:re], wl.data[rb:re], w, 0.0)
        want = orc.mismatch(wl.times[rb:re], model, wl.data[rb:re])
        for lpf in (1, 2, 8):
            out = hs.run(wl.times, wl.data, n_fits=1, n_modes=N, window=(rb, re), t0=0.0, lpf=lpf,
                         omega=w.reshape(1, -1), omega_shared=True)
            if rank == N and s[-1] > 1e-9 * s[0]:
                assert abs(out["mismatch"][0] - want) < 1e-10 and out["status"][0] == 0, (M, lpf)
            elif rank < N:                               # numpy truncates: the device must flag the fit
                assert out["status"][0] != 0, (M, lpf)
    t0s = np.linspace(-3.0, 12.0, 7)
    begin, end = api._window_rows_many(wl.times, t0s, 100 * np.ones(7), "geq")
    out = hs.run(wl.times, wl.data, n_fits=7, n_modes=N, window=(begin, end), t0=t0s, lpf=8,
                 omega=w.reshape(1, -1), omega_shared=True, uniform_weights=1)
    want = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, modes, 0.95, 0.69, t0s)
    np.testing.assert_allclose(out["mismatch"], want, rtol=0, atol=1e-10)


# ---------------------------------------------------------------------------------------
# K1p (csrc/fit_pair.cuh): the lanes of a row slice exchange block columns with shuffles inside
# the block loop, so the harness runs each warp in lock step (tests/hostsim/hostsim_warp.h), which
# also checks that every lane of a warp executes the same sequence of collectives.

def _damped_stack(N, S, K_tot, seed):
    rng = np.random.default_rng(seed)
    times = np.arange(K_tot) * 0.1
    freq = np.linspace(-0.1 * N, 0.1 * N, N) + 0.013 * rng.standard_normal(N) - 1j * (0.02 + 0.06 * rng.random(N))
    C = rng.standard_normal((S, N)) + 1j * rng.standard_normal((S, N))
    data = np.exp(-1j * np.outer(times - times[3], freq)) @ C.T
    data = data.T + 1e-5 * (rng.standard_normal((S, K_tot)) + 1j * rng.standard_normal((S, K_tot)))
    return times, data, freq


@pytest.mark.parametrize("order", [0, 1, 5])
@pytest.mark.parametrize("N,lpf", [(9, 2), (9, 8), (11, 4), (12, 4), (14, 8), (16, 4), (19, 16), (24, 8), (24, 32)])
def test_pair_kernel_emulated_in_lock_step_vs_numpy(N, lpf, order):
    """Seven fits sharing warps: windows of different lengths (lanes run blocks they have no rows
    for), per-fit start times and data series; second pass with model output, fast mismatch and
    the eval-only path, each fit against numpy lstsq on its explicit matrix.  Between collectives
    the lanes run one after another, in ascending, descending and shuffled order: a cross-lane
    shared-memory dependence that no collective orders fails in one of them."""
    times, data, freq = _damped_stack(N, 2, 360, seed=70 + N)
    rng = np.random.default_rng(N + lpf)
    B = 7
    rb = rng.integers(0, 40, B).astype(np.int32)
    re = (rb + rng.integers(3 * N, 300, B)).astype(np.int32)
    re[2] = rb[2] + N + 1                                # barely overdetermined
    t0 = times[rb] - 0.03
    which = rng.integers(0, 2, B).astype(np.int32)
    kw = dict(n_fits=B, n_modes=N, window=(rb, re), t0=t0, lpf=lpf, omega=freq.reshape(1, -1), omega_shared=True,
              series_index=which, pair=True, dt=0.1, order=order)
    out = hs.run(times, data, want_model=True, **kw)
    fast = hs.run(times, data, uniform_weights=1, **kw)
    ev = hs.run(times, data, eval_only=True, C_in=out["C"], **kw)
    checked = 0
    for b in range(B):
        sl = slice(rb[b], re[b])
        K = re[b] - rb[b]
        a, C_ref, res, rank, s, m_ref = orc.lstsq_fit(times[sl], data[which[b], sl], freq, t0[b], None)
        if rank < N:
            assert out["status"][b] & 1, (b, K)          # numpy truncates: the device must say so
            continue
        if out["status"][b] != 0:
            continue                                     # flagged with margin: the host would decide
        checked += 1
        assert np.max(np.abs(out["C"][b] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s), (b, K)
        mm_ref = orc.mismatch(times[sl], m_ref, data[which[b], sl])
        assert abs(out["mismatch"][b] - mm_ref) < 1e-10, (b, K)
        assert abs(fast["mismatch"][b] - out["mismatch"][b]) < 1e-11
        assert abs(ev["mismatch"][b] - out["mismatch"][b]) < 1e-12
        np.testing.assert_allclose(out["model"][b][:K], m_ref, rtol=0, atol=1e-8 * np.max(np.abs(C_ref)))
        Ad = np.column_stack([a, data[which[b], sl]])
        gram, want = out["R"][b].conj().T @ out["R"][b], Ad.conj().T @ Ad
        assert np.max(np.abs(gram[:N] - want[:N])) < 1e-9 * np.max(np.abs(want))
    assert checked >= 3


def test_pair_kernel_emulated_grid_matches_small_kernel():
    """The factored-frequency path (M-chi grid) through K1p at nine columns against K1 on the same
    fits: same rows, same reflections in another distribution over lanes -> agreement to rounding."""
    wl = workloads.config3(res=4)
    import qnmfits_b200 as qf
    modes = [(2, 2, n, 1) for n in range(8)] + [(3, 2, 0, 1)]
    Mf = np.linspace(*wl.Mf_minmax, 4)
    chi = np.linspace(*wl.chif_minmax, 4)
    table, ptr = qf.qnm.constituent_table(modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    kw = dict(n_fits=16, n_modes=9, window=win, t0=0.0, table=table, mode_ptr=ptr, inv_Mf=1.0 / Mf, n_chi=4, n_mf=4,
              dt=0.1, uniform_weights=1)
    pair = hs.run(wl.times, wl.data, lpf=8, pair=True, **kw)
    small = hs.run(wl.times, wl.data, lpf=8, **kw)
    assert np.all(pair["status"] == 0) and np.all(small["status"] == 0)
    np.testing.assert_allclose(pair["mismatch"], small["mismatch"], rtol=0, atol=1e-12)


# ---------------------------------------------------------------------------------------
# K3 (csrc/fit_struct.cuh) and K4 (csrc/fit_panel.cuh): the kernel functions themselves, one emulated
# CTA per fit — all threads as fibers, warp shuffles / mma.sync / __syncthreads emulated, the threads
# resumed in ascending, descending and shuffled order between barriers (the stand-in for racecheck, which
# the GPU pool does not offer).

def _stack(N, L, K_tot, seed, uniform=True):
    rng = np.random.default_rng(seed)
    times = np.arange(K_tot) * 0.1 if uniform else np.cumsum(0.05 + 0.1 * rng.random(K_tot))
    freq = np.linspace(-0.1 * N, 0.1 * N, N) + 0.013 * rng.standard_normal(N) - 1j * (0.02 + 0.06 * rng.random(N))
    coef = rng.standard_normal((L, N)) + 1j * rng.standard_normal((L, N))
    C = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    E = np.exp(-1j * np.outer(times - times[3], freq))
    data = np.stack([E @ (coef[i] * C) for i in range(L)])
    data += 1e-5 * (rng.standard_normal(data.shape) + 1j * rng.standard_normal(data.shape))
    return times, data, freq, coef


@pytest.mark.parametrize("order", [0, 1, 5])
@pytest.mark.parametrize("kernel", ["k3", "k4"])
@pytest.mark.parametrize("N,L,use_coef", [(3, 2, True), (10, 3, True), (12, 1, False), (9, 7, True), (20, 5, True),
                                          (40, 21, True), (44, 21, True)])
def test_struct_kernel_emulated_cta_vs_numpy(N, L, use_coef, kernel, order):
    """Structured two-phase QR — K3, and K4 (blocked, trailing update on mma.sync.m8n8k4.f64, emulated
    with the PTX fragment layout) — on uniform grids (fast mismatch and second pass, model output,
    eval-only) and on a non-uniform grid (direct evaluation), against numpy lstsq on the explicit
    stacked matrix."""
    if kernel == "k3" and N + L > 64:
        pytest.skip("K3 takes at most 64 columns")
    for uniform in (True, False):
        K_tot = 330 if N < 40 else 200
        times, data, freq, coef = _stack(N, L, K_tot, seed=N * 100 + L, uniform=uniform)
        rb, re, t0 = 3, K_tot - 14, float(times[3])      # a row count that is no multiple of the tile height
        K = re - rb
        a, C_ref, res_ref, rank, s, model = orc.lstsq_fit(times[rb:re], data[:, rb:re].reshape(-1), freq, t0,
                                                          coef if (use_coef or L > 1) else None)
        assert rank == N
        mm_ref = orc.multimode_mismatch(times[rb:re], {i: model[i * K:(i + 1) * K] for i in range(L)},
                                        {i: data[i, rb:re] for i in range(L)})
        kw = dict(n_fits=1, n_modes=N, window=(rb, re), t0=t0, omega=freq, coef=coef if (use_coef or L > 1) else None,
                  dt=0.1 if uniform else 0.0, order=order, panel=kernel == "k4")
        out = hs.run_struct(times, data, want_model=True, **kw)
        assert out["status"][0] == 0
        assert np.max(np.abs(out["C"][0] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s), (uniform,)
        assert abs(out["mismatch"][0] - mm_ref) < 1e-10
        np.testing.assert_allclose(out["residual"][0], res_ref[0], rtol=1e-6)
        got = np.concatenate([out["model"][0][i * K:(i + 1) * K] for i in range(L)])
        np.testing.assert_allclose(got, model, rtol=0, atol=1e-8 * np.max(np.abs(C_ref)))
        ev = hs.run_struct(times, data, eval_only=True, C_in=out["C"], **kw)
        assert abs(ev["mismatch"][0] - out["mismatch"][0]) < 1e-12
        if uniform:
            fast = hs.run_struct(times, data, uniform_weights=1, **kw)
            assert abs(fast["mismatch"][0] - mm_ref) < 1e-10


@pytest.mark.parametrize("panel", [False, True])
def test_struct_kernel_emulated_sweep_with_ragged_windows(panel):
    """Several fits, each with its own window and start time (a t0 sweep of a multimode fit)."""
    N, L, B = 10, 3, 6
    times, data, freq, coef = _stack(N, L, 300, seed=7)
    rng = np.random.default_rng(3)
    rb = rng.integers(0, 40, B).astype(np.int32)
    re = (rb + rng.integers(90, 250, B)).astype(np.int32)
    t0 = times[rb] - 0.03
    out = hs.run_struct(times, data, n_fits=B, n_modes=N, window=(rb, re), t0=t0, omega=freq, coef=coef, dt=0.1,
                        panel=panel)
    for b in range(B):
        sl = slice(rb[b], re[b])
        K = re[b] - rb[b]
        a, C_ref, res, rank, s, m_ref = orc.lstsq_fit(times[sl], data[:, sl].reshape(-1), freq, t0[b], coef)
        assert np.max(np.abs(out["C"][b] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s)
        mm_ref = orc.multimode_mismatch(times[sl], {i: m_ref[i * K:(i + 1) * K] for i in range(L)},
                                        {i: data[i, sl] for i in range(L)})
        assert abs(out["mismatch"][b] - mm_ref) < 1e-10


@pytest.mark.parametrize("order", [0, 1, 5])
@pytest.mark.parametrize("N,L", [(3, 2), (10, 3), (12, 1), (17, 3)])
def test_general_kernel_emulated_cta_vs_numpy(N, L, order):
    """K2 (streamed dense Householder, csrc/fit_general.cuh) with a constant mixing table, and with
    per-sample frequencies and mixing coefficients (the dynamic multimode fit, the one case only K2
    takes), against numpy lstsq on the explicit matrix."""
    times, data, freq, coef = _stack(N, L, 260, seed=N * 10 + L)
    rb, re, t0 = 3, 241, float(times[3])
    K = re - rb
    a, C_ref, res_ref, rank, s, model = orc.lstsq_fit(times[rb:re], data[:, rb:re].reshape(-1), freq, t0, coef)
    mm_ref = orc.multimode_mismatch(times[rb:re], {i: model[i * K:(i + 1) * K] for i in range(L)},
                                    {i: data[i, rb:re] for i in range(L)})
    out = hs.run_struct(times, data, n_fits=1, n_modes=N, window=(rb, re), t0=t0, omega=freq, coef=coef, dt=0.0,
                        general=True, order=order)
    assert out["status"][0] == 0
    assert np.max(np.abs(out["C"][0] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s)
    assert abs(out["mismatch"][0] - mm_ref) < 1e-10
    # dynamic spectrum: frequencies and coefficients drift with the sample
    drift = 1.0 + 0.02 * np.exp(-np.arange(len(times)) / 60.0)
    omega_rows = freq[:, None] * drift[None, :]
    coef_rows = coef[:, :, None] * (1.0 + 0.05j * (drift[None, None, :] - 1.0))
    tau = times[rb:re] - t0
    A = np.concatenate([coef_rows[i][:, rb:re].T * np.exp(-1j * omega_rows[:, rb:re].T * tau[:, None]) for i in range(L)])
    d = data[:, rb:re].reshape(-1)
    C_dyn, _, rank, sv = np.linalg.lstsq(A, d, rcond=None)
    m_dyn = A @ C_dyn
    mm_dyn = orc.multimode_mismatch(times[rb:re], {i: m_dyn[i * K:(i + 1) * K] for i in range(L)},
                                    {i: data[i, rb:re] for i in range(L)})
    dyn = hs.run_struct(times, data, n_fits=1, n_modes=N, window=(rb, re), t0=t0, omega=freq, coef=coef, dt=0.0,
                        general=True, order=order, omega_rows=omega_rows, coef_rows=coef_rows)
    assert np.max(np.abs(dyn["C"][0] - C_dyn)) / np.max(np.abs(C_dyn)) < cases.amp_tol(sv)
    assert abs(dyn["mismatch"][0] - mm_dyn) < 1e-10


@pytest.mark.parametrize("order", [0, 1, 5])
@pytest.mark.parametrize("staged", [False, True])
@pytest.mark.parametrize("N,lpf", [(1, 1), (5, 8), (8, 4)])
def test_small_kernel_function_on_emulated_cta(N, lpf, staged, order):
    """K1's kernel function itself (fit_small_kernel: staging of the window, table fill,
    __syncthreads, R-combine behind __syncwarp, butterflies) on an emulated CTA, threads resumed in
    either order: a sweep with ragged windows must reproduce, bit for bit, what the stage-by-stage
    harness above computes, and agree with numpy."""
    times, data, freq = _damped_stack(N, 1, 420, seed=30 + N)
    rng = np.random.default_rng(N * 7 + lpf)
    B = 9
    rb = rng.integers(0, 40, B).astype(np.int32)
    re = (rb + rng.integers(4 * N + 8, 330, B)).astype(np.int32)
    t0 = times[rb] - 0.03
    kw = dict(n_fits=B, n_modes=N, window=(rb, re), t0=t0, lpf=lpf, omega=freq.reshape(1, -1), omega_shared=True, dt=0.1)
    for fast in (0, 1):
        ref = hs.run(times, data[0], uniform_weights=fast, **kw)
        out = hs.run(times, data[0], uniform_weights=fast, cta=True, staged=staged, order=order, **kw)
        assert np.array_equal(out["mismatch"], ref["mismatch"]) and np.array_equal(out["C"], ref["C"])
        assert np.array_equal(out["status"], ref["status"])
    for b in range(B):
        sl = slice(rb[b], re[b])
        a, C_ref, res, rank, s, m_ref = orc.lstsq_fit(times[sl], data[0, sl], freq, t0[b], None)
        assert np.max(np.abs(out["C"][b] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s)
        assert abs(out["mismatch"][b] - orc.mismatch(times[sl], m_ref, data[0, sl])) < 1e-10
