"""Parity of the CUDA path (through the public API -> ctypes -> libqnmfit.so) with the
oracle (numpy lstsq restatement of the reference) and with the golden fixtures the
unmodified reference produced.  Tolerances (BASELINE.json north_star):
amplitudes max(1e-8, 100*cond*eps) relative, mismatch 1e-10 absolute, discrete
choices (frequencies, window rows, rank) exact."""
import warnings

import numpy as np
import pytest

import cases
from oracle import qnmfits_oracle as orc
from qnmfits_b200 import _cabi, workloads

pytestmark = pytest.mark.gpu

MM_TOL = 1e-10


@pytest.fixture(scope="module")
def eng(qf):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from qnmfits_b200._engine import get_engine
    return get_engine()


def test_library_is_loaded_and_device_is_blackwell(eng):
    import torch
    assert torch.cuda.get_device_capability(0)[0] == 10
    assert eng.ctx.handle


def test_ringdown_fit_cases_vs_golden(qf, eng, golden):
    g = golden("cfg1")
    wl, cs = cases.cfg1_cases()
    for name, kw in cs.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fit = qf.ringdown_fit(wl.times, wl.data, **kw)
        assert list(fit.keys()) == ['residual', 'rank', 's', 'mismatch', 'C', 'data', 'model',
                                    'model_times', 't0', 'modes', 'mode_labels', 'frequencies']
        assert np.array_equal(fit["frequencies"], g[name + "__frequencies"]), name
        assert np.array_equal(fit["model_times"], g[name + "__model_times"]), name
        assert int(fit["rank"]) == int(g[name + "__rank"]), name
        assert fit["residual"].shape == g[name + "__residual"].shape, name
        s = g[name + "__s"]
        np.testing.assert_allclose(fit["s"], s, rtol=1e-8, atol=1e-12 * s[0], err_msg=name)
        C_ref = g[name + "__C"]
        if int(g[name + "__rank"]) == len(C_ref):
            err = np.max(np.abs(fit["C"] - C_ref)) / np.max(np.abs(C_ref))
            assert err < cases.amp_tol(s), (name, err)
        else:
            # numpy truncated a singular value: the host completes the minimum-norm
            # solution from the device factor; compare what is well defined
            assert abs(np.linalg.norm(fit["C"]) - np.linalg.norm(C_ref)) < 1e-6 * np.linalg.norm(C_ref)
        assert abs(fit["mismatch"] - float(g[name + "__mismatch"])) < MM_TOL, name
        scale = np.max(np.abs(g[name + "__model"]))
        np.testing.assert_allclose(fit["model"], g[name + "__model"], rtol=0, atol=1e-8 * scale,
                                   err_msg=name)
        if fit["residual"].size:
            np.testing.assert_allclose(fit["residual"], g[name + "__residual"], rtol=1e-7)


def test_g1_injection_recovery(qf, eng, golden):
    g = golden("g1")
    f0 = qf.ringdown_fit(g["times"], g["data"], [(2, 2, 0, 1)], 1, 0.7, 0)
    f10 = qf.ringdown_fit(g["times"], g["data"], [(2, 2, 0, 1)], 1, 0.7, 10)
    assert abs(f0["C"][0] - (1 - 1j)) < 1e-12 and abs(f0["mismatch"]) < 1e-13
    np.testing.assert_allclose(f10["C"], g["C10"], rtol=1e-11)


def test_nonuniform_grid_direct_path(qf, eng, golden):
    g = golden("nonuniform")
    fit = qf.ringdown_fit(g["times"], g["data"], workloads.overtone_modes(5), 0.95, 0.69, 1.0, T=60)
    assert cases.rel_err(fit["C"], g["fit__C"]) < 1e-8
    assert abs(fit["mismatch"] - float(g["fit__mismatch"])) < MM_TOL


def test_t0_sweep_vs_golden_and_oracle(qf, eng, golden, oracle_tables):
    g = golden("cfg2")
    wl = workloads.config2(n_t0=40)
    mm = qf.mismatch_t0_array(wl.times, wl.data, wl.modes, 0.95, 0.69, wl.t0_array)
    assert isinstance(mm, list) and isinstance(mm[0], np.float64)
    np.testing.assert_allclose(mm, g["mismatch"], rtol=0, atol=MM_TOL)
    mmc = qf.mismatch_t0_array(wl.times, wl.data, wl.modes[:4], 0.95, 0.69, wl.t0_array[:10],
                               t0_method="closest", T_array=np.linspace(50, 80, 10))
    np.testing.assert_allclose(mmc, g["mismatch_closest"], rtol=0, atol=MM_TOL)
    # config 2 at full size: 1000 start times vs the oracle
    wl = workloads.config2()
    mm = qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array)
    want = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif,
                                 wl.t0_array)
    np.testing.assert_allclose(mm, want, rtol=0, atol=MM_TOL)


def test_grid_vs_golden(qf, eng, golden):
    g = golden("cfg3")
    wl = workloads.config3(res=12)
    grid = qf.mismatch_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0,
                                  T=wl.T, res=12)
    assert grid.shape == (12, 12) and grid.dtype == np.float64
    np.testing.assert_allclose(grid, g["grid"], rtol=0, atol=MM_TOL)
    gq = qf.mismatch_M_chi_grid(wl.times, wl.data, [(2, 2, 0, 1), (2, 2, 1, 1), (2, 2, 0, 1, 2, 2, 0, 1)],
                                wl.Mf_minmax, wl.chif_minmax, 15.0, T=60, res=5,
                                delta=[0.0, 0.01, 0.0])
    np.testing.assert_allclose(gq, g["grid_quadratic"], rtol=0, atol=MM_TOL)


def test_full_size_grid_properties(qf, eng, oracle_tables):
    """Config 3 at BASELINE size (256 x 256 = 65 536 fits): spot checks against the
    oracle at random grid points, orientation [iMf, ichi], minimum near the truth, and
    index-sharding invariance (slabs computed separately are bit-identical)."""
    wl = workloads.config3(res=256)
    grid = qf.mismatch_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0,
                                  T=wl.T, res=256)
    assert grid.shape == (256, 256) and np.all(np.isfinite(grid))
    rng = np.random.default_rng(5)
    idx = rng.choice(256 * 256, size=96, replace=False)
    want = orc.mismatch_M_chi_grid(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf_minmax,
                                   wl.chif_minmax, wl.t0, T=wl.T, res=256, flat_indices=idx)
    np.testing.assert_allclose(grid.reshape(-1)[idx], want, rtol=0, atol=MM_TOL)
    i, j = np.unravel_index(np.argmin(grid), grid.shape)
    Mf = np.linspace(*wl.Mf_minmax, 256)[i]
    chi = np.linspace(*wl.chif_minmax, 256)[j]
    assert abs(Mf - workloads.MF_TRUE) < 2e-3 and abs(chi - workloads.CHIF_TRUE) < 2e-3
    again = qf.mismatch_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax,
                                   wl.t0, T=wl.T, res=256)
    assert np.array_equal(grid, again)          # deterministic


@pytest.mark.parametrize("res,slabs", [
    (24, ((0, 100), (100, 101), (101, 400), (400, 576))),
    (256, tuple((r * 8192, (r + 1) * 8192) for r in range(8))),          # the 8-GPU plan of the headline grid
    (256, ((0, 32768), (32768, 65536))),                                 # the 2-GPU plan
    (96, ((0, 1), (1, 2305), (2305, 9216))),
])
def test_slab_launches_are_bit_identical_to_one_launch(qf, eng, res, slabs):
    """What each rank of an N-GPU job computes (first_fit offset, plan_fits = the whole sweep)
    equals the corresponding slab of the single launch BIT FOR BIT: the lanes-per-fit split
    (the reduction tree of a fit) is chosen from the sweep, not from the slab
    (the multi-GPU grid is bit-identical to the 1-GPU grid)."""
    import torch
    from qnmfits_b200 import qnmfits as api
    wl = workloads.config3(res=res)
    Mf = np.linspace(*wl.Mf_minmax, res)
    chi = np.linspace(*wl.chif_minmax, res)
    table, ptr = qf.qnm.constituent_table(wl.modes, chi)
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    d = dict(times_d=eng.to_device(wl.times, np.float64),
             data_d=eng.to_device(wl.data.reshape(1, -1), np.complex128),
             omega_tilde_d=eng.to_device(table, np.complex128), mode_ptr_d=eng.to_device(ptr, np.int32),
             inv_Mf_d=eng.to_device(1.0 / Mf, np.float64), n_chi=res, n_mf=res,
             n_constituents=table.shape[1], n_modes=8, row_begin_all=win[0], row_end_all=win[1],
             t0_all=0.0, dt_nominal=0.1)
    n = res * res
    full = eng.empty((n,), torch.float64)
    eng.fit(eng.make_batch(n_fits=n, mismatch_d=full, **d))
    parts = eng.empty((n,), torch.float64)
    lanes, blocks = set(), set()
    for lo, hi in slabs:
        b = eng.make_batch(n_fits=hi - lo, first_fit=lo, plan_fits=n, mismatch_d=parts[lo:hi], **d)
        lanes.add(eng.ctx.plan(b).lanes_per_fit)
        blocks.add(eng.ctx.plan(b).block)
        eng.fit(b)
    eng.synchronize()
    whole = eng.ctx.plan(eng.make_batch(n_fits=n, mismatch_d=full, **d))
    assert lanes == {whole.lanes_per_fit}
    if res == 256 and torch.cuda.get_device_properties(0).multi_processor_count == 148:
        # the block size IS chosen per slab (seven warps fill 147 SMs with an eighth of the grid)
        assert whole.block == 256 and blocks == {224}
    assert torch.equal(full, parts)


def test_general_kernel_matches_small_kernel_and_oracle(qf, eng, oracle_tables):
    """K2 (CTA per fit) forced on a K1-sized problem."""
    import torch
    from qnmfits_b200 import qnmfits as api
    wl = workloads.config1()
    freq = np.array(qf.qnm.omega_list(wl.modes, 0.69, 0.95))
    win = api._window_rows(wl.times, 0.0, 100, "geq")
    d = dict(times_d=eng.to_device(wl.times, np.float64),
             data_d=eng.to_device(wl.data.reshape(1, -1), np.complex128),
             omega_d=eng.to_device(freq.reshape(1, -1), np.complex128), omega_shared=True,
             n_fits=1, n_modes=8, row_begin_all=win[0], row_end_all=win[1], t0_all=0.0)
    outs = {}
    for name, kernel, dt in (("k1", _cabi.KERNEL_SMALL, 0.1), ("k2", _cabi.KERNEL_GENERAL, 0.0)):
        C_d = eng.empty((1, 8), torch.complex128)
        mm_d = eng.empty((1,), torch.float64)
        eng.fit(eng.make_batch(kernel=kernel, dt_nominal=dt, C_d=C_d, mismatch_d=mm_d, **d))
        outs[name] = (eng.to_host(C_d)[0], float(eng.to_host(mm_d)[0]))
    want = orc.ringdown_fit(oracle_tables, wl.times, wl.data, wl.modes, 0.95, 0.69, 0.0)
    for name, (C, mm) in outs.items():
        err = np.max(np.abs(C - want["C"])) / np.max(np.abs(want["C"]))
        assert err < cases.amp_tol(want["s"]), (name, err)
        assert abs(mm - want["mismatch"]) < MM_TOL, name


def test_multimode_vs_golden(qf, eng, golden):
    g = golden("cfg4")
    wl = cases.cfg4_small()
    fit = qf.multimode_ringdown_fit(wl.times, wl.data, cases.MM_MODES, 0.95, 0.69, 5.0, T=80)
    assert list(fit.keys()) == ['residual', 'mismatch', 'C', 'weighted_C', 'data', 'model',
                                'model_times', 't0', 'modes', 'mode_labels', 'frequencies']
    assert np.array_equal(fit["frequencies"], g["frequencies"])
    err = np.max(np.abs(fit["C"] - g["C"])) / np.max(np.abs(g["C"]))
    assert err < 1e-8, err
    assert abs(fit["mismatch"] - float(g["mismatch"])) < MM_TOL
    np.testing.assert_allclose(fit["residual"], g["residual"], rtol=1e-7)
    for lm in cases.MM_SPH:
        m_ref = g[f"model_{lm[0]}_{lm[1]}"]
        np.testing.assert_allclose(fit["model"][lm], m_ref, rtol=0, atol=1e-8 * np.max(np.abs(g["C"])))
        np.testing.assert_allclose(fit["weighted_C"][lm], g[f"weighted_C_{lm[0]}_{lm[1]}"],
                                   rtol=0, atol=1e-8 * np.max(np.abs(g["C"])))
    sub = qf.multimode_ringdown_fit(wl.times, wl.data, cases.MM_MODES[:4], 0.95, 0.69, 5.0, T=80,
                                    spherical_modes=[(2, 2), (3, 2)])
    assert np.max(np.abs(sub["C"] - g["sub_C"])) / np.max(np.abs(g["sub_C"])) < 1e-8
    assert abs(sub["mismatch"] - float(g["sub_mismatch"])) < MM_TOL
    sweep = qf.mismatch_t0_array(wl.times, wl.data, cases.MM_MODES, 0.95, 0.69, wl.t0_array,
                                 T_array=70)
    np.testing.assert_allclose(sweep, g["t0_sweep"], rtol=0, atol=MM_TOL)
    grid = qf.mismatch_M_chi_grid(wl.times, wl.data, cases.MM_MODES, (0.9, 1.0), (0.6, 0.75), 5.0,
                                  T=80, res=4)
    np.testing.assert_allclose(grid, g["grid"], rtol=0, atol=MM_TOL)
    with pytest.raises(ValueError):
        qf.multimode_ringdown_fit(wl.times, wl.data, [(2, 2, 0, 1, 2, 2, 0, 1)], 0.95, 0.69, 5.0)


def test_config4_shape_vs_oracle(qf, eng, oracle_tables):
    """21 spherical modes x 40 QNMs (config 4's shape), three start times vs the oracle."""
    wl = workloads.config4(n_t0=3)
    got = qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array,
                               T_array=wl.T, spherical_modes=wl.spherical_modes)
    want = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif,
                                 wl.t0_array, T_array=wl.T, spherical_modes=wl.spherical_modes)
    np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)


def test_errors_are_reported_not_swallowed(qf, eng):
    import torch
    b = eng.make_batch(times_d=eng.empty((4,), torch.float64), data_d=eng.empty((1, 4), torch.complex128),
                       n_fits=1, n_modes=99, row_begin_all=0, row_end_all=4,
                       mismatch_d=eng.empty((1,), torch.float64))
    with pytest.raises(_cabi.QnmfitError) as e:
        eng.fit(b)
    assert e.value.code == -2 and "n_modes" in str(e.value)
    with pytest.raises(ValueError):
        qf.ringdown_fit(np.linspace(0, 1, 5), np.ones(5, complex), [(2, 2, 0, 1)], 1.0, 0.5, 0.0,
                        t0_method="nearest")


def _synthetic_stack(N, L, K_tot, seed, uniform=True):
    """Well-separated damped sinusoids, random mixing table, injected model + noise."""
    rng = np.random.default_rng(seed)
    if uniform:
        times = np.arange(K_tot) * 0.1
    else:
        times = np.cumsum(0.05 + 0.1 * rng.random(K_tot))
    freq = np.linspace(-0.1 * N, 0.1 * N, N) + 0.013 * rng.standard_normal(N) - 1j * (0.02 + 0.06 * rng.random(N))
    coef = rng.standard_normal((L, N)) + 1j * rng.standard_normal((L, N))
    C = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    E = np.exp(-1j * np.outer(times - times[3], freq))
    data = np.stack([E @ (coef[i] * C) for i in range(L)])
    data += 1e-5 * (rng.standard_normal(data.shape) + 1j * rng.standard_normal(data.shape))
    return times, data, freq, coef


@pytest.mark.parametrize("N,L,use_coef", [(3, 2, True), (12, 1, False), (12, 1, True), (20, 5, True),
                                           (40, 21, True), (63, 1, False), (9, 7, True), (9, 1, False),
                                           (10, 1, False), (11, 1, False), (16, 1, False), (13, 1, False), (14, 1, False),
                                           (15, 1, False), (17, 1, False), (20, 1, False), (24, 1, False), (1, 1, True),
                                           (8, 2, True), (17, 3, True), (44, 21, True), (64, 30, True)])
def test_struct_kernel_vs_oracle_and_general_kernel(qf, eng, N, L, use_coef):
    """K4 (blocked structured QR on DMMA) and K3 (structured two-phase QR), uniform (fast
    mismatch and second pass) and non-uniform grids, against numpy lstsq on the explicit
    stacked matrix and against K2; shapes beyond K3's 64 columns run K4 only."""
    import torch
    K_tot = 700
    for uniform in (True, False):
        times, data, freq, coef = _synthetic_stack(N, L, K_tot, seed=N * 100 + L, uniform=uniform)
        rb, re, t0 = 3, 3 + 613, float(times[3])          # 613 rows: not a multiple of any tile height
        a, C_ref, res_ref, rank, s, model = orc.lstsq_fit(times[rb:re], data[:, rb:re].reshape(-1), freq, t0,
                                                          coef if (use_coef or L > 1) else None)
        assert rank == N
        d_m = {i: data[i, rb:re] for i in range(L)}
        m_m = {i: model[i * (re - rb):(i + 1) * (re - rb)] for i in range(L)}
        mm_ref = orc.multimode_mismatch(times[rb:re], m_m, d_m)
        d = dict(times_d=eng.to_device(times, np.float64), data_d=eng.to_device(data, np.complex128),
                 omega_d=eng.to_device(freq.reshape(1, -1), np.complex128), omega_shared=True,
                 n_fits=1, n_modes=N, n_series=L, row_begin_all=rb, row_end_all=re, t0_all=t0)
        if use_coef or L > 1:
            d.update(coef_d=eng.to_device(coef.reshape(1, L, N), np.complex128), n_coef=1,
                     coef_index_d=eng.to_device(np.zeros(1, np.int32), np.int32))
        tol = cases.amp_tol(s)
        variants = [("k4_second_pass", _cabi.KERNEL_PANEL, False), ("k2", _cabi.KERNEL_GENERAL, False)]
        if N + L <= 64:
            variants.append(("k3_second_pass", _cabi.KERNEL_STRUCT, False))
        if uniform:
            variants.insert(0, ("k4_fast", _cabi.KERNEL_PANEL, True))
            if N + L <= 64:
                variants.insert(0, ("k3_fast", _cabi.KERNEL_STRUCT, True))
        if N <= _cabi.MAX_MODES_SMALL and L == 1 and not use_coef:   # K1 with 2- or 3-row blocks
            variants.append(("k1_second_pass", _cabi.KERNEL_SMALL, False))
            if uniform:
                variants.append(("k1_fast", _cabi.KERNEL_SMALL, True))
        pair = _cabi.MIN_MODES_PAIR <= N <= _cabi.MAX_MODES_PAIR and L == 1 and not use_coef
        if pair:                                                      # K1p: columns split over 2 / 4 lanes
            variants.append(("k1p_second_pass", _cabi.KERNEL_PAIR, False))
            if uniform:
                variants.append(("k1p_fast", _cabi.KERNEL_PAIR, True))
        for name, kernel, fast in variants:
            C_d = eng.empty((1, N), torch.complex128)
            mm_d = eng.empty((1,), torch.float64)
            res_d = eng.empty((1,), torch.float64)
            st_d = eng.empty((1,), torch.int32)
            eng.fit(eng.make_batch(kernel=kernel, dt_nominal=0.1 if uniform else 0.0, uniform_weights=fast,
                                   C_d=C_d, mismatch_d=mm_d, residual_d=res_d, status_d=st_d, **d))
            C = eng.to_host(C_d)[0]
            err = np.max(np.abs(C - C_ref)) / np.max(np.abs(C_ref))
            assert err < tol, (name, uniform, err, tol)
            assert abs(float(eng.to_host(mm_d)[0]) - mm_ref) < MM_TOL, (name, uniform)
            np.testing.assert_allclose(eng.to_host(res_d)[0], res_ref[0], rtol=1e-6, err_msg=name)
            assert int(eng.to_host(st_d)[0]) == 0, name
        plan = eng.ctx.plan(eng.make_batch(kernel=_cabi.KERNEL_AUTO, mismatch_d=mm_d, **d))
        auto_pair = pair
        want = _cabi.KERNEL_PAIR if auto_pair else _cabi.KERNEL_SMALL
        if not auto_pair and (N > _cabi.MAX_MODES_SMALL or L > 1 or use_coef):
            # K4 takes over where K3 ends, and on the large stacked shapes (cfg4) where it is faster
            want = _cabi.KERNEL_STRUCT if N + L <= 64 and not (N >= 36 and L >= 8) else _cabi.KERNEL_PANEL
        assert plan.kernel == want


@pytest.mark.parametrize("kernel", [_cabi.KERNEL_STRUCT, _cabi.KERNEL_PANEL])
def test_struct_kernel_many_fits_windows_and_eval(qf, eng, kernel):
    """K3 / K4 on a sweep: per-fit windows and start times, model output, eval-only path."""
    import torch
    N, L, K_tot, B = 10, 3, 500, 37
    times, data, freq, coef = _synthetic_stack(N, L, K_tot, seed=7)
    rng = np.random.default_rng(3)
    rb = rng.integers(0, 60, B).astype(np.int32)
    re = (rb + rng.integers(200, 400, B)).astype(np.int32)
    t0 = times[rb] - 0.03
    Kmax = int(re.max() - rb.min())     # the library sizes model rows by the union window
    d = dict(times_d=eng.to_device(times, np.float64), data_d=eng.to_device(data, np.complex128),
             omega_d=eng.to_device(freq.reshape(1, -1), np.complex128), omega_shared=True,
             coef_d=eng.to_device(coef.reshape(1, L, N), np.complex128), n_coef=1,
             coef_index_d=eng.to_device(np.zeros(B, np.int32), np.int32),
             n_fits=B, n_modes=N, n_series=L, row_begin_all=int(rb.min()), row_end_all=int(re.max()),
             row_begin_d=eng.to_device(rb, np.int32), row_end_d=eng.to_device(re, np.int32),
             t0_d=eng.to_device(t0, np.float64), dt_nominal=0.1, kernel=kernel)
    C_d = eng.empty((B, N), torch.complex128)
    mm_d = eng.empty((B,), torch.float64)
    model_d = eng.empty((B, L * Kmax), torch.complex128)
    eng.fit(eng.make_batch(C_d=C_d, mismatch_d=mm_d, model_d=model_d, model_stride=L * Kmax, **d))
    mm_fast = eng.empty((B,), torch.float64)
    eng.fit(eng.make_batch(mismatch_d=mm_fast, uniform_weights=True, **d))
    mm_eval = eng.empty((B,), torch.float64)
    eng.evaluate(eng.make_batch(C_d=C_d, mismatch_d=mm_eval, **d))
    C, mm, model = eng.to_host(C_d), eng.to_host(mm_d), eng.to_host(model_d)
    for b in range(B):
        sl = slice(rb[b], re[b])
        K = re[b] - rb[b]
        a, C_ref, res, rank, s, m_ref = orc.lstsq_fit(times[sl], data[:, sl].reshape(-1), freq, t0[b], coef)
        assert np.max(np.abs(C[b] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s)
        mm_ref = orc.multimode_mismatch(times[sl], {i: m_ref[i * K:(i + 1) * K] for i in range(L)},
                                        {i: data[i, sl] for i in range(L)})
        assert abs(mm[b] - mm_ref) < MM_TOL
        got = np.concatenate([model[b, i * K:(i + 1) * K] for i in range(L)])
        np.testing.assert_allclose(got, m_ref, rtol=0, atol=1e-8 * np.max(np.abs(C_ref)))
    np.testing.assert_allclose(eng.to_host(mm_fast), mm, rtol=0, atol=1e-11)
    np.testing.assert_allclose(eng.to_host(mm_eval), mm, rtol=0, atol=1e-12)


@pytest.mark.parametrize("N", [9, 12, 14, 16, 19, 24])
def test_pair_kernel_many_fits_windows_series_and_eval(qf, eng, N):
    """K1p (columns of a row slice split over 2 / 4 lanes) on a sweep of 41 fits sharing warps:
    per-fit windows of different lengths (so lanes run blocks they have no rows for), per-fit
    start times, per-fit data series (``series_index``), model output, the fast mismatch and
    the eval-only path — each fit against numpy lstsq on its explicit matrix."""
    import torch
    K_tot, B, n_series = 520, 41, 3
    times, data, freq, _ = _synthetic_stack(N, n_series, K_tot, seed=50 + N)
    rng = np.random.default_rng(N)
    rb = rng.integers(0, 60, B).astype(np.int32)
    re = (rb + rng.integers(40, 440, B)).astype(np.int32)
    re[3] = rb[3] + N + 1                               # barely overdetermined
    t0 = times[rb] - 0.03
    which = rng.integers(0, n_series, B).astype(np.int32)
    Kmax = int(re.max() - rb.min())
    d = dict(times_d=eng.to_device(times, np.float64), data_d=eng.to_device(data, np.complex128),
             omega_d=eng.to_device(freq.reshape(1, -1), np.complex128), omega_shared=True,
             series_index_d=eng.to_device(which, np.int32), series_stride=K_tot,
             n_fits=B, n_modes=N, n_series=1, row_begin_all=int(rb.min()), row_end_all=int(re.max()),
             row_begin_d=eng.to_device(rb, np.int32), row_end_d=eng.to_device(re, np.int32),
             t0_d=eng.to_device(t0, np.float64), dt_nominal=0.1, kernel=_cabi.KERNEL_PAIR)
    C_d = eng.empty((B, N), torch.complex128)
    mm_d = eng.empty((B,), torch.float64)
    st_d = eng.empty((B,), torch.int32)
    model_d = eng.empty((B, Kmax), torch.complex128)
    R_d = eng.empty((B, N, N + 1), torch.complex128)
    eng.fit(eng.make_batch(C_d=C_d, mismatch_d=mm_d, model_d=model_d, model_stride=Kmax, status_d=st_d, R_d=R_d, **d))
    mm_fast = eng.empty((B,), torch.float64)
    eng.fit(eng.make_batch(mismatch_d=mm_fast, uniform_weights=True, **d))
    mm_eval = eng.empty((B,), torch.float64)
    eng.evaluate(eng.make_batch(C_d=C_d, mismatch_d=mm_eval, **d))
    C, mm, model, st = eng.to_host(C_d), eng.to_host(mm_d), eng.to_host(model_d), eng.to_host(st_d)
    R = eng.to_host(R_d)
    for b in (0, 3, B - 1):                             # the exported factor [R | Q^H d]: what the host repair reads
        sl = slice(rb[b], re[b])
        A = np.exp(-1j * np.outer(times[sl] - t0[b], freq))
        Ad = np.column_stack([A, data[which[b], sl]])
        assert np.allclose(np.tril(R[b, :, :N], -1), 0) and np.allclose(np.diagonal(R[b, :, :N]).imag, 0)
        gram, want = R[b].conj().T @ R[b], Ad.conj().T @ Ad
        assert np.max(np.abs(gram[:N] - want[:N])) < 1e-9 * np.max(np.abs(want))
    checked = 0
    for b in range(B):
        sl = slice(rb[b], re[b])
        K = re[b] - rb[b]
        a, C_ref, res, rank, s, m_ref = orc.lstsq_fit(times[sl], data[which[b], sl], freq, t0[b], None)
        if rank < N or st[b] != 0:
            assert st[b] != 0 or s[-1] > 1e-13 * s[0]      # flagged fits go to the host repair path
            continue
        checked += 1
        assert np.max(np.abs(C[b] - C_ref)) / np.max(np.abs(C_ref)) < cases.amp_tol(s), (b, K)
        mm_ref = orc.mismatch(times[sl], m_ref, data[which[b], sl])
        assert abs(mm[b] - mm_ref) < MM_TOL, (b, K)
        np.testing.assert_allclose(model[b, :K], m_ref, rtol=0, atol=1e-8 * np.max(np.abs(C_ref)))
    assert checked >= B // 2
    ok = st == 0
    np.testing.assert_allclose(eng.to_host(mm_fast)[ok], mm[ok], rtol=0, atol=1e-11)
    np.testing.assert_allclose(eng.to_host(mm_eval)[ok], mm[ok], rtol=0, atol=1e-12)


def test_free_frequency_fit_vs_reference_golden_and_oracle(qf, eng, golden, oracle_tables):
    """free_frequency_fit (reference qnmfits.py:1905-2043): single call and batched
    lock-step search against what the unmodified reference returned.  The optimiser stops
    on xatol = 1e-8 with a mismatch floor ~1e-12, where comparisons are decided by the last
    bits of the objective (the device mismatch differs from numpy's by ~1e-13 = curvature x
    (5e-7)^2), so frequencies agree to that resolution — 1e-6, SURVEY.md 8f N1 — not bitwise;
    the injected truth is only recovered to ~1e-5 at this noise level."""
    g = golden("cfg5")
    tol = 1e-6
    for n_fixed, n_wf in ((2, 6), (1, 3), (0, 3)):
        wl = workloads.config5(n_waveforms=n_wf, n_fixed=n_fixed)
        got, res = qf.free_frequency_fit_batch(wl.times, wl.data, 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif,
                                               return_result=True)
        assert got.shape == (n_wf,) and got.dtype == np.complex128
        np.testing.assert_allclose(got, g[f"fixed{n_fixed}_omega"], rtol=0, atol=tol)
        assert np.all(res.status == 0) and res.launches == res.n_calls
        one = qf.free_frequency_fit(wl.times, wl.data[0], 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
        assert isinstance(one, complex) and abs(one - g[f"fixed{n_fixed}_omega"][0]) < tol
    wl = workloads.config5(n_waveforms=2, n_fixed=1)
    got = qf.free_frequency_fit_batch(wl.times, wl.data, 3.37, modes=wl.modes, Mf=wl.Mf, chif=wl.chif,
                                      t0_method='closest', T=60)
    np.testing.assert_allclose(got, g["closest_omega"], rtol=0, atol=tol)
    # a scipy method other than Nelder-Mead drives the device objective call by call
    w = qf.free_frequency_fit(wl.times, wl.data[0], 3.37, modes=wl.modes, Mf=wl.Mf, chif=wl.chif,
                              t0_method='closest', T=60, min_method='Powell')
    assert abs(w - g["closest_omega"][0]) < 1e-4      # Powell stops on its own, looser, tolerances
    with pytest.raises(ValueError):
        qf.free_frequency_fit(wl.times, wl.data[0], 0.0, t0_method='nearest')


@pytest.mark.parametrize("n_fixed", [2, 9, 11])
def test_free_frequency_objective_matches_oracle_mismatch(qf, eng, oracle_tables, n_fixed):
    """One batched objective call (per-fit data rows, per-fit trial frequency) against the
    numpy objective of the reference (qnmfits.py:2003-2029); 9 and 11 fixed modes run on K1p."""
    from qnmfits_b200 import qnmfits as api
    wl = workloads.config5(n_waveforms=33, n_fixed=n_fixed)
    fixed = np.array(qf.qnm.omega_list(wl.modes, wl.chif, wl.Mf))
    obj = api._FreeFrequencyObjective(wl.times, wl.data, 0.0, fixed, 'geq', 100)
    rng = np.random.default_rng(8)
    idx = np.sort(rng.choice(33, 20, replace=False))
    X = np.column_stack([rng.uniform(0, 2, 20), rng.uniform(-1, 0, 20)])
    X[0] = [2.0, 0.0]          # undamped corner of the box
    got = obj(X, idx)
    sel = orc.window(wl.times, 0.0, 100, 'geq')
    for k, b in enumerate(idx):
        a, C, res, rank, s, model = orc.lstsq_fit(wl.times[sel], wl.data[b][sel],
                                                  np.hstack([fixed, X[k, 0] + 1j * X[k, 1]]), 0.0)
        assert abs(got[k] - orc.mismatch(wl.times[sel], model, wl.data[b][sel])) < MM_TOL


def test_objective_calls_of_changing_size_reuse_the_device_block(qf, eng, oracle_tables):
    """The explicit-frequency launches keep their device block between calls
    (``_ResidentData._mismatches_one_call``): calls with fewer, more (the block grows) and again
    fewer trial points, one of them equal to a fixed mode (flagged, repaired: numpy truncates),
    and calls without per-fit data rows, all against the numpy objective."""
    from qnmfits_b200 import qnmfits as api
    wl = workloads.config5(n_waveforms=12, n_fixed=2)
    fixed = np.array(qf.qnm.omega_list(wl.modes, wl.chif, wl.Mf))
    sel = orc.window(wl.times, 0.0, 100, 'geq')

    def want(b, w):
        a, C, res, rank, s, model = orc.lstsq_fit(wl.times[sel], wl.data[b][sel], np.hstack([fixed, w]), 0.0)
        return orc.mismatch(wl.times[sel], model, wl.data[b][sel])

    rng = np.random.default_rng(21)
    obj = api._FreeFrequencyObjective(wl.times, wl.data[:6], 0.0, fixed, 'geq', 100)
    for idx in (np.array([1, 4]), np.arange(6), np.array([0, 2, 5]), np.array([3])):
        X = np.column_stack([rng.uniform(0.2, 1.5, len(idx)), rng.uniform(-0.9, -0.05, len(idx))])
        if len(idx) == 3:
            X[1] = [fixed[0].real, fixed[0].imag]
        got = obj(X, idx)
        for k, b in enumerate(idx):
            assert abs(got[k] - want(b, X[k, 0] + 1j * X[k, 1])) < MM_TOL, (idx, k)
    assert obj.launches >= 4
    # one data row, no series_index: 3 trial points, then 9 (larger than the block), then 1
    res = api._ResidentData(wl.times, wl.data[7], 0.0, 'geq', 100)
    for n in (3, 9, 1):
        w = rng.uniform(0.2, 1.5, n) + 1j * rng.uniform(-0.9, -0.05, n)
        omega = np.column_stack([np.tile(fixed, (n, 1)), w])
        got = res.mismatches(omega)
        for k in range(n):
            assert abs(got[k] - want(7, w[k])) < MM_TOL, (n, k)


def test_omega_grid_vs_reference_golden(qf, eng, golden):
    """mismatch_omega_grid (reference qnmfits.py:1679-1827): orientation [i_im, i_re], no
    fixed modes, and the 'closest' window that loses one sample per grid point."""
    g = golden("next")
    wl = workloads.config1()
    m2 = wl.modes[:2]
    got = qf.mismatch_omega_grid(wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1), 5.0, T=80, res=7)
    assert got.shape == (7, 7)
    np.testing.assert_allclose(got, g["omega_grid_geq"], rtol=0, atol=MM_TOL)
    got = qf.mismatch_omega_grid(wl.times, wl.data, [], 0.95, 0.69, (0.3, 0.8), (-0.3, -0.05), 20.0, res=4)
    np.testing.assert_allclose(got, g["omega_grid_nofixed"], rtol=0, atol=MM_TOL)
    got = qf.mismatch_omega_grid(wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1), 3.37,
                                 t0_method='closest', T=60, res=5)
    np.testing.assert_allclose(got, g["omega_grid_closest"], rtol=0, atol=MM_TOL)
    with pytest.raises(ValueError):
        qf.mismatch_omega_grid(wl.times, wl.data, m2, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1), 3.37,
                               t0_method='closest', T=60, res=40)
    # a grid with more than 8 columns goes through K3 (7 fixed + 1 free is still K1)
    big = qf.mismatch_omega_grid(wl.times, wl.data, wl.modes, 0.95, 0.69, (0.2, 0.9), (-0.9, -0.1), 5.0, T=80, res=3)
    assert big.shape == (3, 3) and np.all(np.isfinite(big)) and np.all(big < got.max() + 1)


def test_calculate_epsilon_vs_reference_golden(qf, eng, golden):
    """calculate_epsilon (reference qnmfits.py:1418-1594): Nelder-Mead over (Mf, chif),
    single series (with delta and x0) and multimode.  xatol = 1e-6 and a mismatch floor of
    ~1e-12 give best-fit values that agree to ~1e-6."""
    import cases as cs
    g = golden("next")
    wl = workloads.config1()
    eps, Mf, chi = qf.calculate_epsilon(wl.times, wl.data, wl.modes[:4], 0.95, 0.69, 10.0)
    np.testing.assert_allclose([eps, Mf, chi], g["eps_single"], rtol=0, atol=3e-6)
    got = qf.calculate_epsilon(wl.times, wl.data, wl.modes[:3], 0.95, 0.69, 15.0, T=70, delta=[0.0, 0.01, 0.0],
                               x0=[1.0, 0.6])
    np.testing.assert_allclose(got, g["eps_single_x0_delta"], rtol=0, atol=3e-6)
    wl4 = cs.cfg4_small()
    got = qf.calculate_epsilon(wl4.times, wl4.data, cs.MM_MODES, 0.95, 0.69, 5.0, T=80)
    np.testing.assert_allclose(got, g["eps_multimode"], rtol=0, atol=3e-6)
    got = qf.calculate_epsilon(wl4.times, wl4.data, cs.MM_MODES, 0.95, 0.69, 5.0, T=80, x0=[0.97, 0.65])
    np.testing.assert_allclose(got, g["eps_multimode_x0"], rtol=0, atol=3e-6)


def test_dynamic_fits_vs_reference_golden(qf, eng, golden):
    """dynamic_ringdown_fit / dynamic_multimode_ringdown_fit and the dynamic branch of
    mismatch_t0_array (reference qnmfits.py:318-475, 676-911, 1286-1299): per-row frequency
    table through K3, per-row mixing table through K2."""
    g = golden("dynamic")
    wl = workloads.config1()
    Mf_t, chi_t = cases.drift(wl.times)
    fit = qf.dynamic_ringdown_fit(wl.times, wl.data, wl.modes[:5], Mf_t, chi_t, 2.0, T=70)
    assert list(fit.keys()) == ['residual', 'mismatch', 'C', 'data', 'model', 'model_times', 't0', 'modes',
                                'mode_labels', 'frequencies']
    assert np.array_equal(fit["frequencies"], g["single_frequencies"])
    assert np.max(np.abs(fit["C"] - g["single_C"])) / np.max(np.abs(g["single_C"])) < 1e-8
    assert abs(fit["mismatch"] - float(g["single_mismatch"])) < MM_TOL
    np.testing.assert_allclose(fit["residual"], g["single_residual"], rtol=1e-6)
    np.testing.assert_allclose(fit["model"], g["single_model"], rtol=0, atol=1e-8 * np.max(np.abs(g["single_C"])))
    fit = qf.dynamic_ringdown_fit(wl.times, wl.data, wl.modes[:3], 0.95, chi_t, 3.37, t0_method='closest', T=50)
    assert np.max(np.abs(fit["C"] - g["single_closest_C"])) / np.max(np.abs(g["single_closest_C"])) < 1e-8
    assert abs(fit["mismatch"] - float(g["single_closest_mismatch"])) < MM_TOL
    mm = qf.mismatch_t0_array(wl.times, wl.data, wl.modes[:5], Mf_t, chi_t, g["t0s"], T_array=60)
    assert isinstance(mm, list)
    np.testing.assert_allclose(mm, g["sweep_single"], rtol=0, atol=MM_TOL)

    wl4 = cases.cfg4_small()
    Mf4, chi4 = cases.drift(wl4.times)
    fit = qf.dynamic_multimode_ringdown_fit(wl4.times, wl4.data, cases.DYN_MODES, Mf4, chi4, 5.0, T=80,
                                            spherical_modes=cases.DYN_SPH)
    assert list(fit.keys()) == ['residual', 'mismatch', 'C', 'weighted_C', 'data', 'model', 'model_times', 't0',
                                'modes', 'mode_labels', 'frequencies']
    assert tuple(fit["frequencies"].shape) == tuple(g["multi_frequencies_shape"])
    scale = np.max(np.abs(g["multi_C"]))
    assert np.max(np.abs(fit["C"] - g["multi_C"])) / scale < 1e-8
    assert abs(fit["mismatch"] - float(g["multi_mismatch"])) < MM_TOL
    np.testing.assert_allclose(fit["residual"], g["multi_residual"], rtol=1e-6)
    lm = cases.DYN_SPH[1]
    np.testing.assert_allclose(fit["model"][lm], g["multi_model_1"], rtol=0, atol=1e-8 * scale)
    np.testing.assert_allclose(fit["weighted_C"][lm], g["multi_weighted_1"], rtol=0, atol=1e-8 * scale)
    mm = qf.mismatch_t0_array(wl4.times, wl4.data, cases.DYN_MODES, Mf4, chi4, wl4.t0_array, T_array=70,
                              spherical_modes=cases.DYN_SPH)
    np.testing.assert_allclose(mm, g["sweep_multi"], rtol=0, atol=MM_TOL)
    # superset: coefficients that vanish identically (m' != m) — the reference cannot reshape them
    fit = qf.dynamic_multimode_ringdown_fit(wl4.times, wl4.data, cases.MM_MODES, Mf4, chi4, 5.0, T=80)
    assert np.isfinite(fit["mismatch"]) and fit["mismatch"] < 1e-3


def test_real_kerr_tables_device_vs_oracle(qf, eng):
    """The built-in Leaver provider (qnmfits_b200/kerr.py) through the whole device path:
    single fit, Mf-chi grid and a multimode fit against the oracle fed by the same tables
    (real Kerr numbers: other conditioning than the synthetic ladder)."""
    from qnmfits_b200 import kerr
    tables = orc.OracleTables(kerr.modes_cache)
    wl = workloads.config1()
    modes = [(2, 2, n, 1) for n in range(8)]
    try:
        workloads.use_kerr_tables()
        w = np.array(qf.qnm.omega_list(modes, 0.69, 0.95))
        assert np.array_equal(w, np.array(tables.omega_list(modes, 0.69, 0.95)))
        rng = np.random.default_rng(4)
        C = rng.normal(size=8) + 1j * rng.normal(size=8)
        data = np.where(wl.times >= 0, (C[None, :] * np.exp(-1j * w[None, :] * wl.times[:, None])).sum(axis=1), 0)
        data = data + 1e-6 * (rng.normal(size=data.size) + 1j * rng.normal(size=data.size))
        fit = qf.ringdown_fit(wl.times, data, modes, 0.95, 0.69, 0.0)
        ref = orc.ringdown_fit(tables, wl.times, data, modes, 0.95, 0.69, 0.0)
        assert np.array_equal(fit["frequencies"], ref["frequencies"])
        assert int(fit["rank"]) == int(ref["rank"]) == 8
        err = np.max(np.abs(fit["C"] - ref["C"])) / np.max(np.abs(ref["C"]))
        assert err < cases.amp_tol(ref["s"]), (err, ref["s"][0] / ref["s"][-1])
        assert abs(fit["mismatch"] - ref["mismatch"]) < MM_TOL
        grid = qf.mismatch_M_chi_grid(wl.times, data, modes, (0.9, 1.0), (0.6, 0.78), 0.0, res=6)
        gref = orc.mismatch_M_chi_grid(tables, wl.times, data, modes, (0.9, 1.0), (0.6, 0.78), 0.0, res=6)
        assert np.max(np.abs(grid - gref)) < MM_TOL
        sph = [(2, 2), (3, 2), (4, 2)]
        mm_modes = [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1), (2, 2, 0, -1), (4, 2, 0, 1)]
        dd = {lm: data * (0.3 ** i) * np.exp(0.4j * i) for i, lm in enumerate(sph)}
        mfit = qf.multimode_ringdown_fit(wl.times, dd, mm_modes, 0.95, 0.69, 0.0, spherical_modes=sph)
        mref = orc.multimode_ringdown_fit(tables, wl.times, dd, mm_modes, 0.95, 0.69, 0.0, spherical_modes=sph)
        assert abs(mfit["mismatch"] - mref["mismatch"]) < MM_TOL
        scale = np.max(np.abs(mref["C"]))
        assert np.max(np.abs(mfit["C"] - mref["C"])) < 1e-7 * scale
    finally:
        workloads.use_synthetic_tables()


def test_sweeps_on_a_nonuniform_grid_and_ragged_windows(qf, eng, oracle_tables):
    """Sweeps through the direct-evaluation generator (no nominal step) and with windows cut
    short by the end of the series: K1 (8 modes), K3 single series (12 modes) and a multimode
    Mf-chi grid, against the oracle; an empty window raises before any launch."""
    wl = workloads.config2(n_t0=24)
    rng = np.random.default_rng(8)
    times = np.sort(wl.times + rng.uniform(-0.03, 0.03, size=wl.times.size))
    t0s = np.linspace(20.0, 140.0, 24)                     # the last windows hit the end of the data
    for modes in (wl.modes, [(2, 2, n, 1) for n in range(12)]):
        # 12 overtones at these late start times are numerically rank deficient: numpy
        # truncates singular values, the sweep completes the same minimum-norm solution
        # from the device factor (no warning is left over)
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            got = qf.mismatch_t0_array(times, wl.data, modes, wl.Mf, wl.chif, t0s, T_array=60)
        want = orc.mismatch_t0_array(oracle_tables, times, wl.data, modes, wl.Mf, wl.chif, t0s, T_array=60)
        np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)
    w4 = cases.cfg4_small()
    grid = qf.mismatch_M_chi_grid(w4.times, w4.data, w4.modes, (0.9, 1.0), (0.6, 0.78), 5.0, T=70, res=5,
                                  spherical_modes=w4.spherical_modes)
    want = orc.mismatch_M_chi_grid(oracle_tables, w4.times, w4.data, w4.modes, (0.9, 1.0), (0.6, 0.78), 5.0,
                                   T=70, res=5, spherical_modes=w4.spherical_modes)
    np.testing.assert_allclose(grid, want, rtol=0, atol=MM_TOL)
    with pytest.raises(ValueError, match="window is empty"):
        qf.mismatch_t0_array(times, wl.data, wl.modes, wl.Mf, wl.chif, np.array([10.0, 500.0]))
    with pytest.raises(ValueError, match="window is empty"):
        qf.mismatch_M_chi_grid(times, wl.data, wl.modes, (0.9, 1.0), (0.6, 0.7), 500.0, res=3)


def test_rank_deficient_fits_in_sweeps_follow_numpy(qf, eng, oracle_tables):
    """Sweeps whose fits numpy solves by truncating singular values (duplicated labels; many
    overtones at late start times) return numpy's mismatches: uniform grid (fast-mismatch
    path), Mf-chi grid, multimode."""
    wl = workloads.config2(n_t0=16)
    dup = [(2, 2, n, 1) for n in (0, 1, 9, 10)]          # (2,2,9) and (2,2,10) resolve to one sequence
    t0s = np.linspace(0.0, 30.0, 16)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        got = qf.mismatch_t0_array(wl.times, wl.data, dup, wl.Mf, wl.chif, t0s)
        grid = qf.mismatch_M_chi_grid(wl.times, wl.data, dup, (0.9, 1.0), (0.6, 0.78), 5.0, res=4)
        late = qf.mismatch_t0_array(wl.times, wl.data, [(2, 2, n, 1) for n in range(12)], wl.Mf, wl.chif,
                                    np.linspace(25.0, 60.0, 16), T_array=50)
    want = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, dup, wl.Mf, wl.chif, t0s)
    np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)
    gwant = orc.mismatch_M_chi_grid(oracle_tables, wl.times, wl.data, dup, (0.9, 1.0), (0.6, 0.78), 5.0, res=4)
    np.testing.assert_allclose(grid, gwant, rtol=0, atol=MM_TOL)
    lwant = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, [(2, 2, n, 1) for n in range(12)], wl.Mf,
                                  wl.chif, np.linspace(25.0, 60.0, 16), T_array=50)
    np.testing.assert_allclose(late, lwant, rtol=0, atol=MM_TOL)
    w4 = cases.cfg4_small()
    mm_dup = w4.modes + [w4.modes[0]]                     # a duplicated QNM in a multimode fit
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        mgot = qf.mismatch_t0_array(w4.times, w4.data, mm_dup, w4.Mf, w4.chif, w4.t0_array,
                                    spherical_modes=w4.spherical_modes)
    mwant = orc.mismatch_t0_array(oracle_tables, w4.times, w4.data, mm_dup, w4.Mf, w4.chif, w4.t0_array,
                                  spherical_modes=w4.spherical_modes)
    np.testing.assert_allclose(mgot, mwant, rtol=0, atol=MM_TOL)


def test_omega_grid_node_on_a_fixed_mode_follows_numpy(qf, eng, oracle_tables):
    """A frequency-grid node that coincides with one of the fixed modes makes two columns equal:
    numpy truncates the zero singular value; so does the device path (repair of flagged fits)."""
    wl = workloads.config1()
    m2 = wl.modes[:2]
    w0 = complex(qf.qnm.omega_list(m2, 0.69, 0.95)[0])
    d = 0.125                                             # the centre node is w0 to an ulp
    re_mm, im_mm = (w0.real - d, w0.real + d), (w0.imag - d, w0.imag + d)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        got = qf.mismatch_omega_grid(wl.times, wl.data, m2, 0.95, 0.69, re_mm, im_mm, 5.0, T=80, res=5)
    want = orc.mismatch_omega_grid(oracle_tables, wl.times, wl.data, m2, 0.95, 0.69, re_mm, im_mm, 5.0, T=80, res=5)
    assert abs(np.linspace(*re_mm, 5)[2] - w0.real) < 1e-15 and abs(np.linspace(*im_mm, 5)[2] - w0.imag) < 1e-15
    np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)


# --------------------------------------------------------------------------
# BASELINE.json config 4 as stated (quadratic QNMs in the multimode fit) and the full-size
# configurations

def test_quadratic_qnms_in_multimode_fit_vs_oracle(qf, eng, oracle_tables):
    """multimode_ringdown_fit with quadratic labels whose per-series coefficients come from
    the caller (``coef_columns``; semantics of the reference's mapping fit, coef = mu for
    linear labels and alpha for quadratic ones, spatial_mapping_functions.py:202-240)
    against the oracle's ``coef_override`` path: every key of the result dict."""
    wl = workloads.config4(n_t0=3, quadratic=True)
    cols = wl.extra["coef_columns"]
    assert sum(len(m) == 8 for m in wl.modes) == 4
    over = workloads.coef_override(wl.spherical_modes, wl.modes, wl.chif, oracle_tables)
    for t0, T, method in ((0.0, 100, 'geq'), (7.33, 60, 'closest')):
        fit = qf.multimode_ringdown_fit(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, t0, method, T,
                                        wl.spherical_modes, coef_columns=cols)
        ref = orc.multimode_ringdown_fit(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif, t0,
                                         method, T, wl.spherical_modes, coef_override=over)
        assert list(fit.keys()) == list(ref.keys())
        assert np.array_equal(fit["frequencies"], ref["frequencies"])
        assert np.array_equal(fit["model_times"], ref["model_times"])
        s = np.linalg.svd(_stacked_design(ref, over, t0), compute_uv=False)
        scale = np.max(np.abs(ref["C"]))
        assert np.max(np.abs(fit["C"] - ref["C"])) / scale < cases.amp_tol(s)
        assert abs(fit["mismatch"] - ref["mismatch"]) < MM_TOL
        assert fit["residual"].shape == ref["residual"].shape
        for lm in wl.spherical_modes:
            np.testing.assert_allclose(fit["model"][lm], ref["model"][lm], rtol=0, atol=1e-8 * scale)
            np.testing.assert_allclose(fit["weighted_C"][lm], ref["weighted_C"][lm], rtol=0,
                                       atol=cases.amp_tol(s) * scale)
    # the injected amplitudes come back (noise 1e-6)
    fit = qf.multimode_ringdown_fit(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, 0.0, 'geq', 100,
                                    wl.spherical_modes, coef_columns=cols)
    assert np.max(np.abs(fit["C"] - wl.extra["C_true"])) < 1e-3


def _stacked_design(ref, coef, t0):
    tau = ref["model_times"] - t0
    E = np.exp(-1j * np.outer(tau, ref["frequencies"]))
    return np.concatenate([E * coef[i][None, :] for i in range(coef.shape[0])])


def test_config4_full_size_with_quadratic_qnms_sampled_vs_oracle(qf, eng, oracle_tables):
    """Config 4 at BASELINE size — 500 start times x (21 series x 40 QNMs incl. mirror and
    quadratic modes) — through mismatch_t0_array; 32 sampled start times against the oracle."""
    wl = workloads.config4(n_t0=500, quadratic=True)
    got = np.array(qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array,
                                        T_array=wl.T, spherical_modes=wl.spherical_modes,
                                        coef_columns=wl.extra["coef_columns"]))
    assert got.shape == (500,) and np.all(np.isfinite(got))
    over = workloads.coef_override(wl.spherical_modes, wl.modes, wl.chif, oracle_tables)
    idx = np.unique(np.concatenate([[0, 499], np.random.default_rng(4).choice(500, 30, replace=False)]))
    want = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif,
                                 wl.t0_array[idx], T_array=wl.T, spherical_modes=wl.spherical_modes,
                                 coef_override=over)
    np.testing.assert_allclose(got[idx], want, rtol=0, atol=MM_TOL)
    # the linear-label shape of the golden fixtures at full size as well
    wl = workloads.config4(n_t0=500)
    got = np.array(qf.mismatch_t0_array(wl.times, wl.data, wl.modes, wl.Mf, wl.chif, wl.t0_array,
                                        T_array=wl.T, spherical_modes=wl.spherical_modes))
    want = orc.mismatch_t0_array(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf, wl.chif,
                                 wl.t0_array[idx], T_array=wl.T, spherical_modes=wl.spherical_modes)
    np.testing.assert_allclose(got[idx], want, rtol=0, atol=MM_TOL)


def test_multimode_grid_with_spin_dependent_columns_vs_oracle(qf, eng, oracle_tables):
    """mismatch_M_chi_grid over dict data with quadratic columns that depend on the spin:
    every grid point against the oracle (callable coef_override)."""
    wl = workloads.config4(n_t0=1, quadratic=True)
    sph = [(2, 2), (3, 2), (4, 4), (2, 0), (4, -4)]
    modes = [(2, 2, 0, 1), (2, 2, 1, 1), (3, 2, 0, 1), (4, 4, 0, 1), (2, 0, 0, 1), (2, -2, 0, -1),
             (4, -4, 0, 1)] + list(workloads.QUADRATIC_LABELS)
    cols = workloads.quadratic_columns(sph)
    data = {lm: wl.data[lm] for lm in sph}
    grid = qf.mismatch_M_chi_grid(wl.times, data, modes, (0.9, 1.0), (0.6, 0.75), 2.0, T=80, res=5,
                                  spherical_modes=sph, coef_columns=cols)
    want = orc.mismatch_M_chi_grid(
        oracle_tables, wl.times, data, modes, (0.9, 1.0), (0.6, 0.75), 2.0, T=80, res=5, spherical_modes=sph,
        coef_override=lambda chif: workloads.coef_override(sph, modes, chif, oracle_tables))
    np.testing.assert_allclose(grid, want, rtol=0, atol=MM_TOL)


def test_config3_full_grid_512_points_vs_oracle(qf, eng, oracle_tables):
    """The headline grid (256 x 256, 8 overtones) against the oracle on 512 sampled points."""
    wl = workloads.config3(res=256)
    grid = qf.mismatch_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0,
                                  T=wl.T, res=256)
    idx = np.sort(np.random.default_rng(12).choice(256 * 256, 512, replace=False))
    want = orc.mismatch_M_chi_grid(oracle_tables, wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax,
                                   wl.t0, T=wl.T, res=256, flat_indices=idx)
    np.testing.assert_allclose(grid.reshape(-1)[idx], want, rtol=0, atol=MM_TOL)


def test_config5_full_size_sampled_vs_scipy(qf, eng, oracle_tables):
    """Config 5 at BASELINE size: 4096 waveforms searched in lock step; 128 sampled waveforms
    against scipy's Nelder-Mead on the numpy objective (the reference's own code path,
    qnmfits.py:1995-2038).  Agreement is at the optimiser's resolution (see
    test_free_frequency_fit_vs_reference_golden_and_oracle): 1e-6 in the frequency."""
    wl = workloads.config5(n_waveforms=4096, n_fixed=2)
    got = qf.free_frequency_fit_batch(wl.times, wl.data, 0.0, modes=wl.modes, Mf=wl.Mf, chif=wl.chif)
    assert got.shape == (4096,)
    idx = np.sort(np.random.default_rng(5).choice(4096, 128, replace=False))
    want = np.array([orc.free_frequency_fit(oracle_tables, wl.times, wl.data[b], 0.0, modes=wl.modes, Mf=wl.Mf,
                                            chif=wl.chif) for b in idx])
    np.testing.assert_allclose(got[idx], want, rtol=0, atol=1e-6)


# --------------------------------------------------------------------------
# repeated calls (prepared sweeps), repair of exactly the flagged fits, unsorted samples

def test_repeated_calls_reuse_the_prepared_sweep_and_stay_exact(qf, eng, oracle_tables):
    """The second call with the same problem goes through ONE C call (qnmfit_run_host) on the
    prepared sweep; results are bit-identical to a cold call, other data values are honoured,
    other time samples invalidate the entry."""
    from qnmfits_b200 import qnmfits as api
    wl = workloads.config3(res=20)
    args = (wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0)
    qf.clear_sweep_cache()
    cold = qf.mismatch_M_chi_grid(wl.times, wl.data, *args, T=wl.T, res=20)
    assert len(api._sweep_cache) == 1
    launches = eng.ctx.launch_count()
    warm = qf.mismatch_M_chi_grid(wl.times, wl.data, *args, T=wl.T, res=20)
    assert eng.ctx.launch_count() == launches + 1 and len(api._sweep_cache) == 1
    assert np.array_equal(cold, warm)
    other = wl.data * (1.0 + 0.01j) + 1e-4 * np.exp(-0.3j * wl.times)
    got = qf.mismatch_M_chi_grid(wl.times, other, *args, T=wl.T, res=20)
    qf.clear_sweep_cache()
    fresh = qf.mismatch_M_chi_grid(wl.times, other, *args, T=wl.T, res=20)
    assert np.array_equal(got, fresh) and not np.array_equal(got, cold)
    want = orc.mismatch_M_chi_grid(oracle_tables, wl.times, other, *args, T=wl.T, res=20)
    np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)
    shifted = wl.times + 0.003                      # other samples: other window arithmetic
    got = qf.mismatch_M_chi_grid(shifted, wl.data, *args, T=wl.T, res=20)
    want = orc.mismatch_M_chi_grid(oracle_tables, shifted, wl.data, *args, T=wl.T, res=20)
    np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)
    # start-time sweeps and dict data the same way
    wl4 = cases.cfg4_small()
    a4 = (wl4.modes, wl4.Mf, wl4.chif, wl4.t0_array)
    qf.clear_sweep_cache()
    cold = qf.mismatch_t0_array(wl4.times, wl4.data, *a4)
    warm = qf.mismatch_t0_array(wl4.times, wl4.data, *a4)
    assert np.array_equal(cold, warm) and len(api._sweep_cache) == 1
    scaled = {lm: 0.5 * v for lm, v in wl4.data.items()}
    got = qf.mismatch_t0_array(wl4.times, scaled, *a4)
    np.testing.assert_allclose(got, cold, rtol=0, atol=1e-12)      # the mismatch is scale invariant
    qf.clear_sweep_cache()


def test_only_the_flagged_fits_are_repaired(qf, eng, oracle_tables):
    """One rank-deficient fit in a large sweep — a frequency grid whose centre point coincides
    with the fixed mode, so that exactly there two columns are equal and numpy truncates: the
    kernels list the fit, and the host re-fits exactly that fit (one subset launch + one eval
    launch), not the 10 201 of the sweep."""
    wl = workloads.config1()
    w0 = complex(qf.qnm.omega_list([(2, 2, 0, 1)], wl.chif, wl.Mf)[0])
    re_mm = (w0.real - 0.25, w0.real + 0.25)
    im_mm = (w0.imag - 0.05, w0.imag + 0.05)
    res = 101
    assert np.linspace(*re_mm, res)[50] == w0.real and np.linspace(*im_mm, res)[50] == w0.imag
    launches = eng.ctx.launch_count()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = qf.mismatch_omega_grid(wl.times, wl.data, [(2, 2, 0, 1)], wl.Mf, wl.chif, re_mm, im_mm, 0.0, T=100, res=res)
    used = eng.ctx.launch_count() - launches
    assert got.shape == (res, res) and np.all(np.isfinite(got))
    # the reference's loop on the centre point and a few others (qnmfits.py:1785-1803)
    for i_re, i_im in ((50, 50), (49, 50), (50, 51), (0, 0), (100, 37)):
        freq = np.array([w0, np.linspace(*re_mm, res)[i_re] + 1j * np.linspace(*im_mm, res)[i_im]])
        sl = slice(*np.searchsorted(wl.times, [0.0, 100.0]))
        a, C, r, rank, s, model = orc.lstsq_fit(wl.times[sl], wl.data[sl], freq, 0.0)
        assert rank == (1 if (i_re, i_im) == (50, 50) else 2)
        assert abs(got[i_im, i_re] - orc.mismatch(wl.times[sl], model, wl.data[sl])) < MM_TOL, (i_re, i_im)
    assert used == 3, used                           # sweep + refit of the one flagged fit + its evaluation


def test_grid_on_unsorted_samples_follows_the_reference_mask(qf, eng, oracle_tables):
    """mismatch_M_chi_grid with shuffled time samples: the reference's boolean mask works on any
    order (qnmfits.py:233); so does the device path (direct evaluation of every row)."""
    wl = workloads.config3(res=5)
    perm = np.random.default_rng(3).permutation(len(wl.times))
    t, d = wl.times[perm], wl.data[perm]
    got = qf.mismatch_M_chi_grid(t, d, wl.modes[:4], wl.Mf_minmax, wl.chif_minmax, 2.0, T=40, res=5)
    want = orc.mismatch_M_chi_grid(oracle_tables, t, d, wl.modes[:4], wl.Mf_minmax, wl.chif_minmax, 2.0, T=40, res=5)
    np.testing.assert_allclose(got, want, rtol=0, atol=MM_TOL)
