"""Host logic of the public API that needs no GPU: windows, delta, errors, no fallback."""
import numpy as np
import pytest

from qnmfits_b200 import qnmfits as api
from qnmfits_b200 import _dist, _engine, _cabi


def test_window_rows_equal_reference_mask():
    times = np.arange(-500, 1501) * 0.1
    rng = np.random.default_rng(0)
    for t0, T in zip(rng.uniform(-20, 80, 200), rng.uniform(1, 120, 200)):
        mask = (times >= t0) & (times < t0 + T)
        idx = np.nonzero(mask)[0]
        b, e = api._window_rows(times, t0, T, 'geq')
        assert (b, e) == (idx[0], idx[-1] + 1)
        s = api._window(times, t0, T, 'closest')
        assert api._window_rows(times, t0, T, 'closest') == (s.start, max(s.stop, s.start))
    # exact hits on samples
    assert api._window_rows(times, times[600], 10.0, 'geq')[0] == 600


def test_bad_arguments_raise_value_error():
    times = np.linspace(0, 10, 11)
    with pytest.raises(ValueError):
        api._window(times, 0, 1, 'nearest')
    with pytest.raises(ValueError):
        api._delta_factor([0.1, 0.2], 3)
    with pytest.raises(ValueError):
        api._delta_factor("x", 3)
    assert api._delta_factor(0, 3) == 1.0
    assert np.array_equal(api._delta_factor([0.0, 0.5], 2), np.array([1.0, 1.5]))
    with pytest.raises(ValueError):
        api._check_modes([(2, 2, 0)])


def test_nominal_step_detection():
    t = np.arange(-500, 1501) * 0.1
    assert abs(_engine.nominal_step(t, 2.0) - 0.1) < 1e-12
    assert _engine.nominal_step(np.sort(np.random.default_rng(0).random(100)), 2.0) == 0.0
    assert _engine.nominal_step(np.array([0.0, 1.0]), 1.0) == 0.0


def test_rank_and_min_norm_completion_match_numpy():
    """Host completion for rank-deficient fits reproduces numpy.linalg.lstsq from the
    triangular factor alone."""
    rng = np.random.default_rng(1)
    M, N = 200, 5
    A = rng.normal(size=(M, N)) + 1j * rng.normal(size=(M, N))
    A[:, 4] = A[:, 1]                     # exactly duplicated column
    b = rng.normal(size=M) + 1j * rng.normal(size=M)
    Q, R = np.linalg.qr(A)
    Rfull = np.concatenate([R, (Q.conj().T @ b)[:, None]], axis=1)
    x, res, rank, s = np.linalg.lstsq(A, b, rcond=None)
    rk, sv = api._rank_and_singular_values(Rfull, M)
    assert rk == rank == 4
    np.testing.assert_allclose(sv, s, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(api._minimum_norm_from_factor(Rfull, M), x, rtol=1e-10, atol=1e-12)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 65536):
        for ws in (1, 2, 3, 8):
            spans = [_dist.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(hi - lo <= per for lo, hi, per in spans)


def test_no_cpu_fallback_without_gpu(qf):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qnmfits_b200 import workloads
    wl = workloads.config1()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qf.ringdown_fit(wl.times, wl.data, wl.modes, 0.95, 0.69, 0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qf.mismatch_M_chi_grid(wl.times, wl.data, wl.modes, (0.9, 1.0), (0.6, 0.7), 0.0, res=4)


def test_dynamic_spectrum_tables_and_no_cpu_fallback(qf):
    """Per-row tables of the dynamic fits (reference qnmfits.py:432-445, 806-833): scalar
    entries are broadcast along time; without a GPU the fit itself raises."""
    import torch
    from qnmfits_b200 import qnmfits as api
    t = np.linspace(0, 10, 101)
    Mf = np.linspace(0.9, 0.95, 101)
    om, mu = api._row_tables(t, [(2, 2, 0, 1), (3, 2, 0, 1)], Mf, 0.7, [(2, 2), (2, 1)])
    assert om.shape == (2, 101) and mu.shape == (2, 2, 101)
    np.testing.assert_array_equal(om[0], np.array(qf.qnm.omega_list([(2, 2, 0, 1)], 0.7, 1.0))[0] / Mf)
    assert np.all(mu[1] == 0)                      # m' != m: the scalar 0 of qnm.mu, broadcast
    with pytest.raises(ValueError):
        api._row_tables(t, [(2, 2, 0, 1)], Mf[:50], 0.7, None)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            qf.mismatch_t0_array(t, np.ones(101, complex), [(2, 2, 0, 1)], np.ones(101), 0.7, [0.0])


def test_flops_formula():
    assert abs(_cabi.flops_per_fit(1000, 8) - 770378.67) < 1.0


def test_grid_plan_equals_the_two_separate_checks():
    from qnmfits_b200._engine import grid_plan, nominal_step, uniform_weights
    rng = np.random.default_rng(5)
    uniform = np.arange(-500, 1501) * 0.1
    jitter = uniform + rng.normal(scale=3e-13, size=uniform.size)     # recurrence ok, weights not uniform
    rough = np.sort(uniform + rng.uniform(-0.03, 0.03, size=uniform.size))
    for times in (uniform, jitter, rough, uniform[:2], np.array([0.0, 0.0, 0.0])):
        for wmax in (0.5, 40.0):
            dt, uw = grid_plan(times, wmax)
            assert dt == nominal_step(times, wmax)
            assert uw == (uniform_weights(times, dt) if dt > 0 else False)
            assert grid_plan(times, wmax, np.diff(times)) == (dt, uw)
    assert grid_plan(uniform, 2.0) == (grid_plan(uniform, 2.0)[0], True)
    assert grid_plan(jitter, 2.0)[0] > 0 and grid_plan(jitter, 2.0)[1] is False
    assert grid_plan(rough, 2.0) == (0.0, False)


def test_memoised_tables_equal_fresh_tabulation(qf):
    """The memoised grid axes / frequency tables / label lists return what a fresh provider computes."""
    from qnmfits_b200 import qnmfits as api
    from qnmfits_b200.qnm import qnm as Provider
    modes = [(2, 2, n, 1) for n in range(8)] + [(2, 2, 0, 1, 2, 2, 0, 1), (2, 2, 0, -1)]
    chi = np.linspace(0.59, 0.79, 17)
    first = qf.qnm.constituent_table(modes, chi, with_max=True)
    again = qf.qnm.constituent_table(modes, chi, with_max=True)
    fresh = Provider().constituent_table(modes, chi, with_max=True)
    for a, b in ((first, again), (first, fresh)):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert first[2] == np.max(np.abs(first[0]))
    assert not first[0].flags.writeable
    w1 = qf.qnm.omega_list(modes, 0.69, 0.95)
    w2 = qf.qnm.omega_list(modes, 0.69, 0.95)
    w3 = Provider().omega_list(modes, 0.69, 0.95)
    assert w1 == w2 == w3 and isinstance(w2, list)
    idx = [(2, 2) + m for m in modes[:8]] + [(3, 2, 2, 2, 0, -1), (2, 1, 2, 2, 0, 1)]
    m1 = qf.qnm.mu_list(idx, 0.69)
    assert m1 == qf.qnm.mu_list(idx, 0.69) == Provider().mu_list(idx, 0.69)
    assert m1[-1] == 0 and isinstance(m1[-1], int)          # reference qnm.py:336-337
    # array-valued spins are not memoised as scalars and still work
    wa = qf.qnm.omega_list(modes[:2], np.array([0.5, 0.69]), 0.95)
    assert wa[0].shape == (2,) and wa[0][1] == w1[0]
    arr, inv, inv_max = api._linspace(0.85, 1.05, 256)
    assert np.array_equal(arr, np.linspace(0.85, 1.05, 256)) and np.array_equal(inv, 1.0 / arr)
    assert inv_max == np.max(np.abs(inv)) and api._linspace(0.85, 1.05, 256)[0] is arr


def test_pack_layout_keeps_the_counter_in_front_of_the_results():
    """The zero head (a sweep's counter of flagged fits) sits immediately before the result
    region whatever the array sizes — also when an array is large enough to be uploaded
    from its own memory (>= DIRECT_BYTES), which moves it out of the staged group."""
    direct = _engine.Engine.DIRECT_BYTES
    for sizes in ([16008, 32016, None, None, 128, 64],
                  [2400000, 16 * 300000, None, 4004, 4004, 8008, 128],     # 300 000-sample series: data >= 4 MB
                  [direct, direct + 8, 24, None],
                  [None, None],
                  [40]):
        for zero_head in (0, 16):
            offsets, stage_begin, stage_end, out = _engine.pack_layout(sizes, direct, zero_head)
            assert out == stage_end and out % 16 == 0
            spans = sorted((off, off + n) for off, n in zip(offsets, sizes) if n is not None)
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 <= b0                      # no overlap
            for off, n in zip(offsets, sizes):
                if n is None:
                    assert off is None
                    continue
                assert off % 256 == 0
                assert off + n <= out - zero_head    # nothing reaches into the zero head or beyond
                staged = n < direct
                assert (off >= stage_begin) == staged
            # every large array precedes the staged group
            assert all(off + n <= stage_begin for off, n in zip(offsets, sizes) if n is not None and n >= direct)
    with pytest.raises(ValueError):
        _engine.pack_layout([8], direct, zero_head=8)


def test_coef_columns_assemble_like_the_reference_mapping_fit(qf):
    """Caller-supplied coefficient columns (quadratic QNMs in a multimode fit): linear labels
    keep qnm.mu, listed labels take the caller's column — ``coef_lists = mu + alpha`` of the
    reference's spatial_mapping_functions.py:202-210 — for scalar spins and on a spin grid."""
    from qnmfits_b200 import workloads
    spherical, modes = workloads.multimode_labels_quadratic()
    cols = workloads.quadratic_columns(spherical)
    assert len(modes) == 40 and sum(len(m) == 8 for m in modes) == 4
    want = workloads.coef_override(spherical, modes, 0.69)
    lists = api._mu_lists(spherical, modes, 0.69, cols)
    got = np.array([[complex(v) for v in row] for row in lists])
    assert np.array_equal(got, want)
    # array, dict and callable forms of a column are equivalent
    label = workloads.QUADRATIC_LABELS[0]
    alpha = workloads.quadratic_alpha(label, spherical, 0.69)
    as_dict = {lm: alpha[i] for i, lm in enumerate(spherical) if alpha[i] != 0}
    for form in (alpha, list(alpha), as_dict, lambda chif: workloads.quadratic_alpha(label, spherical, chif)):
        lists = api._mu_lists(spherical, [modes[0], label], 0.69, {label: form})
        assert np.array_equal(np.array([row[1] for row in lists]), alpha)
    chis = np.linspace(0.6, 0.75, 5)
    table = api._coef_table(spherical, modes, chis, cols)
    assert table.shape == (5, 21, 40)
    for c, chif in enumerate(chis):
        np.testing.assert_allclose(table[c], workloads.coef_override(spherical, modes, float(chif)), rtol=0, atol=1e-15)
    # without columns the nonlinear label is rejected like in the reference (qnm.py:390)
    with pytest.raises(ValueError):
        api._mu_lists(spherical, modes, 0.69)
    with pytest.raises(ValueError):
        api._coef_table(spherical, modes, chis)
    with pytest.raises(ValueError):
        api._mu_lists(spherical, modes, 0.69, {label: np.ones(3)})


def test_explicit_block_layout_keeps_inputs_and_results_apart():
    """Device block of the explicit-frequency launches (free-frequency objective): whatever the
    number of fits of a call, its inputs end before the result region, which does not move."""
    from qnmfits_b200.qnmfits import explicit_block_layout, FLAG_CAPACITY
    for N in (1, 3, 8, 24):
        for cap in (1, 7, 64, 4096, 65536):
            _, c_off, total = explicit_block_layout(cap, cap, N, True)
            assert c_off % 256 == 0 and total == c_off + 16 + 8 * cap + 8 * FLAG_CAPACITY
            for with_index in (True, False):
                for n in sorted({0, 1, cap // 3, cap - 1, cap} & set(range(cap + 1))):
                    omega_off, c2, t2 = explicit_block_layout(n, cap, N, with_index)
                    assert omega_off % 256 == 0 and (omega_off >= 4 * n if with_index else omega_off == 0)
                    assert omega_off + 16 * n * N <= c2 <= c_off      # inputs end before the counter
                    assert t2 <= total                                 # the block allocated for cap holds it
    with pytest.raises(ValueError):
        explicit_block_layout(5, 4, 3, True)
