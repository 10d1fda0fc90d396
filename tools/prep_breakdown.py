import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import qnmfits_b200 as qf
from qnmfits_b200 import workloads, qnmfits as api, _dist, _engine, _cabi
workloads.use_synthetic_tables()
wl = workloads.config3(res=256)
args = (wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax, wl.t0); kw = dict(T=wl.T, res=256)
for _ in range(5): qf.mismatch_M_chi_grid(*args, **kw)
def T(f, n=500):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e6
eng = _engine.get_engine()
times = np.asarray(wl.times); chi = api._linspace(0.59, 0.79, 256)[0]
table, mode_ptr, tmax = api.qnm.constituent_table(wl.modes, chi, with_max=True)
rows = np.asarray(wl.data, dtype=complex).reshape(1, -1)
inv = api._linspace(0.85, 1.05, 256)[1]
df = np.full(8, 1.0)
host = [np.ascontiguousarray(times, dtype=np.float64), np.ascontiguousarray(rows, dtype=np.complex128), None, None, None, None, None,
        table, mode_ptr, inv, df, api._ZERO]
print('check_modes+asarray', T(lambda: (np.asarray(wl.times), api._check_modes(wl.modes))))
print('linspace x2', T(lambda: (api._linspace(0.85, 1.05, 256), api._linspace(0.59, 0.79, 256))))
print('time_axis', T(lambda: api._time_axis(times, 0.0, 100, 'geq')))
print('series_rows', T(lambda: api._series_rows(wl.data, None)))
print('constituent_table', T(lambda: api.qnm.constituent_table(wl.modes, chi, with_max=True)))
print('delta_factor', T(lambda: np.full(8, api._delta_factor(0.0, 8), dtype=float)))
print('get_engine', T(lambda: _engine.get_engine()))
print('dist.world', T(lambda: _dist.world()))
print('eng.stream', T(lambda: eng.stream()))
print('ascontiguous x6', T(lambda: [np.ascontiguousarray(a) for a in (times, rows, table, mode_ptr, inv, df)]))
st = eng.stream()
print('upload_packed', T(lambda: eng.upload_packed(host, out_bytes=8 * 65536, stream=st)))
print('  torch.empty dev', T(lambda: torch.empty(700000, dtype=torch.uint8, device=eng.device)))
print('  h2d_wait', T(lambda: eng.ctx.h2d_wait()))
keep, ptrs, out = eng.upload_packed(host, out_bytes=8 * 65536, stream=st)
mk = lambda: eng.make_batch(times_d=ptrs[0], data_d=ptrs[1], n_times=2001, series_stride=2001, n_fits=65536, n_modes=8, n_series=1,
    first_fit=0, row_begin_all=500, row_end_all=1500, t0_all=0.0, row_begin_d=None, row_end_d=None, t0_d=None, coef_d=None, coef_index_d=None, n_coef=0,
    dt_nominal=0.1, uniform_weights=True, mismatch_d=out, flagged_d=out - 8, omega_tilde_d=ptrs[7], mode_ptr_d=ptrs[8], inv_Mf_d=ptrs[9], delta_factor_d=ptrs[10], n_chi=256, n_mf=256, n_constituents=8)
print('make_batch', T(mk))
print('prepare total', T(lambda: api._prepare_M_chi_grid(*args, **kw)))
