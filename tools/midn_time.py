import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
from qnmfits_b200 import workloads, _cabi
from qnmfits_b200 import qnmfits as api
workloads.use_synthetic_tables()
wl = workloads.config3(res=128)
for N in (8, 9, 10, 12, 16, 24):
    modes = [(2, 2, n, 1) for n in range(min(N, 12))] + [(3, 2, n, 1) for n in range(max(0, N - 12))]
    sweep, shape = api._prepare_M_chi_grid(wl.times, wl.data, modes, wl.Mf_minmax, wl.chif_minmax, wl.t0, T=wl.T, res=128)
    for _ in range(2): sweep.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sweep.launch_kernel()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    plan = sweep.eng.ctx.plan(sweep.batch)
    fl = _cabi.flops_per_fit(sweep.rows_max, N, 1, True) * 16384
    print(N, 'kernel', plan.kernel, 'ms %.3f' % ms, 'fits/s %.3g' % (16384 / ms * 1e3), 'TF %.2f' % (fl / ms * 1e-9), 'block', plan.block, 'regs', plan.regs_per_thread, flush=True)
