#!/usr/bin/env python
"""Developer tool (GPU box): kernel time of a single-series 128 x 128 Mf-chi grid (16384 fits,
M = 1000) for N = 8 .. 24 columns: K1, K1p (9 .. 16) and K3.  For every N the kernels that take the
shape are timed (AUTO first) and their mismatch grids compared with AUTO's."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from qnmfits_b200 import workloads, _cabi  # noqa: E402
from qnmfits_b200 import qnmfits as api  # noqa: E402

workloads.use_synthetic_tables()
wl = workloads.config3(res=128)
for N in [int(a) for a in sys.argv[1:]] or (8, 9, 10, 11, 12, 16, 24):
    modes = [(2, 2, n, 1) for n in range(min(N, 9))] + [(3, 2, n, 1) for n in range(max(0, N - 9))]
    sweep, shape = api._prepare_M_chi_grid(wl.times, wl.data, modes, wl.Mf_minmax, wl.chif_minmax, wl.t0, T=wl.T, res=128)
    kernels = [_cabi.KERNEL_AUTO]                       # MIDN_ONLY_AUTO=1: only what AUTO picks
    if os.environ.get("MIDN_ONLY_AUTO") != "1":
        if N <= _cabi.MAX_MODES_SMALL:
            kernels.append(_cabi.KERNEL_SMALL)
        if _cabi.MIN_MODES_PAIR <= N <= _cabi.MAX_MODES_PAIR:
            kernels.append(_cabi.KERNEL_PAIR)
        if N >= 9:
            kernels.append(_cabi.KERNEL_STRUCT)
    ref, auto_kernel = None, None
    for kernel in kernels:
        sweep.batch.kernel = kernel
        plan = sweep.eng.ctx.plan(sweep.batch)
        if kernel == _cabi.KERNEL_AUTO:
            auto_kernel = plan.kernel
        elif plan.kernel == auto_kernel:
            continue                                    # already timed as AUTO
        for _ in range(2):
            sweep.launch()
        torch.cuda.synchronize()
        mm = sweep.fetch()[0]
        if ref is None:
            ref = mm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sweep.launch_kernel()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fl = _cabi.flops_per_fit(sweep.rows_max, N, 1, True) * 16384
        print(N, 'kernel', plan.kernel, 'ms %.3f' % ms, 'fits/s %.3g' % (16384 / ms * 1e3), 'TF %.2f' % (fl / ms * 1e-9),
              'block', plan.block, 'lpf', plan.lanes_per_fit, 'regs', plan.regs_per_thread,
              'max|dmm| %.1e' % float(np.max(np.abs(mm - ref))), flush=True)
