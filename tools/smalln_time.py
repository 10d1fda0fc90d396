#!/usr/bin/env python
"""Developer tool (GPU box): K1 kernel time of the 256 x 256 grid for N = 1..8 columns."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from qnmfits_b200 import workloads, _cabi  # noqa: E402
from qnmfits_b200 import qnmfits as api  # noqa: E402

workloads.use_synthetic_tables()
wl = workloads.config3(res=256)
out = {}
for N in range(1, 9):
    sweep, shape = api._prepare_M_chi_grid(wl.times, wl.data, wl.modes[:N], wl.Mf_minmax, wl.chif_minmax, wl.t0,
                                           T=wl.T, res=256)
    for _ in range(3):
        sweep.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sweep.launch_kernel()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    plan = sweep.eng.ctx.plan(sweep.batch)
    flops = _cabi.flops_per_fit(sweep.rows_max, N, 1, True) * 65536
    out[N] = dict(ms=ms, tflops=flops / ms * 1e-9, lpf=plan.lanes_per_fit, grid=plan.grid, regs=plan.regs_per_thread,
                  smem=plan.smem_bytes)
    print(N, json.dumps(out[N]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "smalln_time.json"), "w"), indent=1)
