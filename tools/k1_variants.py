#!/usr/bin/env python
"""Developer tool: build variants of libqnmfit.so with -D switches and time the cfg3 grid
kernel with each (device-resident inputs, CUDA events).  Not part of the product.

    python tools/k1_variants.py build          # here (nvcc cross-compiles)
    python tools/k1_variants.py run            # on the GPU box
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_variants")

VARIANTS = {       # leaf-loop pipelining: H columns of the next block refilled after reflection c + S
    "h0": ["-DQNMFIT_PIPE_H(N)=0"],
    "h2": ["-DQNMFIT_PIPE_H(N)=((N)>=2?2:0)"],
    "h3": ["-DQNMFIT_PIPE_H(N)=((N)>=3?3:0)"],
    "h4": [],
    "h8s0": ["-DQNMFIT_PIPE_H(N)=(N)", "-DQNMFIT_PIPE_S(N)=0"],
}


def build():
    """Every variant is a full library build (all translation units, __graft_entry__.build_library)
    with the variant's -D switches, into tools/_variants/ with its own object directory."""
    import __graft_entry__ as ge
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        ge.build_library(force=True, extra_flags=flags, lib=os.path.join(OUT, f"libqnmfit_{name}.so"),
                         build_dir=os.path.join(ROOT, "build", "variant_" + name))


def run(steps=20):
    import numpy as np
    import torch
    import qnmfits_b200 as qf
    from qnmfits_b200 import _cabi, workloads
    from qnmfits_b200 import qnmfits as api
    from qnmfits_b200._engine import get_engine
    workloads.use_synthetic_tables()
    wl = workloads.config3(res=256)
    eng = get_engine(0)
    sweep, shape = api._prepare_M_chi_grid(wl.times, wl.data, wl.modes, wl.Mf_minmax, wl.chif_minmax,
                                           wl.t0, T=wl.T, res=256)
    stream = torch.cuda.current_stream()
    results = {}
    names = sys.argv[2:] or sorted(f[len('libqnmfit_'):-3] for f in os.listdir(OUT) if f.startswith('libqnmfit_') and f.endswith('.so'))
    for name in names:
        path = os.path.join(OUT, f"libqnmfit_{name}.so") if name != "default" else _cabi.LIB_PATH
        if not os.path.isfile(path):
            continue
        lib = _cabi.load_library(path)
        h = C.c_void_p()
        assert lib.qnmfit_create(0, C.byref(h)) == 0
        for anchor in (0,):
            for uw in (1,):
                sweep.batch.anchor_rows = anchor
                sweep.batch.uniform_weights = uw
                for _ in range(3):
                    rc = lib.qnmfit_fit_batch(h, C.byref(sweep.batch), C.c_void_p(stream.cuda_stream))
                    assert rc == 0, lib.qnmfit_last_error(h)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(steps):
                    lib.qnmfit_fit_batch(h, C.byref(sweep.batch), C.c_void_p(stream.cuda_stream))
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                plan = _cabi.Plan()
                lib.qnmfit_plan_batch(h, C.byref(sweep.batch), C.byref(plan))
                results[f"{name}/anchor{anchor}/fast{uw}"] = dict(
                    ms=round(ms, 4), lpf=plan.lanes_per_fit, grid=plan.grid, block=plan.block,
                    regs=plan.regs_per_thread, smem=plan.smem_bytes)
                print(f"{name:10s} anchor={anchor:4d} fast={uw} {ms:8.4f} ms  lpf={plan.lanes_per_fit} "
                      f"grid={plan.grid} block={plan.block} regs={plan.regs_per_thread}", flush=True)
        lib.qnmfit_destroy(h)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "k1_variants.json"), "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run()
