#!/usr/bin/env python
"""Developer tool (here, after a GPU run): from an ncu report of the headline kernel write
profiles/ncu_k1_<tag>_raw_subset.json (the raw metrics worth keeping), profiles/ncu_k1_<tag>_details.txt
and profiles/ncu_summary.json (dram bytes per launch, read by bench.py for roofline.traffic).
  python tools/ncu_summary.py gpurun_out/prof_k1_r02.ncu-rep r02 "what was captured"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic")
KEEP_PREFIX = ("sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "sm__ops_path_tensor_src_fp64")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    what = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, values = rows[0], rows[1], rows[2]
    subset, dram = {}, 0.0
    for n, u, v in zip(names, units, values):
        if n in KEEP or n.startswith(KEEP_PREFIX) or n == "Kernel Name":
            subset[n] = {"unit": u, "value": v}
        if n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram += float(v.replace(",", "")) * UNIT.get(u, 1.0)
    out = os.path.join(ROOT, "profiles", f"ncu_k1_{tag}_raw_subset.json")
    json.dump(subset, open(out, "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(ROOT, "profiles", f"ncu_k1_{tag}_details.txt"), "w").write(det)
    ms = float(subset["gpu__time_duration.sum"]["value"].replace(",", ""))
    ms *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(subset["gpu__time_duration.sum"]["unit"], 1.0)
    summary = {"dram_bytes_per_launch": int(round(dram)),
               "source": f"profiles/{os.path.basename(out)} (ncu --set full --clock-control none, "
                         f"{subset.get('Kernel Name', {}).get('value', 'fit_small_kernel')}, {what}): "
                         f"dram__bytes_read.sum {subset['dram__bytes_read.sum']['value']} {subset['dram__bytes_read.sum']['unit']} + "
                         f"dram__bytes_write.sum {subset['dram__bytes_write.sum']['value']} {subset['dram__bytes_write.sum']['unit']}",
               "kernel_ms_under_ncu": ms}
    json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"))
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
