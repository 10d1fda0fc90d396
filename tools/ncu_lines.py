#!/usr/bin/env python
"""Developer tool: aggregate the SASS-level samples of an ncu report by CUDA source line.

    tools/ncu_lines.py report.ncu-rep object.o [kernel-regex] [top]

Exports the source page (sass) of the report, disassembles the object with line info
(cuobjdump -xelf + nvdisasm -g) and sums samples / executed instructions per file:line."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_map(obj):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, cur, func = {}, None, None
    for ln in txt.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            func = m.group(1)
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', ln)
        if m and func:
            out.setdefault(func, {})[int(m.group(1), 16)] = cur
    return out


def main():
    rep, obj = sys.argv[1], sys.argv[2]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernel = rows[0][1]
    hdr, data = rows[1], [r for r in rows[2:] if len(r) >= len(rows[1])]
    idx = {h: i for i, h in enumerate(hdr)}
    maps = line_map(obj)
    fn = [k for k in maps if kernel.split("(")[0].split("<")[0] in k][0]
    lm = maps[fn]
    base = int(data[0][idx["Address"]], 16)
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = 0
    for r in data:
        off = int(r[idx["Address"]], 16) - base
        key = lm.get(off)
        n = int(r[idx["# Samples"]] or 0)
        per[key][0] += n
        per[key][1] += int(r[idx["Instructions Executed"]] or 0)
        for h in stalls:
            v = int(r[idx[h]] or 0)
            if v:
                per[key][2][h[6:]] += v
        total += n
    print(f"{kernel}: {total} samples")
    for key, (n, ex, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join(f"{k} {v}" for k, v in st.most_common(3))
        print(f"{str(key):38s} {n:7d} {100 * n / total:5.1f}%  exec {ex:9d}  {why}")


if __name__ == "__main__":
    main()
