#!/usr/bin/env python
"""Developer tool: decode ptxas stall counts from cuobjdump -sass output and report, for
the R-combine loop of K1 (a clean copy of the absorb code), the sum of stall counts
(single-warp issue time) next to the FP64 pipe cycles it needs."""
import collections
import re
import subprocess
import sys


def load(so, fun):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, so], capture_output=True, text=True).stdout.split("\n")
    ins, i = [], 0
    while i < len(txt):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", txt[i])
        if m and i + 1 < len(txt):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", txt[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF))
                i += 2
                continue
        i += 1
    return ins


def op(t):
    return (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]


def loops(ins):
    addr = {a: i for i, (a, _, _) in enumerate(ins)}
    out = []
    for i, (a, t, _) in enumerate(ins):
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                out.append((addr[tgt], i))
    return out


def main():
    so = sys.argv[1]
    fun = sys.argv[2] if len(sys.argv) > 2 else "_Z16fit_small_kernelILi8ELi256ELb1EEv9FitParams"
    ins = load(so, fun)
    print("instructions", len(ins))
    for s, e in sorted(loops(ins), key=lambda x: x[1] - x[0], reverse=True)[:6]:
        seg = ins[s:e + 1]
        c = collections.Counter(op(t) for _, t, _ in seg)
        fp = c["DFMA"] + c["DMUL"] + c["DADD"]
        if fp < 200:
            continue
        stall = sum(max(st, 1) for _, _, st in seg)
        print(f"loop {s}-{e}: n={len(seg)} fp64={fp} sum_stall={stall} cycles/fp64={stall / fp:.2f} "
              f"LDS={c['LDS']} STS={c['STS']} LDL={c['LDL']} STL={c['STL']} MUFU={c['MUFU']}")


if __name__ == "__main__":
    main()
