#!/usr/bin/env python
"""Developer tool: summarise the SASS page of an ncu report (csv from
`ncu -i X.ncu-rep --page source --csv --print-source sass`): instruction mix, stall mix,
and the hot loop (instructions sharing the modal executed count) in windows."""
import collections
import csv
import sys


def op(t):
    p = t.strip().split()
    o = p[1] if p[0].startswith('@') else p[0]
    return o.split('.')[0]


def main():
    path = sys.argv[1]
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(open(path)))
    hdr = rows[1]; rows = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = [int(r[ix['Instructions Executed']]) for r in rows]
    src = [r[ix['Source']].strip() for r in rows]
    smp = [int(r[ix['# Samples']]) for r in rows]
    keys = ['stall_wait', 'stall_selected', 'stall_math', 'stall_not_selected', 'stall_short_sb', 'stall_dispatch',
            'stall_mio', 'stall_long_sb', 'stall_branch_resolving', 'stall_no_inst', 'stall_barrier']
    st = {k: [int(r[ix[k]]) for r in rows] for k in keys}
    T = sum(ex)
    mix = collections.Counter()
    for e, s in zip(ex, src):
        mix[op(s)] += e
    fp = mix['DFMA'] + mix['DMUL'] + mix['DADD']
    print(f"warp instructions {T}  fp64 {fp} ({100 * fp / T:.1f}%)")
    print(' '.join(f"{k}:{100 * v / T:.1f}" for k, v in mix.most_common(12)))
    S = sum(smp)
    print("stalls %:", ' '.join(f"{k[6:]}:{100 * sum(v) / S:.1f}" for k, v in st.items()))
    cnt = collections.Counter()
    for e, s in zip(ex, smp):
        cnt[e] += s
    hot = cnt.most_common(1)[0][0]
    idx = [i for i, e in enumerate(ex) if e == hot]
    nfp = sum(1 for i in idx if op(src[i]) in ('DFMA', 'DMUL', 'DADD'))
    hs = sum(smp[i] for i in idx)
    print(f"hot loop: exec {hot}, {len(idx)} instrs, fp64 {nfp}, samples {hs} ({100 * hs / S:.1f}% of all)")
    print("hot stalls %:", ' '.join(f"{k[6:]}:{100 * sum(st[k][i] for i in idx) / hs:.1f}" for k in keys))
    mixh = collections.Counter(op(src[i]) for i in idx)
    print("hot mix:", dict(mixh.most_common()))
    if W:
        for s0 in range(0, len(idx), W):
            seg = idx[s0:s0 + W]
            tot = sum(smp[i] for i in seg)
            f = sum(1 for i in seg if op(src[i]) in ('DFMA', 'DMUL', 'DADD'))
            print(s0, f"fp64 {f}/{len(seg)} samples/instr {tot / len(seg):.1f}",
                  ' '.join(f"{k[6:10]}:{sum(st[k][i] for i in seg)}" for k in keys[:6]))
    if len(sys.argv) > 3:
        with open(sys.argv[3], 'w') as f:
            for n, i in enumerate(idx):
                f.write(f"{n:5d} {i:6d} {smp[i]:5d} " + ' '.join(f"{st[k][i]:4d}" for k in keys[:6]) + "  " + src[i] + "\n")


if __name__ == '__main__':
    main()
