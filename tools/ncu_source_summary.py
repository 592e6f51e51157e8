#!/usr/bin/env python3
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export: share of stall samples, of executed
warp instructions, active lanes per instruction and the top stall reasons.  usage: ncu_source_summary.py file.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
tot = collections.Counter(); inst = collections.Counter(); tinst = collections.Counter()
stall = collections.defaultdict(collections.Counter); srcs = {}
for k, hi in enumerate(hdr_idx):
    fname = rows[hi - 2][1].split('/')[-1]; hdr = rows[hi]
    end = hdr_idx[k + 1] - 2 if k + 1 < len(hdr_idx) else len(rows)
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr) or r[0] == '':
            continue
        key = (fname, int(r[0]))
        def f(x):
            try:
                return float(r[ci[x]])
            except (ValueError, KeyError):
                return 0.0
        tot[key] += f('# Samples'); inst[key] += f('Instructions Executed'); tinst[key] += f('Thread Instructions Executed'); srcs[key] = r[1]
        for h in hdr:
            if h.startswith('stall_') and 'Not Issued' not in h:
                stall[key][h] += f(h)
T = sum(tot.values()); I = sum(inst.values())
print(f"total stall samples {T:.0f}, warp instructions {I:.4g}, thread instructions {sum(tinst.values()):.4g}")
by = collections.defaultdict(lambda: [0, 0, 0])
for k in tot:
    by[k[0]][0] += tot[k]; by[k[0]][1] += inst[k]; by[k[0]][2] += tinst[k]
for f_, (a, b, c) in sorted(by.items(), key=lambda x: -x[1][0]):
    print(f"{f_:32s} samples {a / T * 100:5.1f}%  inst {b / I * 100:5.1f}%  lanes/inst {c / max(b, 1):4.1f}")
for key, v in tot.most_common(top):
    st = ', '.join(f"{a[6:]}={b / v * 100:.0f}%" for a, b in stall[key].most_common(3))
    print(f"{key[0]}:{key[1]:4d} samp {v / T * 100:5.2f}% inst {inst[key] / I * 100:5.2f}% lanes {tinst[key] / max(inst[key], 1):4.1f} [{st}] {srcs[key].strip()[:84]}")
