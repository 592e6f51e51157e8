#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: opcode mix, SIMT efficiency,
stall reasons and the hottest instructions of the FIRST kernel in the file."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sect = rows[starts[0]:starts[1]] if len(starts) > 1 else rows[starts[0]:]
print(sect[0][1][:100])
hdr = sect[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in sect[2:] if len(r) >= len(hdr) - 1 and r[0].startswith("0x")]


def I(r, h):
    try:
        return int(float(r[ix[h]] or 0))
    except (ValueError, IndexError):
        return 0


tot = sum(I(r, "Instructions Executed") for r in data)
thr = sum(I(r, "Thread Instructions Executed") for r in data)
print(f"warp instructions {tot}, thread instructions {thr}, avg active threads {thr / max(tot, 1):.2f}, SASS lines {len(data)}")
mix, mthr = collections.Counter(), collections.Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] += I(r, "Instructions Executed"); mthr[op] += I(r, "Thread Instructions Executed")
for k, v in mix.most_common(20):
    print(f"  {k:10s} {v:11d} {v / tot:6.1%}  avg threads {mthr[k] / max(v, 1):5.1f}")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(I(r, h) for r in data) for h in st}
s = sum(agg.values()) or 1
print("stalls:", {k: round(v / s, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
if top_n:
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:top_n]:
        print(f"  samples {I(r, '# Samples'):6d} inst {I(r, 'Instructions Executed'):9d} thr {r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']].strip()[:90]}")
