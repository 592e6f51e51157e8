#!/usr/bin/env python3
"""profiles/r2_launch_shares.txt from the ncu launch list (profiles/r2_launches.csv: ncu --metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r2_launches.csv"
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict(); tot = 0.0
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
    ms = float(r[idx["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[idx["Metric Unit"]], 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 3 --no-cpu-baseline (the first 400 launches: scene build, fast "
       "headline, exact side run, per-stage explanation pass; per-launch times are cold-cache and serialised: shares, not absolutes)",
       f"total kernel time {tot:.1f} ms over {len(rows) - 1} launches"]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    out.append(f"{100 * ms / tot:6.2f} % {ms:10.2f} ms {n:5d} x {ms / n:9.3f} ms  {k}")
per = lambda pat: next((ms / n for k, (n, ms) in agg.items() if re.search(pat, k)), 0.0)
f, rg, rs, ft = per(r"ptb_fast::k_chunk_fused"), per(r"ptb_fast::k_chunk_raygen"), per(r"ptb::k_resolve"), per(r"ptb::k_fold_totals")
out.append(f"one timed step (fast) = k_chunk_raygen {rg:.3f} + k_chunk_fused {f:.3f} + k_resolve {rs:.3f} + k_fold_totals {ft:.3f} ms: "
           f"the fused kernel is {100 * f / (f + rg + rs + ft):.1f} % of the step (bench.py share_of_step by CUDA events: see the bench line)")
(ROOT / "profiles" / "r2_launch_shares.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[1:5] + out[-1:]))
