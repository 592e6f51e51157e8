#!/usr/bin/env python3
"""Static SASS statistics of one kernel in an object/cubin/.so: instruction count, code bytes, opcode histogram.
usage: sass_stats.py <file> <substring of the mangled name> [top_n]"""
import collections
import re
import subprocess
import sys

f, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
cur, stats = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        stats[cur][m.group(2).split(".")[0]] += 1
for name, c in stats.items():
    if pat in name:
        n = sum(c.values())
        print(f"{name}: {n} instructions, {n * 16 / 1024:.1f} KB")
        print("  " + ", ".join(f"{k} {v}" for k, v in c.most_common(top)))
