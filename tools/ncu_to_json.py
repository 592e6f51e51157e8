#!/usr/bin/env python3
"""Turns one kernel of an `ncu -i X.ncu-rep --page raw --csv` export into the entry bench.py reads for its roofline block.

usage: ncu_to_json.py raw.csv <key> <segments_per_launch> <source text> [kernel substring]
  key  e.g. "c2:fast:pipeline3" (config : arithmetic mode : pipeline) -> profiles/r2_ncu_summary.json[key]
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
raw, key, segments, source = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
pat = sys.argv[5] if len(sys.argv) > 5 else "fused"
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
row = [d for d in data if pat in d[idx["Kernel Name"]]][0]


def val(name, scale_unit=True):
    v = float(row[idx[name]].replace(",", ""))
    if scale_unit:
        u = units[idx[name]]
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v


warp_inst = val("smsp__inst_executed.sum")
lanes = val("smsp__thread_inst_executed_per_inst_executed.ratio")
stall = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(row[i].replace(",", ""))
         for k, i in idx.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and row[i]}
entry = {
    "kernel": row[idx["Kernel Name"]].split("(")[0], "source": source, "segments_per_launch": segments,
    "time_ms_under_ncu": val("gpu__time_duration.sum"),
    "registers": val("launch__registers_per_thread", False), "grid": val("launch__grid_size", False),
    "occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active", False),
    "issue_slot_util": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False) / 100.0,
    "lanes_per_inst": lanes, "warp_inst": warp_inst, "thread_inst_per_segment": warp_inst * lanes / segments,
    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
    "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
    "l2_throughput_pct": val("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
    "l1_throughput_pct": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", False),
    "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct", False), "l2_hit_pct": val("lts__t_sector_hit_rate.pct", False),
    "stalls_per_issue": {k: round(v, 3) for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]},
}
entry["dram_bytes_per_launch"] = entry["dram_bytes_read"] + entry["dram_bytes_write"]
out = ROOT / "profiles" / "r2_ncu_summary.json"
allv = json.loads(out.read_text()) if out.exists() else {}
allv[key] = entry
out.write_text(json.dumps(allv, indent=1) + "\n")
print(key, json.dumps(entry))
