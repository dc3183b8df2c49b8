#!/usr/bin/env python
"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into a short JSON."""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{r[i]} {units[i]}".strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.3:
                    stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 2)
        d["stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
