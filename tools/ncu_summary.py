#!/usr/bin/env python
"""Reduce `ncu --page raw --csv` dumps to a per-kernel summary (JSON) that is small enough to commit and that bench.py
reads for the roofline block's `traffic` field.

    python tools/ncu_summary.py profiles/r01_ncu_*.csv > profiles/r01_ncu_summary.json
"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0,
        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return re.sub(r"zvx::|<unnamed>::|unnamed>::|\(anonymous namespace\)::", "", name)


def main():
    out = {}
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            k = short(r[col["Kernel Name"]])
            e = out.setdefault(k, {"launches": 0, "source": []})
            e["launches"] += 1
            if path not in e["source"]:
                e["source"].append(path)
            for m, key in WANT.items():
                if m not in col or r[col[m]] in ("", "n/a"):
                    continue
                v = float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
                e.setdefault(key, []).append(v)
    summ = {}
    for k, e in out.items():
        s = {"launches_captured": e["launches"], "source": e["source"]}
        for key in WANT.values():
            if key in e:
                s["avg_" + key] = sum(e[key]) / len(e[key])
        if "dram_read" in e and "dram_write" in e:
            s["avg_dram_traffic_bytes"] = (sum(e["dram_read"]) + sum(e["dram_write"])) / e["launches"]
            if "duration" in e:   # achieved HBM bandwidth of the kernel (DRAM bytes moved / kernel duration), GB/s
                s["avg_dram_gbs"] = (sum(e["dram_read"]) + sum(e["dram_write"])) / sum(e["duration"]) / 1e9
        summ[k] = s
    json.dump(summ, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
