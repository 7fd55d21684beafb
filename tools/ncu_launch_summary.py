#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.

    python tools/ncu_launch_summary.py launches.csv [first_id [last_id]] > summary.md
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            i = int(r["ID"])
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        if lo <= i <= hi:
            rows.append((r["Kernel Name"], ms))
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values()) or 1.0
    print("| kernel | launches | total (ms) | share |")
    print("|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (k[:60], n, ms, 100 * ms / total))
    print("\ntotal: %d launches, %.3f ms" % (len(rows), total))


if __name__ == "__main__":
    main()
