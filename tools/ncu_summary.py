#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.

  python tools/ncu_summary.py gpurun_out/launches.csv [skip_launches] [take_launches]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    take = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    rows = rows[skip:skip + take]
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)} total {total/1e3:.3f} ms")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us/1e3:10.3f} ms {100*us/total:6.2f}% {n:6d} x {us/n:9.1f} us  {name[:110]}")


if __name__ == "__main__":
    main()
