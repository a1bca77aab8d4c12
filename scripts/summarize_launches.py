"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {total/1e3:.3f} ms total device time (serialised, cold cache)")
    print(f"{'share':>7} {'total_ms':>10} {'count':>6} {'avg_us':>10}  kernel")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us/total*100:6.2f}% {us/1e3:10.3f} {cnt:6d} {us/cnt:10.2f}  {name}")


if __name__ == "__main__":
    main(sys.argv[1])
