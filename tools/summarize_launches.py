"""Aggregate an ncu `--csv` launch list (gpu__time_duration.sum) by kernel: count, total, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""), val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, grid, blk, us in rows:
        short = re.sub(r"\(.*", "", name)
        m = re.search(r"gemm_tap_kernel<(\d+), *(\d+)>", name)
        if m:
            short = f"gemm_tap_kernel<{m.group(1)},{m.group(2)}>"
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms total (serialised, cold-cache: compare shares)")
    print(f"{'kernel':60s} {'count':>6s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {n:6d} {us / 1e3:9.3f} {100 * us / total:6.1f}% {us / n:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
