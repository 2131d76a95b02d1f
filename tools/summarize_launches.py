"""Aggregate an ncu `--csv` launch list by kernel: count, total time, share, and (when collected) DRAM
bytes and tensor-pipe activity.  Usage: python tools/summarize_launches.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

SCALE_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
SCALE_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def short_name(name):
    name = name.replace("void ", "").replace("sa::", "")
    m = re.match(r"(\w+)<(.*?)>\(", name)
    if m:
        args = re.sub(r"\((int|bool)\)", "", m.group(2))
        return f"{m.group(1)}<{args}>"
    return re.sub(r"\(.*", "", name)


def main(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = defaultdict(dict)
    order = []
    for r in csv.DictReader(lines):
        k = r["ID"]
        if k not in per:
            order.append(k)
            per[k]["name"] = short_name(r["Kernel Name"])
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        m, u = r["Metric Name"], r.get("Metric Unit", "")
        if m == "gpu__time_duration.sum":
            per[k]["us"] = val * SCALE_US.get(u, 1e-3)
        elif m.startswith("dram__bytes_read"):
            per[k]["rd"] = val * SCALE_MB.get(u, 1e-6)
        elif m.startswith("dram__bytes_write"):
            per[k]["wr"] = val * SCALE_MB.get(u, 1e-6)
        elif m.startswith("sm__pipe_tensor_cycles_active"):
            per[k]["tensor"] = val
        elif m.startswith("sm__pipe_tensor_subpipe_hmma_cycles_active"):
            per[k].setdefault("tensor", val)  # % of active cycles (no elapsed-based metric in this pass)
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for k in order:
        d = per[k]
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d.get("us", 0.0)
        a[2] += d.get("rd", 0.0)
        a[3] += d.get("wr", 0.0)
        a[4] += d.get("tensor", 0.0) * d.get("us", 0.0)
    total = sum(v[1] for v in agg.values())
    print(f"{len(order)} launches, {total / 1e3:.3f} ms total (serialised, cold-cache: compare shares)")
    print(f"{'kernel':50s} {'count':>5s} {'ms':>8s} {'share':>6s} {'avg us':>8s} {'dramRd MB/l':>11s} {'dramWr MB/l':>11s} {'tensor%':>8s}")
    for k, (n, us, rd, wr, tw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:50]:50s} {n:5d} {us / 1e3:8.3f} {100 * us / total:5.1f}% {us / n:8.1f} {rd / n:11.1f} {wr / n:11.1f} "
              f"{(tw / us if us else 0):8.1f}")
    g = [v for k, v in agg.items() if k.startswith("gemm_tap")]
    if g:
        n = sum(v[0] for v in g)
        us = sum(v[1] for v in g)
        print(f"gemm_tap family: {n} launches, {us / 1e3:.3f} ms, DRAM traffic {sum(v[2] + v[3] for v in g) / n:.1f} MB per launch "
              f"(read {sum(v[2] for v in g) / n:.1f} + write {sum(v[3] for v in g) / n:.1f}), time-weighted tensor-pipe activity "
              f"{sum(v[4] for v in g) / us:.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
