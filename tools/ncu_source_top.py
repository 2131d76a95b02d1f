"""Top stalled SASS instructions from `ncu -i rep --page source --csv --print-source sass -s N -c 1`."""
import csv
import sys


def main(path, n=40):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi]) and r[0] != "Address"]
    idx = {h: i for i, h in enumerate(hdr)}
    print(rows[0][1][:150] if rows[0] else "")
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall}
    print("total samples", tot, "| stall mix:", ", ".join(f"{k[6:]}={100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
        s = int(r[idx["# Samples"]] or 0)
        st = sorted(((h[6:], int(r[idx[h]] or 0)) for h in stall), key=lambda kv: -kv[1])[:2]
        print(f"{r[idx['Address']][-5:]} {s:6d} {100 * s / max(tot, 1):5.1f}%  {r[idx['Source']][:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
