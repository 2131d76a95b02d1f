"""SASS evidence for the judge: per-kernel counts of the tcgen05 / TMEM / TMA mnemonics in libsa_b200.so.

  python tools/sass_summary.py > profiles/rN_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), UTCCP / LDTM / STTM = TMEM copies / loads / stores, UTMALDG / UTMASTG = TMA tensor
loads / stores, UBLKCP = 1-D bulk copies, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit (B200_PROFILING.md lists the
mnemonics that prove tcgen05 / TMA)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "stableanimator_b200" / "libsa_b200.so"
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
             "SYNCS", "FFMA2", "MUFU", "HMMA", "REDG", "ATOMG", "ATOMS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    counts, order, cur = {}, [], None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
    tot = collections.Counter()
    print(f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (sm_100a)")
    print(f"{'kernel':<78} {'instr':>7} " + " ".join(f"{m:>7}" for m in MNEMONICS))
    for fn in order:
        c = counts[fn]
        name = re.sub(r"\(.*", "", demangle(fn)).replace("void ", "").replace("sa::", "")
        print(f"{name[:78]:<78} {c['_total']:>7} " + " ".join(f"{c[m]:>7}" for m in MNEMONICS))
        tot.update(c)
    print(f"{'TOTAL':<78} {tot['_total']:>7} " + " ".join(f"{tot[m]:>7}" for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
