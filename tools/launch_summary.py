"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv
python bench.py --steps 2 --warmup 1`): launches, total and mean duration and share per kernel.
usage: python tools/launch_summary.py X.csv > profiles/<name>.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = re.sub(r"<unnamed>::", "", r[4])
        name = re.sub(r"\((Layout|Sweep2Args|SweepArgs|PwArgs).*", "", name).replace("void ", "")
        ns = float(r[14].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel':70s} {'n':>5s} {'total ms':>10s} {'mean us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {v[0]:5d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.1f} {100 * v[1] / tot:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1])
