#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys


def main(path, tail=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict(); seq = []
    for x in csv.DictReader(lines):
        n = x["Kernel Name"]; v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        n = n.replace("<unnamed>::", "").split("(")[0][:70]
        seq.append((n, v)); a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(t for _, t in agg.values())
    print(f"{len(seq)} launches, {tot:.3f} ms total")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.3f} ms {100 * t / tot:5.1f}% {c:5d}  {n}")
    for n, v in seq[-tail:] if tail else []:
        print(f"{v:9.3f} {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
