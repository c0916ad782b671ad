#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share / count / mean duration.
usage: launch_summary.py file.csv [first_launch [last_launch]]"""
import csv
import re
import sys
from collections import defaultdict


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    id_i = hdr.index("ID")
    for r in rd:
        if len(r) <= val_i:
            continue
        v = float(r[val_i].replace(",", ""))
        u = r[unit_i]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        rows.append((int(r[id_i]), r[name_i], us))
    return rows


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("(anonymous namespace)::", "")
    return name.split("::")[-1] if "<" not in name else name.replace("void ", "")


def main():
    rows = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = [r for r in rows if lo <= r[0] < hi]
    agg = defaultdict(lambda: [0, 0.0])
    for _, n, us in rows:
        a = agg[short(n)]
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot:.0f} us over {len(rows)} launches")
    print("| kernel | share | launches | avg us | total us |\n|---|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.003:
            continue
        print(f"| `{n}` | {100 * t / tot:.1f}% | {c} | {t / c:.1f} | {t:.0f} |")


if __name__ == "__main__":
    main()
