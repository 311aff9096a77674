#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid): share of the summed kernel
time, launches, average duration.  Usage: ncu_summarize.py launches.csv [title] > profiles/xxx.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    head = rows[0]
    col = {n: i for i, n in enumerate(head)}
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rows[1:]:
        if len(r) < len(head) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^(void )?(ss::)?(\(anonymous namespace\)::)?", "", r[col["Kernel Name"]])
        name = re.sub(r"\(.*$", "", name)
        unit = r[col["Metric Unit"]]
        v = float(r[col["Metric Value"]].replace(",", ""))
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        key = (name, r[col["Grid Size"]].replace(" ", ""), r[col["Block Size"]].replace(" ", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    print(f"# {title}\n")
    print(f"{n} launches, {total:.0f} us summed kernel time (ncu: cold-cache, serialised -- compare shares, not absolutes).\n")
    print("| share | launches | avg us | total us | kernel | grid | block |")
    print("|---|---|---|---|---|---|---|")
    for (name, grid, block), (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * us / total:.1f}% | {cnt} | {us / cnt:.2f} | {us:.0f} | `{name}` | {grid} | {block} |")
    by_name = collections.defaultdict(lambda: [0, 0.0])
    for (name, _, _), (cnt, us) in agg.items():
        by_name[name][0] += cnt
        by_name[name][1] += us
    print("\n## by kernel\n")
    print("| share | launches | avg us | kernel |")
    print("|---|---|---|---|")
    for name, (cnt, us) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * us / total:.1f}% | {cnt} | {us / cnt:.2f} | `{name}` |")


if __name__ == "__main__":
    main()
