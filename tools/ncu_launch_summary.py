#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/README.md)."""
import collections
import csv
import sys


def main(path, top=25):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    tot, cnt, mx = collections.Counter(), collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if "gpu__time_duration.sum" not in row.get("Metric Name", ""):
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        tot[k] += v; cnt[k] += 1; mx[k] = max(mx[k], v)
    T = sum(tot.values())
    print(f"{'ms':>10} {'share':>6} {'n':>5} {'max ms':>9}  kernel")
    for k, v in tot.most_common(top):
        print(f"{v:10.2f} {100 * v / T:5.1f}% {cnt[k]:5d} {mx[k]:9.2f}  {k[:100]}")
    print(f"{T:10.2f} total over {sum(cnt.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
