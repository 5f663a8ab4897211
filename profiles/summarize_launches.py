#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print(f"total {T:.1f} us over {sum(cnt.values())} launches (cold-cache, serialised: compare shares)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d} avg={v / cnt[k]:8.2f} us  {k[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
