"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (and per grid size) totals."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    by_grid = len(sys.argv) > 2 and sys.argv[2] == "grid"
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path, errors="ignore")):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", d["Kernel Name"])
        if by_grid:
            k += " grid=" + d.get("Grid Size", "") + " blk=" + d.get("Block Size", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print(f"{v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}% n={v[0]:4d} {k[:120]}")
    print(f"total {tot / 1e6:.3f} ms")


if __name__ == "__main__":
    main()
