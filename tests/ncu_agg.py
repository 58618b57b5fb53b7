"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tests/ncu_agg.py file.csv [N]"""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% {n:5d} x {ms / n * 1e3:8.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
