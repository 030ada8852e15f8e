"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last step only).

    python tools/summarize_launches.py gpurun_out/launches.csv LAUNCHES_PER_STEP > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys


def main():
    path, per = sys.argv[1], int(sys.argv[2])
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
    last = list(zip(names[-per:], vals[-per:]))
    tot = sum(v for _, v in last)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in last:
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", n))
        agg[short][0] += 1
        agg[short][1] += v
    print(f"# ncu launch list: last step, {per} launches, {tot * 1e-6:.2f} ms serialised (cold cache, gpu__time_duration.sum)\n")
    print("| kernel | launches | total ms | share | ms / launch |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {v * 1e-6:.3f} | {100 * v / tot:.1f}% | {v * 1e-6 / c:.3f} |")


if __name__ == "__main__":
    main()
