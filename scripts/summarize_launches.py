"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:90]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("ln::"))
    print(f"# ncu launch list summary: {path}\n")
    print(f"{n} launches, {tot / 1e3:.2f} ms of kernel time (per-launch times are cold-cache and serialised under ncu; use the SHARES).")
    print(f"Kernels of this repo (`ln::*`): {100 * ours / tot:.1f}% of kernel time.\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.2f} |")


if __name__ == "__main__":
    main(sys.argv[1])
