"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel tables.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv[.gz] > profiles/rNN_launches.md

Two tables: (1) ONE training step -- the launches between the last two optimizer kernels (ln::adamw_amsgrad_kernel), i.e. one
replay of the step graph, which is what the shares of the step must be read from; (2) the whole capture (which
also holds model initialisation, eager warm-up passes and the graph-capture pass)."""
import collections
import csv
import gzip
import re
import sys


def load(path):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:90]
        launches.append((name, v))
    return launches


def table(launches, top=45):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v in launches:
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("ln::"))
    ours_n = sum(v[0] for k, v in agg.items() if k.startswith("ln::"))
    print(f"{n} launches, {tot / 1e3:.2f} ms of kernel time (per-launch times are cold-cache and serialised under ncu; use the SHARES).")
    print(f"Kernels of this repo (`ln::*`): {ours_n} launches, {100 * ours / tot:.1f}% of kernel time.\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.2f} |")
    print()


def last_step(launches, marker="adamw_amsgrad"):
    hits = [i for i, (n, _) in enumerate(launches) if marker in n]
    groups = []
    for i in hits:
        if groups and i - groups[-1][-1] < 20:
            groups[-1].append(i)
        else:
            groups.append([i])
    if len(groups) < 2:
        return None
    return launches[groups[-2][-1] + 1: groups[-1][-1] + 1]


def main(path):
    launches = load(path)
    print(f"# ncu launch list summary: {path}\n")
    step = last_step(launches)
    if step:
        print("## one training step (one replay of the step graph: forward + loss + backward + AdamW)\n")
        table(step)
    print("## whole capture (initialisation + eager warm-up + graph capture + replays)\n")
    table(launches, top=30)


if __name__ == "__main__":
    main(sys.argv[1])
