"""ncu raw CSV of scripts/ncu_bench_conv.py -> profiles/roofline_traffic.json (DRAM bytes per ln_conv_fwd call =
filter_prep + conv_tc2, last captured pair: warm L2 for the re-laid-out filter is what the step sees too).
    python scripts/roofline_traffic.py gpurun_out/bench_conv_raw.csv profiles/roofline_traffic.json"""
import csv
import json
import sys


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v.replace(",", "")) * mult


def main(src, dst):
    rows = [r for r in csv.reader(open(src, newline="")) if r]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr, units, data = rows[h], rows[h + 1], rows[h + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    per = []
    for r in data:
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        per.append({"kernel": r[col["Kernel Name"]].split("(")[0], "dram_read": rd, "dram_write": wr,
                    "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[col["gpu__time_duration.sum"]]]})
    last_prep = max(i for i, k in enumerate(per) if "filter_prep" in k["kernel"])
    pair = per[last_prep:last_prep + 2]
    out = {"source": "ncu --set full --clock-control none, scripts/ncu_bench_conv.py (B200)", "launches": pair,
           "traffic_bytes_per_call": sum(k["dram_read"] + k["dram_write"] for k in pair)}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
