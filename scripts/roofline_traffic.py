"""ncu raw CSV of scripts/ncu_bench_conv.py -> profiles/roofline_traffic.json (DRAM bytes of one conv_tc3 launch -- the
kernel bench.py's `roofline` times; filter slabs are prepared once per step by a separate batched launch).
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
    convs = [k for k in per if "conv_tc3" in k["kernel"]]
    pair = convs[-1:]                     # the last launch: the prepared slabs and the values sit in L2, as inside the step
    out = {"source": "ncu --set full --clock-control none, scripts/ncu_bench_conv.py (B200)", "launches": pair,
           "traffic_bytes_per_call": sum(k["dram_read"] + k["dram_write"] for k in pair)}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
