#!/bin/bash
# conv v2 (persistent, cp.async) validation + comparison with v1, slice/splat tweaks
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "conv or tensor_core or static" > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv rc=$?"
tail -5 gpurun_out/pytest_conv.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench_ops.py --quick --n 1000000 > gpurun_out/ops_v2.jsonl 2> gpurun_out/ops_v2.err; echo "ops v2 rc=$?"
LN_CONV_TC_V1=1 timeout 600 python bench_ops.py --quick --n 1000000 > gpurun_out/ops_v1.jsonl 2> gpurun_out/ops_v1.err; echo "ops v1 rc=$?"
python - <<'PY'
import json
for tag in ("v2","v1"):
    for l in open(f"gpurun_out/ops_{tag}.jsonl"):
        r=json.loads(l)
        if tag=="v1" and "conv_fwd tc" not in r["op"]: continue
        print(tag, f"{r['op'][:44]:44s} n={r['n']:8d} nv={r['nv']:8d} V={r.get('val_dim','-'):>4} {r['us']:10.1f}us GB/s={r.get('GBps',0):8.1f} hbm={r.get('hbm_frac',0):.3f} TF={r.get('TFLOPs',0):7.2f} ref_us={r.get('ref_us',0):10.1f}")
PY
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2' -o gpurun_out/r01c_conv_tc2 -f python scripts/ncu_ops.py > gpurun_out/ncu_conv2.log 2>&1; echo "ncu rc=$?"
