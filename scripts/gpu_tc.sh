#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?"
cat gpurun_out/tc_debug.log | tail -20
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "tensor_core" > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
tail -15 gpurun_out/pytest_tc.log
timeout 900 python bench_ops.py --quick --n 100000 1000000 > gpurun_out/ops.jsonl 2> gpurun_out/ops.err; echo "ops rc=$?"
tail -3 gpurun_out/ops.err
cat gpurun_out/ops.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l)
    print(f\"{r['op'][:44]:44s} n={r['n']:8d} nv={r['nv']:8d} V={r.get('val_dim','-'):>4} {r['us']:10.1f}us  GB/s={r.get('GBps',0):8.1f} hbm={r.get('hbm_frac',0):.3f} TF={r.get('TFLOPs',0):7.2f} ref_us={r.get('ref_us',0):10.1f} x{r.get('speedup_vs_ref_kernel',0):.2f}\")
"
