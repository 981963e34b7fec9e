#!/bin/bash
# flaky-test diagnosis + ncu --set full of every op kernel at sweep size (kernel-name regex fixed)
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "graphed_step_matches or static_shape" 2>&1 | grep -E "passed|failed|AssertionError" | head -3
done
PREC=1 LOSS=nll timeout 600 python scripts/diag_flaky.py > gpurun_out/diag_flaky_prec1.txt 2>&1; echo "diag rc=$?"
PREC=0 LOSS=nll timeout 600 python scripts/diag_flaky.py > gpurun_out/diag_flaky_prec0.txt 2>&1; echo "diag rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|conv_wgrad_tc|slice_fwd|scatter_rows|splat_build|neighbour_table|slice_classify|gather_fwd|filter_prep' -c 40 -o gpurun_out/r01g_ops -f python scripts/ncu_ops.py > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -3 gpurun_out/ncu_ops.log
ncu -i gpurun_out/r01g_ops.ncu-rep --page raw --csv > gpurun_out/r01g_ops_raw.csv 2>/dev/null
ls -la gpurun_out
