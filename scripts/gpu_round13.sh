#!/bin/bash
# information round: parity, torch small-kernel attribution of the step, ncu --set full of every op kernel at sweep size
# and of the ShapeNet-sized conv the bench quotes, Morton-order sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_fills.py > gpurun_out/profile_fills.txt 2>&1; echo "profile_fills rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ln::' -c 60 -o gpurun_out/r01g_ops -f python scripts/ncu_ops.py > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -3 gpurun_out/ncu_ops.log
ncu -i gpurun_out/r01g_ops.ncu-rep --page raw --csv > gpurun_out/r01g_ops_raw.csv 2>/dev/null
N=2048 V=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|conv_wgrad_tc|filter_prep|group_norm' -c 30 -o gpurun_out/r01g_small -f python scripts/ncu_ops.py > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
ncu -i gpurun_out/r01g_small.ncu-rep --page raw --csv > gpurun_out/r01g_small_raw.csv 2>/dev/null
timeout 600 python bench_ops.py --quick --n 1000000 --vals 32 64 --order morton > gpurun_out/ops_morton.jsonl 2> gpurun_out/ops_morton.err; echo "ops morton rc=$?"
python scripts/show_ops.py gpurun_out/ops_morton.jsonl
ls -la gpurun_out
