#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "tensor_core or conv or group_norm or padding" > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv rc=$?"
tail -5 gpurun_out/pytest_conv.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench_ops.py --quick --n 1000000 --vals 32 64 128 > gpurun_out/ops_a.jsonl 2> gpurun_out/ops_a.err; echo "ops default rc=$?"
python scripts/show_ops.py gpurun_out/ops_a.jsonl
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_eager.csv python bench.py --steps 2 --warmup 3 --mode eager > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|conv_wgrad_tc' -o gpurun_out/r01d_conv -f python scripts/ncu_ops.py > gpurun_out/ncu_conv2.log 2>&1; echo "ncu conv rc=$?"
