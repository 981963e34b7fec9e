#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "tensor_core or conv" > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv rc=$?"
tail -5 gpurun_out/pytest_conv.log
LN_CONV_TRUNC=1 timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "tensor_core" > gpurun_out/pytest_trunc.log 2>&1; echo "pytest trunc rc=$?"
tail -4 gpurun_out/pytest_trunc.log
timeout 600 python bench_ops.py --quick --n 1000000 --vals 32 64 128 > gpurun_out/ops_a.jsonl 2> gpurun_out/ops_a.err; echo "ops default rc=$?"
LN_CONV_LOOKAHEAD=1 timeout 600 python bench_ops.py --quick --n 1000000 --vals 64 > gpurun_out/ops_p1.jsonl 2>/dev/null; echo "ops P=1 rc=$?"
LN_CONV_LOOKAHEAD=2 timeout 600 python bench_ops.py --quick --n 1000000 --vals 64 > gpurun_out/ops_p2.jsonl 2>/dev/null; echo "ops P=2 rc=$?"
LN_CONV_LOOKAHEAD=4 timeout 600 python bench_ops.py --quick --n 1000000 --vals 64 > gpurun_out/ops_p4.jsonl 2>/dev/null; echo "ops P=4 rc=$?"
LN_CONV_TRUNC=1 timeout 600 python bench_ops.py --quick --n 1000000 --vals 64 > gpurun_out/ops_trunc.jsonl 2>/dev/null; echo "ops trunc rc=$?"
python scripts/show_ops.py gpurun_out/ops_a.jsonl
python scripts/show_ops.py gpurun_out/ops_p1.jsonl gpurun_out/ops_p2.jsonl gpurun_out/ops_p4.jsonl gpurun_out/ops_trunc.jsonl
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -3 gpurun_out/pytest_gpu.log
