#!/bin/bash
mkdir -p gpurun_out
LOSS=nll timeout 300 python scripts/diag_flaky.py > gpurun_out/diag_nll.log 2>&1; grep -E "dynamic\]|worst" gpurun_out/diag_nll.log | cut -c1-170 | head -30
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-200
LN_CONV_BWD_FORK=0 timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph_nofork.log 2>&1; echo "bench graph nofork rc=$?"
tail -1 gpurun_out/bench_graph_nofork.log | cut -c1-200
timeout 600 python bench_ops.py --quick --n 1000000 --vals 8 32 64 128 > gpurun_out/ops_a.jsonl 2> gpurun_out/ops_a.err; echo "ops rc=$?"
python scripts/show_ops.py gpurun_out/ops_a.jsonl | grep -v conv
timeout 600 python bench_ops.py --quick --n 1000000 --vals 32 64 --order morton > gpurun_out/ops_morton.jsonl 2> gpurun_out/ops_morton.err; echo "ops morton rc=$?"
python scripts/show_ops.py gpurun_out/ops_morton.jsonl
timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_graph.log 2>&1; echo "ncu launches graph rc=$?"
grep -v '^"' gpurun_out/launches_graph.csv | head -8
