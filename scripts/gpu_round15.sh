#!/bin/bash
# slab + persistent slice/scatter, two-phase splat_build, id-prefetching wgrad: parity + op sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench_ops.py --quick --n 1000000 --vals 8 32 64 128 > gpurun_out/ops_a.jsonl 2> gpurun_out/ops_a.err; echo "ops rc=$?"
python scripts/show_ops.py gpurun_out/ops_a.jsonl | grep -v "SIMT\|reference"
timeout 600 python bench_ops.py --quick --n 1000000 --vals 32 64 --order morton > gpurun_out/ops_morton.jsonl 2> gpurun_out/ops_morton.err; echo "ops morton rc=$?"
python scripts/show_ops.py gpurun_out/ops_morton.jsonl | grep -v "SIMT\|reference\|conv"
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-200
