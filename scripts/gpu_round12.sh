#!/bin/bash
# re-entry check: parity, both bench arms, op sweep, graph launch list, full ncu captures of the op kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -1 gpurun_out/bench_graph.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 python bench_ops.py --quick --n 1000000 --vals 8 32 64 128 > gpurun_out/ops_a.jsonl 2> gpurun_out/ops_a.err; echo "ops rc=$?"
python scripts/show_ops.py gpurun_out/ops_a.jsonl
timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_graph.log 2>&1; echo "ncu launches graph rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_graph.csv > gpurun_out/launches_graph.md 2>&1; head -30 gpurun_out/launches_graph.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|conv_wgrad_tc|slice_fwd|slice_bwd|splat_accumulate|splat_build|neighbour_table|slice_classify|gather_fwd|scatter' -o gpurun_out/r01e_ops -f python scripts/ncu_ops.py > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -3 gpurun_out/ncu_ops.log
ls -la gpurun_out
