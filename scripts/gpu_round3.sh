#!/bin/bash
# GPU box round: parity tests, smoke, both bench arms, op sweep, ncu launch list + host profile.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -1 gpurun_out/bench_ref.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log
timeout 900 python bench_ops.py --quick --n 100000 1000000 > gpurun_out/ops.jsonl 2> gpurun_out/ops.err; echo "ops rc=$?"
tail -3 gpurun_out/ops.err
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_bench.log
