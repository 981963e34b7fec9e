#!/bin/bash
# Runs on the GPU box (via gpurun): golden vectors from the reference kernels, GPU parity tests,
# smoke, bench (both arms) and an ncu launch list.  Every step is bounded by `timeout`.
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python oracle/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
tail -3 gpurun_out/golden.log
cp gpurun_out/golden/* tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -m pytest tests -m "not gpu" -q --no-header -p no:cacheprovider > gpurun_out/pytest_cpu.log 2>&1; echo "pytest cpu rc=$?"
tail -5 gpurun_out/pytest_cpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -3 gpurun_out/bench_ref.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
  tail -2 gpurun_out/ncu_bench.log
fi
