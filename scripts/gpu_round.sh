#!/bin/bash
# Runs on the GPU box (via gpurun): golden vectors from the reference kernels, GPU parity tests,
# smoke, and a short bench.  Every step is bounded by `timeout` so a hung kernel cannot eat the lease.
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python oracle/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
tail -5 gpurun_out/golden.log
cp -n gpurun_out/golden/* tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -15 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -15 gpurun_out/bench.log
