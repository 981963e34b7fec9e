#!/bin/bash
# static-shape / CUDA-graph step: tests, bench in both modes, ncu --set full of the hot kernels at sweep size
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "static or graphed or padding or fused_group_norm" > gpurun_out/pytest_static.log 2>&1; echo "pytest static rc=$?"
tail -15 gpurun_out/pytest_static.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -2 gpurun_out/bench_graph.log | cut -c1-1500
timeout 600 python bench.py --steps 30 --warmup 5 --mode eager > gpurun_out/bench_eager.log 2>&1; echo "bench eager rc=$?"
tail -1 gpurun_out/bench_eager.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ln::' -o gpurun_out/r01b_ops_n1M_v64 -f python scripts/ncu_ops.py > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -3 gpurun_out/ncu_ops.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -5 gpurun_out/pytest_gpu.log
