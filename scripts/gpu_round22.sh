#!/bin/bash
# r01n: flip-robust model-level gradient tests (fresh process x3 each) + the full GPU suite with -x, as the driver runs it
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "lnn_model_matches_cpu_port or graphed_step_matches" 2>&1 | grep -E "passed|failed|AssertionError" | head -4
done
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider --durations=5 > $O/r01n_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01n_pytest_gpu.txt
