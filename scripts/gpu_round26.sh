#!/bin/bash
# r01r: collision / contention parity test + the full GPU suite with -x at HEAD
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "duplicate_points" 2>&1 | grep -E "passed|failed|Error" | head -5
timeout 600 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > $O/r01r_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -2 $O/r01r_pytest_gpu.txt
