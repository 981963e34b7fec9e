#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
if l:
    r=json.loads(l[-1]); print({k:r[k] for k in ('value','ms_per_step','gpu_launches')}, r['e2e'], r['roofline'], r['cpu_baseline'])
else:
    print(open('gpurun_out/bench.log').read()[-3000:])
PY
timeout 600 python bench.py --steps 30 --warmup 5 --conv-precision 0 > gpurun_out/bench_p0.log 2>&1; echo "bench p0 rc=$?"
grep -o '"value": [0-9.]*, "unit": "scans/s", "n_gpus"' gpurun_out/bench_p0.log
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
head -60 gpurun_out/profile_step.log
