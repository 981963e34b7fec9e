#!/bin/bash
# r01l: N-GPU run of bench.py exactly as the driver launches it (default flags: all-reduce captured in the step graph).
# usage: gpurun --gpus N -- bash scripts/gpu_round20.sh N
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 30 --warmup 5 > $O/r01l_bench_${N}gpu.log 2>&1; echo "bench ${N}gpu rc=$?"
grep '^{' $O/r01l_bench_${N}gpu.log | tail -1 > $O/r01l_bench_${N}gpu.json
python - "$O/r01l_bench_${N}gpu.json" <<'PY'
import json, sys
try:
    r = json.load(open(sys.argv[1]))
    print({k: r.get(k) for k in ("value", "n_gpus", "ms_per_step", "gpu_launches")}, r.get("e2e"), r.get("config", {}).get("replicas_bit_identical_after_run"), r.get("config", {}).get("execution", "")[:80], r.get("clocks"))
except Exception as exc:
    print("no JSON line:", exc)
PY
tail -5 $O/r01l_bench_${N}gpu.log | grep -v '^{' | cut -c1-300
