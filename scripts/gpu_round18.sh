#!/bin/bash
# r01j: 2-GPU run of both bench arms exactly as the driver launches them (torchrun, one rank per GPU, NCCL), with and
# without the all-reduce captured inside the step graph.  Run with: gpurun --gpus 2
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
run2() {  # $1 = log name, rest = bench.py flags
  local name=$1; shift
  NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 30 --warmup 5 "$@" > $O/$name.log 2>&1; echo "$name rc=$?"
  grep '^{' $O/$name.log | tail -1 > $O/$name.json
  python - "$O/$name.json" <<'PY'
import json, sys
try:
    r = json.load(open(sys.argv[1]))
    print({k: r.get(k) for k in ("impl", "value", "n_gpus", "ms_per_step", "gpu_launches")}, r.get("e2e"), r.get("config", {}).get("replicas_bit_identical_after_run"), r.get("config", {}).get("execution", "")[:60])
except Exception as exc:
    print("no JSON line:", exc)
PY
}
run2 r01j_bench_2gpu --no-cpu-baseline
tail -4 $O/r01j_bench_2gpu.log | grep -v '^{' | cut -c1-300
run2 r01j_bench_2gpu_captured_allreduce --no-cpu-baseline --capture-collective
tail -4 $O/r01j_bench_2gpu_captured_allreduce.log | grep -v '^{' | cut -c1-300
run2 r01j_bench_2gpu_reference --impl reference --steps 10 --warmup 3
