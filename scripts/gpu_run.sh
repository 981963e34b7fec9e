#!/bin/bash
# One parametrised GPU-box script (replaces the per-round gpu_round*.sh files).
#   scripts/gpu_run.sh <tag> [steps...]     steps: tests | tests-k:<expr> | smoke | bench | bench-ref | ops | scenes | launches | ncu-ops
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
TAG=${1:-r02}; shift
O=gpurun_out; mkdir -p $O
for step in "$@"; do
  case "$step" in
    tests)      timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest_gpu.txt ;;
    tests-x)    timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > $O/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -30 $O/${TAG}_pytest_gpu.txt ;;
    tests-k:*)  timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "${step#tests-k:}" > $O/${TAG}_pytest_k.txt 2>&1; echo "pytest -k rc=$?"; tail -40 $O/${TAG}_pytest_k.txt ;;
    smoke)      timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/${TAG}_smoke.txt ;;
    bench)      timeout 600 python bench.py --steps 30 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err ;;
    bench-quick) timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $O/${TAG}_bench_quick.json 2> $O/${TAG}_bench_quick.err; echo "bench rc=$?"; head -c 700 $O/${TAG}_bench_quick.json; tail -5 $O/${TAG}_bench_quick.err ;;
    bench-ref)  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "bench-ref rc=$?"; head -c 600 $O/${TAG}_bench_reference.json; tail -5 $O/${TAG}_bench_reference.err ;;
    ops)        timeout 900 python bench_ops.py > $O/${TAG}_ops_sweep.jsonl 2> $O/${TAG}_ops.err; echo "ops rc=$?"; tail -5 $O/${TAG}_ops.err ;;
    scenes)     timeout 900 python bench_scenes.py > $O/${TAG}_scenes.jsonl 2> $O/${TAG}_scenes.err; echo "scenes rc=$?"; cat $O/${TAG}_scenes.jsonl | cut -c1-400; tail -5 $O/${TAG}_scenes.err ;;
    launches)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"; python scripts/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.md 2>&1; head -60 $O/${TAG}_launches.md; gzip -f $O/${TAG}_launches.csv ;;
    *) echo "unknown step $step" ;;
  esac
done
