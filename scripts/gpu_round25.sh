#!/bin/bash
# r01q: headline bench + smoke at HEAD (after the slice_classify / GroupNorm changes)
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python __graft_entry__.py --smoke > $O/r01q_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/r01q_smoke.txt
timeout 400 python bench.py --steps 30 --warmup 5 > $O/bench_graph.log 2>&1; echo "bench graph rc=$?"
grep '^{' $O/bench_graph.log | tail -1 > $O/r01q_bench_graph.json; cut -c1-200 $O/r01q_bench_graph.json
tail -3 $O/bench_graph.log | grep -v '^{' | cut -c1-300
