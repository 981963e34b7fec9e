#!/bin/bash
# r01m: final check at HEAD -- parity, smoke, GroupNorm op numbers, scene timings, headline bench (both arms)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider --durations=5 > $O/r01m_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01m_pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > $O/r01m_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/r01m_smoke.txt
timeout 300 python bench_ops.py --gn-only 462744 --vals 32 64 128 > $O/r01m_ops_group_norm.jsonl 2> $O/ops_gn.err; echo "gn ops rc=$?"
python scripts/show_ops.py $O/r01m_ops_group_norm.jsonl | cut -c20-200; tail -2 $O/ops_gn.err
timeout 600 python bench_scenes.py --scene both --impl ours > $O/r01m_scenes_ours.jsonl 2> $O/scenes_ours.err; echo "scenes ours rc=$?"
cut -c1-330 $O/r01m_scenes_ours.jsonl; tail -3 $O/scenes_ours.err
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_graph.log 2>&1; echo "bench graph rc=$?"
grep '^{' $O/bench_graph.log | tail -1 > $O/r01m_bench_graph.json; cut -c1-200 $O/r01m_bench_graph.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.log 2>&1; echo "bench ref rc=$?"
grep '^{' $O/bench_ref.log | tail -1 > $O/r01m_bench_reference.json; cut -c1-200 $O/r01m_bench_reference.json
