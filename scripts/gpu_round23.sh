#!/bin/bash
# r01o: GroupNorm backward with pre-reduced partials -- parity, op numbers, full suite, scenes, headline bench
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider --durations=3 > $O/r01o_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01o_pytest_gpu.txt
timeout 300 python bench_ops.py --gn-only 462744 --vals 32 64 128 > $O/r01o_ops_group_norm.jsonl 2> $O/ops_gn.err; echo "gn ops rc=$?"
python scripts/show_ops.py $O/r01o_ops_group_norm.jsonl | cut -c20-170; tail -2 $O/ops_gn.err
timeout 600 python bench_scenes.py --scene both --impl ours > $O/r01o_scenes_ours.jsonl 2> $O/scenes_ours.err; echo "scenes ours rc=$?"
cut -c1-330 $O/r01o_scenes_ours.jsonl; tail -3 $O/scenes_ours.err
timeout 300 python __graft_entry__.py --smoke > $O/r01o_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/r01o_smoke.txt
