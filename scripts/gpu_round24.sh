#!/bin/bash
# r01p: tiled slice_classify forward + vectorised backward -- parity (full suite, -x), op sweep incl. slice_classify, scenes
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider --durations=3 > $O/r01p_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01p_pytest_gpu.txt; grep -E "^E  |Error" $O/r01p_pytest_gpu.txt | head -5
timeout 400 python bench_ops.py --quick --n 1000000 --vals 32 64 128 > $O/r01p_ops_sweep.jsonl 2> $O/ops.err; echo "ops rc=$?"
python scripts/show_ops.py $O/r01p_ops_sweep.jsonl | grep -v "SIMT\|conv" | cut -c20-170; tail -2 $O/ops.err
timeout 300 python bench_scenes.py --scene both --impl ours > $O/r01p_scenes_ours.jsonl 2> $O/scenes_ours.err; echo "scenes ours rc=$?"
cut -c1-330 $O/r01p_scenes_ours.jsonl; tail -3 $O/scenes_ours.err
