#!/bin/bash
# r01k: row-tiled GroupNorm for scene-sized lattices -- parity, scene timings, KITTI launch list, headline bench
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "group_norm" > $O/r01k_pytest_gn.txt 2>&1; echo "pytest gn rc=$?"
tail -12 $O/r01k_pytest_gn.txt | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=5 > $O/r01k_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01k_pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > $O/r01k_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/r01k_smoke.txt
timeout 600 python bench_scenes.py --scene both --impl ours > $O/r01k_scenes_ours.jsonl 2> $O/scenes_ours.err; echo "scenes ours rc=$?"
cut -c1-420 $O/r01k_scenes_ours.jsonl; tail -3 $O/scenes_ours.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/r01k_launches_kitti_pass.csv python bench_scenes.py --scene kitti --steps 1 --warmup 0 > $O/ncu_kitti.log 2>&1; echo "ncu kitti rc=$?"
python scripts/summarize_launches.py $O/r01k_launches_kitti_pass.csv > $O/r01k_launches_kitti_pass.md 2>$O/summarize.err; sed -n 5,24p $O/r01k_launches_kitti_pass.md | cut -c1-140
gzip -f $O/r01k_launches_kitti_pass.csv
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_graph.log 2>&1; echo "bench graph rc=$?"
grep '^{' $O/bench_graph.log | tail -1 > $O/r01k_bench_graph.json; cut -c1-200 $O/r01k_bench_graph.json
du -sh $O
