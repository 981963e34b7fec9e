#!/bin/bash
# r01i: wide (384/512-channel) layers on the tensor cores via N chunking, scene-sized (KITTI / ScanNet) timings of both
# arms, ncu --set full of every op kernel at sweep size (kernel regex fixed), launch list of a KITTI-sized pass
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "wide_layers or tensor_core or transposed" > $O/r01i_pytest_conv.txt 2>&1; echo "pytest conv rc=$?"
tail -4 $O/r01i_pytest_conv.txt
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=5 > $O/r01i_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01i_pytest_gpu.txt
timeout 600 python bench_scenes.py --scene both --impl ours > $O/r01i_scenes_ours.jsonl 2> $O/scenes_ours.err; echo "scenes ours rc=$?"
cut -c1-420 $O/r01i_scenes_ours.jsonl; tail -3 $O/scenes_ours.err
timeout 600 python bench_scenes.py --scene kitti --impl ours --conv-precision 0 > $O/r01i_scenes_ours_fp32simt.jsonl 2>> $O/scenes_ours.err; echo "scenes ours p0 rc=$?"
cut -c1-420 $O/r01i_scenes_ours_fp32simt.jsonl
timeout 900 python bench_scenes.py --scene both --impl reference --steps 3 --warmup 1 > $O/r01i_scenes_reference.jsonl 2> $O/scenes_ref.err; echo "scenes ref rc=$?"
cut -c1-420 $O/r01i_scenes_reference.jsonl; tail -3 $O/scenes_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|conv_wgrad_tc|slice_fwd|scatter_rows|splat_build|neighbour_table|slice_classify|gather_fwd|filter_prep' -c 40 -o $O/r01i_ops -f python scripts/ncu_ops.py > $O/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -2 $O/ncu_ops.log
ncu -i $O/r01i_ops.ncu-rep --page raw --csv > $O/r01i_ops_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py $O/r01i_ops_raw.csv > $O/r01i_ops_ncu.md 2>$O/summarize.err
grep -c '^## ' $O/r01i_ops_ncu.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/r01i_launches_kitti_pass.csv python bench_scenes.py --scene kitti --steps 1 --warmup 0 > $O/ncu_kitti.log 2>&1; echo "ncu kitti rc=$?"
python scripts/summarize_launches.py $O/r01i_launches_kitti_pass.csv > $O/r01i_launches_kitti_pass.md 2>>$O/summarize.err; sed -n 1,22p $O/r01i_launches_kitti_pass.md
gzip -f $O/r01i_launches_kitti_pass.csv
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_graph.log 2>&1; echo "bench graph rc=$?"
grep '^{' $O/bench_graph.log | tail -1 > $O/r01i_bench_graph.json; cut -c1-200 $O/r01i_bench_graph.json
ls -la $O/*.ncu-rep
for f in $O/*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 30000000 ]; then echo "dropping $f ($s bytes), raw csv kept"; rm -f $f; fi; done
du -sh $O
