#!/bin/bash
# r01h evidence pass at HEAD: parity (twice, flakiness check), smoke, both bench arms, ncu launch list of the
# graph step, ncu --set full of the bench conv call (roofline.traffic) and of every op kernel at sweep size, op sweep
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=8 > $O/r01h_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"
tail -3 $O/r01h_pytest_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > $O/r01h_pytest_gpu_2.txt 2>&1; echo "pytest gpu (2nd) rc=$?"
tail -2 $O/r01h_pytest_gpu_2.txt
timeout 300 python __graft_entry__.py --smoke > $O/r01h_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/r01h_smoke.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.log 2>&1; echo "bench ref rc=$?"
grep '^{' $O/bench_ref.log | tail -1 > $O/r01h_bench_reference.json; cut -c1-260 $O/r01h_bench_reference.json
# traffic first, so that the bench line carries it
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|filter_prep' -c 12 -o $O/r01h_bench_conv -f python scripts/ncu_bench_conv.py > $O/ncu_bench_conv.log 2>&1; echo "ncu bench conv rc=$?"
ncu -i $O/r01h_bench_conv.ncu-rep --page raw --csv > $O/r01h_bench_conv_raw.csv 2>/dev/null
python scripts/roofline_traffic.py $O/r01h_bench_conv_raw.csv profiles/roofline_traffic.json && cp profiles/roofline_traffic.json $O/roofline_traffic.json
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_graph.log 2>&1; echo "bench graph rc=$?"
grep '^{' $O/bench_graph.log | tail -1 > $O/r01h_bench_graph.json; cut -c1-400 $O/r01h_bench_graph.json
tail -3 $O/bench_graph.log | grep -v '^{' | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r01h_launches_graph_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/summarize_launches.py $O/r01h_launches_graph_step.csv > $O/r01h_launches_graph_step.md 2>$O/summarize.err; head -12 $O/r01h_launches_graph_step.md
gzip -f $O/r01h_launches_graph_step.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ln::' -c 40 -o $O/r01h_ops -f python scripts/ncu_ops.py > $O/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -2 $O/ncu_ops.log
ncu -i $O/r01h_ops.ncu-rep --page raw --csv > $O/r01h_ops_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py $O/r01h_ops_raw.csv > $O/r01h_ops_ncu.md 2>>$O/summarize.err
ls -la $O/*.ncu-rep
# keep gpurun_out under the 64 MiB merge limit
for f in $O/*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 25000000 ]; then echo "dropping $f ($s bytes), raw csv kept"; rm -f $f; fi; done
timeout 600 python bench_ops.py --quick --n 1000000 --vals 8 32 64 128 > $O/r01h_ops_sweep.jsonl 2> $O/ops_a.err; echo "ops rc=$?"
python scripts/show_ops.py $O/r01h_ops_sweep.jsonl | grep -v "SIMT" | cut -c20-200
timeout 300 python bench_ops.py --quick --n 1000000 --vals 32 --order morton > $O/r01h_ops_sweep_morton.jsonl 2> $O/ops_m.err; echo "ops morton rc=$?"
python scripts/show_ops.py $O/r01h_ops_sweep_morton.jsonl | grep -v "SIMT\|reference\|conv" | cut -c20-200
du -sh $O
