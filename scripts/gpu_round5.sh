#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_graph.py > gpurun_out/diag_graph.log 2>&1; echo "diag rc=$?"
cat gpurun_out/diag_graph.log | tail -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_fwd_tc|conv_wgrad|slice_|splat_|neighbour_table|gather_|filter_prep' -o gpurun_out/r01b_ops_n1M_v64 -f python scripts/ncu_ops.py > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
tail -3 gpurun_out/ncu_ops.log
ls -la gpurun_out/*.ncu-rep
