#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --cpu-sample-steps 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_final.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_factor_dep -s 4 -c 2 -o gpurun_out/prof_factor_dep -f python bench.py --steps 2 --warmup 3 --cpu-sample-steps 1 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_warp2 -s 10 -c 2 -o gpurun_out/prof_fwd -f python bench.py --steps 2 --warmup 3 --cpu-sample-steps 1 > gpurun_out/ncu_full2.log 2>&1
echo "full2 rc=$?"; ls -la gpurun_out/*.ncu-rep
