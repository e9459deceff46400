#!/bin/bash
# round-2 evidence run (1 GPU): segment timeline, launch lists (serialised, cold-cache: compare SHARES), full ncu captures of the
# dominant kernels.  Outputs go to gpurun_out/; the summaries that are judged are copied into profiles/ by tools/summarise_profiles.py
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/tools
python tools/step_timeline.py > gpurun_out/r02_step_timeline.txt 2>&1
python tools/profile_front.py > gpurun_out/r02_profile_front.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 1400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 30 --warmup 6 --no-secondary --cpu-sample-steps 0 > gpurun_out/r02_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_factor_dep -s 12 -c 1 -o gpurun_out/r02_prof_factor_dep \
    python bench.py --steps 10 --warmup 3 --no-secondary --cpu-sample-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ozaki_syrk -s 2 -c 1 -o gpurun_out/r02_prof_ozaki \
    tools/microbench/ozaki_syrk_tcgen05 2048 4096 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense_solve_flow -s 3 -c 1 -o gpurun_out/r02_prof_dense_solve \
    python -c "import bench_configs as BC; BC.config2(cpu=False, lib=False)" > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02_launches_c2.csv \
    python -c "import bench_configs as BC; BC.config2(cpu=False, lib=False)" > /dev/null 2>&1
python tools/bench_configs.py c5md > gpurun_out/r02_c5md.json 2>&1
ls -la gpurun_out | tail -12
