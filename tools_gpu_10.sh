#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E  " | tail -6
for g in 0 1; do
if [ $g = 0 ]; then export B2_NO_STEP_GRAPH=1; else unset B2_NO_STEP_GRAPH; fi
timeout 300 python bench.py --steps 48 --warmup 6 --cpu-sample-steps 1 > gpurun_out/bench_g$g.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_g$g.json')); print('stepgraph',$g,{k:round(d[k],4) for k in ('value','ms_per_step','ms_per_factorize','ms_per_assemble','ms_per_solve')}, round(d['e2e']['value'],1), d['counters'])"
tail -3 gpurun_out/bench1.err | cut -c1-300
done
