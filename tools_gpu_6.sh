#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E  " | tail -6
for dep in 1 0; do
B2_DEP=$dep timeout 300 python bench.py --steps 24 --warmup 4 --cpu-sample-steps 1 > gpurun_out/bench_dep$dep.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_dep$dep.json')); print('dep',$dep,{k:round(d[k],4) for k in ('value','ms_per_step','ms_per_factorize')}, round(d['e2e']['value'],1), d['counters'])"
tail -3 gpurun_out/bench1.err | cut -c1-300
done
timeout 300 python tools_bench_configs.py c2 c4 c5s 2>&1 | tail -4
