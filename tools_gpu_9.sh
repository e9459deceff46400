#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 48 --warmup 6 --cpu-sample-steps 1 > gpurun_out/bench_m.json 2> gpurun_out/bench1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print({k:round(d[k],4) for k in ('value','ms_per_step','ms_per_factorize')}, round(d['e2e']['value'],1), d['clocks'])"
done
