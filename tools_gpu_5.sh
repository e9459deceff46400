#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 16" "24 16" "32 16" "48 8" "48 32" "32 32" "64 16"; do
set -- $cfg
B2_FUSE_MAX=$1 B2_NEMIN=$2 timeout 600 python bench.py --steps 24 --warmup 4 --cpu-sample-steps 1 > gpurun_out/bench_t.json 2> gpurun_out/bench1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_t.json')); print('fuse',$1,'nemin',$2,{k:round(d[k],4) for k in ('value','ms_per_step','ms_per_factorize')}, round(d['e2e']['value'],1), d['config']['levels'], d['config']['supernodes'], d['config']['max_front'])"
done
