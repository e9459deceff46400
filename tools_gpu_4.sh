#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E  " | tail -4
for fm in 48 128 400; do
B2_FUSE_MAX=$fm timeout 600 python bench.py --steps 24 --warmup 4 --cpu-sample-steps 1 > gpurun_out/bench_f$fm.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_f$fm.json')); print('fuse',$fm,{k:d[k] for k in ('value','ms_per_step','ms_per_factorize','ms_per_assemble')}, d['e2e']['value'], d['counters'])"
done
