#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E  " | tail -8
python tools_profile_front.py 2>&1 | tail -22
timeout 900 python bench.py --steps 24 --warmup 4 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench1.json')); print({k:d[k] for k in ('value','ms_per_step','ms_per_factorize','ms_per_assemble')}, d['e2e']['value'], d['counters'])"
tail -3 gpurun_out/bench1.err
