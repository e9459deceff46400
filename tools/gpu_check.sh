#!/bin/bash
# one GPU session: parity tests, headline bench, per-config measurements
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^E  " | tail -5
timeout 300 python bench.py --steps 48 --warmup 6 --cpu-sample-steps 2 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_latest.json')); print({k:round(d[k],4) for k in ('value','ms_per_step','ms_per_factorize','ms_per_assemble','ms_per_solve')}, round(d['e2e']['value'],1), d['counters'])"
tail -2 gpurun_out/bench_latest.err | cut -c1-300
timeout 500 python tools/bench_configs.py ${CONFIGS:-c2 c2eq c5s} 2>/dev/null | tee -a gpurun_out/configs.jsonl | cut -c1-900
