#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E  " | tail -6
timeout 400 python tools_bench_configs.py c5s c5m 2>&1 | tail -3
