#!/bin/bash
# A/B of the dense look-ahead schedule: whole-panel trsm/update on the chain (NEAR=0), near-diagonal kernels on the chain (NEAR=1),
# and the chain kernels launched programmatically dependent (PDL=1)
export PYTHONPATH=$PWD:$PWD/tools:$PWD/oracle
for v in "0 0 0" "1 0 0" "1 1 0" "1 1 1" "1 0 1"; do
  set -- $v
  B2_DENSE_NEAR=$1 B2_DENSE_PDL=$2 B2_DENSE_INV_SIDE=$3 timeout 200 python -c "
import bench_configs as BC, json
for neq in (0,):
    r = BC.config2(n_eq=neq, cpu=False, lib=False)
    print('NEAR=$1 PDL=$2 INVSIDE=$3', json.dumps({k: r[k] for k in ('config','inertia','residual','ms_factorize','factor_tflops','ms_solve')}))
" 2>&1 | tail -3
done
