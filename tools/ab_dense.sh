#!/bin/bash
# A/B of the dense look-ahead schedule knobs (sparse_ldl.cu: enqueue_dense_factor_lookahead); one process per setting
# usage: tools/ab_dense.sh "NEAR PDL INVSIDE RELAX EARLY_RESERVED RESERVED" ...
export PYTHONPATH=$PWD:$PWD/tools:$PWD/oracle
[ $# -eq 0 ] && set -- "0 0 0 0 8 1" "1 0 0 0 8 1"
for v in "$@"; do
  set -- $v
  B2_DENSE_NEAR=$1 B2_DENSE_PDL=$2 B2_DENSE_INV_SIDE=$3 B2_DENSE_RELAX=$4 B2_DENSE_EARLY_RESERVED=$5 B2_DENSE_RESERVED_SMS=$6 timeout 200 python -c "
import bench_configs as BC, json
r = BC.config2(n_eq=0, cpu=False, lib=False)
print('NEAR=$1 PDL=$2 INVSIDE=$3 RELAX=$4 EARLY=$5 RES=$6', json.dumps({k: r[k] for k in ('inertia','residual','ms_factorize','factor_tflops')}))
" 2>&1 | tail -1
done
