#!/bin/bash
# parameter sweep of the headline workload (analysis options are read from the environment by bench.py); one line per setting
for cfg in ${SWEEP:-"B2_FUSE_MAX=2" "B2_FUSE_MAX=4" "B2_FUSE_MAX=6" "B2_FUSE_MAX=8" "B2_FUSE_MAX=10" "B2_FUSE_MAX=12" "B2_FUSE_MAX=8 B2_NEMIN=12" "B2_FUSE_MAX=8 B2_NEMIN=20"}; do
  env $cfg python bench.py --steps 48 --warmup 6 --no-secondary --cpu-sample-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$cfg', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'fac',round(d['ms_per_factorize'],4),'sol',round(d['ms_per_solve'],4),'asm',round(d['ms_per_assemble'],4),'lev',d['solver']['levels'],'sn',d['solver']['supernodes'],'nnzl',d['solver']['nnz_l'],'solves/fact',round(d['solver']['refinement_solves_per_factorization'],2),'cnt',d['counters'])"
done
