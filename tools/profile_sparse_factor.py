#!/usr/bin/env python
"""One factorisation + solve of the augmented 3-D grid KKT (configs[4] family) inside a cudaProfilerStart/Stop window:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_sparse_factor.py 40
then tools/summarise_launches.py out.csv."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import madnlp_jl_b200 as pkg
from madnlp_jl_b200 import kkt as K
from madnlp_jl_b200.linear_solvers import B200SparseSolver, DeviceCSC
W = pkg.workloads
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N, n_tot, m, I, J, V = W.augmented_grid_kkt(nx, nx, nx)
cp, rv, mp = K.coo_to_csc(I, J, N, N)
plan = K._transfer_plan(mp, len(rv))
nz = torch.zeros(len(rv), dtype=torch.float64, device="cuda"); Vd = torch.from_numpy(V).cuda()
pkg.capi.check(pkg.capi.lib.b2_transfer(plan.h, nz.data_ptr(), Vd.data_ptr(), None))
M = B200SparseSolver(DeviceCSC(N, N, cp, rv, nz), B200SparseSolver.default_options(kkt_n_primal=n_tot))
M.factorize(); torch.cuda.synchronize()
torch.cuda.profiler.start()
M.factorize(); torch.cuda.synchronize()
x = torch.ones(N, dtype=torch.float64, device="cuda")
M.solve_linear_system(x); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("inertia", M.inertia(), M.stats())
