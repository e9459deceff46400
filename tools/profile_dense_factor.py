#!/usr/bin/env python
"""One dense factorisation (configs[1]: N = 4096) inside a cudaProfilerStart/Stop window, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_dense_factor.py
and a per-kernel breakdown of the blocked LDL^T chain (tools/summarise_launches.py out.csv)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import madnlp_jl_b200 as pkg
from madnlp_jl_b200.linear_solvers import B200DenseSolver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
G = rng.standard_normal((N, N + 64))
A = torch.from_numpy(G @ G.T / N + np.eye(N)).cuda()
M = B200DenseSolver(A)
M.factorize(); torch.cuda.synchronize()
torch.cuda.profiler.start()
M.factorize(); torch.cuda.synchronize()
x = torch.ones(N, dtype=torch.float64, device="cuda")
M.solve_linear_system(x); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("inertia", M.inertia())
