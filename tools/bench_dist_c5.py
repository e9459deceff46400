#!/usr/bin/env python
"""configs[4] sweep over GPUs: factorize + solve of the augmented 3-D grid KKT with the subtree-sharded solver.
  1 GPU : python tools/bench_dist_c5.py 64
  N GPUs: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/bench_dist_c5.py 64
Timing: CUDA events per rank around each call (barrier before), MAX over ranks, median of the repeats.  Rank 0 prints
one JSON line (inertia and residual are checked against the full matrix on every run)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import torch.distributed as dist
import madnlp_oracle as o, madnlp_jl_b200 as pkg
from madnlp_jl_b200 import kkt as K
from madnlp_jl_b200.linear_solvers import B200SparseSolver, DeviceCSC
from madnlp_jl_b200.parallel import DistributedSparseSolver

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = pkg.workloads
N, n_tot, m, I, J, V = W.augmented_grid_kkt(nx, nx, nx)
cp, rv, mp = K.coo_to_csc(I, J, N, N)
plan = K._transfer_plan(mp, len(rv))
nz = torch.zeros(len(rv), dtype=torch.float64, device="cuda"); Vd = torch.from_numpy(V).cuda()
pkg.capi.check(pkg.capi.lib.b2_transfer(plan.h, nz.data_ptr(), Vd.data_ptr(), None))
csc = DeviceCSC(N, N, cp, rv, nz)
t0 = time.perf_counter()
if world > 1:
    M = DistributedSparseSolver(csc, DistributedSparseSolver.default_options(kkt_n_primal=n_tot), rank=rank, world=world)
else:
    M = B200SparseSolver(csc, B200SparseSolver.default_options(kkt_n_primal=n_tot))
t_an = time.perf_counter() - t0


def timed(fn):
    ts = []
    for r in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if r:
            ts.append(float(t.item()))
    return float(np.median(ts))


t_fac = timed(M.factorize)
inertia = M.inertia()
b = torch.from_numpy(np.random.default_rng(5).standard_normal(N)).cuda()
t_sol = timed(lambda: M.solve_linear_system(b.clone()))
x = M.solve_linear_system(b.clone())
if rank == 0:
    Kf = o.tril_to_full(cp, rv, nz.cpu().numpy(), N)
    xh = x.cpu().numpy(); bh = b.cpu().numpy()
    res = float(np.abs(Kf @ xh - bh).max() / (abs(Kf).max() * np.abs(xh).max() + np.abs(bh).max()))
    st = M.stats()
    print(json.dumps(dict(config="C5 augmented 3-D grid %d^3, subtree-sharded LDL^T" % nx, n_gpus=world, N=N, nnz_l=st["nnz_l"], flops=st["flops"],
                          inertia=list(inertia), expected_inertia=[n_tot, 0, m], residual=res, analysis_s=t_an, ms_factorize=t_fac,
                          factor_tflops=st["flops"] / t_fac / 1e9, ms_solve=t_sol)))
if world > 1:
    dist.destroy_process_group()
