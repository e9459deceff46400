"""N>1 path on CPU: two and four `gloo` ranks run the subtree-sharded factor/solve PROTOCOL of madnlp.jl_b200/parallel.py
(local phase -> all-reduce of the exchange region -> replicated top tree -> ...) with the kernels' arithmetic replayed in
numpy over the symbolic structure and buffer layout exported by the C ABI.  What this pins without a GPU: every rank
derives the same partition, the exchange regions line up across ranks, each exchanged block/vector has exactly one
contributor, the inertia reduction and the masked all-reduce of x are complete."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import madnlp_oracle as o
    import madnlp_jl_b200 as pkg
    from mf_emulator import Symbolic, PhasedReplay
    import ctypes as C
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        W = pkg.workloads
        model, st = W.acopf_case("case300_synth")
        cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
        k = o.SparseCondensedKKTSystem(cb)
        it = W.ipm_iterates(model, st, 1, seed=21, y_scale=1e3, eq_box=(1e-1, 1.0))[0]     # indefinite: inertia has negatives
        k.initialize(); k.jac[:] = it.jac; k.hess[:] = it.hess; k.compress_jacobian(); k.compress_hessian()
        k.reg[:] = 1e-8; k.du_diag[:] = 0.0; k.l_diag[:] = it.l_diag; k.u_diag[:] = it.u_diag
        k.l_lower[:] = it.l_lower; k.u_lower[:] = it.u_lower
        o.set_aug_diagonal_(k); k.build_kkt()
        n = k.n
        S = Symbolic(n, k.aug_colptr, k.aug_rowval, n_parts=world, part_rank=rank)
        # every rank must hold the identical partition / layout
        sig = torch.tensor([int(S.exch_cb), int(S.exch_cbv), int(S.ns), int((S.owner == -1).sum()), int(S.perm[:50].sum())])
        lo = sig.clone(); hi = sig.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert (lo == hi).all()
        assert S.exch_cb > 0 and S.exch_cbv > 0
        R = PhasedReplay(S, rank)
        R.factor_phase(k.aug_nz, 0)
        exch = torch.from_numpy(R.ws[:S.exch_cb])
        nz_contrib = (exch != 0).to(torch.int32)
        dist.all_reduce(nz_contrib)
        assert int(nz_contrib.max()) <= 1                      # exactly one contributor per exchanged entry
        dist.all_reduce(exch)                                  # in place on R.ws (shared memory with numpy)
        R.factor_phase(k.aug_nz, 1)
        loc = torch.tensor([R.neg[0]]); dist.all_reduce(loc)
        neg = int(loc.item()) + R.neg[1]
        mask = np.zeros(n, dtype=np.uint8)
        pkg.capi.check(pkg.capi.lib.b2_owned_mask(S.h, mask.ctypes.data))
        msum = torch.from_numpy(mask.astype(np.int32)); dist.all_reduce(msum)
        assert (msum == 1).all()                               # every row of x is finalised by exactly one rank
        b = np.random.default_rng(3).standard_normal(n)
        x = b[S.perm].copy()
        R.fwd_phase(x, 0)
        xv = torch.from_numpy(R.cbv[:S.exch_cbv]); dist.all_reduce(xv)
        R.fwd_phase(x, 1); R.bwd_phase(x, 1); R.bwd_phase(x, 0)
        out = np.zeros(n); out[S.perm] = x
        out *= mask
        t = torch.from_numpy(out); dist.all_reduce(t)
        Kf = o.tril_to_full(k.aug_colptr, k.aug_rowval, k.aug_nz, n).toarray()
        ev = np.linalg.eigvalsh(Kf)
        res = np.abs(Kf @ out - b).max() / (np.abs(Kf).max() * np.abs(out).max() + np.abs(b).max())
        q.put((rank, neg, int((ev < 0).sum()), float(res)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_protocol_with_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    results = sorted(q.get(timeout=5) for _ in range(world))
    for rank, neg, neg_true, res in results:
        assert neg == neg_true and neg_true > 0
        assert res < 1e-10
