#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: IPM iterations/sec (+ ms/factorize) on the OPF-10k condensed KKT (fp64).

A "step" = one pass of the hot path over one synthetic interior-point iterate, in the call order of MadNLP's
`regular!` (src/IPM/solver.jl:216-298): compress_jacobian!/compress_hessian! -> set_aug_diagonal! -> build_kkt! ->
factorize! -> inertia -> [regularise + refactor while the inertia is wrong] -> Richardson(solve_kkt! + KKT mat-vec).
Workload: synthetic AC-OPF with the (nbus, nbranch, ngen) counts of pglib case10000_goc (no pglib data offline),
SparseCondensedKKTSystem, 24 distinct iterates (mu: 1e-1 -> 1e-9) cycled.

  value : steps/sec with the iterate's inputs already resident in HBM (device-to-device staging only)
  e2e   : same metric through the host-facing path: inputs in pinned HOST memory, H2D of (jac, hess, reg, du_diag,
          l_diag, u_diag, l_lower, u_lower, rhs) and D2H of the step direction d INSIDE the timed region, every step
  --impl reference : the CPU oracle (numpy assembly + SuperLU standing in for UMFPACK, 1 core) on the same workload

Timing: every step is bracketed by CUDA events on the launching stream; between steps (untimed) L2 is flushed by
writing a 256 MiB buffer; the K steps are bracketed by barrier + synchronize; multi-GPU = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "ipm_iters_per_sec"
UNIT = "iter/s"
N_ITERATES = 24
FIELDS = ("jac", "hess", "reg", "du_diag", "l_diag", "u_diag", "l_lower", "u_lower", "rhs")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="case10000_goc")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--cpu-sample-steps", type=int, default=8)
    return ap.parse_args()


def make_workload(name):
    import madnlp_jl_b200 as pkg
    W = pkg.workloads
    model, st = W.acopf_case(name)
    its = W.ipm_iterates(model, st, N_ITERATES, seed=0)
    return model, st, its


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)"""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

        def run():
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.02)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's CPU path restated (oracle): numpy assembly, SuperLU (UMFPACK stand-in) factor/solve, Richardson.
    Rank 0 only; bounded sample per the --steps/--warmup given."""
    if rank != 0:
        return
    import madnlp_oracle as o
    model, st, its = make_workload(args.workload)
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    kkt = o.SparseCondensedKKTSystem(cb, o.UmfpackStandInSolver)
    kkt.initialize()
    steps = min(args.steps, max(1, args.cpu_sample_steps))
    warm = min(args.warmup, 1)

    def step(it):
        kkt.get_jacobian()[:] = it.jac; kkt.get_hessian()[:] = it.hess
        kkt.reg[:] = it.reg; kkt.du_diag[:] = it.du_diag
        kkt.l_diag[:] = it.l_diag; kkt.u_diag[:] = it.u_diag; kkt.l_lower[:] = it.l_lower; kkt.u_lower[:] = it.u_lower
        kkt.compress_jacobian(); kkt.compress_hessian()
        o.set_aug_diagonal_(kkt)
        kkt.build_kkt()
        t0 = time.perf_counter()
        kkt.linear_solver.factorize()
        tf = time.perf_counter() - t0
        b = o.UnreducedKKTVector.for_kkt(kkt); b.full()[:] = it.rhs
        x = o.UnreducedKKTVector.for_kkt(kkt); w = o.UnreducedKKTVector.for_kkt(kkt)
        o.solve_refine(x, kkt, b, w)
        return tf
    for i in range(warm):
        step(its[i % len(its)])
    t0 = time.perf_counter()
    tfs = [step(its[(warm + i) % len(its)]) for i in range(steps)]
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * dt / steps, "ms_per_factorize": 1e3 * float(np.mean(tfs)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"acopf_{args.workload}_synthetic_condensed_kkt", "n": st.nvar, "m": st.ncon, "iterates": N_ITERATES},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{steps} IPM steps of the same workload: numpy assembly + SuperLU (UMFPACK stand-in) + Richardson"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import madnlp_jl_b200 as pkg
    from madnlp_jl_b200 import kkt as K
    from madnlp_jl_b200.ipm import IPMLinearAlgebra

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, st, its = make_workload(args.workload)

    class CB:
        pass
    cb = CB()
    cb.nvar, cb.ncon = st.nvar, st.ncon
    cb.jac_I, cb.jac_J, cb.hess_I, cb.hess_J = st.jac_I, st.jac_J, st.hess_I, st.hess_J
    cb.ind_ineq, cb.ind_lb, cb.ind_ub = st.ind_ineq, st.ind_lb, st.ind_ub

    if world > 1:
        from madnlp_jl_b200.parallel import DistributedSparseSolver
        solver_cls = lambda csc, opt: DistributedSparseSolver(csc, opt, rank=rank, world=world)   # noqa: E731
    else:
        solver_cls = None
    opt = pkg.capi.default_options()
    if os.environ.get("B2_FUSE_MAX"):
        opt.fuse_max_fronts = int(os.environ["B2_FUSE_MAX"])
    if os.environ.get("B2_DEP"):
        opt.dep_schedule = int(os.environ["B2_DEP"])
    if os.environ.get("B2_NEMIN"):
        opt.nemin = int(os.environ["B2_NEMIN"])
    kkt = K.create_kkt_system(K.SparseCondensedKKTSystem, cb, solver_cls, opt)
    kkt.initialize()
    la = IPMLinearAlgebra(kkt, use_cuda_graph=(world == 1 and not os.environ.get("B2_NO_STEP_GRAPH")))
    stats = kkt.linear_solver.stats()

    host = [{k: torch.from_numpy(np.ascontiguousarray(getattr(it, k))).pin_memory() for k in FIELDS} for it in its]
    devit = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d_bytes = sum(v.numel() * 8 for v in host[0].values())
    d_host = torch.zeros(la.d.values.numel(), dtype=torch.float64).pin_memory()
    d2h_bytes = d_host.numel() * 8
    flush_buf = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def time_phase(fn, reps=10):
        """median CUDA-event time of one call, L2 flushed before each repeat"""
        ts = []
        for _ in range(reps):
            if not args.no_flush:
                flush_buf.fill_(1.0)
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(stream); fn(); a1.record(stream); a1.synchronize()
            ts.append(a0.elapsed_time(a1))
        return float(np.median(ts))

    def one_step(i, e2e, record):
        it = (host if e2e else devit)[i % N_ITERATES]
        if not args.no_flush:
            flush_buf.fill_(1.0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        la.load_iterate(it)
        ok = la.step(mu=its[i % N_ITERATES].mu)
        if e2e:
            d_host.copy_(la.d.values, non_blocking=True)
        e1.record(stream)
        e1.synchronize()
        assert ok
        return e0.elapsed_time(e1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(e2e, record=False):
        for i in range(args.warmup):
            one_step(i, e2e, False)
        barrier()
        t0 = time.perf_counter()
        ms = [one_step(args.warmup + i, e2e, record) for i in range(args.steps)]
        barrier()
        wall = time.perf_counter() - t0
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), wall

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    dev_ms, dev_wall = timed_run(False, record=True)
    e2e_ms, e2e_wall = timed_run(True)
    # phase timings of the hot path's three metrics, measured separately from the step loop
    def assemble():
        kkt.compress_jacobian(); kkt.compress_hessian(); kkt.set_aug_diagonal_(); kkt.build_kkt()
    xsol = torch.randn(kkt.n, dtype=torch.float64, device=dev)
    asm_ms = time_phase(assemble)
    fac_ms = time_phase(kkt.linear_solver.factorize)
    sol_ms = time_phase(lambda: kkt.linear_solver.solve_linear_system(xsol))
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # algorithmic bytes of one numeric factorisation (SURVEY.md 8d, A9 sparse): 8*(nnz K + nnz L)
        alg_bytes = 8.0 * (stats["nnz_a"] + stats["nnz_l"])
        achieved = alg_bytes / (fac_ms * 1e-3) / 1e9 if fac_ms else None
        value = args.steps / (dev_ms * 1e-3)
        e2e_val = args.steps / (e2e_ms * 1e-3)
        n_launch = stats["n_factor_launches"] + 2 + 3      # factor graph kernels + diag/condensed assembly + transfers/diag
        solves = la.cnt["backsolves"] / max(1, la.cnt["factorizations"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"acopf_{args.workload}_synthetic_condensed_kkt", "kkt": "SparseCondensedKKTSystem",
                       "n": st.nvar, "m": st.ncon, "nnz_kkt": stats["nnz_a"], "nnz_l": stats["nnz_l"], "factor_flops": stats["flops"],
                       "supernodes": stats["n_supernodes"], "levels": stats["n_levels"], "max_front": stats["max_front"],
                       "iterates": N_ITERATES, "l2": "flushed between steps (256 MiB write, untimed)" if not args.no_flush else "not flushed",
                       "parallelism": f"subtree-sharded x{world}" if world > 1 else "single GPU",
                       "refinement_solves_per_factorization": solves},
            "ms_per_factorize": fac_ms, "ms_per_assemble": asm_ms, "ms_per_solve": sol_ms,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps},
            # own kernels per step: iterate load (1, device-resident run) + assembly (5) + numeric factorisation + start of the
            # refinement (||b||, w = b: 2) + per refinement step: the triangular sweeps + pre1/pre2/post1/finish/update/mul (6)
            "gpu_launches": int((1 + 5 + stats["n_factor_launches"] + 2 + solves * (stats["n_solve_launches"] + 6)) * args.steps),
            "roofline": {"kernel": "k_factor_dep: numeric multifrontal LDL^T of the whole elimination tree in one launch", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                         "traffic": 11.39e6 if (args.workload == "case10000_goc" and world == 1) else None,
                         "traffic_source": "profiles/r01_prof_factor_dep_summary.txt (ncu --set full: dram read 11.32 MB + write 0.07 MB per launch)",
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                         "note": "latency-bound: %d fronts of order <= %d in %d levels, %.3g Mflop" % (
                             stats["n_supernodes"], stats["max_front"], stats["n_levels"], stats["flops"] / 1e6)},
            "clocks": clocks,
            "wall_s": {"device_resident": dev_wall, "e2e": e2e_wall},
            "counters": la.cnt,
        }
        line["cpu_baseline"] = cpu_baseline_sample(args, st, its)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_sample(args, st, its):
    """oracle ('port') timed on this box's host cores on a bounded sample of the same workload"""
    import madnlp_oracle as o
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    kkt = o.SparseCondensedKKTSystem(cb, o.UmfpackStandInSolver)
    kkt.initialize()
    nstep = max(1, min(args.cpu_sample_steps, 6))
    t0 = time.perf_counter()
    for i in range(nstep):
        it = its[i % len(its)]
        kkt.get_jacobian()[:] = it.jac; kkt.get_hessian()[:] = it.hess
        kkt.reg[:] = it.reg; kkt.du_diag[:] = it.du_diag
        kkt.l_diag[:] = it.l_diag; kkt.u_diag[:] = it.u_diag; kkt.l_lower[:] = it.l_lower; kkt.u_lower[:] = it.u_lower
        kkt.compress_jacobian(); kkt.compress_hessian(); o.set_aug_diagonal_(kkt); kkt.build_kkt()
        kkt.linear_solver.factorize()
        b = o.UnreducedKKTVector.for_kkt(kkt); b.full()[:] = it.rhs
        x = o.UnreducedKKTVector.for_kkt(kkt); w = o.UnreducedKKTVector.for_kkt(kkt)
        o.solve_refine(x, kkt, b, w)
    dt = time.perf_counter() - t0
    return {"value": nstep / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{nstep} IPM steps of the same workload through the numpy/SuperLU oracle"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
