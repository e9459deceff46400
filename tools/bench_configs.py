#!/usr/bin/env python
"""Per-configuration measurements beyond the headline bench (BASELINE.json configs[1], [2], [4]):
ms/assemble, ms/factorize, ms/solve with CUDA events (median of 20 after 3 warm-ups, L2 flushed between repeats),
achieved GFLOP/s / GB/s against the measured peaks, next to (a) the reference's GPU library path where torch exposes
the same library routine (cuBLAS GEMM, cuSOLVER sytrf/potrf) and (b) the CPU oracle.  Prints one JSON line per config."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import madnlp_oracle as o, madnlp_jl_b200 as pkg
from madnlp_jl_b200 import kkt as K
from madnlp_jl_b200.linear_solvers import B200SparseSolver, DeviceCSC
W = pkg.workloads
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def dgemm_peak():
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: torch.matmul(a, b), reps=5, warm=2)
    return 2 * n ** 3 / (ms * 1e-3) / 1e12


def config2(n=4096, m=2048, n_eq=0):
    qp = W.dense_qp(n=n, m=m, n_eq=n_eq, seed=1)
    it = W.dense_qp_iterate(qp, mu=1e-3, seed=2)
    cb = o.Callback(qp.n, qp.m, [], [], [], [], qp.ind_ineq, qp.ind_lb, qp.ind_ub)
    kg = K.DenseCondensedKKTSystem(cb); kg.initialize(); kg.set_dense(hess_np=qp.P, jac_np=qp.A)
    for name in ("reg", "du_diag", "l_diag", "u_diag", "l_lower", "u_lower"):
        getattr(kg, name).copy_(dev(it[name]))
    kg.set_aug_diagonal_()
    ns = m - n_eq; N = n + n_eq
    t_asm = timeit(kg.build_kkt)
    t_fac = timeit(kg.linear_solver.factorize)
    inertia = kg.linear_solver.inertia()
    x = torch.randn(N, dtype=torch.float64, device="cuda")
    t_sol = timeit(lambda: kg.linear_solver.solve_linear_system(x.clone()))
    # residual property at full size
    b = torch.randn(N, dtype=torch.float64, device="cuda"); xs = kg.linear_solver.solve_linear_system(b.clone())
    A = kg.aug_com.t(); Af = torch.tril(A) + torch.tril(A, -1).t()
    res = float((Af @ xs - b).abs().max() / (Af.abs().max() * xs.abs().max() + b.abs().max()))
    # library bars (what the reference's GPU path calls): cuBLAS for J'DJ, cuSOLVER sytrf / potrf
    J = kg.jac.t()[kg._ind_ineq_d]; D = kg.diag_buffer
    t_cublas = timeit(lambda: (J.t() * D) @ J)
    t_potrf = timeit(lambda: torch.linalg.cholesky(Af)) if n_eq == 0 else None
    t_sytrf = timeit(lambda: torch.linalg.ldl_factor(Af), reps=5)
    # CPU oracle (LAPACK dsytrf through scipy, all threads OpenBLAS gives)
    kc = o.DenseCondensedKKTSystem(cb); kc.initialize(); kc.hess[:] = qp.P; kc.jac[:] = qp.A
    for name in ("reg", "du_diag", "l_diag", "u_diag", "l_lower", "u_lower"):
        getattr(kc, name)[:] = it[name]
    o.set_aug_diagonal_(kc)
    t0 = time.perf_counter(); kc.build_kkt(); c_asm = time.perf_counter() - t0
    t0 = time.perf_counter(); kc.linear_solver.factorize(); c_fac = time.perf_counter() - t0
    assert kc.linear_solver.inertia() == inertia, (kc.linear_solver.inertia(), inertia)
    syrk_flop = n * (n + 1) * ns; fac_flop = N ** 3 / 3
    return dict(config="C2 DenseCondensedKKT n=%d m=%d n_eq=%d fp64" % (n, m, n_eq), inertia=inertia, residual=res,
                ms_assemble=t_asm, assemble_tflops=syrk_flop / t_asm / 1e9, ms_factorize=t_fac, factor_tflops=fac_flop / t_fac / 1e9,
                ms_solve=t_sol, solve_gbs=8.0 * N * N / t_sol / 1e6,
                lib_ms_cublas_gemm=t_cublas, lib_ms_cusolver_potrf=t_potrf, lib_ms_cusolver_sytrf=t_sytrf,
                cpu_ms_assemble=1e3 * c_asm, cpu_ms_factorize=1e3 * c_fac, cpu_threads=os.cpu_count())


def config_sparse_opf(case):
    model, st = W.acopf_case(case)
    it = W.ipm_iterates(model, st, 1, seed=3)[0]
    class CB: pass
    cb = CB(); cb.nvar, cb.ncon = st.nvar, st.ncon
    cb.jac_I, cb.jac_J, cb.hess_I, cb.hess_J = st.jac_I, st.jac_J, st.hess_I, st.hess_J
    cb.ind_ineq, cb.ind_lb, cb.ind_ub = st.ind_ineq, st.ind_lb, st.ind_ub
    kg = K.SparseCondensedKKTSystem(cb); kg.initialize()
    for name in ("reg", "du_diag", "l_diag", "u_diag", "l_lower", "u_lower"):
        getattr(kg, name).copy_(dev(getattr(it, name)))
    kg.get_jacobian().copy_(dev(it.jac)); kg.get_hessian().copy_(dev(it.hess))
    def asm():
        kg.compress_jacobian(); kg.compress_hessian(); kg.set_aug_diagonal_(); kg.build_kkt()
    t_asm = timeit(asm); t_fac = timeit(kg.linear_solver.factorize)
    x = torch.randn(kg.n, dtype=torch.float64, device="cuda")
    t_sol = timeit(lambda: kg.linear_solver.solve_linear_system(x))
    stt = kg.linear_solver.stats(); ps = kg.plan_sizes()
    asm_bytes = 8 * (len(it.jac) + len(it.hess)) * 2 + 8 * (kg.n_tot + kg.m) + 20 * stt["nnz_a"] + 16 * ps["jptr"]
    return dict(config="C3/C4 SparseCondensedKKT %s" % case, n=kg.n, m=kg.m, nnz_kkt=stt["nnz_a"], nnz_l=stt["nnz_l"], flops=stt["flops"],
                levels=stt["n_levels"], max_front=stt["max_front"], inertia=kg.linear_solver.inertia(),
                ms_assemble=t_asm, assemble_gbs=asm_bytes / t_asm / 1e6, ms_factorize=t_fac,
                factor_gbs=8.0 * (stt["nnz_a"] + stt["nnz_l"]) / t_fac / 1e6, factor_gflops=stt["flops"] / t_fac / 1e6,
                ms_solve=t_sol, solve_gbs=24.0 * stt["nnz_l"] / t_sol / 1e6, hbm_peak_gbs=HBM)


def config5(nx):
    N, n_tot, m, I, J, V = W.augmented_grid_kkt(nx, nx, nx)
    t0 = time.perf_counter()
    cp, rv, mp = K.coo_to_csc(I, J, N, N)
    plan = K._transfer_plan(mp, len(rv))
    nz = torch.zeros(len(rv), dtype=torch.float64, device="cuda"); Vd = dev(V)
    pkg.capi.check(pkg.capi.lib.b2_transfer(plan.h, nz.data_ptr(), Vd.data_ptr(), None))
    csc = DeviceCSC(N, N, cp, rv, nz)
    M = B200SparseSolver(csc, B200SparseSolver.default_options(kkt_n_primal=n_tot))
    t_an = time.perf_counter() - t0
    stt = M.stats()
    t_asm = timeit(lambda: pkg.capi.lib.b2_transfer(plan.h, nz.data_ptr(), Vd.data_ptr(), None))
    t_fac = timeit(M.factorize, reps=5, warm=1)
    inertia = M.inertia()
    b = torch.randn(N, dtype=torch.float64, device="cuda")
    t_sol = timeit(lambda: M.solve_linear_system(b.clone()), reps=5, warm=1)
    x = M.solve_linear_system(b.clone())
    import scipy.sparse as sp
    Kf = o.tril_to_full(cp, rv, nz.cpu().numpy(), N)
    xh = x.cpu().numpy(); bh = b.cpu().numpy()
    res = float(np.abs(Kf @ xh - bh).max() / (abs(Kf).max() * np.abs(xh).max() + np.abs(bh).max()))
    return dict(config="C5 SparseKKT augmented 3-D grid %d^3" % nx, N=N, nnz_kkt=stt["nnz_a"], nnz_l=stt["nnz_l"], flops=stt["flops"],
                max_front=stt["max_front"], levels=stt["n_levels"], big_fronts=stt["n_big_fronts"], inertia=inertia,
                expected_inertia=(n_tot, 0, m), residual=res, analysis_s=t_an, ms_assemble=t_asm,
                assemble_gbs=(16.0 * len(V) + 12.0 * len(rv)) / t_asm / 1e6, ms_factorize=t_fac, factor_tflops=stt["flops"] / t_fac / 1e9,
                ms_solve=t_sol, solve_gbs=24.0 * stt["nnz_l"] / t_sol / 1e6,
                factor_bytes=stt["factor_bytes"], workspace_bytes=stt["workspace_bytes"])


if __name__ == "__main__":
    which = sys.argv[1:] or ["peak", "c2", "c2eq", "c3", "c4", "c5s"]
    for wch in which:
        if wch == "peak": r = dict(config="cuBLAS DGEMM 8192^3 (fp64 roofline denominator)", tflops=dgemm_peak())
        elif wch == "c2": r = config2()
        elif wch == "c2eq": r = config2(n_eq=256)
        elif wch == "c3": r = config_sparse_opf("case1354_pegase")
        elif wch == "c4": r = config_sparse_opf("case10000_goc")
        elif wch == "c5s": r = config5(40)
        elif wch == "c5m": r = config5(64)
        elif wch == "c5": r = config5(89)
        print(json.dumps(r), flush=True)
