"""Pins the CPU oracle (oracle/madnlp_oracle.py) against everything the reference's own tests hold for this path
 before it is trusted as the checker for the CUDA path."""
import numpy as np
import pytest

import madnlp_oracle as o


def test_2x2_known_answer_vector():
    """lib/MadNLPTests/src/MadNLPTests.jl:24-51 (test_linear_solver): the reference's only explicit KAT."""
    row, col, val = np.array([0, 1, 1]), np.array([0, 0, 1]), np.array([1.0, 0.1, 2.0])
    b = np.array([1.0, 3.0])
    sol = np.array([0.8542713567839195, 1.4572864321608041])
    A = np.zeros((2, 2), order="F")
    A[row, col] = val
    M = o.LapackCPUSolver(A).factorize()
    assert M.inertia() == (2, 0, 0)
    x = M.solve(b.copy())
    assert np.abs(x - sol).max() < 1e-4 or np.abs(x - sol).max() / np.abs(sol).max() < 1e-4   # solcmp, MadNLPTests.jl:18-22
    assert np.abs(x - sol).max() < 1e-15
    # same matrix through the sparse stand-in (matrix_test.jl:33-39 runs the KAT for Umfpack too)
    cp, rv, mp = o.coo_to_csc(row, col, 2, 2)
    nz = np.zeros(len(rv)); o.transfer(nz, val, mp)
    U = o.UmfpackStandInSolver(cp, rv, nz, 2).factorize()
    assert np.abs(U.solve(b.copy()) - sol).max() < 1e-14


HS15_EXPECTED = np.array([0.24987493746873435, 0.00497512437810945, -1.0, -0.7501250625312657,
                          -0.9989999999999999, -0.7493749374687343, -1.001, -1.0007501250625312,
                          0.9997501250625312])


@pytest.mark.parametrize("kind", ["sparse", "sparse_umfpack", "condensed", "dense_condensed"])
def test_hs15_kkt_identity(kind):
    """test/kkt_test.jl:27-48 -> test_kkt_system (MadNLPTests.jl:53-110): K * solve_kkt(K, 1) == 1 and inertia."""
    cb = o.HS15Model.callback()
    if kind == "sparse":
        kkt = o.SparseKKTSystem(cb, o.DenseLDLInertiaSolver)
    elif kind == "sparse_umfpack":
        kkt = o.SparseKKTSystem(cb, o.UmfpackStandInSolver)
    elif kind == "condensed":
        kkt = o.SparseCondensedKKTSystem(cb)
    else:
        kkt = o.DenseCondensedKKTSystem(cb)
    x, y, inertia = o.test_kkt_system(kkt, o.HS15Model, dense=(kind == "dense_condensed"))
    assert np.allclose(y.full(), 1.0, rtol=np.sqrt(np.finfo(float).eps), atol=0)
    assert np.abs(x.full() - HS15_EXPECTED).max() < 1e-14
    if inertia is not None:
        assert kkt.is_inertia_correct(*inertia)
    if kind == "sparse":
        assert inertia == (4, 0, 2)
        assert len(kkt.aug_I) == 15 and len(kkt.aug_rowval) == 13       # 2 duplicate COO positions (Appendix A)
        dense = o.tril_to_full(kkt.aug_colptr, kkt.aug_rowval, kkt.aug_nz, 6).toarray()
        assert np.allclose(np.diag(dense), [2.999, 201.0, 0.999, 0.999, 0.0, 0.0])
        assert dense[4, 2] == -1.0 and dense[5, 0] == 1.0 and dense[5, 3] == -1.0
    if kind == "condensed":
        assert inertia == (2, 0, 0)
        dense = o.tril_to_full(kkt.aug_colptr, kkt.aug_rowval, kkt.aug_nz, 2).toarray()
        assert np.allclose(dense, [[3.998, 0.0], [0.0, 201.0]])


def test_bunch_kaufman_inertia_with_2x2_blocks():
    """num_neg_ev (src/LinearSolvers/lapack.jl:247-268) against eigenvalue signs on matrices that force 2x2 pivots."""
    rng = np.random.default_rng(0)
    for n in (4, 9, 30):
        A = rng.standard_normal((n, n)); A = A + A.T
        A[np.diag_indices(n)] = 0.0                     # zero diagonal => 2x2 pivots
        M = o.LapackCPUSolver(np.asfortranarray(A)).factorize()
        ev = np.linalg.eigvalsh(A)
        assert (M.ipiv < 0).any()
        assert M.inertia() == (int((ev > 0).sum()), 0, int((ev < 0).sum()))


def test_condensed_matches_full_system():
    """Every formulation eliminates the same unreduced system (SURVEY.md fact 2 / test/madnlp_dense.jl:44-48)."""
    import madnlp_jl_b200 as pkg
    W = pkg.workloads
    model, st = W.acopf_case("case30_synth")
    it = W.ipm_iterates(model, st, 1, seed=3)[0]
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    sols = []
    for K in (o.SparseKKTSystem, o.SparseCondensedKKTSystem):
        kkt = K(cb, o.DenseLDLInertiaSolver)
        kkt.initialize()
        kkt.get_jacobian()[:] = it.jac; kkt.get_hessian()[:] = it.hess
        kkt.compress_jacobian(); kkt.compress_hessian()
        kkt.reg[:] = 1e-8; kkt.du_diag[:] = 0.0
        kkt.l_diag[:] = it.l_diag; kkt.u_diag[:] = it.u_diag; kkt.l_lower[:] = it.l_lower; kkt.u_lower[:] = it.u_lower
        o.set_aug_diagonal_(kkt); kkt.build_kkt(); kkt.linear_solver.factorize()
        b = o.UnreducedKKTVector.for_kkt(kkt); b.full()[:] = it.rhs
        x = o.UnreducedKKTVector.for_kkt(kkt); w = o.UnreducedKKTVector.for_kkt(kkt)
        ok, nit, ratio = o.solve_refine(x, kkt, b, w)
        assert ok and ratio < 1e-8
        sols.append(x.full().copy())
    assert np.abs(sols[0] - sols[1]).max() / np.abs(sols[0]).max() < 1e-6


def test_committed_golden_fixture_matches_oracle():
    """tests/golden/hs15_kkt.json (written by tests/golden/make_golden.py) pins the oracle across refactors and is what
    the GPU tests compare against on the box (no reference tree there)."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hs15_kkt.json")))
    assert g["kat_2x2"]["x"] == [0.8542713567839195, 1.4572864321608041]
    cb = o.HS15Model.callback()
    kkt = o.SparseKKTSystem(cb, o.DenseLDLInertiaSolver)
    x, y, inertia = o.test_kkt_system(kkt, o.HS15Model)
    assert np.abs(x.full() - np.array(g["hs15_sparse"]["solve_kkt_of_ones"])).max() < 1e-14
    assert list(inertia) == g["hs15_sparse"]["inertia"] == [4, 0, 2]
    assert (kkt.aug_colptr == np.array(g["hs15_sparse"]["colptr"])).all()
    assert np.abs(kkt.aug_nz - np.array(g["hs15_sparse"]["nzval"])).max() == 0.0
    kc = o.SparseCondensedKKTSystem(cb)
    xc, yc, ic = o.test_kkt_system(kc, o.HS15Model)
    assert np.abs(xc.full() - np.array(g["hs15_condensed"]["solve_kkt_of_ones"])).max() < 1e-14
    assert np.abs(kc.aug_nz - np.array(g["hs15_condensed"]["nzval"])).max() == 0.0


def test_ipm_reduction_restatements_known_answers():
    """Hand-derived values for the restated line-search scalars (src/IPM/kernels.jl:113-130,263-388,675-695): two variables,
    x = (0.5, 1), bounds 0 <= x1 <= 1, x2 <= 3 (ind_lb = [0], ind_ub = [0, 1])."""
    import numpy as np
    x = np.array([0.5, 1.0]); xl = np.array([0.0, -np.inf]); xu = np.array([1.0, 3.0])
    dx = np.array([-1.0, 4.0]); f = np.array([2.0, -1.0]); jacl = np.array([0.25, 0.5])
    zl = np.array([0.2, 0.0]); zu = np.array([0.4, 0.1])
    ind_lb = np.array([0]); ind_ub = np.array([0, 1])
    mu, tau = 0.1, 0.99
    assert o.get_alpha_max(x, xl, xu, dx, tau) == min(1.0, 0.5 * tau / 1.0, 2.0 * tau / 4.0)
    assert o.get_alpha_z(zl[ind_lb], zu[ind_ub], np.array([-0.4]), np.array([0.3, -0.05]), tau) == min(1.0, 0.2 * tau / 0.4, 0.1 * tau / 0.05)
    assert np.isclose(o.get_varphi(1.0, x[ind_lb], xl[ind_lb], xu[ind_ub], x[ind_ub], mu), 1.0 - mu * (np.log(0.5) + np.log(0.5) + np.log(2.0)), rtol=0, atol=1e-15)
    assert o.get_varphi(1.0, np.array([-1.0]), np.array([0.0]), np.array([]), np.array([]), mu) == np.inf        # infeasible slack
    # (f - mu/(x-xl) + mu/(xu-x)) * dx : x1: (2 - 0.2 + 0.2) * -1 = -2 ; x2: (-1 - 0 + 0.05) * 4 = -3.8
    assert np.isclose(o.get_varphi_d(f, x, xl, xu, dx, mu), -5.8, rtol=0, atol=1e-14)
    assert np.isclose(o.get_inf_du(f, zl, zu, jacl, 2.0), max(abs(2 - 0.2 + 0.4 + 0.25), abs(-1 - 0 + 0.1 + 0.5)) / 2.0)
    # complementarity products: lb: 0.5*0.2 = 0.1 ; ub: 0.5*0.4 = 0.2, 2*0.1 = 0.2
    assert np.isclose(o.get_inf_compl(x[ind_lb], xl[ind_lb], zl[ind_lb], xu[ind_ub], x[ind_ub], zu[ind_ub], mu, 4.0), 0.1 / 4.0)
    assert np.isclose(o.get_average_complementarity(x[ind_lb], xl[ind_lb], zl[ind_lb], x[ind_ub], xu[ind_ub], zu[ind_ub]), 0.5 / 3)
    assert np.isclose(o.get_min_complementarity(x[ind_lb], xl[ind_lb], zl[ind_lb], x[ind_ub], xu[ind_ub], zu[ind_ub]), 0.1)
    assert o.get_average_complementarity(*(np.array([]),) * 6) == 0.0 and o.get_min_complementarity(*(np.array([]),) * 6) == np.inf
    assert o.get_rel_search_norm(x, dx) == 4.0 / 2.0
    assert o.get_sd(np.array([3.0, -5.0]), zl[ind_lb], zu[ind_ub], 100.0) == 1.0 and o.get_sc(np.array([600.0]), np.array([]), 100.0) == 6.0
    p = o.set_aug_rhs(x, xl, xu, f, zl, zu, jacl, np.array([0.7]), mu, ind_lb, ind_ub)
    assert np.allclose(p, [-2 + 0.2 - 0.4 - 0.25, 1 + 0 - 0.1 - 0.5, -0.7, (0 - 0.5) * 0.2 + mu, (1 - 0.5) * 0.4 - mu, (3 - 1) * 0.1 - mu], rtol=0, atol=1e-15)


# ------------------------------------------------------------------------------------------------ C part of the oracle
def _standalone_workloads():
    """workloads.py is plain numpy: load it by path so that oracle-only tests do not need the CUDA library"""
    import importlib.util, os, sys
    if "b2_workloads_standalone" in sys.modules:
        return sys.modules["b2_workloads_standalone"]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "madnlp.jl_b200", "workloads.py")
    spec = importlib.util.spec_from_file_location("b2_workloads_standalone", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2_workloads_standalone"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_c_oracle_is_loaded_and_pinned_to_the_2x2_kat():
    """oracle/kkt_oracle.c (LDLSolver = src/LinearSolvers/ldl.jl) against the reference's KAT (MadNLPTests.jl:24-51)."""
    assert o.clib is not None, "oracle/libkkt_oracle.so missing (make -C oracle)"
    row, col, val = np.array([0, 1, 1]), np.array([0, 0, 1]), np.array([1.0, 0.1, 2.0])
    cp, rv, mp = o.coo_to_csc(row, col, 2, 2)
    nz = np.zeros(len(rv)); o.transfer(nz, val, mp)
    M = o.LDLSolver(cp, rv, nz, 2).factorize()
    assert M.inertia() == (2, 0, 0)
    x = M.solve(np.array([1.0, 3.0]))
    assert np.abs(x - np.array([0.8542713567839195, 1.4572864321608041])).max() < 1e-15


def test_c_loops_are_bit_identical_to_the_numpy_statements():
    """_transfer!, _build_condensed_aug_coord!, _set_aug_diagonal!, reduce_rhs!/finish_aug_solve!/_kktmul! in C ==
    the numpy add.at statements (same sequential order of additions); mat-vecs to rounding."""
    W = _standalone_workloads()
    model, st = W.acopf_case("case300_synth")
    it = W.ipm_iterates(model, st, 3, seed=7)[1]
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    out = {}
    try:
        for use_c in (True, False):
            o.USE_C = use_c
            kkt = o.SparseCondensedKKTSystem(cb, o.DenseLDLInertiaSolver)
            kkt.initialize()
            kkt.get_jacobian()[:] = it.jac; kkt.get_hessian()[:] = it.hess
            kkt.reg[:] = it.reg; kkt.du_diag[:] = it.du_diag
            kkt.l_diag[:] = it.l_diag; kkt.u_diag[:] = it.u_diag; kkt.l_lower[:] = it.l_lower; kkt.u_lower[:] = it.u_lower
            kkt.compress_jacobian(); kkt.compress_hessian(); o.set_aug_diagonal_(kkt); kkt.build_kkt()
            kkt.linear_solver.factorize()
            x = o.UnreducedKKTVector.for_kkt(kkt); x.full()[:] = it.rhs
            w = o.UnreducedKKTVector.for_kkt(kkt); w.full()[:] = 1.0
            kkt.mul(w, x, -0.5, 2.0)
            r = x.copy(); o.reduce_rhs(kkt, r)
            f = x.copy(); o.finish_aug_solve(kkt, f)
            out[use_c] = dict(aug=kkt.aug_nz.copy(), jt=kkt.jt_nz.copy(), h=kkt.hess_nz.copy(), pr=kkt.pr_diag.copy(),
                              r=r.full().copy(), f=f.full().copy(), w=w.full().copy())
    finally:
        o.USE_C = True
    for k in ("aug", "jt", "h", "pr", "r", "f"):
        assert (out[True][k] == out[False][k]).all(), k
    assert np.abs(out[True]["w"] - out[False]["w"]).max() <= 1e-14 * np.abs(out[False]["w"]).max()


def test_ldl_solver_inertia_and_solution_against_lapack():
    """LDLSolver (Davis LDL) vs dsytrf-based truth: same inertia on an indefinite quasi-definite KKT, same solution."""
    rng = np.random.default_rng(5)
    n, m = 60, 25
    H = rng.standard_normal((n, n)); H = H @ H.T + n * np.eye(n)
    J = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.2)
    K = np.block([[H, J.T], [J, -1e-3 * np.eye(m)]])
    Kl = np.tril(K)
    I, Jc = np.nonzero(Kl)
    cp, rv, mp = o.coo_to_csc(I, Jc, n + m, n + m)
    nz = np.zeros(len(rv)); o.transfer(nz, Kl[I, Jc], mp)
    M = o.LDLSolver(cp, rv, nz, n + m, perm=np.arange(n + m)).factorize()     # natural order: primal block first
    T = o.DenseLDLInertiaSolver(cp, rv, nz, n + m).factorize()
    assert M.inertia() == T.inertia() == (n, 0, m)
    b = rng.standard_normal(n + m)
    assert np.abs(M.solve(b.copy()) - T.solve(b.copy())).max() < 1e-10
    M2 = o.LDLSolver(cp, rv, nz, n + m).factorize()                           # minimum-degree stand-in ordering
    assert sorted(M2.perm.tolist()) == list(range(n + m))
    assert np.abs(M2.solve(b.copy()) - T.solve(b.copy())).max() < 1e-9
