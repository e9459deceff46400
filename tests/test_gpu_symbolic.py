"""Device-side symbolic phase: b2_coo_to_csc_device (device sorts, like lib/MadNLPGPU/src/KKT/gpu_sparse.jl:260-302)
must reproduce the host construction b2_coo_to_csc / the oracle's coo_to_csc (src/matrixtools.jl:55-95) EXACTLY: same colptr, rowval,
COO->CSC map -- integer work, bit-exact bar."""
import ctypes as C

import numpy as np
import pytest

import madnlp_oracle as o
import madnlp_jl_b200 as pkg

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _device_coo_to_csc(I, J, m, n):
    lib, check = pkg.capi.lib, pkg.capi.check
    nnz = len(I)
    Id = torch.from_numpy(np.ascontiguousarray(I, dtype=np.int32)).cuda()
    Jd = torch.from_numpy(np.ascontiguousarray(J, dtype=np.int32)).cuda()
    colptr = torch.full((n + 1,), -7, dtype=torch.int32, device="cuda")
    rowval = torch.full((max(nnz, 1),), -7, dtype=torch.int32, device="cuda")
    cmap = torch.full((max(nnz, 1),), -7, dtype=torch.int64, device="cuda")
    ncsc = C.c_int64(-1)
    check(lib.b2_coo_to_csc_device(m, n, nnz, Id.data_ptr() if nnz else None, Jd.data_ptr() if nnz else None, colptr.data_ptr(),
                                   rowval.data_ptr(), cmap.data_ptr(), C.byref(ncsc), torch.cuda.current_stream().cuda_stream))
    return colptr.cpu().numpy(), rowval.cpu().numpy()[:ncsc.value], cmap.cpu().numpy()[:nnz], ncsc.value


@pytest.mark.parametrize("m,n,nnz,seed", [(1, 1, 0, 0), (5, 7, 1, 1), (301, 257, 5000, 2), (40000, 40000, 600000, 3)])
def test_device_coo_to_csc_matches_host_and_oracle(m, n, nnz, seed):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from madnlp_jl_b200 import kkt as K
    rng = np.random.default_rng(seed)
    I = rng.integers(0, m, nnz); J = rng.integers(0, n, nnz)
    if nnz > 1000:                                  # duplicates, an empty column range, a full column
        I[:400] = I[400:800]; J[:400] = J[400:800]
        J[J == n // 2] = n // 2 + 1
    cp_d, rv_d, mp_d, ncsc = _device_coo_to_csc(I, J, m, n)
    cp_h, rv_h, mp_h = K.coo_to_csc(I, J, m, n)
    assert ncsc == len(rv_h)
    assert (cp_d == cp_h).all() and (rv_d == rv_h).all() and (mp_d == mp_h).all()
    if nnz:
        cp_o, rv_o, mp_o = o.coo_to_csc(I, J, m, n)
        assert (cp_d == cp_o).all() and (rv_d == rv_o).all() and (mp_d == mp_o).all()


def test_device_coo_to_csc_on_the_headline_kkt_pattern():
    """the augmented COO of the OPF-10k SparseKKTSystem ([pr_diag | hess | jac | -1 | du_diag] positions, augmented.jl:77-107)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    W = pkg.workloads
    model, st = W.acopf_case("case1354_pegase", relax_equality=False)
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    kc = o.SparseKKTSystem(cb, o.UmfpackStandInSolver)
    cp_d, rv_d, mp_d, ncsc = _device_coo_to_csc(kc.aug_I, kc.aug_J, kc.N, kc.N)
    assert (cp_d == kc.aug_colptr).all() and (rv_d == kc.aug_rowval).all() and (mp_d == kc.aug_csc_map).all()


@pytest.mark.parametrize("case", ["hs15", "case300_synth", "case1354_pegase"])
def test_device_condensed_symbolic_equals_host_plan(case):
    """b2_condensed_symbolic_device == b2_condensed_symbolic (build_condensed_aug_symbolic, condensed.jl:201-301): same pattern, same map
    sizes, and -- because the per-slot source order is the same -- a BIT-IDENTICAL assembled matrix."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from madnlp_jl_b200 import kkt as K
    lib, check = pkg.capi.lib, pkg.capi.check
    W = pkg.workloads
    if case == "hs15":
        cb = o.HS15Model.callback()
        jac = o.HS15Model.jac_coord(np.array([0.3, 0.7])); hess = o.HS15Model.hess_coord(np.array([0.3, 0.7]), np.array([0.5, -0.2]))
    else:
        model, st = W.acopf_case(case)
        cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
        it = W.ipm_iterates(model, st, 1, seed=3)[0]
        jac, hess = it.jac, it.hess
    kg = K.SparseCondensedKKTSystem(cb)
    n, m = kg.n, kg.m
    dev32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()
    hcp, hrv = dev32(kg.hess_com.colptr), dev32(kg.hess_com.rowval if len(kg.hess_com.rowval) else np.zeros(1))
    jcp, jrv = dev32(kg.jt_csc.colptr), dev32(kg.jt_csc.rowval if len(kg.jt_csc.rowval) else np.zeros(1))
    h = C.c_void_p(); nnz = C.c_int64(0)
    check(lib.b2_condensed_symbolic_device(n, m, hcp.data_ptr(), hrv.data_ptr(), jcp.data_ptr(), jrv.data_ptr(), C.byref(h), C.byref(nnz),
                                           torch.cuda.current_stream().cuda_stream))
    try:
        assert nnz.value == kg.aug_com.nnz
        cp = np.zeros(n + 1, dtype=np.int32); rv = np.zeros(nnz.value, dtype=np.int32)
        check(lib.b2_condensed_pattern(h, cp.ctypes.data, rv.ctypes.data))
        assert (cp == kg.aug_com.colptr).all() and (rv == kg.aug_com.rowval).all()
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.b2_condensed_plan_sizes(h, C.byref(a), C.byref(b), C.byref(c)))
        assert dict(dptr=a.value, hptr=b.value, jptr=c.value) == kg.plan_sizes()
        # numeric: assemble through both plans
        kg.initialize()
        kg.get_jacobian().copy_(torch.from_numpy(np.ascontiguousarray(jac)).cuda()); kg.get_hessian().copy_(torch.from_numpy(np.ascontiguousarray(hess)).cuda())
        kg.compress_jacobian(); kg.compress_hessian()
        rng = np.random.default_rng(1)
        kg.pr_diag.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, len(kg.pr_diag))).cuda()); kg.du_diag.copy_(torch.from_numpy(-rng.uniform(1e-6, 1e-3, m)).cuda())
        kg.build_kkt()
        ref = kg.aug_com.nzval.clone()
        out = torch.full_like(ref, 7.0); dbuf = torch.zeros(m, dtype=torch.float64, device="cuda")
        check(lib.b2_condensed_assemble(h, out.data_ptr(), kg.pr_diag.data_ptr(), kg.du_diag.data_ptr(), kg.hess_com.nzval.data_ptr(),
                                        kg.jt_csc.nzval.data_ptr(), dbuf.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
    finally:
        lib.b2_condensed_plan_destroy(h)


def test_kkt_system_built_with_device_symbolic_is_identical(monkeypatch):
    """the host mirror with B2_DEVICE_SYMBOLIC=1 (all sorts on the device, like MadNLPGPU) == the default host constructions"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from madnlp_jl_b200 import kkt as K
    W = pkg.workloads
    model, st = W.acopf_case("case300_synth")
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    it = W.ipm_iterates(model, st, 1, seed=4)[0]
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("B2_DEVICE_SYMBOLIC", flag)
        kg = K.SparseCondensedKKTSystem(cb); kg.initialize()
        for k, v in ((kg.get_jacobian(), it.jac), (kg.get_hessian(), it.hess), (kg.reg, it.reg + 1e-8), (kg.du_diag, it.du_diag - 1e-7),
                     (kg.l_diag, it.l_diag), (kg.u_diag, it.u_diag), (kg.l_lower, it.l_lower), (kg.u_lower, it.u_lower)):
            k.copy_(torch.from_numpy(np.ascontiguousarray(v)).cuda())
        kg.compress_jacobian(); kg.compress_hessian(); kg.set_aug_diagonal_(); kg.build_kkt()
        outs.append((kg.aug_com.colptr.copy(), kg.aug_com.rowval.copy(), kg.aug_com.nzval.cpu().numpy(), kg.jt_csc.nzval.cpu().numpy()))
    for a, b in zip(outs[0], outs[1]):
        assert (a == b).all()
