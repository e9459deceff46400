"""Host-side mirror of MadNLP's AbstractKKTSystem interface (src/KKT/KKTsystem.jl:86-205) with all storage on the
device (torch tensors = device memory only) and every numeric operation a C-ABI call into the CUDA library.

Same names / fields / argument meaning as the reference so tests read like test/kkt_test.jl:
    create_kkt_system(KKT, cb, linear_solver) ; initialize ; get_jacobian / get_hessian (aliasing views the callbacks
    write into) ; compress_jacobian / compress_hessian ; build_kkt ; factorize_kkt (via kkt.linear_solver) ;
    solve_kkt(w) ; mul(w, x, alpha, beta) ; regularize_diagonal ; set_aug_diagonal_ ; is_inertia_correct ;
    should_regularize_dual ; num_variables ; fields reg, pr_diag, du_diag, l_diag, u_diag, l_lower, u_lower,
    ind_lb, ind_ub, hess, jac, aug_com, linear_solver.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import lib, check, ptr
from .linear_solvers import B200DenseSolver, B200SparseSolver, DeviceCSC

_DEV = "cuda"


def _dz(n):
    return torch.zeros(int(n), dtype=torch.float64, device=_DEV)


def _sp(stream=None):
    return capi.stream_ptr(stream)


def force_lower_triangular(I, J):
    """src/matrixtools.jl:129-137 (host, one-time, on the sparsity pattern)."""
    sw = J > I
    tmp = J[sw].copy()
    J[sw] = I[sw]
    I[sw] = tmp


def coo_to_csc(I, J, m, n):
    """src/matrixtools.jl:55-95 via b2_coo_to_csc: (colptr, rowval, map), 0-based."""
    I32 = np.ascontiguousarray(I, dtype=np.int32)
    J32 = np.ascontiguousarray(J, dtype=np.int32)
    nnz = len(I32)
    colptr = np.zeros(n + 1, dtype=np.int32)
    rowval = np.zeros(max(nnz, 1), dtype=np.int32)
    cmap = np.zeros(max(nnz, 1), dtype=np.int64)
    ncsc = C.c_int64(0)
    check(lib.b2_coo_to_csc(m, n, nnz, I32.ctypes.data, J32.ctypes.data, colptr.ctypes.data, rowval.ctypes.data,
                            cmap.ctypes.data, C.byref(ncsc)))
    return colptr, rowval[:ncsc.value].copy(), cmap[:nnz].copy()


def coo_to_csc_device(I, J, m, n):
    """The same construction with device sorts (lib/MadNLPGPU/src/KKT/gpu_sparse.jl:260-302) via b2_coo_to_csc_device; returns host
    copies of (colptr, rowval, map) -- identical to coo_to_csc's (tests/test_gpu_symbolic.py)."""
    nnz = len(I)
    Id = torch.from_numpy(np.ascontiguousarray(I, dtype=np.int32)).to(_DEV)
    Jd = torch.from_numpy(np.ascontiguousarray(J, dtype=np.int32)).to(_DEV)
    colptr = torch.zeros(n + 1, dtype=torch.int32, device=_DEV)
    rowval = torch.zeros(max(nnz, 1), dtype=torch.int32, device=_DEV)
    cmap = torch.zeros(max(nnz, 1), dtype=torch.int64, device=_DEV)
    ncsc = C.c_int64(0)
    check(lib.b2_coo_to_csc_device(m, n, nnz, ptr(Id) if nnz else None, ptr(Jd) if nnz else None, ptr(colptr), ptr(rowval), ptr(cmap),
                                   C.byref(ncsc), _sp()))
    return colptr.cpu().numpy(), rowval.cpu().numpy()[:ncsc.value].copy(), cmap.cpu().numpy()[:nnz].copy()


def _symbolic_on_device():
    """B2_DEVICE_SYMBOLIC=1: build the COO->CSC maps and the condensed pattern/maps with device sorts, like the
    reference's GPU path; default: the host constructions (identical results)."""
    import os
    return os.environ.get("B2_DEVICE_SYMBOLIC") == "1"


class _Plan:
    """owns a native plan handle and frees it"""

    def __init__(self, handle, destroy):
        self.h = handle
        self._destroy = destroy

    def __del__(self):
        if getattr(self, "h", None):
            self._destroy(self.h)
            self.h = None


def _transfer_plan(cmap, nnz_csc):
    h = C.c_void_p()
    cm = np.ascontiguousarray(cmap, dtype=np.int64)
    check(lib.b2_transfer_plan_create(len(cm), int(nnz_csc), cm.ctypes.data, C.byref(h)))
    return _Plan(h, lib.b2_transfer_plan_destroy)


def _spmv_plan(nrow, ncol, colptr, rowval):
    h = C.c_void_p()
    cp = np.ascontiguousarray(colptr, dtype=np.int32)
    rv = np.ascontiguousarray(rowval, dtype=np.int32)
    check(lib.b2_spmv_plan_create(nrow, ncol, cp.ctypes.data, rv.ctypes.data if len(rv) else None, C.byref(h)))
    return _Plan(h, lib.b2_spmv_plan_destroy)


def _bounds(n_tot, ind_lb, ind_ub):
    h = C.c_void_p()
    lb = np.ascontiguousarray(ind_lb, dtype=np.int64)
    ub = np.ascontiguousarray(ind_ub, dtype=np.int64)
    check(lib.b2_bounds_create(n_tot, len(lb), len(ub), lb.ctypes.data if len(lb) else None,
                               ub.ctypes.data if len(ub) else None, C.byref(h)))
    return _Plan(h, lib.b2_bounds_destroy)


class UnreducedKKTVector:
    """src/KKT/rhs.jl:90-129: one contiguous device buffer [x (n_tot) | y (m) | zl (nlb) | zu (nub)] with views."""

    def __init__(self, n, m, nlb, nub):
        self.n, self.m, self.nlb, self.nub = int(n), int(m), int(nlb), int(nub)
        self.values = _dz(n + m + nlb + nub)

    @classmethod
    def for_kkt(cls, kkt):
        return cls(len(kkt.pr_diag), len(kkt.du_diag), len(kkt.l_diag), len(kkt.u_diag))

    def full(self):
        return self.values

    def primal(self):
        return self.values[: self.n]

    def dual(self):
        return self.values[self.n: self.n + self.m]

    def primal_dual(self):
        return self.values[: self.n + self.m]

    def dual_lb(self):
        return self.values[self.n + self.m: self.n + self.m + self.nlb]

    def dual_ub(self):
        return self.values[self.n + self.m + self.nlb:]

    def copy(self):
        o = UnreducedKKTVector(self.n, self.m, self.nlb, self.nub)
        o.values.copy_(self.values)
        return o


class _KKTBase:
    stream = None

    # ---- generic pieces (src/KKT/KKTsystem.jl:210-256, src/IPM/kernels.jl) ----
    def _init_common(self, cb, n_tot, m):
        nlb, nub = len(cb.ind_lb), len(cb.ind_ub)
        self.reg = _dz(n_tot)
        self.l_diag = _dz(nlb); self.u_diag = _dz(nub)
        self.l_lower = _dz(nlb); self.u_lower = _dz(nub)
        self.ind_lb = np.asarray(cb.ind_lb, dtype=np.int64)
        self.ind_ub = np.asarray(cb.ind_ub, dtype=np.int64)
        self.ind_ineq = np.asarray(cb.ind_ineq, dtype=np.int64)
        self._bounds = _bounds(n_tot, self.ind_lb, self.ind_ub)
        self._n_tot, self._m = int(n_tot), int(m)

    def set_aug_diagonal_(self):
        """_set_aug_diagonal!  (src/IPM/kernels.jl:22-27)."""
        check(lib.b2_set_aug_diagonal(self._bounds.h, ptr(self.reg), ptr(self.l_lower), ptr(self.l_diag),
                                      ptr(self.u_lower), ptr(self.u_diag), ptr(self.pr_diag), _sp(self.stream)))

    def regularize_diagonal(self, primal, dual):
        """src/KKT/KKTsystem.jl:222-226."""
        check(lib.b2_regularize_diagonal(self._n_tot, self._m, float(primal), float(dual), ptr(self.reg),
                                         ptr(self.pr_diag), ptr(self.du_diag), _sp(self.stream)))

    def reduce_rhs(self, w):
        check(lib.b2_reduce_rhs(self._bounds.h, self._m, ptr(self.l_diag), ptr(self.u_diag), ptr(w.values), _sp(self.stream)))

    def finish_aug_solve(self, w):
        check(lib.b2_finish_aug_solve(self._bounds.h, self._m, ptr(self.l_lower), ptr(self.u_lower), ptr(self.l_diag),
                                      ptr(self.u_diag), ptr(w.values), _sp(self.stream)))

    def _kktmul(self, w, x, alpha, beta):
        check(lib.b2_kktmul(self._bounds.h, self._m, ptr(self.reg), ptr(self.du_diag), ptr(self.l_lower), ptr(self.u_lower),
                            ptr(self.l_diag), ptr(self.u_diag), float(alpha), float(beta), ptr(x.values), ptr(w.values),
                            _sp(self.stream)))

    def factorize_kkt(self):
        return self.linear_solver.factorize()

    def get_kkt(self):
        return self.aug_com

    def is_inertia_correct(self, num_pos, num_zero, num_neg):
        """src/KKT/KKTsystem.jl:242-244."""
        return num_zero == 0 and num_pos == self.num_variables()

    def should_regularize_dual(self, num_pos, num_zero, num_neg):
        """src/KKT/KKTsystem.jl:252-254."""
        return num_zero != 0

    def _initialize_common(self):
        self.reg.fill_(1.0); self.pr_diag.fill_(1.0); self.du_diag.zero_(); self.hess.zero_()


# ======================================================================================================
class SparseKKTSystem(_KKTBase):
    """src/KKT/Sparse/augmented.jl: augmented system as COO value vector
    V = [pr_diag(n_tot) | hess(nnzh) | jac(nnzj) | slack -1 (ns) | du_diag(m)] (aliasing views) -> lower CSC."""

    def __init__(self, cb, linear_solver=B200SparseSolver, opt_linear_solver=None):
        n, m = cb.nvar, cb.ncon
        ns = len(cb.ind_ineq)
        hI = np.array(cb.hess_I, dtype=np.int64); hJ = np.array(cb.hess_J, dtype=np.int64)
        force_lower_triangular(hI, hJ)                               # augmented.jl:65
        jI = np.asarray(cb.jac_I, dtype=np.int64); jJ = np.asarray(cb.jac_J, dtype=np.int64)
        n_jac, n_hess = len(jI), len(hI)
        n_tot = n + ns
        self.n, self.m, self.ns, self.n_tot = n, m, ns, n_tot
        L = n_tot + m + n_hess + n_jac + ns                          # augmented.jl:75
        o1 = n_tot; o2 = o1 + n_hess; o3 = o2 + n_jac; o4 = o3 + ns
        I = np.empty(L, dtype=np.int64); J = np.empty(L, dtype=np.int64)
        ineq = np.asarray(cb.ind_ineq, dtype=np.int64)
        I[:o1] = np.arange(n_tot); J[:o1] = np.arange(n_tot)
        I[o1:o2] = hI; J[o1:o2] = hJ
        I[o2:o3] = jI + n_tot; J[o2:o3] = jJ
        I[o3:o4] = ineq + n_tot; J[o3:o4] = np.arange(n, n + ns)
        I[o4:] = np.arange(n_tot, n_tot + m); J[o4:] = np.arange(n_tot, n_tot + m)
        self.V = _dz(L)
        self.pr_diag = self.V[:o1]
        self.hess = self.V[o1:o2]
        self.jac = self.V[o2:o4]
        self.jac_callback = self.V[o2:o3]
        self.du_diag = self.V[o4:]
        self._init_common(cb, n_tot, m)
        N = n_tot + m
        self.N = N
        cp, rv, mp = coo_to_csc(I, J, N, N)
        self.aug_com = DeviceCSC(N, N, cp, rv, _dz(len(rv)))
        self._aug_plan = _transfer_plan(mp, len(rv))
        self.aug_csc_map = mp
        jI2 = np.concatenate([jI, ineq]); jJ2 = np.concatenate([jJ, np.arange(n, n + ns)])
        cp, rv, mp = coo_to_csc(jI2, jJ2, m, n_tot)
        self.jac_com = DeviceCSC(m, n_tot, cp, rv, _dz(len(rv)))
        self._jac_plan = _transfer_plan(mp, len(rv))
        self._jac_spmv = _spmv_plan(m, n_tot, cp, rv)
        cp, rv, mp = coo_to_csc(hI, hJ, n_tot, n_tot)
        self.hess_com = DeviceCSC(n_tot, n_tot, cp, rv, _dz(len(rv)))
        self._hess_plan = _transfer_plan(mp, len(rv))
        self._hess_spmv = _spmv_plan(n_tot, n_tot, cp, rv)
        if opt_linear_solver is None and hasattr(linear_solver, "default_options"):
            opt_linear_solver = linear_solver.default_options()
        if opt_linear_solver is not None and getattr(opt_linear_solver, "kkt_n_primal", None) == 0:
            opt_linear_solver.kkt_n_primal = n_tot     # zero (2,2) block: dual rows follow a primal neighbour
        self.linear_solver = linear_solver(self.aug_com, opt_linear_solver)

    def num_variables(self):
        return len(self.pr_diag)

    def initialize(self):
        """Sparse/utils.jl:52-62."""
        self._initialize_common()
        self.l_lower.zero_(); self.u_lower.zero_(); self.l_diag.fill_(1.0); self.u_diag.fill_(1.0)
        self.hess_com.nzval.zero_()

    def get_jacobian(self):
        return self.jac_callback

    def get_hessian(self):
        return self.hess

    def compress_jacobian(self):
        """Sparse/utils.jl:36-40."""
        if self.ns:
            check(lib.b2_fill(self.ns, -1.0, ptr(self.jac[-self.ns:]), _sp(self.stream)))
        check(lib.b2_transfer(self._jac_plan.h, ptr(self.jac_com.nzval), ptr(self.jac), _sp(self.stream)))

    def compress_hessian(self):
        """Sparse/utils.jl:48-50."""
        check(lib.b2_transfer(self._hess_plan.h, ptr(self.hess_com.nzval), ptr(self.hess), _sp(self.stream)))

    def build_kkt(self):
        """augmented.jl:146-148: transfer!(aug_com, aug_raw, aug_csc_map)."""
        check(lib.b2_transfer(self._aug_plan.h, ptr(self.aug_com.nzval), ptr(self.V), _sp(self.stream)))

    def solve_kkt(self, w: UnreducedKKTVector):
        """src/IPM/factorization.jl:41-46."""
        self.reduce_rhs(w)
        self.linear_solver.solve_linear_system(w.primal_dual())
        self.finish_aug_solve(w)
        return w

    def mul(self, w, x, alpha=1.0, beta=0.0):
        """src/IPM/factorization.jl:231-237."""
        sp = _sp(self.stream)
        check(lib.b2_spmv_symlower(self._hess_spmv.h, ptr(self.hess_com.nzval), ptr(x.values), ptr(w.values), alpha, beta, sp))
        check(lib.b2_spmv_t(self._jac_spmv.h, ptr(self.jac_com.nzval), ptr(x.dual()), ptr(w.values), alpha, 1.0, sp))
        check(lib.b2_spmv_n(self._jac_spmv.h, ptr(self.jac_com.nzval), ptr(x.values), ptr(w.dual()), alpha, beta, sp))
        self._kktmul(w, x, alpha, beta)
        return w

    def jtprod(self, y, x):
        """Sparse/utils.jl:28-30."""
        check(lib.b2_spmv_t(self._jac_spmv.h, ptr(self.jac_com.nzval), ptr(x), ptr(y), 1.0, 0.0, _sp(self.stream)))


# ======================================================================================================
class SparseCondensedKKTSystem(_KKTBase):
    """src/KKT/Sparse/condensed.jl: n x n condensed system H + Sigma_x + J' D J (all constraints inequalities)."""

    def __init__(self, cb, linear_solver=B200SparseSolver, opt_linear_solver=None):
        n, m = cb.nvar, cb.ncon
        ns = len(cb.ind_ineq)
        if ns != m:
            raise ValueError("SparseCondensedKKTSystem does not support equality constrained NLPs.")   # condensed.jl:68-70
        hI = np.array(cb.hess_I, dtype=np.int64); hJ = np.array(cb.hess_J, dtype=np.int64)
        force_lower_triangular(hI, hJ)
        jI = np.asarray(cb.jac_I, dtype=np.int64); jJ = np.asarray(cb.jac_J, dtype=np.int64)
        self.n, self.m, self.ns, self.n_tot = n, m, ns, n + ns
        self.pr_diag = _dz(n + ns); self.du_diag = _dz(m)
        self._init_common(cb, n + ns, m)
        self.buffer = _dz(m); self.buffer2 = _dz(m); self.diag_buffer = _dz(m)
        self.hess = _dz(len(hI)); self.jac = _dz(len(jI))
        dev_sym = _symbolic_on_device()
        coo_to_csc = coo_to_csc_device if dev_sym else globals()["coo_to_csc"]
        cp, rv, mp = coo_to_csc(jJ, jI, n, m)                        # jt_coo: I = jac_J, J = jac_I (condensed.jl:105-110)
        self.jt_csc = DeviceCSC(n, m, cp, rv, _dz(len(rv)))
        self._jt_plan = _transfer_plan(mp, len(rv))
        self._jt_spmv = _spmv_plan(n, m, cp, rv)
        hcp, hrv, hmp = coo_to_csc(hI, hJ, n, n)
        self.hess_com = DeviceCSC(n, n, hcp, hrv, _dz(len(hrv)))
        self._hess_plan = _transfer_plan(hmp, len(hrv))
        self._hess_spmv = _spmv_plan(n, n, hcp, hrv)
        h = C.c_void_p(); nnz_aug = C.c_int64(0)
        if dev_sym:
            d32 = lambda a: torch.from_numpy(np.ascontiguousarray(a if len(a) else np.zeros(1), dtype=np.int32)).to(_DEV)
            pats = [d32(hcp), d32(hrv), d32(cp), d32(rv)]
            check(lib.b2_condensed_symbolic_device(n, m, ptr(pats[0]), ptr(pats[1]), ptr(pats[2]), ptr(pats[3]), C.byref(h), C.byref(nnz_aug), _sp()))
        else:
            check(lib.b2_condensed_symbolic(n, m, hcp.ctypes.data, hrv.ctypes.data if len(hrv) else None, cp.ctypes.data,
                                            rv.ctypes.data if len(rv) else None, C.byref(h), C.byref(nnz_aug)))
        self._cond = _Plan(h, lib.b2_condensed_plan_destroy)
        acp = np.zeros(n + 1, dtype=np.int32); arv = np.zeros(nnz_aug.value, dtype=np.int32)
        check(lib.b2_condensed_pattern(h, acp.ctypes.data, arv.ctypes.data))
        self.aug_com = DeviceCSC(n, n, acp, arv, _dz(nnz_aug.value))
        self.N = n
        self.linear_solver = linear_solver(self.aug_com, opt_linear_solver)

    def plan_sizes(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.b2_condensed_plan_sizes(self._cond.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(dptr=a.value, hptr=b.value, jptr=c.value)

    def num_variables(self):
        return len(self.pr_diag)

    def initialize(self):
        self._initialize_common()
        self.l_lower.zero_(); self.u_lower.zero_(); self.l_diag.fill_(1.0); self.u_diag.fill_(1.0)
        self.hess_com.nzval.zero_()

    def get_jacobian(self):
        return self.jac

    def get_hessian(self):
        return self.hess

    def compress_jacobian(self):
        """condensed.jl:145-148."""
        check(lib.b2_transfer(self._jt_plan.h, ptr(self.jt_csc.nzval), ptr(self.jac), _sp(self.stream)))

    def compress_hessian(self):
        check(lib.b2_transfer(self._hess_plan.h, ptr(self.hess_com.nzval), ptr(self.hess), _sp(self.stream)))

    def build_kkt(self):
        """condensed.jl:354-366 (+ :328-345)."""
        check(lib.b2_condensed_assemble(self._cond.h, ptr(self.aug_com.nzval), ptr(self.pr_diag), ptr(self.du_diag),
                                        ptr(self.hess_com.nzval), ptr(self.jt_csc.nzval), ptr(self.diag_buffer),
                                        _sp(self.stream)))

    def is_inertia_correct(self, num_pos, num_zero, num_neg):
        """condensed.jl:138-140."""
        return num_zero == 0 and num_pos == self.n

    def should_regularize_dual(self, num_pos, num_zero, num_neg):
        return True                                                    # condensed.jl:141

    def solve_kkt(self, w: UnreducedKKTVector):
        """src/IPM/factorization.jl:143-167."""
        sp = _sp(self.stream)
        check(lib.b2_condensed_solve_pre(self._bounds.h, self._jt_spmv.h, self.n, self.m, ptr(self.jt_csc.nzval),
                                         ptr(self.pr_diag), ptr(self.diag_buffer), ptr(self.l_diag), ptr(self.u_diag),
                                         ptr(self.buffer), ptr(w.values), sp))
        self.linear_solver.solve_linear_system(w.values[: self.n])
        check(lib.b2_condensed_solve_post(self._bounds.h, self._jt_spmv.h, self.n, self.m, ptr(self.jt_csc.nzval),
                                          ptr(self.pr_diag), ptr(self.diag_buffer), ptr(self.l_lower), ptr(self.u_lower),
                                          ptr(self.l_diag), ptr(self.u_diag), ptr(self.buffer), ptr(w.values), sp))
        return w

    def mul(self, w, x, alpha=1.0, beta=0.0):
        """src/IPM/factorization.jl:303-324."""
        check(lib.b2_condensed_kkt_mul(self._bounds.h, self._hess_spmv.h, self._jt_spmv.h, self.n, self.m,
                                       ptr(self.hess_com.nzval), ptr(self.jt_csc.nzval), ptr(self.reg), ptr(self.du_diag),
                                       ptr(self.l_lower), ptr(self.u_lower), ptr(self.l_diag), ptr(self.u_diag),
                                       float(alpha), float(beta), ptr(x.values), ptr(w.values), _sp(self.stream)))
        return w

    def mul_norm(self, w, x, alpha, beta, norm_out):
        """mul! that also accumulates ||w||_inf of the result into the (zeroed) device scalar `norm_out`"""
        check(lib.b2_condensed_kkt_mul_norm(self._bounds.h, self._hess_spmv.h, self._jt_spmv.h, self.n, self.m,
                                            ptr(self.hess_com.nzval), ptr(self.jt_csc.nzval), ptr(self.reg), ptr(self.du_diag),
                                            ptr(self.l_lower), ptr(self.u_lower), ptr(self.l_diag), ptr(self.u_diag),
                                            float(alpha), float(beta), ptr(x.values), ptr(w.values), ptr(norm_out), _sp(self.stream)))
        return w

    def jtprod(self, y, x):
        """condensed.jl:150-156."""
        check(lib.b2_spmv_n(self._jt_spmv.h, ptr(self.jt_csc.nzval), ptr(x), ptr(y), 1.0, 0.0, _sp(self.stream)))
        y[self.n:] = -x


# ======================================================================================================
class DenseCondensedKKTSystem(_KKTBase):
    """src/KKT/Dense/condensed.jl.  Dense matrices are torch tensors whose MEMORY is the column-major matrix
    (tensor[j, i] = M[i, j]), so pointers can be handed to the kernels exactly as Julia would hand them."""

    def __init__(self, cb, linear_solver=B200DenseSolver, opt_linear_solver=None):
        n, m = cb.nvar, cb.ncon
        ind_ineq = np.asarray(cb.ind_ineq, dtype=np.int64)
        ind_eq = np.setdiff1d(np.arange(m), ind_ineq).astype(np.int64)
        ns = len(ind_ineq); n_eq = m - ns
        self.n, self.m, self.ns, self.n_eq = n, m, ns, n_eq
        N = n + n_eq
        self.N = N
        self.aug_com = torch.zeros((N, N), dtype=torch.float64, device=_DEV)
        self.hess = torch.zeros((n, n), dtype=torch.float64, device=_DEV)      # memory: column-major n x n
        self.jac = torch.zeros((n, m), dtype=torch.float64, device=_DEV)       # memory: column-major m x n
        self.pr_diag = _dz(n + ns); self.du_diag = _dz(m)
        self._init_common(cb, n + ns, m)
        self.l_diag.fill_(1.0); self.u_diag.fill_(1.0)
        self.pd_buffer = _dz(N); self.diag_buffer = _dz(ns); self.buffer = _dz(m)
        self.ind_eq = ind_eq
        self._ind_ineq_d = torch.from_numpy(ind_ineq).to(_DEV)
        self._ind_eq_d = torch.from_numpy(ind_eq).to(_DEV)
        h = C.c_void_p()
        ii = np.ascontiguousarray(ind_ineq, dtype=np.int64)
        check(lib.b2d_kkt_create(n, m, ns, ii.ctypes.data if ns else None, C.byref(h)))
        self._dk = _Plan(h, lib.b2d_kkt_destroy)
        # J' D J on the 5th-generation tensor cores (tcgen05 int8 digits + TMA, csrc/ozaki_kernels.cuh) when the contraction is big
        # enough to pay for the digit split; B2_OZAKI=0/1 forces the DMMA kernel / the tensor-core kernel
        import os
        want = os.environ.get("B2_OZAKI")
        use = (n >= 512 and 256 <= ns <= 16384) if want is None else (want != "0" and 0 < ns <= 16384)
        self._ozaki = None
        if use:
            hz = C.c_void_p()
            check(lib.b2d_ozaki_plan_create(n, ns, C.byref(hz)))
            self._ozaki = _Plan(hz, lib.b2d_ozaki_plan_destroy)
        self.linear_solver = linear_solver(self.aug_com, opt_linear_solver)

    def num_variables(self):
        return self.n

    def initialize(self):
        self._initialize_common()

    def get_jacobian(self):
        return self.jac

    def get_hessian(self):
        return self.hess

    def set_dense(self, hess_np=None, jac_np=None):
        """upload column-major host matrices (what hess_dense!/jac_dense! would have written)."""
        if hess_np is not None:
            self.hess.copy_(torch.from_numpy(np.ascontiguousarray(hess_np.T)))
        if jac_np is not None:
            self.jac.copy_(torch.from_numpy(np.ascontiguousarray(jac_np.T)))

    def compress_jacobian(self):
        pass

    def compress_hessian(self):
        pass

    def build_kkt(self):
        """Dense/condensed.jl:157-186 as diag-buffer + ONE contraction kernel with fused scaling/epilogue + equality rows; the
        contraction runs on tcgen05 (int8 Ozaki digits, TMA) when self._ozaki is set, else on the fp64 DMMA path."""
        if self._ozaki is not None:
            check(lib.b2d_condensed_assemble_ozaki(self._ozaki.h, self.n, self.m, self.ns, self.n_eq, ptr(self._ind_ineq_d), ptr(self._ind_eq_d),
                                                   ptr(self.hess), ptr(self.jac), ptr(self.pr_diag), ptr(self.du_diag),
                                                   ptr(self.diag_buffer), ptr(self.aug_com), _sp(self.stream)))
            return
        check(lib.b2d_condensed_assemble(self.n, self.m, self.ns, self.n_eq, ptr(self._ind_ineq_d), ptr(self._ind_eq_d),
                                         ptr(self.hess), ptr(self.jac), ptr(self.pr_diag), ptr(self.du_diag),
                                         ptr(self.diag_buffer), ptr(self.aug_com), _sp(self.stream)))

    def tensor_core_status(self):
        """True if the tcgen05 assembly is active and none of its (bounded) pipeline waits ever timed out"""
        if self._ozaki is None:
            return None
        t = C.c_int32(0)
        check(lib.b2d_ozaki_plan_status(self._ozaki.h, C.byref(t), _sp(self.stream)))
        return t.value == 0

    def is_inertia_correct(self, num_pos, num_zero, num_neg):
        """Dense/condensed.jl:189-191."""
        return num_zero == 0 and num_neg == self.n_eq

    def _gemv(self, trans, x, y, alpha, beta):
        """y = alpha*op(jac)*x + beta*y with jac the m x n column-major device matrix (own kernels, no cuBLAS)"""
        fn = lib.b2d_gemv_t if trans else lib.b2d_gemv_n
        check(fn(self.m, self.n, self.m, ptr(self.jac), ptr(x), ptr(y), float(alpha), float(beta), _sp(self.stream)))

    def solve_kkt(self, w: UnreducedKKTVector):
        """src/IPM/factorization.jl:190-229: own kernels around the dense solve (b2d_kkt_solve_pre / _post)."""
        sp = _sp(self.stream)
        check(lib.b2d_kkt_solve_pre(self._dk.h, self._bounds.h, ptr(self.jac), ptr(self.pr_diag), ptr(self.diag_buffer),
                                    ptr(self.l_diag), ptr(self.u_diag), ptr(self.buffer), ptr(self.pd_buffer), ptr(w.values), sp))
        self.linear_solver.solve_linear_system(self.pd_buffer)
        check(lib.b2d_kkt_solve_post(self._dk.h, self._bounds.h, ptr(self.jac), ptr(self.pr_diag), ptr(self.diag_buffer),
                                     ptr(self.l_lower), ptr(self.u_lower), ptr(self.l_diag), ptr(self.u_diag), ptr(self.buffer),
                                     ptr(self.pd_buffer), ptr(w.values), sp))
        return w

    def mul(self, w, x, alpha=1.0, beta=0.0):
        """src/IPM/factorization.jl:303-324 (AbstractDenseKKTSystem): symv + 2 gemv + one fused tail kernel (b2d_kkt_mul)."""
        check(lib.b2d_kkt_mul(self._dk.h, self._bounds.h, ptr(self.hess), ptr(self.jac), ptr(self.reg), ptr(self.du_diag),
                              ptr(self.l_lower), ptr(self.u_lower), ptr(self.l_diag), ptr(self.u_diag), float(alpha), float(beta),
                              ptr(x.values), ptr(w.values), _sp(self.stream)))
        return w

    def jtprod(self, y, x):
        """src/KKT/Dense/utils.jl:12-23: y[1:n] = jac' x ; y[n + k] = -x[ind_ineq[k]] (not on the per-iteration solve path)."""
        self._gemv(True, x, y[: self.n], 1.0, 0.0)
        y[self.n:] = -x[self._ind_ineq_d]


def create_kkt_system(kkt_type, cb, linear_solver=None, opt_linear_solver=None):
    """src/IPM/IPM.jl:157-165 -> create_kkt_system(::Type{K}, cb, linear_solver; opt_linear_solver)."""
    if linear_solver is None:
        linear_solver = B200DenseSolver if kkt_type is DenseCondensedKKTSystem else B200SparseSolver
    return kkt_type(cb, linear_solver, opt_linear_solver)
