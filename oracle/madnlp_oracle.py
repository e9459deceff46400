"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy/scipy restatement of MadNLP's per-iteration KKT hot path (assembly ->
symmetric-indefinite factorisation + inertia -> solve), used ONLY as the checker
in ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  Nothing in the product package (``madnlp.jl_b200/``)
imports this file; the product path fails loudly when the CUDA library is missing.

Every function cites the reference file:line (relative to the MadNLP.jl tree
@ e1028096, v0.10.1) it follows.  Indices are 0-based here (the reference is
1-based); value layouts are identical.

Pinned against (tests/test_oracle_golden.py):
  * the reference's only known-answer vector for factor+solve
    (lib/MadNLPTests/src/MadNLPTests.jl:24-51): 2x2 system, x = [0.85427.., 1.45728..],
    inertia (2,0,0);
  * the reference's self-consistency identity ``K * solve_kkt(K, 1) == 1`` +
    ``is_inertia_correct`` on HS15 for every KKT formulation
    (lib/MadNLPTests/src/MadNLPTests.jl:53-110, test/kkt_test.jl:27-48);
  * the HS15 worked example of SURVEY.md Appendix A.
Beyond these the reference holds no golden numbers for this path (factor/solve
arithmetic lives in LAPACK/UMFPACK binaries outside the tree): step directions on
the larger configurations are "parity unpinned" by the reference and defined by
this oracle (LAPACK ``dsytrf/dsytrs`` through scipy -- the same routine
``LapackCPUSolver`` calls, src/LinearSolvers/lapack.jl:164-172 -- SuperLU
standing in for UMFPACK's unsymmetric LU, src/LinearSolvers/umfpack.jl:28-55 -- and
``LDLSolver``: the reference's src/LinearSolvers/ldl.jl over a C restatement of the
published LDL algorithm that LDLFactorizations.jl translates, oracle/kkt_oracle.c).

C part (oracle/kkt_oracle.c -> oracle/libkkt_oracle.so, built by oracle/Makefile): the
reference's sequential scalar loops (``_transfer!``, ``_build_condensed_aug_coord!``, CSC
mat-vecs) and the sparse LDL^T.  When the library is present the loops below run through
it (same order of additions as the numpy ``add.at`` statements they replace -- checked
bit-for-bit in tests/test_oracle_golden.py); ``USE_C = False`` forces the numpy path.
"""
from __future__ import annotations

import ctypes as _C
import os as _os

import numpy as np
import scipy.linalg.lapack as _lapack
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_CLIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "libkkt_oracle.so")
USE_C = True


def _load_clib():
    if not _os.path.exists(_CLIB_PATH):
        return None
    lib = _C.CDLL(_CLIB_PATH)
    p, i32, i64, f64 = _C.c_void_p, _C.c_int32, _C.c_int64, _C.c_double
    lib.okkt_transfer.argtypes = [i64, p, i64, p, p]
    lib.okkt_condensed_coord.argtypes = [i64, p, p, p, p, p, i64, p, i64, p, i64, p]
    lib.okkt_csc_mul_n.argtypes = [i32, i32, p, p, p, p, p, f64, f64]
    lib.okkt_csc_mul_t.argtypes = [i32, i32, p, p, p, p, p, f64, f64]
    lib.okkt_csc_symv_lower.argtypes = [i32, p, p, p, p, p, f64, f64]
    lib.okkt_ldl_symbolic.restype = p
    lib.okkt_ldl_symbolic.argtypes = [i32, p, p, p]
    lib.okkt_ldl_free.argtypes = [p]
    lib.okkt_ldl_nnz.restype = i64
    lib.okkt_ldl_nnz.argtypes = [p]
    lib.okkt_ldl_numeric.restype = i32
    lib.okkt_ldl_numeric.argtypes = [p, p, p, p]
    lib.okkt_ldl_solve.argtypes = [p, p]
    lib.okkt_ldl_inertia.argtypes = [p, _C.POINTER(i64), _C.POINTER(i64), _C.POINTER(i64)]
    lib.okkt_ldl_D.restype = _C.POINTER(f64)
    lib.okkt_ldl_D.argtypes = [p]
    lib.okkt_set_aug_diagonal.argtypes = [i64, i64, i64] + [p] * 8
    lib.okkt_reduce_rhs.argtypes = [i64, i64] + [p] * 7
    lib.okkt_finish_aug_solve.argtypes = [i64, i64] + [p] * 9
    lib.okkt_kktmul.argtypes = [i64, i64, i64, i64] + [p] * 8 + [f64, f64, p, p]
    for f in (lib.okkt_set_aug_diagonal, lib.okkt_reduce_rhs, lib.okkt_finish_aug_solve, lib.okkt_kktmul):
        f.restype = None
    for f in (lib.okkt_transfer, lib.okkt_condensed_coord, lib.okkt_csc_mul_n, lib.okkt_csc_mul_t, lib.okkt_csc_symv_lower,
              lib.okkt_ldl_free, lib.okkt_ldl_solve, lib.okkt_ldl_inertia):
        f.restype = None
    return lib


clib = _load_clib()


def _c_ok(*arrays):
    return USE_C and clib is not None and all(a.flags.c_contiguous for a in arrays)


def _ptr(a):
    return a.ctypes.data


# --------------------------------------------------------------------------
# matrix tools  (src/matrixtools.jl)
# --------------------------------------------------------------------------
def force_lower_triangular(I, J):
    """src/matrixtools.jl:129-137 -- swap (i,j) so that i >= j, in place."""
    sw = J > I
    tmp = J[sw].copy()
    J[sw] = I[sw]
    I[sw] = tmp


def coo_to_csc(I, J, m, n):
    """src/matrixtools.jl:55-95.

    Returns (colptr, rowval, map) with ``map[k]`` = position in the CSC value
    vector of COO entry k.  Duplicates share a slot (``sparse()`` sums them) and
    rows are sorted inside each column, exactly as Julia's ``sparse(I,J,V)``.
    """
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    key = J * m + I
    ukey, inv = np.unique(key, return_inverse=True)
    rowval = (ukey % m).astype(np.int32)
    cols = ukey // m
    colptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr).astype(np.int32)
    return colptr, rowval, inv.astype(np.int64)


def transfer(nz, V, cmap):
    """src/matrixtools.jl:79-88 -- ``nz .= 0; nz[map[k]] += V[k]`` in COO order."""
    if _c_ok(nz, V, cmap) and nz.dtype == np.float64 and V.dtype == np.float64 and cmap.dtype == np.int64:
        clib.okkt_transfer(len(nz), _ptr(nz), len(V), _ptr(V), _ptr(cmap))
        return nz
    nz[:] = 0.0
    np.add.at(nz, cmap, V)
    return nz


def csc_to_scipy(colptr, rowval, nz, shape):
    return sp.csc_matrix((nz, rowval, colptr), shape=shape)


def tril_to_full(colptr, rowval, nz, n):
    """src/matrixtools.jl:40-46 / umfpack.jl:28-35 -- symmetric expansion."""
    L = csc_to_scipy(colptr, rowval, nz, (n, n))
    return (L + sp.tril(L, -1).T).tocsc()


# --------------------------------------------------------------------------
# callback stand-in: only the fields the KKT constructors read
# (src/Callbacks/nlpmodels.jl:369-406)
# --------------------------------------------------------------------------
class Callback:
    def __init__(self, nvar, ncon, jac_I, jac_J, hess_I, hess_J,
                 ind_ineq, ind_lb, ind_ub):
        self.nvar = int(nvar)
        self.ncon = int(ncon)
        self.jac_I = np.asarray(jac_I, dtype=np.int64)
        self.jac_J = np.asarray(jac_J, dtype=np.int64)
        self.hess_I = np.asarray(hess_I, dtype=np.int64)
        self.hess_J = np.asarray(hess_J, dtype=np.int64)
        self.ind_ineq = np.asarray(ind_ineq, dtype=np.int64)
        all_c = np.arange(self.ncon)
        self.ind_eq = np.setdiff1d(all_c, self.ind_ineq)
        self.ind_lb = np.asarray(ind_lb, dtype=np.int64)
        self.ind_ub = np.asarray(ind_ub, dtype=np.int64)
        self.nnzj = len(self.jac_I)
        self.nnzh = len(self.hess_I)


# --------------------------------------------------------------------------
# UnreducedKKTVector  (src/KKT/rhs.jl:90-129)
# --------------------------------------------------------------------------
class UnreducedKKTVector:
    """values = [x (n_tot) | y (m) | zl (nlb) | zu (nub)], all views of one buffer."""

    def __init__(self, n, m, nlb, nub, ind_lb, ind_ub):
        self.values = np.zeros(n + m + nlb + nub)
        self.n, self.m, self.nlb, self.nub = n, m, nlb, nub
        self.ind_lb, self.ind_ub = ind_lb, ind_ub

    @classmethod
    def for_kkt(cls, kkt):
        return cls(len(kkt.pr_diag), len(kkt.du_diag), len(kkt.l_diag),
                   len(kkt.u_diag), kkt.ind_lb, kkt.ind_ub)

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
        o = UnreducedKKTVector(self.n, self.m, self.nlb, self.nub, self.ind_lb, self.ind_ub)
        o.values[:] = self.values
        return o


# --------------------------------------------------------------------------
# generic KKT pieces  (src/KKT/KKTsystem.jl:210-256, src/IPM/kernels.jl)
# --------------------------------------------------------------------------
def set_aug_diagonal_(kkt):
    """src/IPM/kernels.jl:22-27  (_set_aug_diagonal!)."""
    if _c_ok(kkt.pr_diag, kkt.reg, kkt.ind_lb, kkt.ind_ub):
        clib.okkt_set_aug_diagonal(len(kkt.pr_diag), len(kkt.ind_lb), len(kkt.ind_ub), _ptr(kkt.ind_lb), _ptr(kkt.ind_ub),
                                   _ptr(kkt.reg), _ptr(kkt.l_lower), _ptr(kkt.l_diag), _ptr(kkt.u_lower), _ptr(kkt.u_diag),
                                   _ptr(kkt.pr_diag))
        return
    kkt.pr_diag[:] = kkt.reg
    kkt.pr_diag[kkt.ind_lb] -= kkt.l_lower / kkt.l_diag
    kkt.pr_diag[kkt.ind_ub] -= kkt.u_lower / kkt.u_diag


def regularize_diagonal(kkt, primal, dual):
    """src/KKT/KKTsystem.jl:222-226."""
    kkt.reg += primal
    kkt.pr_diag += primal
    kkt.du_diag -= dual


def reduce_rhs(kkt, d):
    """src/IPM/kernels.jl:182-195."""
    xp = d.primal()
    if _c_ok(d.values, kkt.ind_lb, kkt.ind_ub):
        clib.okkt_reduce_rhs(len(kkt.ind_lb), len(kkt.ind_ub), _ptr(kkt.ind_lb), _ptr(kkt.ind_ub), _ptr(kkt.l_diag),
                             _ptr(kkt.u_diag), _ptr(xp), _ptr(d.dual_lb()), _ptr(d.dual_ub()))
        return
    xp[kkt.ind_lb] -= d.dual_lb() / kkt.l_diag
    xp[kkt.ind_ub] -= d.dual_ub() / kkt.u_diag


def finish_aug_solve(kkt, d):
    """src/IPM/kernels.jl:198-204."""
    xp = d.primal()
    dlb = d.dual_lb()
    dub = d.dual_ub()
    if _c_ok(d.values, kkt.ind_lb, kkt.ind_ub):
        clib.okkt_finish_aug_solve(len(kkt.ind_lb), len(kkt.ind_ub), _ptr(kkt.ind_lb), _ptr(kkt.ind_ub), _ptr(kkt.l_lower),
                                   _ptr(kkt.u_lower), _ptr(kkt.l_diag), _ptr(kkt.u_diag), _ptr(xp), _ptr(dlb), _ptr(dub))
        return
    dlb[:] = (-dlb + kkt.l_lower * xp[kkt.ind_lb]) / kkt.l_diag
    dub[:] = (dub - kkt.u_lower * xp[kkt.ind_ub]) / kkt.u_diag


def kktmul_(w, x, kkt, alpha, beta):
    """src/IPM/kernels.jl:161-180  (_kktmul!)."""
    if _c_ok(w.values, x.values, kkt.ind_lb, kkt.ind_ub, kkt.reg, kkt.du_diag):
        clib.okkt_kktmul(len(kkt.reg), len(kkt.du_diag), len(kkt.ind_lb), len(kkt.ind_ub), _ptr(kkt.ind_lb), _ptr(kkt.ind_ub),
                         _ptr(kkt.reg), _ptr(kkt.du_diag), _ptr(kkt.l_lower), _ptr(kkt.u_lower), _ptr(kkt.l_diag),
                         _ptr(kkt.u_diag), float(alpha), float(beta), _ptr(x.values), _ptr(w.values))
        return
    w.primal()[:] += alpha * kkt.reg * x.primal()
    w.dual()[:] += alpha * kkt.du_diag * x.dual()
    wp = w.primal()
    wp[kkt.ind_lb] -= alpha * x.dual_lb()
    wp[kkt.ind_ub] += alpha * x.dual_ub()
    xp = x.primal()
    w.dual_lb()[:] = beta * w.dual_lb() + alpha * (xp[kkt.ind_lb] * kkt.l_lower - x.dual_lb() * kkt.l_diag)
    w.dual_ub()[:] = beta * w.dual_ub() + alpha * (xp[kkt.ind_ub] * kkt.u_lower + x.dual_ub() * kkt.u_diag)


def is_inertia_correct_default(kkt, num_pos, num_zero, num_neg):
    """src/KKT/KKTsystem.jl:242-244."""
    return num_zero == 0 and num_pos == kkt.num_variables()


# --------------------------------------------------------------------------
# linear solvers (CPU oracles for A9/A10/A11)
# --------------------------------------------------------------------------
def num_neg_ev(n, D, ipiv):
    """src/LinearSolvers/lapack.jl:247-268 -- Bunch-Kaufman inertia from ipiv.

    ``ipiv`` is LAPACK's 1-based pivot vector (negative entries mark 2x2 blocks).
    """
    numneg = 0
    t = 0.0
    for k in range(n):
        d = D[k, k]
        if ipiv[k] < 0:
            if t == 0:
                t = abs(D[k + 1, k])
                d = (d / t) * D[k + 1, k + 1] - t
            else:
                d = t
                t = 0.0
        if d < 0:
            numneg += 1
        if d == 0:
            return -1
    return numneg


class LapackCPUSolver:
    """src/LinearSolvers/lapack.jl + lapack_common.jl, BUNCHKAUFMAN algorithm.

    ``A`` is the dense KKT matrix kept BY REFERENCE (lapack.jl:40); only its lower
    triangle is read (``dsytrf('L')``, lapack.jl:164-167).
    """
    input_type = "dense"

    def __init__(self, A):
        self.A = A
        self.n = A.shape[0]
        self.fact = None
        self.ipiv = None
        self.info = 0

    def factorize(self):
        fact = np.array(self.A, order="F", copy=True)      # lapack_common.jl:28 transfer_matrix!
        ldu, ipiv, info = _lapack.dsytrf(fact, lower=1, overwrite_a=1)
        self.fact, self.ipiv, self.info = ldu, ipiv, info
        return self

    def is_inertia(self):
        return True

    def inertia(self):
        """lapack.jl:240-245; returns (num_pos, num_zero, num_neg)."""
        # scipy returns 0-based ipiv with 2x2 blocks flagged by repeated negative entries
        # in the LAPACK convention shifted; rebuild LAPACK's 1-based signed vector.
        ipiv_l = _to_lapack_ipiv(self.ipiv, self.fact)
        numneg = num_neg_ev(self.n, self.fact, ipiv_l)
        numzero = 1 if self.info > 0 else 0
        numpos = self.n - numneg - numzero
        return (numpos, numzero, numneg)

    def solve(self, x):
        """lapack_common.jl:75-81 + lapack.jl:169-172; in place."""
        sol, info = _lapack.dsytrs(self.fact, self.ipiv, x, lower=1)
        x[:] = sol
        return x

    def improve(self):
        return False

    def introduce(self):
        return "Lapack-CPU (BUNCHKAUFMAN) [scipy dsytrf/dsytrs]"


def _to_lapack_ipiv(ipiv, fact):
    """scipy's f2py wrapper returns LAPACK's ipiv unchanged (1-based, negative = 2x2)."""
    return np.asarray(ipiv)


class UmfpackStandInSolver:
    """Stand-in for src/LinearSolvers/umfpack.jl: expand tril->full (umfpack.jl:28-35),
    unsymmetric sparse LU (here SuperLU; UMFPACK is absent from this image), no inertia
    (umfpack.jl:57-58).  The matrix is kept by reference through (colptr,rowval,nzval).
    """
    input_type = "csc"

    def __init__(self, colptr, rowval, nzval, n):
        self.colptr, self.rowval, self.nzval, self.n = colptr, rowval, nzval, n
        self.lu = None

    def factorize(self):
        full = tril_to_full(self.colptr, self.rowval, self.nzval, self.n)
        try:
            # UMFPACK picks its "symmetric strategy" for such matrices: AMD on A+A' and diagonal pivots preferred with
            # sym_pivot_tolerance = 0.001; SuperLU's closest setting is MMD(A'+A) + SymmetricMode + diag threshold 1e-3.
            self.lu = spla.splu(full, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=1e-3,
                                options=dict(SymmetricMode=True))
        except RuntimeError:            # singular: soft failure (umfpack.jl:41-44)
            self.lu = None
        return self

    def is_inertia(self):
        return False

    def inertia(self):
        raise RuntimeError("InertiaException")     # umfpack.jl:57-58

    def solve(self, x):
        if self.lu is None:                        # umfpack.jl:46-55: leave rhs unchanged
            return x
        x[:] = self.lu.solve(x)
        return x

    def improve(self):
        return False

    def introduce(self):
        return "umfpack stand-in (SuperLU)"


class LDLSolver:
    """src/LinearSolvers/ldl.jl:5-62 (LDLSolver over LDLFactorizations.jl): tril -> full (ldl.jl:21), symbolic analysis at
    construction, ``factorize`` = ``full.nzval .= tril_to_full_view; ldl_factorize!`` (ldl.jl:29-33), in-place ``ldiv!``
    (ldl.jl:35-40), inertia = signs of D (ldl.jl:43-58).  The numeric kernel is oracle/kkt_oracle.c (Davis' LDL, the
    algorithm LDLFactorizations.jl translates).  Ordering: the reference calls AMD.jl; AMD is not in this image, so the
    minimum-degree ordering SuperLU computes on the same pattern (MMD on A'+A) stands in -- the ordering changes fill and
    rounding, not the mathematics.  Sequential, like the reference."""
    input_type = "csc"

    def __init__(self, colptr, rowval, nzval, n, perm=None):
        if clib is None:
            raise RuntimeError("oracle/libkkt_oracle.so missing: run `make -C oracle`")
        self.colptr, self.rowval, self.nzval, self.n = colptr, rowval, nzval, n
        # get_tril_to_full (src/matrixtools.jl:17-46): full pattern + a view that maps tril values into it
        nnz = len(rowval)
        cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
        rows = np.asarray(rowval, dtype=np.int64)
        off = rows != cols
        I = np.concatenate([rows, cols[off]]); J = np.concatenate([cols, rows[off]])
        src = np.concatenate([np.arange(nnz), np.arange(nnz)[off]])
        order = np.lexsort((I, J))
        self._full_rowval = np.ascontiguousarray(I[order], dtype=np.int32)
        cp = np.zeros(n + 1, dtype=np.int64); np.add.at(cp, J + 1, 1)
        self._full_colptr = np.ascontiguousarray(np.cumsum(cp), dtype=np.int32)
        self._tril_to_full = src[order]
        self._full_nz = np.zeros(len(order))
        if perm is None:
            pat = sp.csc_matrix((np.ones(len(order)), self._full_rowval, self._full_colptr), shape=(n, n))
            pat = pat + sp.diags(np.asarray(abs(pat).sum(axis=0)).ravel() + 1.0)       # values irrelevant: ordering only
            perm = spla.splu(pat.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                             options=dict(SymmetricMode=True)).perm_c
            # SuperLU: column perm_c[j] of A becomes column j... perm_c maps old -> new position; we need new -> old
            perm = np.argsort(perm)
        self.perm = np.ascontiguousarray(perm, dtype=np.int32)
        self._h = clib.okkt_ldl_symbolic(n, _ptr(self._full_colptr), _ptr(self._full_rowval), _ptr(self.perm))
        self.nnz_l = int(clib.okkt_ldl_nnz(self._h))
        self.ok = False

    def __del__(self):
        if getattr(self, "_h", None) and clib is not None:
            clib.okkt_ldl_free(self._h)
            self._h = None

    def factorize(self):
        np.take(self.nzval, self._tril_to_full, out=self._full_nz)               # ldl.jl:30
        r = clib.okkt_ldl_numeric(self._h, _ptr(self._full_colptr), _ptr(self._full_rowval), _ptr(self._full_nz))
        self.ok = (r == self.n)
        return self

    def is_inertia(self):
        return True

    def inertia(self):
        a, b, c = _C.c_int64(), _C.c_int64(), _C.c_int64()
        clib.okkt_ldl_inertia(self._h, _C.byref(a), _C.byref(b), _C.byref(c))
        if not self.ok:                          # stopped at an exactly-zero pivot: the remaining D entries are stale
            return (0, self.n, 0)
        return (a.value, b.value, c.value)

    def solve(self, x):
        if not self.ok:                          # ldl.jl:37-39: failed factorisation leaves the rhs unchanged
            return x
        assert x.flags.c_contiguous and x.dtype == np.float64
        clib.okkt_ldl_solve(self._h, _ptr(x))
        return x

    def improve(self):
        return False

    def introduce(self):
        return "LDL (Davis Alg. 849, C restatement of LDLFactorizations.jl)"


def csc_mul_n(colptr, rowval, nz, shape, x, y, alpha, beta):
    """y = alpha*A*x + beta*y for a CSC matrix (SparseArrays mul!)."""
    if _c_ok(nz, x, y):
        clib.okkt_csc_mul_n(shape[0], shape[1], _ptr(colptr), _ptr(rowval), _ptr(nz), _ptr(x), _ptr(y), alpha, beta)
    else:
        y[:] = alpha * (csc_to_scipy(colptr, rowval, nz, shape) @ x) + (beta * y if beta != 0.0 else 0.0)
    return y


def csc_mul_t(colptr, rowval, nz, shape, x, y, alpha, beta):
    """y = alpha*A'*x + beta*y."""
    if _c_ok(nz, x, y):
        clib.okkt_csc_mul_t(shape[0], shape[1], _ptr(colptr), _ptr(rowval), _ptr(nz), _ptr(x), _ptr(y), alpha, beta)
    else:
        y[:] = alpha * (csc_to_scipy(colptr, rowval, nz, shape).T @ x) + (beta * y if beta != 0.0 else 0.0)
    return y


def csc_symv_lower(colptr, rowval, nz, n, x, y, alpha, beta):
    """y = alpha*Symmetric(A,:L)*x + beta*y."""
    if _c_ok(nz, x, y):
        clib.okkt_csc_symv_lower(n, _ptr(colptr), _ptr(rowval), _ptr(nz), _ptr(x), _ptr(y), alpha, beta)
    else:
        H = csc_to_scipy(colptr, rowval, nz, (n, n))
        y[:] = alpha * ((H + sp.tril(H, -1).T) @ x) + (beta * y if beta != 0.0 else 0.0)
    return y


class DenseLDLInertiaSolver:
    """Inertia *truth* for sparse matrices small enough to densify: eigenvalue signs.
    Used to pin the product's inertia triple (A10); factor/solve through dsytrf."""
    input_type = "csc"

    def __init__(self, colptr, rowval, nzval, n):
        self.colptr, self.rowval, self.nzval, self.n = colptr, rowval, nzval, n
        self.inner = None

    def factorize(self):
        full = tril_to_full(self.colptr, self.rowval, self.nzval, self.n).toarray()
        self.dense = full
        self.inner = LapackCPUSolver(full).factorize()
        return self

    def is_inertia(self):
        return True

    def inertia(self):
        return self.inner.inertia()

    def solve(self, x):
        return self.inner.solve(x)

    def improve(self):
        return False


# --------------------------------------------------------------------------
# SparseKKTSystem  (src/KKT/Sparse/augmented.jl)
# --------------------------------------------------------------------------
class SparseKKTSystem:
    """Augmented (reduced) KKT in COO -> lower CSC; value vector
    V = [pr_diag(n_tot) | hess(nnzh) | jac(nnzj) | slack -1 (ns) | du_diag(m)]
    with aliasing views (augmented.jl:77-107)."""

    def __init__(self, cb: Callback, linear_solver=UmfpackStandInSolver):
        n, m = cb.nvar, cb.ncon
        ns = len(cb.ind_ineq)
        hI, hJ = cb.hess_I.copy(), cb.hess_J.copy()
        force_lower_triangular(hI, hJ)                      # augmented.jl:65
        n_jac, n_hess = cb.nnzj, len(hI)
        n_tot = n + ns
        self.n, self.m, self.ns, self.n_tot = n, m, ns, n_tot
        L = n_tot + m + n_hess + n_jac + ns                 # augmented.jl:75
        I = np.zeros(L, dtype=np.int64)
        J = np.zeros(L, dtype=np.int64)
        self.V = np.zeros(L)
        o1 = n_tot
        o2 = o1 + n_hess
        o3 = o2 + n_jac
        o4 = o3 + ns
        I[:o1] = np.arange(n_tot); J[:o1] = np.arange(n_tot)
        I[o1:o2] = hI; J[o1:o2] = hJ
        I[o2:o3] = cb.jac_I + n_tot; J[o2:o3] = cb.jac_J
        I[o3:o4] = cb.ind_ineq + n_tot; J[o3:o4] = np.arange(n, n + ns)
        I[o4:] = np.arange(n_tot, n_tot + m); J[o4:] = np.arange(n_tot, n_tot + m)
        self.aug_I, self.aug_J = I, J
        self.pr_diag = self.V[:o1]
        self.hess = self.V[o1:o2]
        self.jac = self.V[o2:o4]            # callback part + slack part
        self.jac_callback = self.V[o2:o3]
        self.du_diag = self.V[o4:]
        nlb, nub = len(cb.ind_lb), len(cb.ind_ub)
        self.reg = np.zeros(n_tot)
        self.l_diag = np.zeros(nlb); self.u_diag = np.zeros(nub)
        self.l_lower = np.zeros(nlb); self.u_lower = np.zeros(nub)
        self.ind_ineq, self.ind_lb, self.ind_ub = cb.ind_ineq, cb.ind_lb, cb.ind_ub
        N = n_tot + m
        self.N = N
        self.aug_colptr, self.aug_rowval, self.aug_csc_map = coo_to_csc(I, J, N, N)
        self.aug_nz = np.zeros(len(self.aug_rowval))
        # jac_raw (m x n_tot) and hess_raw (n_tot x n_tot)   augmented.jl:109-122
        self.jac_I = np.concatenate([cb.jac_I, cb.ind_ineq])
        self.jac_J = np.concatenate([cb.jac_J, np.arange(n, n + ns)])
        self.jac_colptr, self.jac_rowval, self.jac_csc_map = coo_to_csc(self.jac_I, self.jac_J, m, n_tot)
        self.jac_nz = np.zeros(len(self.jac_rowval))
        self.hess_colptr, self.hess_rowval, self.hess_csc_map = coo_to_csc(hI, hJ, n_tot, n_tot)
        self.hess_nz = np.zeros(len(self.hess_rowval))
        self.linear_solver = linear_solver(self.aug_colptr, self.aug_rowval, self.aug_nz, N)

    def num_variables(self):
        return len(self.pr_diag)

    def initialize(self):
        """Sparse/utils.jl:52-62."""
        self.reg[:] = 1.0; self.pr_diag[:] = 1.0; self.du_diag[:] = 0.0
        self.hess[:] = 0.0; self.l_lower[:] = 0.0; self.u_lower[:] = 0.0
        self.l_diag[:] = 1.0; self.u_diag[:] = 1.0; self.hess_nz[:] = 0.0

    def get_jacobian(self):
        return self.jac_callback

    def get_hessian(self):
        return self.hess

    def compress_jacobian(self):
        """Sparse/utils.jl:36-40."""
        if self.ns:
            self.jac[-self.ns:] = -1.0
        transfer(self.jac_nz, self.jac, self.jac_csc_map)

    def compress_hessian(self):
        """Sparse/utils.jl:48-50."""
        transfer(self.hess_nz, self.hess, self.hess_csc_map)

    def build_kkt(self):
        """augmented.jl:146-148."""
        transfer(self.aug_nz, self.V, self.aug_csc_map)

    def is_inertia_correct(self, p, z, n):
        return is_inertia_correct_default(self, p, z, n)

    def jac_com(self):
        return csc_to_scipy(self.jac_colptr, self.jac_rowval, self.jac_nz, (self.m, self.n_tot))

    def hess_com(self):
        return csc_to_scipy(self.hess_colptr, self.hess_rowval, self.hess_nz, (self.n_tot, self.n_tot))

    def solve_kkt(self, w: UnreducedKKTVector):
        """src/IPM/factorization.jl:41-46."""
        reduce_rhs(self, w)
        self.linear_solver.solve(w.primal_dual())
        finish_aug_solve(self, w)
        return w

    def mul(self, w, x, alpha=1.0, beta=0.0):
        """src/IPM/factorization.jl:231-237."""
        H = self.hess_com()
        Hs = H + sp.tril(H, -1).T
        Jc = self.jac_com()
        w.primal()[:] = alpha * (Hs @ x.primal()) + beta * w.primal()
        w.primal()[:] += alpha * (Jc.T @ x.dual())
        w.dual()[:] = alpha * (Jc @ x.primal()) + beta * w.dual()
        kktmul_(w, x, self, alpha, beta)
        return w


# --------------------------------------------------------------------------
# SparseCondensedKKTSystem  (src/KKT/Sparse/condensed.jl)
# --------------------------------------------------------------------------
def build_condensed_aug_symbolic(H_colptr, H_rowval, n, Jt_colptr, Jt_rowval, m):
    """src/KKT/Sparse/condensed.jl:201-301.

    Pattern of tril(H) U diag U tril(Jt Jt') as lower CSC, plus
      dptr[(dst, src)]            -- pr_diag[src] -> nz[dst]
      hptr[(dst, src)]            -- H.nz[src]    -> nz[dst]
      jptr[(dst, (col, k, l))]    -- D[col]*Jt.nz[k]*Jt.nz[l] -> nz[dst]
    The reference sorts the (row,col) key list with ``sortperm`` (stable) -- the order
    of sources inside one destination slot is therefore: diag, hess (by index), then
    Jt triples in (col, j, k) enumeration order.  We reproduce that order exactly since
    it fixes the floating-point summation order of _build_condensed_aug_coord!.
    """
    nnzH = len(H_rowval)
    # counts per Jt column  (condensed.jl:158-165)
    cnts = np.diff(Jt_colptr).astype(np.int64)
    nnzjtsj = int(np.sum(cnts * (cnts + 1) // 2))
    tot = n + nnzH + nnzjtsj
    kind = np.empty(tot, dtype=np.int64)     # -1 diag, 0 hess, >0: Jt column + 1
    s1 = np.empty(tot, dtype=np.int64)
    s2 = np.empty(tot, dtype=np.int64)
    row = np.empty(tot, dtype=np.int64)
    col = np.empty(tot, dtype=np.int64)
    # diag entries (condensed.jl:231-240)
    kind[:n] = -1; s1[:n] = np.arange(n); s2[:n] = 0
    row[:n] = np.arange(n); col[:n] = np.arange(n)
    # hess entries (condensed.jl:167-175)
    hcols = np.repeat(np.arange(n), np.diff(H_colptr))
    kind[n:n + nnzH] = 0; s1[n:n + nnzH] = np.arange(nnzH); s2[n:n + nnzH] = 0
    row[n:n + nnzH] = H_rowval; col[n:n + nnzH] = hcols
    # Jt pairs (condensed.jl:177-190): for column i, for j in col range, for k>=j
    p = n + nnzH
    for i in range(m):
        a, b = int(Jt_colptr[i]), int(Jt_colptr[i + 1])
        c = b - a
        if c == 0:
            continue
        jj, kk = np.triu_indices(c)          # j<=k, row-major == reference loop order
        cnt = len(jj)
        kind[p:p + cnt] = i + 1
        s1[p:p + cnt] = a + jj
        s2[p:p + cnt] = a + kk
        col[p:p + cnt] = Jt_rowval[a + jj]    # c1
        row[p:p + cnt] = Jt_rowval[a + kk]    # c2
        p += cnt
    assert p == tot
    # sort by (col,row), stable  (condensed.jl:251)
    order = np.lexsort((row, col))            # lexsort is stable; last key primary
    kind, s1, s2, row, col = kind[order], s1[order], s2[order], row[order], col[order]
    newslot = np.ones(tot, dtype=bool)
    newslot[1:] = (row[1:] != row[:-1]) | (col[1:] != col[:-1])
    guide = np.cumsum(newslot) - 1            # 0-based slot id
    rowval = row[newslot].astype(np.int32)
    ucol = col[newslot]
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, ucol + 1, 1)
    colptr = np.cumsum(colptr).astype(np.int32)
    b = kind == -1
    dptr = np.stack([guide[b], s1[b]], axis=1)
    b = kind == 0
    hptr = np.stack([guide[b], s1[b]], axis=1)
    b = kind > 0
    jptr = np.stack([guide[b], kind[b] - 1, s1[b], s2[b]], axis=1)
    return colptr, rowval, dptr, hptr, jptr


def build_condensed_aug_coord(nz, pr_diag, H_nz, Jt_nz, diag_buffer, dptr, hptr, jptr):
    """src/KKT/Sparse/condensed.jl:328-345 -- same three accumulation passes in the
    same order (hess, diag, JtDJ), sequential adds inside each pass."""
    if _c_ok(nz, pr_diag, H_nz, Jt_nz, diag_buffer, dptr, hptr, jptr) and dptr.dtype == hptr.dtype == jptr.dtype == np.int64:
        clib.okkt_condensed_coord(len(nz), _ptr(nz), _ptr(pr_diag), _ptr(H_nz), _ptr(Jt_nz), _ptr(diag_buffer),
                                  len(dptr), _ptr(dptr), len(hptr), _ptr(hptr), len(jptr), _ptr(jptr))
        return nz
    nz[:] = 0.0
    np.add.at(nz, hptr[:, 0], H_nz[hptr[:, 1]])
    np.add.at(nz, dptr[:, 0], pr_diag[dptr[:, 1]])
    np.add.at(nz, jptr[:, 0], diag_buffer[jptr[:, 1]] * Jt_nz[jptr[:, 2]] * Jt_nz[jptr[:, 3]])
    return nz


class SparseCondensedKKTSystem:
    """src/KKT/Sparse/condensed.jl:8-153, 354-366."""

    def __init__(self, cb: Callback, linear_solver=None):
        n, m = cb.nvar, cb.ncon
        ns = len(cb.ind_ineq)
        if ns != m:
            raise ValueError("SparseCondensedKKTSystem does not support equality constrained NLPs.")  # condensed.jl:68-70
        hI, hJ = cb.hess_I.copy(), cb.hess_J.copy()
        force_lower_triangular(hI, hJ)
        self.n, self.m, self.ns, self.n_tot = n, m, ns, n + ns
        nlb, nub = len(cb.ind_lb), len(cb.ind_ub)
        self.reg = np.zeros(n + ns); self.pr_diag = np.zeros(n + ns); self.du_diag = np.zeros(m)
        self.l_diag = np.zeros(nlb); self.u_diag = np.zeros(nub)
        self.l_lower = np.zeros(nlb); self.u_lower = np.zeros(nub)
        self.buffer = np.zeros(m); self.buffer2 = np.zeros(m); self.diag_buffer = np.zeros(m)
        self.hess = np.zeros(len(hI)); self.jac = np.zeros(cb.nnzj)
        self.ind_ineq, self.ind_lb, self.ind_ub = cb.ind_ineq, cb.ind_lb, cb.ind_ub
        # jt_coo = (n x m) with I=jac_J, J=jac_I  (condensed.jl:105-110)
        self.jt_colptr, self.jt_rowval, self.jt_csc_map = coo_to_csc(cb.jac_J, cb.jac_I, n, m)
        self.jt_nz = np.zeros(len(self.jt_rowval))
        self.hess_colptr, self.hess_rowval, self.hess_csc_map = coo_to_csc(hI, hJ, n, n)
        self.hess_nz = np.zeros(len(self.hess_rowval))
        (self.aug_colptr, self.aug_rowval, self.dptr, self.hptr, self.jptr) = build_condensed_aug_symbolic(
            self.hess_colptr, self.hess_rowval, n, self.jt_colptr, self.jt_rowval, m)
        self.aug_nz = np.zeros(len(self.aug_rowval))
        self.N = n
        if linear_solver is None:
            linear_solver = DenseLDLInertiaSolver
        self.linear_solver = linear_solver(self.aug_colptr, self.aug_rowval, self.aug_nz, n)

    def num_variables(self):
        return len(self.pr_diag)

    def initialize(self):
        self.reg[:] = 1.0; self.pr_diag[:] = 1.0; self.du_diag[:] = 0.0
        self.hess[:] = 0.0; self.l_lower[:] = 0.0; self.u_lower[:] = 0.0
        self.l_diag[:] = 1.0; self.u_diag[:] = 1.0; self.hess_nz[:] = 0.0

    def get_jacobian(self):
        return self.jac

    def get_hessian(self):
        return self.hess

    def compress_jacobian(self):
        """condensed.jl:145-148."""
        transfer(self.jt_nz, self.jac, self.jt_csc_map)

    def compress_hessian(self):
        transfer(self.hess_nz, self.hess, self.hess_csc_map)

    def build_kkt(self):
        """condensed.jl:354-366."""
        n, m = self.n, self.m
        Ss = self.pr_diag[n:n + m]
        Sd = self.du_diag
        self.diag_buffer[:] = Ss / (1.0 - Sd * Ss)
        build_condensed_aug_coord(self.aug_nz, self.pr_diag, self.hess_nz, self.jt_nz,
                                  self.diag_buffer, self.dptr, self.hptr, self.jptr)

    def is_inertia_correct(self, p, z, ng):
        """condensed.jl:138-140."""
        return z == 0 and p == self.n

    def should_regularize_dual(self, p, z, ng):
        """condensed.jl:141."""
        return True

    def jt_csc(self):
        return csc_to_scipy(self.jt_colptr, self.jt_rowval, self.jt_nz, (self.n, self.m))

    def hess_com(self):
        return csc_to_scipy(self.hess_colptr, self.hess_rowval, self.hess_nz, (self.n, self.n))

    def solve_kkt(self, w: UnreducedKKTVector):
        """src/IPM/factorization.jl:143-167."""
        n, m = self.n, self.m
        full = w.full()
        wx = full[:n]; ws = full[n:n + m]; wz = full[n + m:n + 2 * m]
        Ss = self.pr_diag[n:n + m]
        reduce_rhs(self, w)
        self.buffer[:] = self.diag_buffer * (wz + ws / Ss)
        jt = (self.jt_colptr, self.jt_rowval, self.jt_nz, (n, m))
        csc_mul_n(*jt, self.buffer, wx, 1.0, 1.0)                   # mul!(wx, jt_csc, buffer, 1, 1)
        self.linear_solver.solve(wx)
        csc_mul_t(*jt, wx, self.buffer2, 1.0, 0.0)                  # mul!(buffer2, jt_csc', wx)
        wz[:] = -self.buffer + self.diag_buffer * self.buffer2
        ws[:] = (ws + wz) / Ss
        finish_aug_solve(self, w)
        return w

    def mul(self, w, x, alpha=1.0, beta=0.0):
        """src/IPM/factorization.jl:303-324."""
        n, m = self.n, self.m
        xf, wf = x.full(), w.full()
        xx = xf[:n]; xs = xf[n:n + m]; xz = xf[n + m:n + 2 * m]
        wx = wf[:n]; ws = wf[n:n + m]; wz = wf[n + m:n + 2 * m]
        jt = (self.jt_colptr, self.jt_rowval, self.jt_nz, (n, m))
        csc_symv_lower(self.hess_colptr, self.hess_rowval, self.hess_nz, n, xx, wx, alpha, beta)   # Symmetric(hess_com,:L)
        csc_mul_n(*jt, xz, wx, alpha, 1.0)
        csc_mul_t(*jt, xx, wz, alpha, beta)
        wz[:] -= alpha * xs
        ws[:] = beta * ws - alpha * xz
        kktmul_(w, x, self, alpha, beta)
        return w


# --------------------------------------------------------------------------
# DenseCondensedKKTSystem  (src/KKT/Dense/condensed.jl)
# --------------------------------------------------------------------------
class DenseCondensedKKTSystem:
    """Dense/condensed.jl:10-191.  hess: n x n (both triangles), jac: m x n."""

    def __init__(self, cb: Callback, linear_solver=LapackCPUSolver):
        n, m = cb.nvar, cb.ncon
        ns = len(cb.ind_ineq)
        n_eq = m - ns
        self.n, self.m, self.ns, self.n_eq = n, m, ns, n_eq
        nlb, nub = len(cb.ind_lb), len(cb.ind_ub)
        N = n + n_eq
        self.N = N
        self.aug_com = np.zeros((N, N), order="F")
        self.hess = np.zeros((n, n), order="F")
        self.jac = np.zeros((m, n), order="F")
        self.jac_ineq = np.zeros((ns, n), order="F")
        self.reg = np.zeros(n + ns); self.pr_diag = np.zeros(n + ns); self.du_diag = np.zeros(m)
        self.l_diag = np.ones(nlb); self.u_diag = np.ones(nub)
        self.l_lower = np.zeros(nlb); self.u_lower = np.zeros(nub)
        self.pd_buffer = np.zeros(N); self.diag_buffer = np.zeros(ns); self.buffer = np.zeros(m)
        self.ind_eq, self.ind_ineq = cb.ind_eq, cb.ind_ineq
        self.ind_lb, self.ind_ub = cb.ind_lb, cb.ind_ub
        self.linear_solver = linear_solver(self.aug_com)

    def num_variables(self):
        return self.n

    def initialize(self):
        """KKTsystem.jl:210-216."""
        self.reg[:] = 1.0; self.pr_diag[:] = 1.0; self.du_diag[:] = 0.0; self.hess[:] = 0.0

    def get_jacobian(self):
        return self.jac

    def get_hessian(self):
        return self.hess

    def compress_jacobian(self):
        pass

    def compress_hessian(self):
        pass

    def build_kkt(self):
        """Dense/condensed.jl:157-186 (+ helpers :120-155)."""
        n, ns, n_eq = self.n, self.ns, self.n_eq
        self.aug_com[:] = 0.0
        Ss = self.pr_diag[n:n + ns]
        Sd = self.du_diag[self.ind_ineq]
        self.diag_buffer[:] = Ss / (1.0 - Sd * Ss)
        self.jac_ineq[:] = self.jac[self.ind_ineq, :] * np.sqrt(self.diag_buffer)[:, None]
        W = self.jac_ineq.T @ self.jac_ineq
        self.aug_com[:n, :n] = W
        # _build_condensed_kkt_system!  (:120-143)
        self.aug_com[:n, :n] += self.hess
        self.aug_com[np.arange(n), np.arange(n)] += self.pr_diag[:n]
        if n_eq:
            Je = self.jac[self.ind_eq, :]
            self.aug_com[n:, :n] = Je
            self.aug_com[:n, n:] = Je.T
            self.aug_com[n + np.arange(n_eq), n + np.arange(n_eq)] = self.du_diag[self.ind_eq]

    def is_inertia_correct(self, p, z, ng):
        """Dense/condensed.jl:189-191."""
        return z == 0 and ng == self.n_eq

    def solve_kkt(self, w: UnreducedKKTVector):
        """src/IPM/factorization.jl:190-229."""
        n, ns, n_eq, m = self.n, self.ns, self.n_eq, self.m
        full = w.full()
        wx = full[:n]; ws = full[n:n + ns]
        i_eq = self.ind_eq + n + ns
        i_in = self.ind_ineq + n + ns
        x = self.pd_buffer
        Ss = self.pr_diag[n:n + ns]
        reduce_rhs(self, w)
        self.buffer[:] = 0.0
        self.buffer[self.ind_ineq] = self.diag_buffer * (full[i_in] + ws / Ss)
        x[:n] = self.jac.T @ self.buffer
        x[:n] += wx
        x[n:] = full[i_eq]
        self.linear_solver.solve(x)
        wx[:] = x[:n]
        dual = w.dual()
        wz_old = full[i_in].copy()
        dual[:] = self.jac @ wx
        full[i_eq] = x[n:]
        # wz .*= diag_buffer  acts on the freshly written J*wx entries of the ineq rows
        full[i_in] *= self.diag_buffer
        dual[:] -= self.buffer
        ws[:] = (ws + full[i_in]) / Ss
        finish_aug_solve(self, w)
        return w

    def mul(self, w, x, alpha=1.0, beta=0.0):
        """src/IPM/factorization.jl:326-344 (AbstractDenseKKTSystem)."""
        n, ns, m = self.n, self.ns, self.m
        wp, xp = w.primal(), x.primal()
        wx = wp[:n]; ws = wp[n:]
        xx = xp[:n]; xs = xp[n:]
        wy, xy = w.dual(), x.dual()
        Hs = np.tril(self.hess) + np.tril(self.hess, -1).T      # _symv!('L', ...)
        wx[:] = alpha * (Hs @ xx) + beta * wx
        if m > 0:
            wx[:] += alpha * (self.jac.T @ xy)
            wy[:] = alpha * (self.jac @ xx) + beta * wy
        ws[:] = beta * ws - alpha * xy[self.ind_ineq]
        wy[self.ind_ineq] -= alpha * xs
        kktmul_(w, x, self, alpha, beta)
        return w


# --------------------------------------------------------------------------
# Richardson refinement  (src/LinearSolvers/backsolve.jl:27-76)
# --------------------------------------------------------------------------
def solve_refine(x, kkt, b, w, tol=1e-8, max_iter=10):
    """Returns (ok, n_iter, residual_ratio); richardson_tol = tol^(5/4),
    acceptable = tol^(5/8) (backsolve.jl:25)."""
    r_tol = tol ** (5 / 4)
    r_acc = tol ** (5 / 8)
    norm_b = np.linalg.norm(b.full(), np.inf)
    ratio = 0.0
    x.full()[:] = 0.0
    it = 0
    if norm_b != 0:
        w.full()[:] = b.full()
        while True:
            kkt.solve_kkt(w)
            x.full()[:] += w.full()
            w.full()[:] = b.full()
            kkt.mul(w, x, -1.0, 1.0)
            norm_w = np.linalg.norm(w.full(), np.inf)
            norm_x = np.linalg.norm(x.full(), np.inf)
            ratio = norm_w / (min(norm_x, 1e6 * norm_b) + norm_b)
            it += 1
            if it >= max_iter or ratio < r_tol:
                break
    return ratio < r_acc, it, ratio


# --------------------------------------------------------------------------
# fixtures restated from lib/MadNLPTests
# --------------------------------------------------------------------------
class HS15Model:
    """lib/MadNLPTests/src/Instances/hs15.jl:1-104."""
    nvar, ncon = 2, 2
    x0 = np.zeros(2)
    y0 = np.zeros(2)
    lvar = np.array([-np.inf, -np.inf]); uvar = np.array([0.5, np.inf])
    lcon = np.array([1.0, 0.0]); ucon = np.array([np.inf, np.inf])
    jac_I = np.array([0, 0, 1, 1]); jac_J = np.array([0, 1, 0, 1])
    hess_I = np.array([0, 1, 1]); hess_J = np.array([0, 0, 1])

    @staticmethod
    def jac_coord(x):
        return np.array([x[1], x[0], 1.0, 2 * x[1]])

    @staticmethod
    def hess_coord(x, y, obj_weight=1.0):
        H = np.array([obj_weight * (-400.0 * x[1] + 1200.0 * x[0] ** 2 + 2.0),
                      obj_weight * (-400.0 * x[0]),
                      obj_weight * 200.0])
        H[1] += y[0] * 1.0
        H[2] += y[1] * 2.0
        return H

    @staticmethod
    def jac_dense(x):
        return np.array([[x[1], x[0]], [1.0, 2 * x[1]]])

    @staticmethod
    def hess_dense(x, y, obj_weight=1.0):
        h = HS15Model.hess_coord(x, y, obj_weight)
        return np.array([[h[0], h[1]], [h[1], h[2]]])

    @classmethod
    def callback(cls):
        """Index sets as create_callback derives them (nlpmodels.jl:369-406): both
        constraints are inequalities -> 2 slacks with bounds lcon/ucon."""
        ind_ineq = np.array([0, 1])
        xl = np.concatenate([cls.lvar, cls.lcon[ind_ineq]])
        xu = np.concatenate([cls.uvar, cls.ucon[ind_ineq]])
        ind_lb = np.where(np.isfinite(xl))[0]
        ind_ub = np.where(np.isfinite(xu))[0]
        return Callback(2, 2, cls.jac_I, cls.jac_J, cls.hess_I, cls.hess_J, ind_ineq, ind_lb, ind_ub)


def test_kkt_system(kkt, model, dense=False):
    """lib/MadNLPTests/src/MadNLPTests.jl:53-110 restated; returns (x, y, inertia)."""
    kkt.initialize()
    x0, y0 = model.x0, model.y0
    if dense:
        kkt.get_jacobian()[:] = model.jac_dense(x0)
        kkt.get_hessian()[:] = model.hess_dense(x0, y0)
    else:
        kkt.get_jacobian()[:] = model.jac_coord(x0)
        kkt.get_hessian()[:] = model.hess_coord(x0, y0)
    kkt.compress_jacobian()
    kkt.compress_hessian()
    kkt.l_lower[:] = 1e-3
    kkt.u_lower[:] = 1e-3
    set_aug_diagonal_(kkt)
    kkt.build_kkt()
    kkt.linear_solver.factorize()
    x = UnreducedKKTVector.for_kkt(kkt)
    x.full()[:] = 1.0
    kkt.solve_kkt(x)
    y = x.copy()
    y.full()[:] = 0.0
    kkt.mul(y, x)
    inertia = kkt.linear_solver.inertia() if kkt.linear_solver.is_inertia() else None
    return x, y, inertia


# ------------------------------------------------------------------------------------------------------------------
# IPM reductions and set_aug_rhs!: scalar restatements of src/IPM/kernels.jl, used to check the single-pass
# device reductions.  Arguments as in the reference: x, xl, xu, f, zl, zu, jacl, dx are full primal vectors (+-inf for
# absent bounds); ind_lb / ind_ub are the 0-based index sets of the finite bounds; dzl, dzu, l are compressed.
# ------------------------------------------------------------------------------------------------------------------
def _jl_min(*vals):
    """Julia's min: NaN-propagating"""
    out = np.inf
    for v in vals:
        if np.isnan(v) or np.isnan(out):
            out = np.nan
        else:
            out = min(out, v)
    return out


def get_alpha_max(x, xl, xu, dx, tau):
    """src/IPM/kernels.jl:356-371"""
    alpha = 1.0
    for i in range(len(x)):
        a = (-x[i] + xl[i]) * tau / dx[i] if dx[i] < 0 else np.inf
        c = (-x[i] + xu[i]) * tau / dx[i] if dx[i] > 0 else np.inf
        alpha = _jl_min(alpha, a, c)
    return alpha


def get_alpha_z(zl_r, zu_r, dzl, dzu, tau):
    """src/IPM/kernels.jl:373-388"""
    alpha = 1.0
    for i in range(len(zl_r)):
        alpha = _jl_min(alpha, (-zl_r[i]) * tau / dzl[i] if dzl[i] < 0 else np.inf)
    for i in range(len(zu_r)):
        alpha = _jl_min(alpha, (-zu_r[i]) * tau / dzu[i] if dzu[i] < 0 else np.inf)
    return alpha


def get_varphi(obj_val, x_lr, xl_r, xu_r, x_ur, mu):
    """src/IPM/kernels.jl:263-283"""
    def one(a, b):
        d = a - b
        return np.inf if d < 0 else (-mu * np.log(d) if d > 0 else np.inf)
    v = obj_val
    for a, b in zip(x_lr, xl_r):
        v += one(a, b)
    for a, b in zip(xu_r, x_ur):
        v += one(a, b)
    return v


def get_varphi_d(f, x, xl, xu, dx, mu):
    """src/IPM/kernels.jl:341-354"""
    return float(np.sum((f - mu / (x - xl) + mu / (xu - x)) * dx))


def get_inf_du(f, zl, zu, jacl, sd):
    """src/IPM/kernels.jl:285-291"""
    return float(np.max(np.abs(f - zl + zu + jacl), initial=0.0)) / sd


def get_inf_compl(x_lr, xl_r, zl_r, xu_r, x_ur, zu_r, mu, sc):
    """src/IPM/kernels.jl:293-303"""
    a = np.max(np.abs((x_lr - xl_r) * zl_r - mu), initial=0.0)
    b = np.max(np.abs((xu_r - x_ur) * zu_r - mu), initial=0.0)
    return float(max(a, b)) / sc


def get_average_complementarity(x_lr, xl_r, zl_r, x_ur, xu_r, zu_r):
    """src/IPM/kernels.jl:305-314"""
    n = len(x_lr) + len(x_ur)
    if n == 0:
        return 0.0
    return float((np.dot(x_lr, zl_r) - np.dot(xl_r, zl_r) + np.dot(xu_r, zu_r) - np.dot(x_ur, zu_r)) / n)


def get_min_complementarity(x_lr, xl_r, zl_r, x_ur, xu_r, zu_r):
    """src/IPM/kernels.jl:322-333"""
    return float(min(np.min((x_lr - xl_r) * zl_r, initial=np.inf), np.min((xu_r - x_ur) * zu_r, initial=np.inf)))


def get_rel_search_norm(x, dx):
    """src/IPM/kernels.jl:675-681"""
    return float(np.max(np.abs(dx) / (1.0 + np.abs(x)), initial=0.0))


def get_sd(l, zl_r, zu_r, s_max):
    """src/IPM/kernels.jl:684-689"""
    return max(s_max, (np.abs(l).sum() + np.abs(zl_r).sum() + np.abs(zu_r).sum()) / max(1, len(l) + len(zl_r) + len(zu_r))) / s_max


def get_sc(zl_r, zu_r, s_max):
    """src/IPM/kernels.jl:690-695"""
    return max(s_max, (np.abs(zl_r).sum() + np.abs(zu_r).sum()) / max(1, len(zl_r) + len(zu_r))) / s_max


def set_aug_rhs(x, xl, xu, f, zl, zu, jacl, c, mu, ind_lb, ind_ub):
    """src/IPM/kernels.jl:113-130 -> [px | py | pzl | pzu]"""
    px = -f + zl - zu - jacl
    py = -c
    pzl = (xl[ind_lb] - x[ind_lb]) * zl[ind_lb] + mu
    pzu = (xu[ind_ub] - x[ind_ub]) * zu[ind_ub] - mu
    return np.concatenate([px, py, pzl, pzu])


# --------------------------------------------------------------------------
# replay of one IPM iteration's linear algebra on the CPU (the reference-arm counterpart of
# madnlp.jl_b200/ipm.py::IPMLinearAlgebra): regular! src/IPM/solver.jl:216-298 with
# inertia_correction!(InertiaBased) src/IPM/solver.jl:611-670 and the perturbation schedule of
# src/IPM/options.jl:168-175
# --------------------------------------------------------------------------
class IPMLinearAlgebraCPU:
    first_hessian_perturbation = 1e-4
    min_hessian_perturbation = 1e-20
    max_hessian_perturbation = 1e20
    perturb_inc_fact_first = 1e2
    perturb_inc_fact = 8.0
    perturb_dec_fact = 1 / 3
    jacobian_regularization_value = 1e-8
    jacobian_regularization_exponent = 0.25

    def __init__(self, kkt, tol=1e-8):
        self.kkt = kkt
        self.tol = tol
        self.d = UnreducedKKTVector.for_kkt(kkt)
        self.p = UnreducedKKTVector.for_kkt(kkt)
        self.w = UnreducedKKTVector.for_kkt(kkt)
        self.del_w_last = 0.0
        self.cnt = dict(factorizations=0, backsolves=0, regularized=0, failed=0)
        self.t_factorize = 0.0

    def load_iterate(self, it):
        k = self.kkt
        g = (lambda name: it[name]) if isinstance(it, dict) else (lambda name: getattr(it, name))
        k.get_jacobian()[:] = g("jac"); k.get_hessian()[:] = g("hess")
        k.reg[:] = g("reg"); k.du_diag[:] = g("du_diag")
        k.l_diag[:] = g("l_diag"); k.u_diag[:] = g("u_diag"); k.l_lower[:] = g("l_lower"); k.u_lower[:] = g("u_lower")
        self.p.full()[:] = g("rhs")

    def _factorize_wrapper(self):
        import time as _t
        self.kkt.build_kkt()
        t0 = _t.perf_counter()
        self.kkt.linear_solver.factorize()
        self.t_factorize += _t.perf_counter() - t0
        self.cnt["factorizations"] += 1

    def _solve_refine_wrapper(self):
        ok, nit, _ = solve_refine(self.d, self.kkt, self.p, self.w, tol=self.tol)
        self.cnt["backsolves"] += nit
        return ok

    def step(self, mu=1e-2):
        k = self.kkt
        k.compress_jacobian(); k.compress_hessian()
        set_aug_diagonal_(k)
        self._factorize_wrapper()
        n_trial = 0
        del_w = del_c = del_w_prev = del_c_prev = 0.0
        inertia = k.linear_solver.inertia()
        ok = self._solve_refine_wrapper() if k.is_inertia_correct(*inertia) else False
        while not ok:
            if n_trial == 0:
                del_w = self.first_hessian_perturbation if self.del_w_last == 0.0 else max(
                    self.min_hessian_perturbation, self.perturb_dec_fact * self.del_w_last)
            else:
                del_w *= self.perturb_inc_fact_first if self.del_w_last == 0.0 else self.perturb_inc_fact
                if del_w > self.max_hessian_perturbation:
                    self.cnt["failed"] += 1
                    return False
            should_dual = getattr(k, "should_regularize_dual", lambda *a: a[1] != 0)(*inertia)
            del_c = self.jacobian_regularization_value * mu ** self.jacobian_regularization_exponent if should_dual else 0.0
            regularize_diagonal(k, del_w - del_w_prev, del_c - del_c_prev)
            del_w_prev, del_c_prev = del_w, del_c
            self._factorize_wrapper()
            inertia = k.linear_solver.inertia()
            ok = self._solve_refine_wrapper() if k.is_inertia_correct(*inertia) else False
            n_trial += 1
            self.cnt["regularized"] += 1
        if del_w != 0.0:
            self.del_w_last = del_w
        self.last_inertia = inertia
        return True
