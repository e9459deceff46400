"""numpy replay of the multifrontal arithmetic on the symbolic structure exported by the C ABI
(b2_symbolic_export).  TEST INFRASTRUCTURE: lets the CPU-only test-suite validate ordering, supernodes, front row
structures, relative indices, the A->front scatter map and the level schedule without a GPU.  It mirrors what the
CUDA kernels do front by front (front_kernels.cuh / solve_kernels.cuh) but shares no code with them."""
import ctypes as C

import numpy as np

import madnlp_jl_b200 as pkg

capi = pkg.capi
lib = capi.lib


class Symbolic:
    def __init__(self, n, colptr, rowval, **opts):
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int32)
        self.rowval = np.ascontiguousarray(rowval, dtype=np.int32)
        self.opt = capi.default_options(**opts)
        self.h = C.c_void_p()
        capi.check(lib.b2_create_symbolic_only(n, int(self.colptr[-1]), self.colptr.ctypes.data, self.rowval.ctypes.data,
                                               C.byref(self.opt), None, C.byref(self.h)))
        sz = capi.SymbolicSizes()
        capi.check(lib.b2_symbolic_query(self.h, C.byref(sz)))
        self.n = n
        ns = sz.n_supernodes
        self.ns = ns
        a = lambda k, dt: np.zeros(k, dtype=dt)
        self.perm = a(n, np.int32); self.sn_first = a(ns + 1, np.int32); self.sn_parent = a(ns, np.int32)
        self.sn_level = a(ns, np.int32); self.rows_ptr = a(ns + 1, np.int64); self.rows = a(sz.n_rows, np.int32)
        self.lp_off = a(ns + 1, np.int64); self.cb_off = a(ns + 1, np.int64); self.rel_ptr = a(ns + 1, np.int64)
        self.rel = a(max(sz.n_rel, 1), np.int32); self.amap_ptr = a(ns + 1, np.int64)
        self.amap_src = a(max(sz.n_amap, 1), np.int64); self.amap_dst = a(max(sz.n_amap, 1), np.int64)
        p = lambda x: x.ctypes.data
        capi.check(lib.b2_symbolic_export(self.h, p(self.perm), p(self.sn_first), p(self.sn_parent), p(self.sn_level),
                                          p(self.rows_ptr), p(self.rows), p(self.lp_off), p(self.cb_off), p(self.rel_ptr),
                                          p(self.rel), p(self.amap_ptr), p(self.amap_src), p(self.amap_dst)))
        self.owner = a(ns, np.int32)
        capi.check(lib.b2_symbolic_owner(self.h, p(self.owner)))
        st = capi.Stats()
        capi.check(lib.b2_get_stats(self.h, C.byref(st)))
        self.stats = st.as_dict()
        self.lval_size = sz.lval_size
        self.n_levels = sz.n_levels

    def __del__(self):
        if getattr(self, "h", None):
            lib.b2_destroy(self.h)
            self.h = None

    def children(self):
        ch = [[] for _ in range(self.ns)]
        for s in range(self.ns):
            if self.sn_parent[s] >= 0:
                ch[self.sn_parent[s]].append(s)
        return ch

    # ---- numeric replay -------------------------------------------------------------------------------
    def factorize(self, nzval, eps=1e-13):
        ns = self.ns
        L = np.zeros(self.lval_size)
        d = np.zeros(self.n)
        cbs = [None] * ns
        ch = self.children()
        neg = pert = 0
        order = np.lexsort((np.arange(ns), self.sn_level))     # level by level, ascending id inside a level
        for s in order:
            w = self.sn_first[s + 1] - self.sn_first[s]
            f = int(self.rows_ptr[s + 1] - self.rows_ptr[s])
            F = np.zeros((f, f))
            a0, a1 = self.amap_ptr[s], self.amap_ptr[s + 1]
            dst = self.amap_dst[a0:a1] - self.lp_off[s]
            F[dst % f, dst // f] = nzval[self.amap_src[a0:a1]]
            for c in ch[s]:                                     # ascending child id
                rl = self.rel[self.rel_ptr[c]:self.rel_ptr[c + 1]]
                F[np.ix_(rl, rl)] += cbs[c]
                cbs[c] = None
            F = np.tril(F)
            for k in range(w):
                dk = F[k, k]
                if not (abs(dk) >= eps):
                    dk = -eps if dk < 0 else eps
                    pert += 1
                elif dk < 0:
                    neg += 1
                F[k, k] = dk
                u = F[k + 1:, k].copy()
                F[k + 1:, k] = u / dk
                F[k + 1:, k + 1:] -= np.tril(np.outer(F[k + 1:, k], u))
            L[self.lp_off[s]:self.lp_off[s] + f * w] = F[:, :w].T.ravel()   # column-major f x w
            d[self.sn_first[s]:self.sn_first[s + 1]] = np.diag(F)[:w]
            cbs[s] = F[w:, w:] + np.tril(F[w:, w:], -1).T - 0.0
            cbs[s] = np.tril(F[w:, w:]) + np.tril(F[w:, w:], -1).T          # symmetric copy for ix_ add
        self.L, self.d = L, d
        return (self.n - neg - pert, pert, neg)

    def solve(self, b):
        ns = self.ns
        x = b[self.perm].astype(float).copy()
        cbv = [None] * ns
        ch = self.children()
        order = np.lexsort((np.arange(ns), self.sn_level))
        for s in order:                                                      # forward
            w = self.sn_first[s + 1] - self.sn_first[s]
            f = int(self.rows_ptr[s + 1] - self.rows_ptr[s])
            P = self.L[self.lp_off[s]:self.lp_off[s] + f * w].reshape(w, f).T   # f x w
            y = np.zeros(f)
            y[:w] = x[self.sn_first[s]:self.sn_first[s + 1]]
            for c in ch[s]:
                rl = self.rel[self.rel_ptr[c]:self.rel_ptr[c + 1]]
                y[rl] += cbv[c]
            for k in range(w):
                y[k + 1:] -= P[k + 1:, k] * y[k]
            x[self.sn_first[s]:self.sn_first[s + 1]] = y[:w]
            cbv[s] = y[w:]
        for s in order[::-1]:                                                # backward (with D^-1)
            w = self.sn_first[s + 1] - self.sn_first[s]
            f = int(self.rows_ptr[s + 1] - self.rows_ptr[s])
            P = self.L[self.lp_off[s]:self.lp_off[s] + f * w].reshape(w, f).T
            rows = self.rows[self.rows_ptr[s]:self.rows_ptr[s + 1]]
            xx = np.zeros(f)
            c0 = self.sn_first[s]
            xx[:w] = x[c0:c0 + w] / self.d[c0:c0 + w]
            xx[w:] = x[rows[w:]]
            for k in range(w - 1, -1, -1):
                xx[k] -= P[k + 1:, k] @ xx[k + 1:]
            x[c0:c0 + w] = xx[:w]
        out = np.zeros(self.n)
        out[self.perm] = x
        return out
