"""Multi-GPU (one process per GPU) sparse LDL^T: elimination-tree subtrees are owned by ranks, the shared top tree is
replicated, and the ONLY exchange is the sum of the subtree roots' update blocks -- a single NCCL all-reduce per
factorisation and one per solve.  Every rank analyses the same pattern (deterministic), so the
partition needs no communication.

    factorize:  b2_factorize_local -> all_reduce(exchange blocks) -> b2_factorize_top
    solve:      b2_solve_fwd_local -> all_reduce(exchange vector) -> b2_solve_top -> b2_solve_bwd_local
                -> all_reduce(x) (each row is finalised by exactly one rank, others contribute zeros)
    inertia:    all_reduce of the owned-subtree counts + the (replicated) top-tree counts once

Same AbstractLinearSolver surface as B200SparseSolver.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .capi import lib, check


class _CudaView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def device_view(ptr, n):
    """zero-copy float64 torch view of library-owned device memory"""
    return torch.as_tensor(_CudaView(ptr, n), device="cuda")


class DistributedSparseSolver:
    input_type = "csc"

    def _on_stream(self):
        """the b2_* kernels run on `self.stream`; torch.distributed orders its collectives against torch's CURRENT stream --
        make the two the same for the duration of a call (a raw cudaStream_t cannot be made current: refuse it)"""
        import contextlib
        if self.stream is None:
            return contextlib.nullcontext()
        if not isinstance(self.stream, torch.cuda.Stream):
            raise TypeError("DistributedSparseSolver needs a torch.cuda.Stream (or None = current stream)")
        return torch.cuda.stream(self.stream)

    def __init__(self, csc, opt=None, rank=None, world=None, group=None, stream=None):
        capi.require_device()
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.csc = csc
        self.opt = opt if opt is not None else capi.default_options()
        self.opt.n_parts = self.world
        self.opt.part_rank = self.rank
        self._h = C.c_void_p()
        self.colptr = np.ascontiguousarray(csc.colptr, dtype=np.int32)
        self.rowval = np.ascontiguousarray(csc.rowval, dtype=np.int32)
        check(lib.b2_create(csc.n, int(self.colptr[-1]), self.colptr.ctypes.data, self.rowval.ctypes.data,
                            csc.nzval.data_ptr(), C.byref(self.opt), None, C.byref(self._h)))
        self.n = csc.n
        self.stream = stream
        buf = C.c_void_p(); nf = C.c_int64(); nsol = C.c_int64()
        check(lib.b2_exchange_buffer(self._h, C.byref(buf), C.byref(nf), C.byref(nsol)))
        self._xf = device_view(buf.value, nf.value) if nf.value > 0 else None
        vbuf = C.c_void_p(); nv = C.c_int64()
        check(lib.b2_exchange_vector(self._h, C.byref(vbuf), C.byref(nv)))
        self._xv = device_view(vbuf.value, nv.value) if nv.value > 0 else None
        self._cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
        self.exchange_bytes = dict(factor=8 * nf.value, solve=8 * nv.value, x=8 * self.n)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:
            lib.b2_destroy(h)
            self._h = None

    @staticmethod
    def default_options(**kw):
        return capi.default_options(**kw)

    def introduce(self):
        return f"b200kkt multifrontal LDL^T, subtree-sharded over {self.world} GPUs"

    def is_async(self):
        return True

    def factorize(self):
        with self._on_stream():
            sp = capi.stream_ptr(self.stream)
            check(lib.b2_factorize_local(self._h, sp))
            if self._xf is not None and self.world > 1:
                dist.all_reduce(self._xf, group=self.group)
            check(lib.b2_factorize_top(self._h, sp))
        return self

    def solve_linear_system(self, x):
        with self._on_stream():
            sp = capi.stream_ptr(self.stream)
            check(lib.b2_solve_fwd_local(self._h, x.data_ptr(), sp))
            if self._xv is not None and self.world > 1:
                dist.all_reduce(self._xv, group=self.group)
            check(lib.b2_solve_top(self._h, x.data_ptr(), sp))
            check(lib.b2_solve_bwd_local(self._h, x.data_ptr(), sp))
            if self.world > 1:
                dist.all_reduce(x, group=self.group)
        return x

    def is_inertia(self):
        return True

    def inertia(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.b2_inertia_parts(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), capi.stream_ptr(self.stream)))
        loc = torch.tensor([a.value, b.value], dtype=torch.int64, device="cuda")
        if self.world > 1:
            with self._on_stream():
                dist.all_reduce(loc, group=self.group)
        neg = int(loc[0].item()) + c.value
        zero = int(loc[1].item()) + d.value
        return (self.n - neg - zero, zero, neg)

    def improve(self):
        ch = C.c_int32(0)
        check(lib.b2_improve(self._h, C.byref(ch)))
        return bool(ch.value)

    def stats(self):
        st = capi.Stats()
        check(lib.b2_get_stats(self._h, C.byref(st)))
        return st.as_dict()
