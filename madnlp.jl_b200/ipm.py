"""Replay of the linear-algebra call order of one MadNLP IPM iteration (`regular!`, src/IPM/solver.jl:216-298):

    eval_jac_wrapper!  -> compress_jacobian!      (values arrive in kkt.jac)
    eval_lag_hess_wrapper! -> compress_hessian!   (values arrive in kkt.hess)
    set_aug_diagonal!                             (src/IPM/kernels.jl:4-27)
    inertia_correction!(InertiaBased)             (src/IPM/solver.jl:611-670):
        factorize_wrapper! = build_kkt! + factorize!   ; inertia ; is_inertia_correct
        solve_refine_wrapper! (Richardson)             ; on failure improve! and retry (factorization.jl:1-19)
        while !ok: regularize_diagonal!(dw, dc) ; factorize_wrapper! ; inertia ; solve_refine

The model callbacks themselves are out of scope: an iterate supplies their outputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .kkt import UnreducedKKTVector
from .richardson import RichardsonIterator


@dataclass
class InertiaOptions:
    # src/IPM/options.jl:168-175
    first_hessian_perturbation: float = 1e-4
    min_hessian_perturbation: float = 1e-20
    max_hessian_perturbation: float = 1e20
    perturb_inc_fact_first: float = 1e2
    perturb_inc_fact: float = 8.0
    perturb_dec_fact: float = 1 / 3
    jacobian_regularization_value: float = 1e-8
    jacobian_regularization_exponent: float = 0.25


class IPMLinearAlgebra:
    """Owns the work vectors of MadNLPSolver that the hot path touches (d, p, _w4) and drives one iteration."""

    def __init__(self, kkt, tol=1e-8, use_cuda_graph=True, speculate=True):
        self.kkt = kkt
        self.use_cuda_graph = use_cuda_graph
        self.speculate = speculate     # first refinement step queued before the inertia is known (see step())
        self._prologue_graph = None
        self.iterator = RichardsonIterator(kkt, tol=tol, use_cuda_graph=use_cuda_graph)
        self.d = UnreducedKKTVector.for_kkt(kkt)
        self.p = UnreducedKKTVector.for_kkt(kkt)
        self.w = UnreducedKKTVector.for_kkt(kkt)
        self.opt = InertiaOptions()
        self.del_w_last = 0.0
        self.cnt = dict(factorizations=0, backsolves=0, regularized=0, failed=0)

    def load_iterate(self, it, non_blocking=True):
        """Copy one iterate's callback outputs / diagonal inputs into the KKT buffers (H2D when `it` holds pinned
        host tensors, D2D when it holds device tensors)."""
        k = self.kkt
        pairs = ((k.get_jacobian(), it["jac"]), (k.get_hessian(), it["hess"]), (k.reg, it["reg"]), (k.du_diag, it["du_diag"]),
                 (k.l_diag, it["l_diag"]), (k.u_diag, it["u_diag"]), (k.l_lower, it["l_lower"]), (k.u_lower, it["u_lower"]),
                 (self.p.values, it["rhs"]))
        if all(src.is_cuda for _, src in pairs):
            # device-resident producer: one launch for the nine vectors
            import ctypes as C
            from .capi import lib, check, stream_ptr
            cnt = len(pairs)
            src = (C.c_void_p * cnt)(*[s_.data_ptr() for _, s_ in pairs])
            dst = (C.c_void_p * cnt)(*[d_.data_ptr() for d_, _ in pairs])
            ns = (C.c_int64 * cnt)(*[d_.numel() for d_, _ in pairs])
            assert all(d_.numel() == s_.numel() and s_.dtype == torch.float64 and s_.is_contiguous() for d_, s_ in pairs)
            check(lib.b2_copy_many(cnt, src, dst, ns, stream_ptr(getattr(k, "stream", None))))
        else:
            for d_, s_ in pairs:
                d_.copy_(s_, non_blocking=non_blocking)

    def load_iterate_host(self, host_iterates, idx):
        """Host-facing staging of one iterate (the role of SparseWrapperModel's pinned buffers, lib/MadNLPGPU/src/wrappers.jl:
        173-196): `host_iterates[idx]` holds PINNED host tensors.  The copies run on a dedicated copy stream in the order the
        step consumes them; the compute stream waits for the assembly inputs before the prologue and for the right-hand side
        only before the refinement, so the H2D of `rhs` overlaps assembly + factorisation."""
        k = self.kkt
        it = host_iterates[idx]
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._ev_inputs = torch.cuda.Event()
            self._ev_rhs = torch.cuda.Event()
            self._ev_free = torch.cuda.Event()
        main = torch.cuda.current_stream()
        self._ev_free.record(main)                       # the previous step has finished reading the buffers
        cs = self._copy_stream
        cs.wait_event(self._ev_free)
        with torch.cuda.stream(cs):
            for dst, name in ((k.get_jacobian(), "jac"), (k.get_hessian(), "hess"), (k.reg, "reg"), (k.du_diag, "du_diag"),
                              (k.l_diag, "l_diag"), (k.u_diag, "u_diag"), (k.l_lower, "l_lower"), (k.u_lower, "u_lower")):
                dst.copy_(it[name], non_blocking=True)
            self._ev_inputs.record(cs)
            self.p.values.copy_(it["rhs"], non_blocking=True)
            self._ev_rhs.record(cs)
        main.wait_event(self._ev_inputs)
        self._rhs_pending = True

    def _wait_rhs(self):
        if getattr(self, "_rhs_pending", False):
            torch.cuda.current_stream().wait_event(self._ev_rhs)
            self._rhs_pending = False

    def _prologue(self):
        k = self.kkt
        k.compress_jacobian()
        k.compress_hessian()
        k.set_aug_diagonal_()
        k.build_kkt()
        k.linear_solver.factorize()

    def _factorize_wrapper(self):
        self.kkt.build_kkt()
        self.kkt.linear_solver.factorize()
        self.cnt["factorizations"] += 1

    def _solve_refine_wrapper(self):
        ok = self.iterator.solve_refine(self.d, self.p, self.w)
        if not ok and self.kkt.linear_solver.improve():
            # improve!() changed a factorisation parameter (pivot threshold) that the captured prologue has baked in:
            # drop the captured graph so that every later step factorises with the new setting
            self._prologue_graph = None
            self.kkt.linear_solver.factorize()
            ok = self.iterator.solve_refine(self.d, self.p, self.w)
        self.cnt["backsolves"] += self.iterator.ir
        return ok

    def step(self, mu=1e-2, after_prologue=None):
        """One `regular!` linear-algebra pass; returns True when a step direction was obtained.  `after_prologue` (optional
        callable) runs on the host right after assembly + factorisation have been queued and before the host blocks for the
        inertia: host-side work placed there (e.g. queueing the next iterate's H2D copies, HostIteratePipeline) is hidden
        behind ~0.2 ms of device work instead of sitting between two steps."""
        k = self.kkt
        # compress_* + set_aug_diagonal! + the first factorize_wrapper! of inertia_correction!: fixed launch sequence,
        # replayed as one CUDA graph from the third step on (eager, capture, replay)
        if not self.use_cuda_graph or self._prologue_graph is None:
            self._prologue()
            if self.use_cuda_graph:
                self._prologue_graph = False
        elif self._prologue_graph is False:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                self._prologue()
            self._prologue_graph = g
            g.replay()
        else:
            self._prologue_graph.replay()
        self.cnt["factorizations"] += 1
        self._wait_rhs()
        if after_prologue is not None:
            after_prologue()
        # inertia_correction!(InertiaBased)
        o = self.opt
        n_trial = 0
        del_w = del_c = del_w_prev = del_c_prev = 0.0
        ls = k.linear_solver
        if self.speculate and hasattr(ls, "inertia_enqueue"):
            # queue the inertia read AND the first refinement step behind the factorisation, block once for both
            ls.inertia_enqueue()
            self.iterator.start(self.d, self.p, self.w)
            torch.cuda.current_stream().synchronize()
            inertia = ls.inertia_fetch()
            if k.is_inertia_correct(*inertia):
                ok = self._solve_refine_wrapper()
            else:
                self.iterator.discard()
                ok = False
        else:
            inertia = ls.inertia()
            ok = self._solve_refine_wrapper() if k.is_inertia_correct(*inertia) else False
        while not ok:
            if n_trial == 0:
                del_w = o.first_hessian_perturbation if self.del_w_last == 0.0 else max(
                    o.min_hessian_perturbation, o.perturb_dec_fact * self.del_w_last)
            else:
                del_w *= o.perturb_inc_fact_first if self.del_w_last == 0.0 else o.perturb_inc_fact
                if del_w > o.max_hessian_perturbation:
                    self.cnt["failed"] += 1
                    return False
            del_c = (o.jacobian_regularization_value * mu ** o.jacobian_regularization_exponent
                     if k.should_regularize_dual(*inertia) else 0.0)
            k.regularize_diagonal(del_w - del_w_prev, del_c - del_c_prev)
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

class HostIteratePipeline:
    """Double-buffered host<->device staging around `IPMLinearAlgebra.step` (the role of SparseWrapperModel's pinned buffers,
    lib/MadNLPGPU/src/wrappers.jl:173-196, made asynchronous): while step i computes, a copy stream moves the callback outputs of
    iterate i+1 from PINNED host memory into the idle one of two device staging sets, and a second copy stream returns step i-1's
    direction to pinned host memory (the two DMA directions run concurrently).  `load()` then moves the staged iterate into the KKT
    buffers with ONE device-to-device launch (`b2_copy_many`).  Every byte still crosses PCIe once per step; only the waiting is
    hidden.  Events order all buffer reuse, so the caller never synchronises for the copies (`drain()` at the end of a run)."""

    def __init__(self, la, fields):
        k = la.kkt
        self.la = la
        self.fields = tuple(fields)
        self._dst = dict(jac=k.get_jacobian(), hess=k.get_hessian(), reg=k.reg, du_diag=k.du_diag, l_diag=k.l_diag,
                         u_diag=k.u_diag, l_lower=k.l_lower, u_lower=k.u_lower, rhs=la.p.values)
        dev = la.d.values.device
        self.stage = [{f: torch.empty_like(self._dst[f]) for f in self.fields} for _ in range(2)]
        self.d_stage = [torch.empty_like(la.d.values) for _ in range(2)]
        self.d_host = [torch.zeros(la.d.values.numel(), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.ev_staged = [torch.cuda.Event() for _ in range(2)]      # H2D of the slot complete
        self.ev_consumed = [torch.cuda.Event() for _ in range(2)]    # the D2D out of the slot complete (slot may be refilled)
        self.ev_d_ready = [torch.cuda.Event() for _ in range(2)]     # d copied into d_stage[slot]
        self.ev_d_out = [torch.cuda.Event() for _ in range(2)]       # d_stage[slot] is in host memory (slot may be rewritten)
        self.h2d_bytes = sum(self._dst[f].numel() * 8 for f in self.fields)
        self.d2h_bytes = la.d.values.numel() * 8
        self._n_pref = 0
        self._n_out = 0

    def prefetch(self, host_iterate):
        """queue the H2D of one iterate (pinned host tensors) into the next staging slot; returns the slot"""
        slot = self._n_pref & 1
        if self._n_pref >= 2:
            self.h2d.wait_event(self.ev_consumed[slot])
        with torch.cuda.stream(self.h2d):
            for f in self.fields:
                self.stage[slot][f].copy_(host_iterate[f], non_blocking=True)
            self.ev_staged[slot].record(self.h2d)
        self._n_pref += 1
        return slot

    def load(self, slot):
        """compute stream: wait for the slot, move it into the KKT buffers (one launch)"""
        main = torch.cuda.current_stream()
        main.wait_event(self.ev_staged[slot])
        self.la.load_iterate(self.stage[slot])
        self.ev_consumed[slot].record(main)

    def push_result(self):
        """queue the D2H of the step direction `d` behind the step just issued; returns the index of the pinned host buffer"""
        import ctypes as C
        from .capi import lib, check, stream_ptr
        slot = self._n_out & 1
        main = torch.cuda.current_stream()
        if self._n_out >= 2:
            main.wait_event(self.ev_d_out[slot])
        d = self.la.d.values
        src = (C.c_void_p * 1)(d.data_ptr()); dst = (C.c_void_p * 1)(self.d_stage[slot].data_ptr()); ns = (C.c_int64 * 1)(d.numel())
        check(lib.b2_copy_many(1, src, dst, ns, stream_ptr(getattr(self.la.kkt, "stream", None))))
        self.ev_d_ready[slot].record(main)
        self.d2h.wait_event(self.ev_d_ready[slot])
        with torch.cuda.stream(self.d2h):
            self.d_host[slot].copy_(self.d_stage[slot], non_blocking=True)
            self.ev_d_out[slot].record(self.d2h)
        self._n_out += 1
        return slot

    def drain(self):
        """compute stream waits for every queued result copy (call before the closing event of a timed region)"""
        main = torch.cuda.current_stream()
        for slot in range(min(2, self._n_out)):
            main.wait_event(self.ev_d_out[slot])

