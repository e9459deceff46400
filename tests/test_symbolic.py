"""Host logic of the sparse LDL^T: ordering, supernodes, front structures, scatter maps, level schedule and the
subtree-to-rank partition -- validated WITHOUT a GPU by replaying the multifrontal arithmetic in numpy on the
structures the C ABI exports (tests/mf_emulator.py), against dense LAPACK / eigenvalue truth from the oracle."""
import numpy as np
import pytest

import madnlp_oracle as o
import madnlp_jl_b200 as pkg
from mf_emulator import Symbolic

W = pkg.workloads


def _well_conditioned_values(colptr, rowval, n, rng, n_neg=0):
    """random values on a given lower-CSC pattern: diagonally dominant, last n_neg diagonal entries negative
    (quasi-definite => LDL^T exists for every ordering, inertia = (n - n_neg, 0, n_neg))."""
    nz = rng.uniform(-1, 1, len(rowval))
    cols = np.repeat(np.arange(n), np.diff(colptr))
    deg = np.zeros(n)
    off = rowval != cols
    np.add.at(deg, rowval[off], 1.0); np.add.at(deg, cols[off], 1.0)
    sign = np.ones(n); sign[n - n_neg:] = -1.0 if n_neg else 1.0
    diag = rowval == cols
    nz[diag] = sign[rowval[diag]] * (deg[rowval[diag]] + 1.0 + rng.random(diag.sum()))
    return nz


def _check(n, colptr, rowval, nz, expect_neg, tol=1e-9, **opts):
    S = Symbolic(n, colptr, rowval, **opts)
    inertia = S.factorize(nz)
    full = o.tril_to_full(colptr, rowval, nz, n).toarray()
    b = np.random.default_rng(5).standard_normal(n)
    x = S.solve(b)
    xr = np.linalg.solve(full, b)
    assert np.abs(x - xr).max() / np.abs(xr).max() < tol
    assert inertia == (n - expect_neg, 0, expect_neg)
    # structural invariants
    assert sorted(S.perm.tolist()) == list(range(n))
    assert S.sn_first[0] == 0 and S.sn_first[-1] == n and (np.diff(S.sn_first) > 0).all()
    for s in range(S.ns):
        p = S.sn_parent[s]
        assert p == -1 or (p > s and S.sn_level[p] > S.sn_level[s])
        rows = S.rows[S.rows_ptr[s]:S.rows_ptr[s + 1]]
        w = S.sn_first[s + 1] - S.sn_first[s]
        assert (rows[:w] == np.arange(S.sn_first[s], S.sn_first[s + 1])).all()
        assert (np.diff(rows[w:]) > 0).all() and (len(rows) == w or rows[w] >= S.sn_first[s + 1])
    return S


def _condensed_pattern(case):
    model, st = W.acopf_case(case)
    cb = o.Callback(st.nvar, st.ncon, st.jac_I, st.jac_J, st.hess_I, st.hess_J, st.ind_ineq, st.ind_lb, st.ind_ub)
    k = o.SparseCondensedKKTSystem(cb)
    return k.n, k.aug_colptr, k.aug_rowval


@pytest.mark.parametrize("ordering", [0, 1, 2])
@pytest.mark.parametrize("nemin", [1, 16, 40])
def test_condensed_opf_pattern(ordering, nemin):
    n, cp, rv = _condensed_pattern("case30_synth")
    rng = np.random.default_rng(ordering * 10 + nemin)
    _check(n, cp, rv, _well_conditioned_values(cp, rv, n, rng), 0, ordering=ordering, nemin=nemin)


def test_condensed_opf_300_and_negative_pivots():
    n, cp, rv = _condensed_pattern("case300_synth")
    rng = np.random.default_rng(7)
    S = _check(n, cp, rv, _well_conditioned_values(cp, rv, n, rng, n_neg=37), 37)
    assert S.stats["nnz_l"] >= len(rv) and S.stats["n_levels"] == S.n_levels


@pytest.mark.parametrize("chain_merge_f", [32, 64])
def test_only_children_are_absorbed_when_the_latency_rule_is_on(chain_merge_f):
    """b2_options.chain_merge_f: a supernode with exactly one child absorbs it while the merged front order stays <= chain_merge_f
    (<= 64), whatever the explicit zeros -- fewer tree levels; the numeric replay must still reproduce the dense solution, and no
    chain link may be left that the rule could still merge."""
    n, cp, rv = _condensed_pattern("case300_synth")
    rng = np.random.default_rng(chain_merge_f)
    nz = _well_conditioned_values(cp, rv, n, rng)
    S0 = _check(n, cp, rv, nz, 0, chain_merge_f=0)
    S1 = _check(n, cp, rv, nz, 0, chain_merge_f=chain_merge_f)
    assert S1.ns < S0.ns and S1.sn_level.max() <= S0.sn_level.max()
    nchild = np.bincount(S1.sn_parent[S1.sn_parent >= 0], minlength=S1.ns)
    f1 = np.diff(S1.rows_ptr); w1 = np.diff(S1.sn_first)
    for s_ in range(S1.ns):
        p_ = S1.sn_parent[s_]
        if p_ >= 0 and nchild[p_] == 1:
            assert w1[s_] + f1[p_] > min(chain_merge_f, 64)    # (an only child with w_child + f_parent <= bound would have been absorbed)


def test_augmented_kkt_pattern_quasi_definite():
    """SparseKKTSystem-style [[H+S, J'],[J, -dI]] (config 5 generator, small): inertia must be (n_tot, 0, m)."""
    N, n_tot, m, I, J, V = W.augmented_grid_kkt(5, 4, 6)
    cp, rv, mp = o.coo_to_csc(I, J, N, N)
    nz = np.zeros(len(rv)); o.transfer(nz, V, mp)
    # make the (2,2) block clearly negative for a well-conditioned check
    cols = np.repeat(np.arange(N), np.diff(cp))
    nz[(rv == cols) & (rv >= n_tot)] = -1.0
    for ordering in (0, 1):
        _check(N, cp, rv, nz, m, tol=1e-8, ordering=ordering)


def test_hs15_augmented_6x6():
    """C1: HS15 SparseKKTSystem matrix of SURVEY.md Appendix A through the natural ordering (zero (2,2) block is
    reached only after its neighbours are eliminated) -> inertia (4,0,2) and the reference solution."""
    cb = o.HS15Model.callback()
    kkt = o.SparseKKTSystem(cb, o.DenseLDLInertiaSolver)
    o.test_kkt_system(kkt, o.HS15Model)
    S = Symbolic(6, kkt.aug_colptr, kkt.aug_rowval, ordering=2)
    assert S.factorize(kkt.aug_nz) == (4, 0, 2)
    b = np.array([0.0, 1.0, 0.0, 0.0, 1.0, 1.0])                      # reduced rhs of Appendix A
    x = S.solve(b)
    assert np.abs(x - np.array([0.24987493746873435, 0.00497512437810945, -1.0, -0.7501250625312657,
                                -0.9989999999999999, -0.7493749374687343])).max() < 1e-13


def test_tiny_and_diagonal_matrices():
    cp = np.array([0, 1], dtype=np.int32); rv = np.array([0], dtype=np.int32)
    S = Symbolic(1, cp, rv)
    assert S.factorize(np.array([-2.0])) == (0, 0, 1)
    assert np.allclose(S.solve(np.array([4.0])), [-2.0])
    n = 7
    cp = np.arange(n + 1, dtype=np.int32); rv = np.arange(n, dtype=np.int32)
    S = Symbolic(n, cp, rv)
    d = np.array([1.0, -2, 3, 4, -5, 6, 7])
    assert S.factorize(d) == (5, 0, 2)
    assert np.allclose(S.solve(np.ones(n)), 1 / d)


def test_static_pivot_perturbation_is_reported_as_zero():
    """A structurally singular pivot must come back as num_zero > 0 (MUMPS/MA57 convention, mumps.jl:248-250)."""
    cp = np.array([0, 1, 2], dtype=np.int32); rv = np.array([0, 1], dtype=np.int32)
    S = Symbolic(2, cp, rv)
    assert S.factorize(np.array([1.0, 0.0])) == (1, 1, 0)


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_subtree_partition(parts):
    """multi-GPU sharding: every supernode is owned by exactly one rank or by the shared top tree; subtrees are
    closed (a child of an owned node has the same owner); the top tree is an ancestor-closed set."""
    n, cp, rv = _condensed_pattern("case300_synth")
    S = Symbolic(n, cp, rv, n_parts=parts, part_rank=0)
    own = S.owner
    assert own.min() >= -1 and own.max() < parts
    assert set(own[own >= 0].tolist()) == set(range(parts))
    for s in range(S.ns):
        p = S.sn_parent[s]
        if p >= 0:
            assert own[p] == -1 or own[p] == own[s]
            if own[s] == -1:
                assert own[p] == -1
    work = np.zeros(parts)
    for s in range(S.ns):
        if own[s] >= 0:
            w = S.sn_first[s + 1] - S.sn_first[s]; f = S.rows_ptr[s + 1] - S.rows_ptr[s]
            work[own[s]] += sum((f - k) ** 2 for k in range(w)) + 2000
    assert work.max() <= 2.0 * work.mean()
    # numerics are unaffected by the partition bookkeeping
    rng = np.random.default_rng(3)
    nz = _well_conditioned_values(cp, rv, n, rng)
    S.factorize(nz)
    full = o.tril_to_full(cp, rv, nz, n).toarray()
    b = rng.standard_normal(n)
    assert np.abs(S.solve(b) - np.linalg.solve(full, b)).max() < 1e-9


# ---- randomised patterns (hypothesis): arbitrary symmetric sparsity incl. disconnected graphs, dense rows, empty columns
from hypothesis import given, settings, strategies as st_  # noqa: E402


@settings(max_examples=150, deadline=None)
@given(n=st_.integers(min_value=1, max_value=90), density=st_.floats(min_value=0.0, max_value=0.35),
       n_neg=st_.integers(min_value=0, max_value=10), ordering=st_.sampled_from([0, 1, 2]),
       nemin=st_.sampled_from([1, 8, 32]), parts=st_.sampled_from([1, 2, 3]), seed=st_.integers(min_value=0, max_value=10**6))
def test_random_patterns_factor_and_solve(n, density, n_neg, ordering, nemin, parts, seed):
    """Any symmetric pattern, any ordering / amalgamation / partition setting: the exported symbolic structure replayed in
    numpy reproduces the dense solution and the exact inertia, and satisfies the structural invariants."""
    rng = np.random.default_rng(seed)
    n_neg = min(n_neg, n)
    mask = np.tril(rng.random((n, n)) < density, -1)
    if n > 3 and rng.random() < 0.3:
        mask[-1, :-1] = True                                  # one dense row (arrowhead)
    mask |= np.eye(n, dtype=bool)
    rows, cols = np.nonzero(mask.T)                           # walk column by column
    cols_, rows_ = rows, cols                                  # (mask.T nonzero gives (col, row) pairs sorted by col)
    order = np.lexsort((rows_, cols_))
    cols_, rows_ = cols_[order], rows_[order]
    colptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(colptr, cols_ + 1, 1)
    colptr = np.cumsum(colptr).astype(np.int32)
    rowval = rows_.astype(np.int32)
    assert (rowval >= np.repeat(np.arange(n), np.diff(colptr))).all()          # lower triangle, sorted rows
    nz = _well_conditioned_values(colptr, rowval, n, rng, n_neg=n_neg)
    _check(n, colptr, rowval, nz, n_neg, tol=1e-8, ordering=ordering, nemin=nemin, n_parts=parts, part_rank=0)


@pytest.mark.parametrize("ordering", [0, 1])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_duals_sharing_a_primal_neighbour_get_distinct_partners(ordering, seed):
    """Augmented KKT with a ZERO (2,2) block where two equality rows touch the SAME two variables: a fill-reducing
    ordering eliminates the (degree-2) duals first; moving both to just after their common earliest neighbour u leaves the
    rank-one Schur block -(a_i a_j)/d_u and the second dual pivot cancels exactly.  The ordering constraint therefore matches
    every dual with a DISTINCT preceding primal neighbour -> inertia exactly (n_tot, 0, m) with 1 x 1 static pivots, as
    Bunch-Kaufman (dsytrf) reports."""
    rng = np.random.default_rng(seed)
    n_tot, m = 8, 2
    R = rng.standard_normal((n_tot, n_tot))
    H = R @ R.T + n_tot * np.eye(n_tot)              # dense SPD block: every primal has a high degree
    rows, cols, vals = [], [], []
    for j in range(n_tot):
        for i in range(j, n_tot):
            rows.append(i); cols.append(j); vals.append(H[i, j])
    for c in range(m):                               # both constraints couple variables 0 and 1 only
        for j in (0, 1):
            rows.append(n_tot + c); cols.append(j); vals.append(rng.uniform(0.5, 2.0) * (1.0 if (c + j) % 2 else -1.0))
        rows.append(n_tot + c); cols.append(n_tot + c); vals.append(0.0)
    N = n_tot + m
    cp, rv, mp = o.coo_to_csc(np.array(rows), np.array(cols), N, N)
    nz = np.zeros(len(rv)); o.transfer(nz, np.array(vals), mp)
    truth = o.DenseLDLInertiaSolver(cp, rv, nz, N).factorize().inertia()
    assert truth == (n_tot, 0, m)
    S = Symbolic(N, cp, rv, ordering=ordering, kkt_n_primal=n_tot)
    assert S.factorize(nz) == (n_tot, 0, m)
    # structural statement of the matching: a system of distinct representatives among PRECEDING primal neighbours exists
    iperm = np.empty(N, dtype=np.int64); iperm[S.perm] = np.arange(N)
    full = o.tril_to_full(cp, rv, np.ones(len(rv)), N).toarray() != 0
    used = set()
    for v in sorted(range(n_tot, N), key=lambda q: iperm[q]):
        cand = [u for u in range(n_tot) if full[v, u] and iperm[u] < iperm[v] and u not in used]
        assert cand, "dual %d has no unused preceding primal neighbour" % v
        used.add(min(cand, key=lambda u: iperm[u]))
    b = rng.standard_normal(N)
    x = S.solve(b)
    xr = np.linalg.solve(o.tril_to_full(cp, rv, nz, N).toarray(), b)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-9
