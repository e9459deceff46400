// Latency-oriented kernels for small fronts (order f <= 64): ONE WARP PER FRONT.
//
// On AC-OPF-like KKT systems every front is tiny (<= ~50) and the factorisation holds only ~2e7 flops: wall time is
// the dependent-latency chain, not bandwidth or flops.  These kernels therefore
//   * keep a front private to one warp (shared memory slice + registers), so a pivot step costs one __syncwarp
//     instead of two __syncthreads;
//   * take the reciprocal of each pivot with rcp.approx + 2 Newton steps (the only division on the critical path);
//   * carry the triangular-solve recurrences through warp shuffles (chain per pivot: SHFL + DFMA) with the factor
//     entries prefetched from shared memory, independent of the recurrence;
//   * run through a STAGED schedule: a CTA owns a list of stages (local levels of an elimination subtree); its warps
//     sweep the fronts of a stage, __syncthreads(), next stage.  Whole bottom subtrees thus execute inside one
//     kernel launch; data handed from child to parent front travels through global memory (same SM, ordered by
//     the barrier), so no pointer in here is const/__restrict__ for buffers written by these kernels.
#pragma once
#include "common.cuh"
#include "front_kernels.cuh"
#include "solve_kernels.cuh"

namespace b2 {

constexpr int FW_WARPS = 4;

struct ChildRec {          // one per (parent, child) edge, contiguous per parent, ascending child id
    int64_t cb_off;        // child's update block in the workspace (ld = rc)
    int64_t rel_off;       // child's relative indices into the parent front
    int64_t cbv_off;       // child's contribution vector (solve)
    int32_t rc;            // order of the child's update block
    int32_t sn;
};
static_assert(sizeof(ChildRec) == 32, "ChildRec must be 32 bytes");

struct WarpSched {
    const int32_t* cta_ptr;     // [nCTA+1] stage ranges
    const int32_t* stage_off;   // [nstage] offset into list
    const int32_t* stage_cnt;   // [nstage]
    const int32_t* list;        // supernode ids
};

__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// ------------------------------------------------------------------------------------------------ factor
// A front is owned by a TEAM of NW warps (NW = 1: order <= 32, NW = 2: order <= 64); thread `tid` of the team owns
// ROW tid of the front and keeps it in registers as a sliding window: after pivot k, a[j] holds column k+1+j.
// Per pivot the only shared-memory traffic is the broadcast of the pivot column (double-buffered, one team barrier).
// Global loads are issued in independent batches (memory-level parallelism instead of one L2 round trip per element).
#define B2_STAMP(idx) do { if (prof) { team_sync<NW>(team); if (tid == 0) prof[idx] = clock64(); } } while (0)

template <int NW>
__device__ __forceinline__ void team_sync(int team) {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(NW * 32) : "memory");
}

// Split team barrier of the pivot loop (front_factor_team): every warp ARRIVES on its own named barrier and WAITS on its
// partner's, ids alternating with the pivot parity -- bar.arrive + bar.sync by disjoint warps, 64 threads per barrier.  A warp
// can only arrive for pivot k+2 after its partner has left the wait of pivot k, so two ids per warp are enough.
// ids: 1..2 = team_sync; 3 + team*4 + warp*2 + parity (<= 10 of the 16 hardware barriers).
template <int NW>
__device__ __forceinline__ void team_arrive(int team, int warp, int parity) {
    static_assert(NW <= 2, "split barrier is written for one- and two-warp teams");
    if (NW == 2) asm volatile("bar.arrive %0, 64;" ::"r"(3 + team * 4 + warp * 2 + parity) : "memory");
}
template <int NW>
__device__ __forceinline__ void team_wait(int team, int warp, int parity) {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, 64;" ::"r"(3 + team * 4 + (warp ^ 1) * 2 + parity) : "memory");
}

// shared-memory slice of one team:
//   F[maxf*maxf] assembly area, later the finished panel | colbuf[4][FMAX+4] | rel[MAXC][FMAX] ints | recs[MAXC] |
//   stage[STAGE] children's update blocks landed by cp.async
constexpr int MAXC = 8;                                // children staged per round
template <int NW>
struct TeamSmem {
    static constexpr int FMAX = 32 * NW;
    static constexpr int STAGE = (NW == 1) ? 1024 : 4096;   // doubles; one child always fits (rc^2 <= (FMAX-1)^2)
    static __host__ __device__ int fsize(int maxf) { return (maxf * maxf + 1) & ~1; }   // keeps `stage` 16-byte aligned
    static __host__ __device__ int doubles(int maxf) {
        return fsize(maxf) + 4 * (FMAX + 4) + (MAXC * FMAX) / 2 + MAXC * 4 + STAGE;
    }
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {      // L2-only: data written by other CTAs
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// dependency flags of the single-launch ("dependency-driven") schedule
__device__ __forceinline__ void flag_wait(const int* flag, int* err) {
    int v = 0;
    unsigned it = 0;
    do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v) break;
        __nanosleep(32);
    } while (++it < (1u << 24));
    if (!v) atomicExch(err, 1);          // bounded spin: never hang the device, report instead
}
__device__ __forceinline__ void flag_set(int* flag) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One pivot of front_factor_team's software-pipelined loop, for a window of NB live 8-column blocks: the latency chain of pivot k
// (d_k -> reciprocal (approx + 2 Newton steps) -> l_k -> column k+1 -> publish -> arrive) and the pending row update of pivot k-1,
// as one branch-free block.  The publish is a predicated st.shared + bar.arrive in ONE asm without a memory clobber, so that most
// of the pending update (which reads the OTHER column buffers) follows the arrive and overlaps the partner warp's barrier latency.
// The empty volatile asm statements (B2_TIE) keep NVVM from sinking the whole chain below the update; ptxas then issues the six tied
// FMA pairs under the latency of the d_k load and runs the chain contiguously.  (Measured, profiles/r02_pivot_loop.txt: forcing a
// finer interleave with real data dependencies -- one LOP3 per link -- costs more issue slots than the latency it hides.)
struct PivotCtx { const double* cb; const double* pb; double* nb; double* Fk; double eps; int tid, k, f, team; };
#define B2_TIE(c, x, y) asm volatile("" : "+d"(c), "+d"(x), "+d"(y))
template <int NW, int P>
__device__ __forceinline__ void pending_pair(double (&av)[32 * NW + 1], double lp, const double2* up) {
    const double2 u = up[P];
    av[2 * P - 1] = fma(-lp, u.x, av[2 * P]);
    av[2 * P] = fma(-lp, u.y, av[2 * P + 1]);
}
template <int NW, int P0, int P1>
struct PendingRange {
    static __device__ __forceinline__ void run(double (&av)[32 * NW + 1], double lp, const double2* up) {
        if constexpr (P0 < P1) { pending_pair<NW, P0>(av, lp, up); PendingRange<NW, P0 + 1, P1>::run(av, lp, up); }
    }
};
template <int NW, int NB>
__device__ __forceinline__ void pivot_iter(double (&av)[32 * NW + 1], double& lp, const PivotCtx& c, int& nneg, int& npert) {
    constexpr int NP = 4 * NB;                          // pending pairs 1 .. NP-1 (pair 0 is column k+1, handled with the chain)
    double dk = c.cb[1];
    const double u0 = c.cb[2];
    const double2* up = reinterpret_cast<const double2*>(c.pb + 2);
    const bool tiny = !(fabs(dk) >= c.eps), neg = dk < 0.0;
    dk = tiny ? (neg ? -c.eps : c.eps) : dk;
    npert += (c.tid == 0 && tiny) ? 1 : 0;
    nneg += (c.tid == 0 && !tiny && neg) ? 1 : 0;
    double a1 = fma(-lp, up[0].y, av[1]);               // column k+1 with pivot k-1 applied
    double r, e;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dk));
    if constexpr (NP > 1) { pending_pair<NW, 1>(av, lp, up); B2_TIE(r, av[1], av[2]); }
    e = fma(-dk, r, 1.0);
    if constexpr (NP > 2) { pending_pair<NW, 2>(av, lp, up); B2_TIE(e, av[3], av[4]); }
    r = fma(r, e, r);
    if constexpr (NP > 3) { pending_pair<NW, 3>(av, lp, up); B2_TIE(r, av[5], av[6]); }
    e = fma(-dk, r, 1.0);
    if constexpr (NP > 4) { pending_pair<NW, 4>(av, lp, up); B2_TIE(e, av[7], av[8]); }
    r = fma(r, e, r);
    if constexpr (NP > 5) { pending_pair<NW, 5>(av, lp, up); B2_TIE(r, av[9], av[10]); }
    double l = av[0] * r;
    if constexpr (NP > 6) { pending_pair<NW, 6>(av, lp, up); B2_TIE(l, av[11], av[12]); }
    // column k+1 with pivot k applied -- the value the next pivot waits for (junk, unused, when k+1 = f)
    double nx = fma(-l, u0, a1);
    double* dst = c.nb + (c.tid - c.k);                // row tid of column k+1 at nb[tid - (k+1) + 1]
    if (NW == 1) {
        if (c.tid > c.k) *dst = nx;
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\t@p st.shared.f64 [%0], %1;\n\tbar.arrive %4, 64;\n\t}"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst)), "d"(nx), "r"(c.tid), "r"(c.k),
                       "r"(3 + c.team * 4 + (c.tid >> 5) * 2 + ((c.k + 1) & 1)));
    }
    if (c.tid < c.f) c.Fk[c.tid] = (c.tid == c.k) ? dk : l;   // finished column k of the panel (rows < k: scratch)
    av[0] = nx;
    PendingRange<NW, 7, NP>::run(av, lp, up);
    lp = l;
}

template <int NW, bool DEP = false>
__device__ __forceinline__ void front_factor_team(const FactorArgs& a, const ChildRec* childrec, int s, double* sm_team,
                                                  int tid, int team, int maxf, int& nneg, int& npert, long long* prof = nullptr,
                                                  int* done = nullptr, int* err = nullptr) {
    constexpr int FMAX = 32 * NW, TEAM = 32 * NW, STAGE = TeamSmem<NW>::STAGE;
    double* F = sm_team;
    constexpr int CBS = FMAX + 4;                      // four pivot-column buffers
    double* colbuf = F + TeamSmem<NW>::fsize(maxf);    // [4][CBS]
    int* relst = (int*)(colbuf + 4 * CBS);             // [MAXC][FMAX]
    ChildRec* recs = (ChildRec*)(relst + MAXC * FMAX); // [MAXC]
    double* stage = (double*)(recs + MAXC);            // [STAGE]
    B2_STAMP(0);
    if (a.ftrace && tid == 0) a.ftrace[3 * (size_t)s] = global_ns();
    const FrontDesc d = a.desc[s];
    const int f = d.f, w = d.w, r = f - w;
    B2_STAMP(1);
    for (int i = tid; i < f * f; i += TEAM) F[i] = 0.0;
    for (int i = tid; i < 4 * CBS; i += TEAM) colbuf[i] = 0.0;
    team_sync<NW>(team);
    B2_STAMP(2);
    {   // original matrix entries: panel layout pos + col*f == assembly-area layout
        const int32_t* src = a.amap_src + d.amap_off;
        const int32_t* dst = a.amap_dst + d.amap_off;
        for (int t0 = 0; t0 < d.amap_cnt; t0 += 4 * TEAM) {
            int sd[4], dd[4]; double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + tid + u * TEAM;
                sd[u] = (t < d.amap_cnt) ? src[t] : -1;
                dd[u] = (t < d.amap_cnt) ? dst[t] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (sd[u] >= 0) ? __ldg(a.A + sd[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (sd[u] >= 0) F[dd[u]] = v[u];
        }
    }
    team_sync<NW>(team);
    B2_STAMP(3);
    // ---- extend-add of the children (ascending id): rounds of <= MAXC children whose update blocks fit the stage;
    //      every block of a round is landed in shared memory by cp.async in ONE memory round trip.
    // Dependency-driven schedule: the children are still added in ascending id (the summation order is part of the result), but a
    // round does not wait for all of them -- it waits for the next child and takes along those behind it that have ALREADY finished,
    // so a parent whose last child is late has staged and added the others by then (the tree's critical path pays one round, not
    // the whole extend-add: tools/trace_sparse.py).
    int* nready_sh = relst + FMAX - 1;                 // (row 0 of relst holds at most FMAX - 1 indices: this slot is free)
    for (int c0 = 0; c0 < d.nchild;) {
        const int cb0 = (c0 / MAXC) * MAXC;            // children are fetched in blocks of MAXC records
        const int ro = c0 - cb0;                       // first record of this round inside the block
        if (ro == 0) {
            if (tid < min(MAXC, d.nchild - cb0)) recs[tid] = childrec[d.child_off + cb0 + tid];
            team_sync<NW>(team);
        }
        int nrec = min(MAXC, d.nchild - cb0) - ro;
        if (DEP) {
            if (tid < 32) {
                if (tid == 0) flag_wait(done + recs[ro].sn, err);          // child front finished (its update block is in L2)
                __syncwarp();
                int v = (tid == 0) ? 1 : 0;
                if (tid > 0 && tid < nrec) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(done + recs[ro + tid].sn) : "memory");
                const unsigned m = __ballot_sync(0xffffffffu, v != 0);
                if (tid == 0) *nready_sh = __ffs(~m) - 1;                  // length of the finished prefix (>= 1)
            }
            team_sync<NW>(team);
            nrec = *nready_sh;
        }
        const ChildRec* rrec = recs + ro;
        int nc = 0, tot = 0;
        while (nc < nrec) {
            const int sz = (rrec[nc].rc * rrec[nc].rc + 1) & ~1;
            if (nc > 0 && tot + sz > STAGE) break;
            tot += sz; ++nc;
        }
        int off = 0;
        for (int c = 0; c < nc; ++c) {
            const int rc = rrec[c].rc;
            const double* CB = a.ws + rrec[c].cb_off;          // 16-byte aligned, padded to an even count
            const int sz = (rc * rc + 1) & ~1;
            if (DEP) { for (int e = 2 * tid; e < sz; e += 2 * TEAM) cp_async16_cg(stage + off + e, CB + e); }
            else { for (int e = tid; e < rc * rc; e += TEAM) cp_async8(stage + off + e, CB + e); }
            if (tid < rc) cp_async4(relst + c * FMAX + tid, a.rel + rrec[c].rel_off + tid);
            off += sz;
        }
        cp_async_wait_all();
        team_sync<NW>(team);
        off = 0;
        for (int c = 0; c < nc; ++c) {
            const int rc = rrec[c].rc;
            const int* rl = relst + c * FMAX;
            const double* cs = stage + off;
            if (tid < rc) {
                const int ri = rl[tid];
                for (int j0 = 0; j0 <= tid; j0 += 8) {         // row tid of the child block, 8 columns per batch
                    double v[8], g[8]; int ix[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int j = j0 + u; ix[u] = (j <= tid) ? ri + rl[j] * f : -1; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int j = j0 + u; v[u] = (ix[u] >= 0) ? cs[j * rc + tid] : 0.0; g[u] = (ix[u] >= 0) ? F[ix[u]] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (ix[u] >= 0) F[ix[u]] = g[u] + v[u];
                }
            }
            team_sync<NW>(team);
            off += (rc * rc + 1) & ~1;
        }
        c0 += nc;
    }
    B2_STAMP(4);
    if (a.ftrace && tid == 0) a.ftrace[3 * (size_t)s + 1] = global_ns();
    // row `tid` of the front into registers: a[j] = F(tid, j)
    double av[FMAX + 1];
#pragma unroll
    for (int j = 0; j < FMAX; ++j) av[j] = (j < f && tid < f) ? F[tid + j * f] : 0.0;
    av[FMAX] = 0.0;
    // Pivot loop, software-pipelined by one pivot.  The pivot column is published WINDOW-RELATIVE in one of four buffers --
    // cb[1] = d_k, cb[2 + j] = u(k+1+j) -- so that the register window lines up with 16-byte pairs whatever the parity of k: one
    // LDS.128 feeds two FMAs (shared-memory issue, not the FP64 pipe, bounds a lone warp: tools/microbench/fp64_pipes.cu).
    // A pivot has a latency chain (barrier -> d_k -> reciprocal -> l_k -> the one FMA that makes column k+1 -> publish) and a
    // throughput part (the other f-k-2 FMAs of the row).  Iteration k runs the chain of pivot k TOGETHER with the throughput
    // part of pivot k-1 (`pending`: multiplier lp, buffer pb), as straight-line code per number of live 8-column blocks so that
    // ptxas interleaves the two; the team barrier is split (arrive after the publish, wait at the top of the next iteration).
    // Every element still receives the pivots' updates in ascending order, one fma each: bit-identical to the plain loop.
    //   top of iteration k:  av[0] = column k (final), av[j] = column k+j with pivot k-1's update pending (j >= 1)
    //   four buffers: a warp past wait(k) reads (k-1)%4 and k%4 and writes (k+1)%4; its partner, at most one pivot ahead (it has
    //   this warp's arrive(k+1) but not arrive(k+2)), writes (k+2)%4.
    colbuf[tid + 1] = av[0];                           // column 0
    team_arrive<NW>(team, tid >> 5, 0);
    double lp = 0.0;                                   // multiplier of the pending pivot (none yet: the pass only shifts)
    int kb = 0, pbi = 3;
    for (int k = 0; k < w; ++k) {
        const double* cb = colbuf + kb * CBS;
        const double* pb = colbuf + pbi * CBS;
        pbi = kb;
        kb = (kb + 1) & 3;
        double* nb = colbuf + kb * CBS;
        double* Fk = F + k * f;
        const int nblk = (f - k + 7) >> 3;              // live 8-column blocks of the window (team-uniform, >= 1)
        team_wait<NW>(team, tid >> 5, k & 1);
        PivotCtx c{cb, pb, nb, Fk, a.eps, tid, k, f, team};
        switch (nblk) {
            case 1: pivot_iter<NW, 1>(av, lp, c, nneg, npert); break;
            case 2: pivot_iter<NW, 2>(av, lp, c, nneg, npert); break;
            case 3: pivot_iter<NW, 3>(av, lp, c, nneg, npert); break;
            case 4: pivot_iter<NW, 4>(av, lp, c, nneg, npert); break;
            case 5: pivot_iter<NW, (NW > 1 ? 5 : 4)>(av, lp, c, nneg, npert); break;
            case 6: pivot_iter<NW, (NW > 1 ? 6 : 4)>(av, lp, c, nneg, npert); break;
            case 7: pivot_iter<NW, (NW > 1 ? 7 : 4)>(av, lp, c, nneg, npert); break;
            default: pivot_iter<NW, (NW > 1 ? 8 : 4)>(av, lp, c, nneg, npert); break;
        }
    }
    team_wait<NW>(team, tid >> 5, w & 1);              // pairs with the last iteration's (unconditional) arrive
    {   // the last pivot's pending update (no shift): av[j] = column w+j, j >= 1; av[0] = column w is final
        const double2* up = reinterpret_cast<const double2*>(colbuf + pbi * CBS + 2);
        av[1] = fma(-lp, up[0].y, av[1]);
#pragma unroll
        for (int j0 = 0; j0 < FMAX; j0 += 8) {
            if (w + j0 < f) {
#pragma unroll
                for (int j = (j0 ? j0 : 2); j < j0 + 8; j += 2) {
                    const double2 u = up[j >> 1];
                    av[j] = fma(-lp, u.x, av[j]);
                    av[j + 1] = fma(-lp, u.y, av[j + 1]);
                }
            }
        }
    }
    B2_STAMP(5);
    // update block first (it is all the parent waits for): av[j] holds column w+j of row tid -- registers only, no barrier needed
    if (tid >= w && tid < f) {
        double* CBo = a.ws + d.cb_off;
        const int ir = tid - w;
#pragma unroll
        for (int j = 0; j < FMAX; ++j) if (j <= ir) CBo[(size_t)j * r + ir] = av[j];
    }
    // hand-off: the team barrier orders every thread's stores before thread 0's st.release.gpu (release is cumulative over the
    // barrier's synchronises-with edge -- the CUTLASS semaphore pattern), so no team-wide __threadfence() is needed.  The same
    // barrier completes the pivot loop's writes of the panel columns into F.
    team_sync<NW>(team);
    if (DEP && tid == 0) flag_set(done + s);
    if (a.ftrace && tid == 0) a.ftrace[3 * (size_t)s + 2] = global_ns();
    B2_STAMP(6);
    // panel, off the tree's critical path: column-major (forward solve, parent-independent) and row-major copy (backward solve)
    {
        double* Lp = a.L + d.lp_off;
        for (int e = tid; e < f * w; e += TEAM) Lp[e] = F[e];
        if (tid < f) {
            double* Lt = a.Lt + d.lp_off + (size_t)tid * w;
            for (int k = 0; k < w; ++k) Lt[k] = F[tid + k * f];
        }
        if (tid < w) a.dvec[d.col0 + tid] = F[tid + tid * f];
    }
    team_sync<NW>(team);                               // (a team of the staged kernels reuses F for its next front)
    B2_STAMP(7);
}

// debug: re-factor ONE front with clock64() stamps at the phase boundaries (children's update blocks must be valid)
template <int NW>
__global__ void k_factor_team_profile(FactorArgs a, const ChildRec* childrec, int sn, int maxf, long long* prof, int reps) {
    extern __shared__ __align__(16) double sm[];
    int nneg = 0, npert = 0;
    for (int r = 0; r < reps; ++r) front_factor_team<NW>(a, childrec, sn, sm, threadIdx.x, 0, maxf, nneg, npert, prof + 8 * r);
}

// teams per CTA: FW_WARPS one-warp teams, or ONE two-warp team (its staging area is large)
template <int NW> struct TeamsPerCta { static constexpr int value = (NW == 1) ? FW_WARPS : 1; };

template <int NW>
__global__ void __launch_bounds__(TeamsPerCta<NW>::value * NW * 32) k_factor_warp(FactorArgs a, const ChildRec* childrec, WarpSched ws, int maxf) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NTEAM = TeamsPerCta<NW>::value;
    const int team = threadIdx.x / (32 * NW), tid = threadIdx.x % (32 * NW);
    double* smt = sm + (size_t)team * TeamSmem<NW>::doubles(maxf);
    const int s0 = ws.cta_ptr[blockIdx.x], s1 = ws.cta_ptr[blockIdx.x + 1];
    int nneg = 0, npert = 0;
    for (int st = s0; st < s1; ++st) {
        const int off = ws.stage_off[st], cnt = ws.stage_cnt[st];
        for (int q = team; q < cnt; q += NTEAM) front_factor_team<NW>(a, childrec, ws.list[off + q], smt, tid, team, maxf, nneg, npert);
        if (s1 - s0 > 1) __syncthreads();
    }
    if (tid == 0) {
        if (nneg) atomicAdd(a.counters + 0, nneg);
        if (npert) atomicAdd(a.counters + 1, npert);
    }
}

// ------------------------------------------------------------------------------------------------ solves
// Programmatic dependent launch: a level's kernel is launched while the previous level still runs.  Everything that is
// CONSTANT during a solve (descriptors, index lists, the factor panels) is fetched before pdl_wait(); only the values the
// previous levels produce (xp, cbv) are read after it.  Both calls are no-ops for a launch without the attribute.
// (pdl_trigger / pdl_wait: common.cuh)

// A front's solve is three memory round trips, whatever its number of children or pivots:
//   (1) descriptor;  (2) child records + the whole panel (cp.async into shared memory, in flight while (3) runs) + own
//   right-hand side;  (3) every child's relative indices and contribution vector at once.
// Forward:  thread = row i of the front; y_i in a register; the recurrence y_i -= L(i,k) y_k reads L from the staged
//           column-major panel and y_k from a shuffle (one-warp teams) or a double-buffered shared slot (two-warp teams).
// Backward: thread = pivot column j; t_j in a register; L(i,j) comes from the staged ROW-major copy `Lt`.
template <int NW>
struct SolveSmem {
    static constexpr int FMAX = 32 * NW;
    // ys/xs [FMAX] | slots [16] | recs [MAXC] (4 doubles each) | panel [FMAX*FMAX]
    static constexpr int doubles = FMAX + 16 + 4 * MAXC + FMAX * FMAX;
};

template <int NW, bool DEP = false>
__device__ __forceinline__ void front_fwd_team(const SolveArgs& a, const ChildRec* childrec, int s, double* sm_team, int tid, int team,
                                               int* done = nullptr, int* err = nullptr) {
    constexpr int FMAX = 32 * NW, TEAM = 32 * NW;
    double* ys = sm_team;                              // [FMAX] assembly of the front's rhs
    double* yb = ys + FMAX;                            // [2][8] broadcast slots
    ChildRec* recs = (ChildRec*)(yb + 16);             // [MAXC]
    double* P = (double*)(recs + MAXC);                // panel, column-major, ld f
    const FrontDesc d = a.desc[s];
    const int f = d.f, w = d.w;
    {
        const double* Lp = a.L + d.lp_off;
        for (int e = tid; e < f * w; e += TEAM) cp_async8(P + e, Lp + e);
    }
    for (int c0 = 0; c0 < max(d.nchild, 1); c0 += MAXC) {
        const int nc = min(MAXC, d.nchild - c0);
        if (tid < nc) {
            recs[tid] = childrec[d.child_off + c0 + tid];
            if (DEP) flag_wait(done + recs[tid].sn, err);
        }
        team_sync<NW>(team);
        int tg[MAXC]; double vv[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) tg[c] = (c < nc && tid < recs[c].rc) ? a.rel[recs[c].rel_off + tid] : -1;
        if (c0 == 0) {                                  // (constant data is in flight; now the values of the levels below)
            if (!DEP) pdl_wait();
            ys[tid] = (tid < w) ? a.xp[d.col0 + tid] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            vv[c] = (tg[c] >= 0) ? (DEP ? __ldcg(a.cbv + recs[c].cbv_off + tid) : a.cbv[recs[c].cbv_off + tid]) : 0.0;
        if (c0 == 0) team_sync<NW>(team);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {               // ascending child order: deterministic sums
            if (c < nc) {
                if (tg[c] >= 0) ys[tg[c]] += vv[c];
                team_sync<NW>(team);
            }
        }
    }
    cp_async_wait_all();
    team_sync<NW>(team);
    double y = ys[tid];
    if (NW == 1) {
        for (int k0 = 0; k0 < w; k0 += 8) {
            double l[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int k = k0 + u; l[u] = (k < w && tid > k && tid < f) ? P[k * f + tid] : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                if (k < w) y = fma(-l[u], __shfl_sync(0xffffffffu, y, k), y);      // team-uniform
            }
        }
    } else {
        // two-warp team, blocked by warp: warp 0 eliminates pivots 0..31 among its own rows with shuffles and publishes
        // them; warp 1 applies them in one parallel pass, then eliminates pivots 32.. among its rows -- two barriers per
        // front instead of one per pivot
        const int w0 = min(w, 32);
        if (tid < 32) {
            for (int k0 = 0; k0 < w0; k0 += 8) {
                double l[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int k = k0 + u; l[u] = (k < w0 && tid > k && tid < f) ? P[k * f + tid] : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 + u;
                    if (k < w0) y = fma(-l[u], __shfl_sync(0xffffffffu, y, k), y);
                }
            }
            if (tid < w0) ys[tid] = y;
        }
        team_sync<NW>(team);
        if (tid >= 32) {
            if (tid < f) {
                double acc = 0.0;
#pragma unroll 8
                for (int k = 0; k < w0; ++k) acc = fma(P[k * f + tid], ys[k], acc);
                y -= acc;
            }
            for (int k0 = 32; k0 < w; k0 += 8) {
                double l[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int k = k0 + u; l[u] = (k < w && tid > k && tid < f) ? P[k * f + tid] : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 + u;
                    if (k < w) y = fma(-l[u], __shfl_sync(0xffffffffu, y, k - 32), y);
                }
            }
        }
    }
    if (tid < f) { if (tid < w) a.xp[d.col0 + tid] = y; else a.cbv[a.cbv_off[s] + tid - w] = y; }
    // hand-off: the team barrier orders every thread's stores before thread 0's st.release.gpu (release is cumulative over the
    // barrier's synchronises-with edge -- the CUTLASS semaphore pattern), so no team-wide __threadfence() is needed
    team_sync<NW>(team);
    if (DEP && tid == 0) flag_set(done + s);
}

template <int NW, bool DEP = false>
__device__ __forceinline__ void front_bwd_team(const SolveArgs& a, int s, double* sm_team, int tid, int team,
                                               int* done = nullptr, int* err = nullptr, const int32_t* parent = nullptr) {
    constexpr int FMAX = 32 * NW, TEAM = 32 * NW;
    double* xs = sm_team;                              // [FMAX] gathered ancestor values
    double* xb = xs + FMAX;                            // [2][8]
    double* P = xb + 16 + 4 * MAXC;                    // row-major f x w panel (same slice layout as the forward sweep)
    const FrontDesc d = a.desc[s];
    const int f = d.f, w = d.w, r = f - w;
    {
        const double* Lt = a.Lt + d.lp_off;
        for (int e = tid; e < f * w; e += TEAM) cp_async8(P + e, Lt + e);
    }
    const int32_t* rows = a.rows + d.rows_off + w;
    const int myrow = (tid < r) ? rows[tid] : 0;
    const double dinv = (tid < w) ? fast_rcp(a.dvec[d.col0 + tid]) : 0.0;
    if (DEP) {      // all ancestors are final once the parent is
        if (tid == 0) { const int p = parent[s]; if (p >= 0) flag_wait(done + p, err); }
        team_sync<NW>(team);
    } else pdl_wait();
    if (tid < r) xs[tid] = DEP ? __ldcg(a.xp + myrow) : a.xp[myrow];
    double t = (tid < w) ? a.xp[d.col0 + tid] * dinv : 0.0;
    cp_async_wait_all();
    team_sync<NW>(team);
    {   // t_j -= sum_{i >= w} L(i,j) x_i : no recurrence
        double acc = 0.0;
        if (tid < w) {
#pragma unroll 8
            for (int i = 0; i < r; ++i) acc = fma(P[(w + i) * w + tid], xs[i], acc);
        }
        t -= acc;
    }
    // back-substitution with L11': x_k final -> t_j -= L(k,j) x_k for j < k
    if (NW == 1) {
        for (int k0 = w - 1; k0 >= 1; k0 -= 8) {
            double l[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int k = k0 - u; l[u] = (k >= 1 && tid < k) ? P[k * w + tid] : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 - u;
                if (k >= 1) t = fma(-l[u], __shfl_sync(0xffffffffu, t, k), t);
            }
        }
    } else {
        // blocked by warp (see the forward sweep): warp 1 finishes columns 32.. with shuffles and publishes them in
        // xs[32..] (the gathered ancestors occupy xs[0 .. f-w) with f - w < 32 here); warp 0 applies them in one pass
        if (w > 32) {
            if (tid >= 32) {
                for (int k0 = w - 1; k0 >= 33; k0 -= 8) {
                    double l[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int k = k0 - u; l[u] = (k >= 33 && tid < k) ? P[k * w + tid] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = k0 - u;
                        if (k >= 33) t = fma(-l[u], __shfl_sync(0xffffffffu, t, k - 32), t);
                    }
                }
                if (tid < w) xs[tid] = t;
            }
            team_sync<NW>(team);
            if (tid < 32) {
                double acc = 0.0;
#pragma unroll 8
                for (int k = 32; k < w; ++k) acc = fma(P[k * w + tid], xs[k], acc);
                t -= acc;
            }
        }
        if (tid < 32) {
            for (int k0 = min(w, 32) - 1; k0 >= 1; k0 -= 8) {
                double l[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int k = k0 - u; l[u] = (k >= 1 && tid < k) ? P[k * w + tid] : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 - u;
                    if (k >= 1) t = fma(-l[u], __shfl_sync(0xffffffffu, t, k), t);
                }
            }
        }
    }
    if (tid < w) a.xp[d.col0 + tid] = t;
    // hand-off: the team barrier orders every thread's stores before thread 0's st.release.gpu (release is cumulative over the
    // barrier's synchronises-with edge -- the CUTLASS semaphore pattern), so no team-wide __threadfence() is needed
    team_sync<NW>(team);
    if (DEP && tid == 0) flag_set(done + s);
}

// NTEAM teams per CTA.  Measured on OPF-10k (profiles/r02_sweep.txt): sweeping the fused bottom subtrees with 8 one-warp teams per
// CTA instead of 4 is SLOWER (0.139 -> 0.164 ms per solve: fewer resident CTAs, wider barriers), while smaller subtrees
// (fuse_max_fronts 16 -> 8) are faster (0.150 -> 0.139 ms) -- so the fused launches keep TeamsPerCta teams.
constexpr int SOLVE_FUSED_TEAMS = 4;
template <int NW, int NTEAM = TeamsPerCta<NW>::value>
__global__ void __launch_bounds__(NTEAM * NW * 32) k_fwd_warp2(SolveArgs a, const ChildRec* childrec, WarpSched ws) {
    extern __shared__ __align__(16) double smd[];
    double (*sm)[SolveSmem<NW>::doubles] = (double (*)[SolveSmem<NW>::doubles])smd;
    const int team = threadIdx.x / (32 * NW), tid = threadIdx.x % (32 * NW);
    pdl_trigger();
    const int s0 = ws.cta_ptr[blockIdx.x], s1 = ws.cta_ptr[blockIdx.x + 1];
    for (int st = s0; st < s1; ++st) {
        const int off = ws.stage_off[st], cnt = ws.stage_cnt[st];
        for (int q = team; q < cnt; q += NTEAM) front_fwd_team<NW>(a, childrec, ws.list[off + q], sm[team], tid, team);
        if (s1 - s0 > 1) __syncthreads();
    }
    pdl_wait();                                         // (idle teams too: this grid's completion must imply its predecessor's)
}

template <int NW, int NTEAM = TeamsPerCta<NW>::value>
__global__ void __launch_bounds__(NTEAM * NW * 32) k_bwd_warp2(SolveArgs a, WarpSched ws) {
    extern __shared__ __align__(16) double smd[];
    double (*sm)[SolveSmem<NW>::doubles] = (double (*)[SolveSmem<NW>::doubles])smd;
    const int team = threadIdx.x / (32 * NW), tid = threadIdx.x % (32 * NW);
    pdl_trigger();
    const int s0 = ws.cta_ptr[blockIdx.x], s1 = ws.cta_ptr[blockIdx.x + 1];
    for (int st = s1 - 1; st >= s0; --st) {
        const int off = ws.stage_off[st], cnt = ws.stage_cnt[st];
        for (int q = team; q < cnt; q += NTEAM) front_bwd_team<NW>(a, ws.list[off + q], sm[team], tid, team);
        if (s1 - s0 > 1) __syncthreads();
    }
    pdl_wait();
}


// ------------------------------------------------------------------------------------------------ single-launch schedule
// The whole (team-class) elimination tree in ONE launch per sweep: CTAs are issued in topological order, a front waits on
// its children's completion flags (acquire loads on global memory, bounded spin) instead of on a kernel boundary.  A CTA
// of 128 threads runs a GROUP of tasks: four one-warp teams (fronts of order <= 32) or one two-warp team (order <= 64).
struct DepSched {
    const int32_t* grp_type;   // 1 or 2 (warps per team)
    const int32_t* grp_ptr;    // [ngroup+1] into tasks
    const int32_t* tasks;      // supernode ids in topological (level, id) order
    int ngroup;
};

// Forward progress: a CTA only ever waits for groups that come EARLIER in the topological order.  Groups are therefore claimed
// through an atomic ticket instead of blockIdx.x: whichever CTA the hardware starts first takes the earliest unclaimed group, so a
// running CTA never waits for work that no running (or finished) CTA owns -- independent of the block dispatch order, which the
// programming model does not specify (the bounded spin in flag_wait stays as a second line of defence).
__device__ __forceinline__ int claim_group(int* ticket, int ngroup) {
    __shared__ int g_sh;
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        if (t == ngroup - 1) atomicExch(ticket, 0);     // last claim of this launch: re-arm the counter for the next one
        g_sh = t;
    }
    __syncthreads();
    return g_sh;
}

__global__ void __launch_bounds__(128) k_factor_dep(FactorArgs a, const ChildRec* childrec, DepSched ds, int maxf1, int maxf2,
                                                    int* done, int* err, int* ticket) {
    extern __shared__ __align__(16) double sm[];
    const int g = claim_group(ticket, ds.ngroup);
    const int type = ds.grp_type[g], t0 = ds.grp_ptr[g], n = ds.grp_ptr[g + 1] - t0;
    int nneg = 0, npert = 0;
    if (type == 1) {
        const int team = threadIdx.x >> 5, tid = threadIdx.x & 31;
        if (team < n) front_factor_team<1, true>(a, childrec, ds.tasks[t0 + team], sm + (size_t)team * TeamSmem<1>::doubles(maxf1),
                                                 tid, team, maxf1, nneg, npert, nullptr, done, err);
        if (tid == 0) { if (nneg) atomicAdd(a.counters + 0, nneg); if (npert) atomicAdd(a.counters + 1, npert); }
    } else {
        const int team = threadIdx.x >> 6, tid = threadIdx.x & 63;
        if (team < n) front_factor_team<2, true>(a, childrec, ds.tasks[t0 + team], sm, tid, team, maxf2, nneg, npert, nullptr, done, err);
        if (tid == 0 && team < n) { if (nneg) atomicAdd(a.counters + 0, nneg); if (npert) atomicAdd(a.counters + 1, npert); }
    }
}

__global__ void __launch_bounds__(128) k_fwd_dep(SolveArgs a, const ChildRec* childrec, DepSched ds, int* done, int* err, int* ticket) {
    extern __shared__ __align__(16) double smd[];       // max(4 one-warp slices, 1 two-warp slice)
    double (*sm)[SolveSmem<1>::doubles] = (double (*)[SolveSmem<1>::doubles])smd;
    const int g = claim_group(ticket, ds.ngroup);
    const int type = ds.grp_type[g], t0 = ds.grp_ptr[g], n = ds.grp_ptr[g + 1] - t0;
    if (type == 1) {
        const int team = threadIdx.x >> 5, tid = threadIdx.x & 31;
        if (team < n) front_fwd_team<1, true>(a, childrec, ds.tasks[t0 + team], sm[team], tid, team, done, err);
    } else {
        const int team = threadIdx.x >> 6, tid = threadIdx.x & 63;
        if (team < n) front_fwd_team<2, true>(a, childrec, ds.tasks[t0 + team], smd, tid, team, done, err);
    }
}

__global__ void __launch_bounds__(128) k_bwd_dep(SolveArgs a, DepSched ds, const int32_t* parent, int* done, int* err, int* ticket) {
    extern __shared__ __align__(16) double smd[];
    double (*sm)[SolveSmem<1>::doubles] = (double (*)[SolveSmem<1>::doubles])smd;
    const int g = ds.ngroup - 1 - claim_group(ticket, ds.ngroup);   // reverse topological order
    const int type = ds.grp_type[g], t0 = ds.grp_ptr[g], n = ds.grp_ptr[g + 1] - t0;
    if (type == 1) {
        const int team = threadIdx.x >> 5, tid = threadIdx.x & 31;
        if (team < n) front_bwd_team<1, true>(a, ds.tasks[t0 + team], sm[team], tid, team, done, err, parent);
    } else {
        const int team = threadIdx.x >> 6, tid = threadIdx.x & 63;
        if (team < n) front_bwd_team<2, true>(a, ds.tasks[t0 + team], smd, tid, team, done, err, parent);
    }
}

}  // namespace b2
