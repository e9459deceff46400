// Multi-CTA triangular solves for HBM-resident fronts (order > 64) and the dense solver.
//
// The pivot columns of a front are cut into blocks of BS = 128.  After the factorisation, k_big_inv inverts every
// 128 x 128 unit-lower diagonal block once (in shared memory).  A sweep is then one launch per block:
//   forward  block b: every CTA forms y_b = Linv_b * rhs_b redundantly (128x128 GEMV out of shared memory), CTA 0 stores it,
//                     and each CTA updates its own 64 rows below:  y_i -= sum_k L(i,k) y_b(k)      (coalesced stream of L)
//   backward block b: every CTA forms x_b = Linv_b' * t_b, CTA 0 stores it, each CTA updates its 256 earlier pivot
//                     columns:  t_j -= sum_i L(b_i, j) x_b(i)      (right-looking: no cross-CTA reduction, deterministic)
// so a sweep over N = 4096 is 32 launches of ~60 CTAs instead of one CTA streaming 64 MB alone.
#pragma once
#include "solve_kernels.cuh"

namespace b2 {

constexpr int BS = 128;

struct BigSolveArgs {
    SolveArgs s;
    const double* Linv;          // inverted diagonal blocks, BS*BS doubles each (column-major, ld BS)
    const int64_t* linv_off;     // per supernode: offset of its first block in Linv (-1: not a big front)
    double* side;                // [n] solved pivot blocks are parked here: CTA 0 must not overwrite the block's right-hand
                                 // side in xp while the other CTAs of the same launch are still reading it
};

// ---- inversion of the unit-lower diagonal blocks (once per factorisation); grid (max blocks, nfronts), 128 threads
__global__ void __launch_bounds__(BS) k_big_inv(const FrontDesc* desc, const int32_t* __restrict__ list, const double* __restrict__ L,
                                                double* __restrict__ Linv, const int64_t* __restrict__ linv_off) {
    extern __shared__ double T[];                 // [BS][BS+1], T[i*(BS+1)+j] = row i, col j
    const int s = list[blockIdx.y];
    const FrontDesc d = desc[s];
    const int kb = blockIdx.x * BS;
    if (kb >= d.w) return;
    const int nb = min(BS, d.w - kb), f = d.f, tid = threadIdx.x;
    const double* Lp = L + d.lp_off;
    for (int c = 0; c < nb; ++c) {                // coalesced over rows
        if (tid < nb) T[tid * (BS + 1) + c] = (tid > c) ? Lp[(size_t)(kb + c) * f + kb + tid] : (tid == c ? 1.0 : 0.0);
    }
    __syncthreads();
    // in-place inversion, columns from the last to the first: X(i,j) = -( L(i,j) + sum_{j<k<i} X(i,k) L(k,j) )
    for (int j = nb - 2; j >= 0; --j) {
        double v = 0.0;
        if (tid > j && tid < nb) {
            v = T[tid * (BS + 1) + j];
            for (int k = j + 1; k < tid; ++k) v = fma(T[tid * (BS + 1) + k], T[k * (BS + 1) + j], v);
        }
        __syncthreads();
        if (tid > j && tid < nb) T[tid * (BS + 1) + j] = -v;
        __syncthreads();
    }
    double* out = Linv + linv_off[s] + (size_t)blockIdx.x * BS * BS;
    for (int c = 0; c < BS; ++c) out[(size_t)c * BS + tid] = (tid < nb && c < nb) ? T[tid * (BS + 1) + c] : 0.0;
}

// location of entry i (0..f) of the front's solve vector: pivots live in xp, the rest in the contribution vector
__device__ __forceinline__ double* yptr(const SolveArgs& a, const FrontDesc& d, int64_t cbv0, int i) {
    return (i < d.w) ? a.xp + d.col0 + i : a.cbv + cbv0 + (i - d.w);
}

// ---- forward init: zero the contribution vector, pull the children (ascending id); one CTA per front
__global__ void __launch_bounds__(1024) k_bs_fwd_init(BigSolveArgs b, const int32_t* __restrict__ list) {
    const SolveArgs& a = b.s;
    const int s = list[blockIdx.x];
    const FrontDesc d = a.desc[s];
    const int64_t cbv0 = a.cbv_off[s];
    const int r = d.f - d.w, tid = threadIdx.x;
    for (int i = tid; i < r; i += 1024) a.cbv[cbv0 + i] = 0.0;
    __syncthreads();
    for (int c = 0; c < d.nchild; ++c) {
        const int cs = a.child_idx[d.child_off + c];
        const FrontDesc dc = a.desc[cs];
        const int rc = dc.f - dc.w;
        const int32_t* rl = a.rel + dc.rel_off;
        const double* cv = a.cbv + a.cbv_off[cs];
        for (int i = tid; i < rc; i += 1024) *yptr(a, d, cbv0, rl[i]) += cv[i];
        __syncthreads();
    }
}

constexpr int BSF_ROWS = 64;     // rows per CTA in the forward update (x 4 k-groups = 256 threads)
__global__ void __launch_bounds__(256) k_bs_fwd(BigSolveArgs b, const int32_t* __restrict__ list, int blk) {
    extern __shared__ double sm[];                 // Linv block [BS][BS+1] | rhs[BS] | y[BS] | part[4][64]
    const SolveArgs& a = b.s;
    const int s = list[blockIdx.y];
    const FrontDesc d = a.desc[s];
    const int kb = blk * BS;
    if (kb >= d.w) return;
    const int nb = min(BS, d.w - kb), f = d.f, tid = threadIdx.x;
    const int row0 = kb + nb + blockIdx.x * BSF_ROWS;
    if (row0 >= f && blockIdx.x > 0) return;
    double* T = sm; double* rhs = T + BS * (BS + 1); double* yk = rhs + BS; double* part = yk + BS;
    const double* Li = b.Linv + b.linv_off[s] + (size_t)blk * BS * BS;
    for (int e = tid; e < BS * BS; e += 256) { const int c = e / BS, i = e - c * BS; T[i * (BS + 1) + c] = Li[e]; }
    if (tid < BS) rhs[tid] = (tid < nb) ? a.xp[d.col0 + kb + tid] : 0.0;
    __syncthreads();
    if (tid < BS) {
        double v = 0.0;
        for (int c = 0; c <= tid; ++c) v = fma(T[tid * (BS + 1) + c], rhs[c], v);
        yk[tid] = v;
        if (blockIdx.x == 0 && tid < nb) b.side[d.col0 + kb + tid] = v;
    }
    __syncthreads();
    const int ir = tid & (BSF_ROWS - 1), kg = tid / BSF_ROWS;       // 4 groups of 32 pivots
    const int i = row0 + ir;
    double acc = 0.0;
    if (i < f) {
        const double* col = a.L + d.lp_off + (size_t)(kb + kg * 32) * f + i;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) if (kg * 32 + k < nb) acc = fma(col[(size_t)k * f], yk[kg * 32 + k], acc);
    }
    part[kg * BSF_ROWS + ir] = acc;
    __syncthreads();
    if (kg == 0 && i < f) {
        const double tot = part[ir] + part[BSF_ROWS + ir] + part[2 * BSF_ROWS + ir] + part[3 * BSF_ROWS + ir];
        *yptr(a, d, a.cbv_off[s], i) -= tot;
    }
}

// ---- backward init: t_j = y_j / d_j - sum_{i >= w} L(i,j) x(rows_i)  (y_j parked in `side` by the forward sweep);
//      one warp per pivot column (lanes stride the rows: coalesced stream of the column), grid (ceil(w/8), nfronts)
__global__ void __launch_bounds__(256) k_bs_bwd_init(BigSolveArgs b, const int32_t* __restrict__ list) {
    const SolveArgs& a = b.s;
    const int s = list[blockIdx.y];
    const FrontDesc d = a.desc[s];
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= d.w) return;
    const int f = d.f;
    const double* __restrict__ col = a.L + d.lp_off + (size_t)j * f;
    const int32_t* __restrict__ rows = a.rows + d.rows_off;
    double acc = 0.0;
    int i = d.w + lane;
    for (; i + 96 < f; i += 128) {
        double l[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { l[u] = col[i + 32 * u]; x[u] = a.xp[rows[i + 32 * u]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = fma(l[u], x[u], acc);
    }
    for (; i < f; i += 32) acc = fma(col[i], a.xp[rows[i]], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) a.xp[d.col0 + j] = b.side[d.col0 + j] / a.dvec[d.col0 + j] - acc;
}

constexpr int BSB_COLS = 256;    // earlier pivot columns per CTA in the backward update
__global__ void __launch_bounds__(256) k_bs_bwd(BigSolveArgs b, const int32_t* __restrict__ list, int blk) {
    extern __shared__ double sm[];                 // Linv block [BS][BS+1] | t[BS] | x[BS]
    const SolveArgs& a = b.s;
    const int s = list[blockIdx.y];
    const FrontDesc d = a.desc[s];
    const int kb = blk * BS;
    if (kb >= d.w) return;
    const int nb = min(BS, d.w - kb), f = d.f, tid = threadIdx.x;
    const int col0 = blockIdx.x * BSB_COLS;
    if (col0 >= kb && blockIdx.x > 0) return;
    double* T = sm; double* tk = T + BS * (BS + 1); double* xk = tk + BS;
    const double* Li = b.Linv + b.linv_off[s] + (size_t)blk * BS * BS;
    for (int e = tid; e < BS * BS; e += 256) { const int c = e / BS, i = e - c * BS; T[i * (BS + 1) + c] = Li[e]; }
    if (tid < BS) tk[tid] = (tid < nb) ? a.xp[d.col0 + kb + tid] : 0.0;
    __syncthreads();
    if (tid < BS) {                                // x = Linv' * t : x_t = sum_{c >= t} Linv(c,t) t_c
        double v = 0.0;
        for (int c = tid; c < nb; ++c) v = fma(T[c * (BS + 1) + tid], tk[c], v);
        xk[tid] = v;
        if (blockIdx.x == 0 && tid < nb) b.side[d.col0 + kb + tid] = v;
    }
    __syncthreads();
    const int j = col0 + tid;
    if (j < kb) {
        const double* col = a.L + d.lp_off + (size_t)j * f + kb;
        double acc = 0.0;
#pragma unroll 8
        for (int i = 0; i < BS; ++i) if (i < nb) acc = fma(col[i], xk[i], acc);
        a.xp[d.col0 + j] -= acc;
    }
}

// ---- end of a front's backward sweep: solved pivots from `side` back into xp; grid (ceil(w/256), nfronts)
__global__ void __launch_bounds__(256) k_bs_bwd_finish(BigSolveArgs b, const int32_t* __restrict__ list) {
    const FrontDesc d = b.s.desc[list[blockIdx.y]];
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < d.w) b.s.xp[d.col0 + j] = b.side[d.col0 + j];
}

}  // namespace b2
