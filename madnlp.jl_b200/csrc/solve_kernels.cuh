// Level-scheduled triangular solves on the supernodal factor (replaces cuDSS "solve",
// lib/MadNLPGPU/ext/MadNLPGPUCUDAExt/cudss.jl:171-181, and dsytrs, src/LinearSolvers/lapack.jl:169-172).
//
// Multifrontal formulation: the forward sweep passes a contribution vector (length r) from each front to its
// parent (pull + fixed child order => deterministic, no atomics); the backward sweep gathers already-final
// ancestor values.  x lives in permuted order in `xp`.
#pragma once
#include "front_kernels.cuh"

namespace b2 {

struct SolveArgs {
    const FrontDesc* desc;
    const int32_t* rows;
    const int32_t* child_idx;
    const int32_t* rel;
    const int64_t* cbv_off;   // per supernode offset of its contribution vector
    const double* L;
    const double* Lt;         // row-major panel copies (warp-class fronts)
    const double* dvec;
    double* xp;
    double* cbv;
};

constexpr int SOLVE_WARPS = 4;   // fronts per CTA in the warp-per-front kernels

// ---- forward, warp per front (f <= 64 or so; panel staged in shared memory)
__global__ void __launch_bounds__(SOLVE_WARPS * 32) k_fwd_warp(SolveArgs a, const int32_t* __restrict__ list, int nlist,
                                                               int smem_per_warp) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * SOLVE_WARPS + warp;
    if (idx >= nlist) return;
    const FrontDesc d = a.desc[list[idx]];
    const int f = d.f, w = d.w, r = f - w;
    double* P = sm + (size_t)warp * smem_per_warp;   // panel f*w
    double* y = P + f * w;                           // f
    const double* Lp = a.L + d.lp_off;
    for (int i = lane; i < f * w; i += 32) P[i] = Lp[i];
    for (int i = lane; i < f; i += 32) y[i] = (i < w) ? a.xp[d.col0 + i] : 0.0;
    __syncwarp();
    for (int c = 0; c < d.nchild; ++c) {
        const int cs = a.child_idx[d.child_off + c];
        const FrontDesc dc = a.desc[cs];
        const int rc = dc.f - dc.w;
        const int32_t* rl = a.rel + dc.rel_off;
        const double* cv = a.cbv + a.cbv_off[cs];
        for (int i = lane; i < rc; i += 32) y[rl[i]] += cv[i];
        __syncwarp();
    }
    for (int k = 0; k < w; ++k) {
        const double yk = y[k];
        for (int i = k + 1 + lane; i < f; i += 32) y[i] -= P[i + k * f] * yk;
        __syncwarp();
    }
    for (int i = lane; i < w; i += 32) a.xp[d.col0 + i] = y[i];
    double* cv = a.cbv + a.cbv_off[list[idx]];
    for (int i = lane; i < r; i += 32) cv[i] = y[w + i];
}

// ---- backward (with the D^{-1} scaling folded in), warp per front
__global__ void __launch_bounds__(SOLVE_WARPS * 32) k_bwd_warp(SolveArgs a, const int32_t* __restrict__ list, int nlist,
                                                               int smem_per_warp) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * SOLVE_WARPS + warp;
    if (idx >= nlist) return;
    const FrontDesc d = a.desc[list[idx]];
    const int f = d.f, w = d.w;
    double* P = sm + (size_t)warp * smem_per_warp;
    double* x = P + f * w;
    const double* Lp = a.L + d.lp_off;
    const int32_t* rows = a.rows + d.rows_off;
    for (int i = lane; i < f * w; i += 32) P[i] = Lp[i];
    for (int i = lane; i < f; i += 32)
        x[i] = (i < w) ? a.xp[d.col0 + i] / a.dvec[d.col0 + i] : a.xp[rows[i]];
    __syncwarp();
    for (int k = w - 1; k >= 0; --k) {
        double s = 0.0;
        for (int i = k + 1 + lane; i < f; i += 32) s = fma(P[i + k * f], x[i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) x[k] -= s;
        __syncwarp();
    }
    for (int i = lane; i < w; i += 32) a.xp[d.col0 + i] = x[i];
}

// ---- forward, CTA per front (any size; panel streamed from HBM, y in shared memory)
constexpr int SOLVE_CTA = 1024;   // one CTA streams the whole panel: many threads x unrolled independent loads = MLP
__global__ void __launch_bounds__(SOLVE_CTA) k_fwd_cta(SolveArgs a, const int32_t* __restrict__ list) {
    extern __shared__ double y[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = list[blockIdx.x];
    const FrontDesc d = a.desc[s];
    const int f = d.f, w = d.w, r = f - w;
    const double* Lp = a.L + d.lp_off;
    for (int i = tid; i < f; i += SOLVE_CTA) y[i] = (i < w) ? a.xp[d.col0 + i] : 0.0;
    __syncthreads();
    for (int c = 0; c < d.nchild; ++c) {
        const int cs = a.child_idx[d.child_off + c];
        const FrontDesc dc = a.desc[cs];
        const int rc = dc.f - dc.w;
        const int32_t* rl = a.rel + dc.rel_off;
        const double* cv = a.cbv + a.cbv_off[cs];
        for (int i = tid; i < rc; i += SOLVE_CTA) y[rl[i]] += cv[i];
        __syncthreads();
    }
    for (int k0 = 0; k0 < w; k0 += 32) {
        const int nb = min(32, w - k0);
        if (warp == 0) {               // triangular part of this block of pivots
            for (int k = k0; k < k0 + nb; ++k) {
                const double yk = y[k];
                const int i = k + 1 + lane;
                if (i < k0 + nb) y[i] -= Lp[(size_t)k * f + i] * yk;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = k0 + nb + tid; i < f; i += SOLVE_CTA) {
            double acc = y[i];
            const double* col = Lp + (size_t)k0 * f + i;
            if (nb == 32) {
#pragma unroll
                for (int k = 0; k < 32; ++k) acc = fma(-col[(size_t)k * f], y[k0 + k], acc);
            } else {
                for (int k = 0; k < nb; ++k) acc = fma(-col[(size_t)k * f], y[k0 + k], acc);
            }
            y[i] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < w; i += SOLVE_CTA) a.xp[d.col0 + i] = y[i];
    double* cv = a.cbv + a.cbv_off[s];
    for (int i = tid; i < r; i += SOLVE_CTA) cv[i] = y[w + i];
}

__global__ void __launch_bounds__(SOLVE_CTA) k_bwd_cta(SolveArgs a, const int32_t* __restrict__ list) {
    extern __shared__ double x[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = SOLVE_CTA / 32;
    const int s = list[blockIdx.x];
    const FrontDesc d = a.desc[s];
    const int f = d.f, w = d.w;
    const double* Lp = a.L + d.lp_off;
    const int32_t* rows = a.rows + d.rows_off;
    for (int i = tid; i < f; i += SOLVE_CTA)
        x[i] = (i < w) ? a.xp[d.col0 + i] / a.dvec[d.col0 + i] : a.xp[rows[i]];
    __syncthreads();
    const int nblk = (w + 31) / 32;
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * 32, nb = min(32, w - k0);
        // x[k] -= sum_{i >= k0+nb} L(i,k) x[i]   (one warp per pivot column, coalesced over i)
        for (int k = k0 + warp; k < k0 + nb; k += NW) {
            double sacc = 0.0;
            const double* col = Lp + (size_t)k * f;
            int i = k0 + nb + lane;
            for (; i + 7 * 32 < f; i += 8 * 32) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = col[i + 32 * u];
#pragma unroll
                for (int u = 0; u < 8; ++u) sacc = fma(v[u], x[i + 32 * u], sacc);
            }
            for (; i < f; i += 32) sacc = fma(col[i], x[i], sacc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) x[k] -= sacc;
        }
        __syncthreads();
        if (warp == 0) {               // triangular part, last pivot of the block first
            for (int k = k0 + nb - 1; k >= k0; --k) {
                const int i = k + 1 + lane;
                double v = (i < k0 + nb) ? Lp[(size_t)k * f + i] * x[i] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) x[k] -= v;
                __syncwarp();
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < w; i += SOLVE_CTA) a.xp[d.col0 + i] = x[i];
}

__global__ void k_perm_in(int n, const int32_t* __restrict__ perm, const double* __restrict__ x, double* __restrict__ xp) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) xp[i] = x[perm[i]];
}
__global__ void k_perm_out(int n, const int32_t* __restrict__ perm, const double* __restrict__ xp, double* __restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[perm[i]] = xp[i];
}
// masked variant for the multi-GPU back-substitution: only rows this rank finalises are written, others zeroed
__global__ void k_perm_out_masked(int n, const int32_t* __restrict__ perm, const uint8_t* __restrict__ mask_p,
                                  const double* __restrict__ xp, double* __restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        x[perm[i]] = mask_p[i] ? xp[i] : 0.0;
}

}  // namespace b2
