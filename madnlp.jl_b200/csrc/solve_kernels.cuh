// Triangular solves on the supernodal factor (replaces cuDSS "solve",
// lib/MadNLPGPU/ext/MadNLPGPUCUDAExt/cudss.jl:171-181, and dsytrs, src/LinearSolvers/lapack.jl:169-172): shared argument
// block + the permutation kernels.  The sweeps themselves live in warp_kernels.cuh (fronts of order <= 64) and
// bigsolve_kernels.cuh (larger fronts, dense solver).
//
// Multifrontal formulation: the forward sweep passes a contribution vector (length r) from each front to its
// parent (pull + fixed child order => deterministic, no atomics); the backward sweep gathers already-final
// ancestor values.  x lives in permuted order in `xp`.
#pragma once
#include "common.cuh"
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

__global__ void k_perm_in(int n, const int32_t* __restrict__ perm, const double* __restrict__ x, double* __restrict__ xp) {
    pdl_sync();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) xp[i] = x[perm[i]];
}
__global__ void k_perm_out(int n, const int32_t* __restrict__ perm, const double* __restrict__ xp, double* __restrict__ x) {
    pdl_sync();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[perm[i]] = xp[i];
}
// masked variant for the multi-GPU back-substitution: only rows this rank finalises are written, others zeroed
__global__ void k_perm_out_masked(int n, const int32_t* __restrict__ perm, const uint8_t* __restrict__ mask_p,
                                  const double* __restrict__ xp, double* __restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        x[perm[i]] = mask_p[i] ? xp[i] : 0.0;
}

}  // namespace b2
