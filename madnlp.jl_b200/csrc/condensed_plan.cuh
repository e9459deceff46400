// b2_condensed_plan: the one-time product of build_condensed_aug_symbolic (src/KKT/Sparse/condensed.jl:201-301) -- shared by the
// host construction (assembly.cu: b2_condensed_symbolic) and the device construction (symbolic_device.cu:
// b2_condensed_symbolic_device); both produce identical contents.
#pragma once
#include <vector>

#include "common.cuh"

struct b2_condensed_plan {
    int32_t n = 0, m = 0;
    int64_t nnz_aug = 0, n_dptr = 0, n_hptr = 0, n_jptr = 0;
    std::vector<int32_t> colptr, rowval;
    b2::DevBuf<int32_t> hsrc, dsrc, tptr;   // per slot: H.nz index or -1, pr_diag index or -1, triple range
    b2::DevBuf<int4> trip;                  // (col, k, l, 0): D[col]*Jt.nz[k]*Jt.nz[l]
};
