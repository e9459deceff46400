/*
 * CPU ORACLE (C part) -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's *sequential CPU loops* on the KKT hot path, and of the sparse LDL^T the
 * reference's `LDLSolver` calls.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs load this library
 * (through oracle/madnlp_oracle.py); nothing under madnlp.jl_b200/ does.
 *
 * Functions and what they follow (paths relative to MadNLP.jl @ e1028096, v0.10.1):
 *   okkt_transfer            src/matrixtools.jl:79-88          _transfer!:  nz .= 0; nz[map[k]] += V[k]
 *   okkt_condensed_coord     src/KKT/Sparse/condensed.jl:328-345  _build_condensed_aug_coord! (hess, diag, JtDJ passes)
 *   okkt_csc_mul_n / _t      SparseArrays mul!(y, A, x, a, b) / mul!(y, A', x, a, b) as called from
 *                            src/IPM/factorization.jl:231-237,303-324
 *   okkt_csc_symv_lower      mul!(y, Symmetric(A, :L), x, a, b)   (same call sites)
 *   okkt_ldl_*               the numeric kernel behind src/LinearSolvers/ldl.jl:16-40 (LDLSolver): LDLFactorizations.jl
 *                            v0.10 (Project.toml:26; NOT vendored in the reference tree), which is a Julia translation of
 *                            T. Davis, "Algorithm 849: a concise sparse Cholesky factorization package" (LDL, ACM TOMS 31(4),
 *                            2005): elimination tree + column counts (ldl_symbolic), up-looking numeric LDL^T without
 *                            pivoting (ldl_numeric), L / D / L' solves.  This file restates that published algorithm.
 *                            The fill-reducing permutation is an input (the reference gets it from AMD.jl).
 *
 * Compile:  gcc -O2 -ffp-contract=off -shared -fPIC -o libkkt_oracle.so kkt_oracle.c     (see oracle/Makefile)
 * -ffp-contract=off: sums must round exactly like Julia's scalar loops (no FMA contraction).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ assembly loops */
void okkt_transfer(int64_t nnz_csc, double* nz, int64_t L, const double* V, const int64_t* map) {
    for (int64_t i = 0; i < nnz_csc; ++i) nz[i] = 0.0;
    for (int64_t k = 0; k < L; ++k) nz[map[k]] += V[k];
}

/* hptr[(dst,src)] x nh, dptr[(dst,src)] x nd, jptr[(dst,col,k,l)] x nj; all int64 row-major */
void okkt_condensed_coord(int64_t nnz, double* nz, const double* pr_diag, const double* H_nz, const double* Jt_nz,
                          const double* diag_buffer, int64_t nd, const int64_t* dptr, int64_t nh, const int64_t* hptr,
                          int64_t nj, const int64_t* jptr) {
    for (int64_t i = 0; i < nnz; ++i) nz[i] = 0.0;
    for (int64_t q = 0; q < nh; ++q) nz[hptr[2 * q]] += H_nz[hptr[2 * q + 1]];
    for (int64_t q = 0; q < nd; ++q) nz[dptr[2 * q]] += pr_diag[dptr[2 * q + 1]];
    for (int64_t q = 0; q < nj; ++q) {
        const int64_t* t = jptr + 4 * q;
        nz[t[0]] += diag_buffer[t[1]] * Jt_nz[t[2]] * Jt_nz[t[3]];
    }
}

/* y = alpha*A*x + beta*y, A: nrow x ncol CSC */
void okkt_csc_mul_n(int32_t nrow, int32_t ncol, const int32_t* colptr, const int32_t* rowval, const double* nz,
                    const double* x, double* y, double alpha, double beta) {
    if (beta == 0.0) for (int32_t i = 0; i < nrow; ++i) y[i] = 0.0;
    else if (beta != 1.0) for (int32_t i = 0; i < nrow; ++i) y[i] *= beta;
    for (int32_t j = 0; j < ncol; ++j) {
        const double axj = alpha * x[j];
        for (int32_t p = colptr[j]; p < colptr[j + 1]; ++p) y[rowval[p]] += nz[p] * axj;
    }
}

/* y = alpha*A'*x + beta*y */
void okkt_csc_mul_t(int32_t nrow, int32_t ncol, const int32_t* colptr, const int32_t* rowval, const double* nz,
                    const double* x, double* y, double alpha, double beta) {
    (void)nrow;
    for (int32_t j = 0; j < ncol; ++j) {
        double t = 0.0;
        for (int32_t p = colptr[j]; p < colptr[j + 1]; ++p) t += nz[p] * x[rowval[p]];
        y[j] = (beta == 0.0) ? alpha * t : alpha * t + beta * y[j];
    }
}

/* y = alpha*Symmetric(A,:L)*x + beta*y, A: n x n lower CSC */
void okkt_csc_symv_lower(int32_t n, const int32_t* colptr, const int32_t* rowval, const double* nz, const double* x,
                         double* y, double alpha, double beta) {
    if (beta == 0.0) for (int32_t i = 0; i < n; ++i) y[i] = 0.0;
    else if (beta != 1.0) for (int32_t i = 0; i < n; ++i) y[i] *= beta;
    for (int32_t j = 0; j < n; ++j) {
        const double axj = alpha * x[j];
        double t = 0.0;
        for (int32_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            const int32_t i = rowval[p];
            y[i] += nz[p] * axj;
            if (i != j) t += nz[p] * x[i];
        }
        y[j] += alpha * t;
    }
}

/* ------------------------------------------------------------------------------------------------ sparse LDL^T (Davis, Alg. 849)
 * A: n x n symmetric, FULL pattern in CSC (the reference expands tril -> full, ldl.jl:21,31); only entries in the upper
 * triangle of P A P' are used.  P[k] = original index of the k-th pivot, Pinv its inverse. */
typedef struct {
    int32_t n;
    int32_t *P, *Pinv, *Parent, *Lnz, *Flag, *Pattern;
    int64_t* Lp;
    int32_t* Li;
    double *Lx, *D, *Y;
    int32_t ok;     /* last numeric factorisation reached every pivot (no exact zero) */
} okkt_ldl;

void okkt_ldl_free(okkt_ldl* F) {
    if (!F) return;
    free(F->P); free(F->Pinv); free(F->Parent); free(F->Lnz); free(F->Flag); free(F->Pattern);
    free(F->Lp); free(F->Li); free(F->Lx); free(F->D); free(F->Y);
    free(F);
}

/* ldl_symbolic: elimination tree and column counts of L */
okkt_ldl* okkt_ldl_symbolic(int32_t n, const int32_t* Ap, const int32_t* Ai, const int32_t* perm) {
    okkt_ldl* F = (okkt_ldl*)calloc(1, sizeof(okkt_ldl));
    F->n = n;
    F->P = (int32_t*)malloc(sizeof(int32_t) * n); F->Pinv = (int32_t*)malloc(sizeof(int32_t) * n);
    F->Parent = (int32_t*)malloc(sizeof(int32_t) * n); F->Lnz = (int32_t*)malloc(sizeof(int32_t) * n);
    F->Flag = (int32_t*)malloc(sizeof(int32_t) * n); F->Pattern = (int32_t*)malloc(sizeof(int32_t) * n);
    F->Lp = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
    F->D = (double*)malloc(sizeof(double) * n); F->Y = (double*)malloc(sizeof(double) * n);
    for (int32_t k = 0; k < n; ++k) { F->P[k] = perm ? perm[k] : k; }
    for (int32_t k = 0; k < n; ++k) F->Pinv[F->P[k]] = k;
    for (int32_t k = 0; k < n; ++k) {
        F->Parent[k] = -1; F->Flag[k] = k; F->Lnz[k] = 0;
        const int32_t kk = F->P[k];
        for (int32_t p = Ap[kk]; p < Ap[kk + 1]; ++p) {
            int32_t i = F->Pinv[Ai[p]];
            if (i < k) {
                for (; F->Flag[i] != k; i = F->Parent[i]) {
                    if (F->Parent[i] == -1) F->Parent[i] = k;
                    F->Lnz[i]++;
                    F->Flag[i] = k;
                }
            }
        }
    }
    F->Lp[0] = 0;
    for (int32_t k = 0; k < n; ++k) F->Lp[k + 1] = F->Lp[k] + F->Lnz[k];
    F->Li = (int32_t*)malloc(sizeof(int32_t) * (size_t)(F->Lp[n] > 0 ? F->Lp[n] : 1));
    F->Lx = (double*)malloc(sizeof(double) * (size_t)(F->Lp[n] > 0 ? F->Lp[n] : 1));
    return F;
}

int64_t okkt_ldl_nnz(const okkt_ldl* F) { return F->Lp[F->n]; }

/* ldl_numeric: up-looking; returns n on success, else the index of the first zero pivot */
int32_t okkt_ldl_numeric(okkt_ldl* F, const int32_t* Ap, const int32_t* Ai, const double* Ax) {
    const int32_t n = F->n;
    int32_t *Parent = F->Parent, *Lnz = F->Lnz, *Flag = F->Flag, *Pattern = F->Pattern, *Li = F->Li;
    int64_t* Lp = F->Lp;
    double *Lx = F->Lx, *D = F->D, *Y = F->Y;
    F->ok = 0;
    for (int32_t k = 0; k < n; ++k) {
        Y[k] = 0.0;
        int32_t top = n;
        Flag[k] = k;
        Lnz[k] = 0;
        const int32_t kk = F->P[k];
        for (int32_t p = Ap[kk]; p < Ap[kk + 1]; ++p) {
            int32_t i = F->Pinv[Ai[p]];
            if (i <= k) {
                Y[i] += Ax[p];
                int32_t len = 0;
                for (; Flag[i] != k; i = Parent[i]) { Pattern[len++] = i; Flag[i] = k; }
                while (len > 0) Pattern[--top] = Pattern[--len];
            }
        }
        D[k] = Y[k];
        Y[k] = 0.0;
        for (; top < n; ++top) {
            const int32_t i = Pattern[top];
            const double yi = Y[i];
            Y[i] = 0.0;
            const int64_t p2 = Lp[i] + Lnz[i];
            int64_t p;
            for (p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
            const double l_ki = yi / D[i];
            D[k] -= l_ki * yi;
            Li[p] = k;
            Lx[p] = l_ki;
            Lnz[i]++;
        }
        if (D[k] == 0.0) return k;
    }
    F->ok = 1;
    return n;
}

/* x <- A^{-1} x:  b = P x; L b; D b; L' b; x = P' b   (ldiv! of LDLFactorizations) */
void okkt_ldl_solve(const okkt_ldl* F, double* x) {
    const int32_t n = F->n;
    double* b = F->Y;
    for (int32_t k = 0; k < n; ++k) b[k] = x[F->P[k]];
    for (int32_t j = 0; j < n; ++j) {
        const int64_t p2 = F->Lp[j] + F->Lnz[j];
        const double bj = b[j];
        for (int64_t p = F->Lp[j]; p < p2; ++p) b[F->Li[p]] -= F->Lx[p] * bj;
    }
    for (int32_t j = 0; j < n; ++j) b[j] /= F->D[j];
    for (int32_t j = n - 1; j >= 0; --j) {
        const int64_t p2 = F->Lp[j] + F->Lnz[j];
        double bj = b[j];
        for (int64_t p = F->Lp[j]; p < p2; ++p) bj -= F->Lx[p] * b[F->Li[p]];
        b[j] = bj;
    }
    for (int32_t k = 0; k < n; ++k) x[F->P[k]] = b[k];
    for (int32_t k = 0; k < n; ++k) b[k] = 0.0;
}

void okkt_ldl_inertia(const okkt_ldl* F, int64_t* pos, int64_t* zero, int64_t* neg) {   /* ldl.jl:43-58 */
    int64_t p = 0, z = 0, m = 0;
    for (int32_t i = 0; i < F->n; ++i) {
        const double d = F->D[i];
        if (d > 0) ++p; else if (d == 0) ++z; else ++m;
    }
    *pos = p; *zero = z; *neg = m;
}

const double* okkt_ldl_D(const okkt_ldl* F) { return F->D; }
int32_t okkt_ldl_ok(const okkt_ldl* F) { return F->ok; }

/* ------------------------------------------------------------------------------------------------ IPM vector loops
 * src/IPM/kernels.jl:22-27 (_set_aug_diagonal!), :182-195 (reduce_rhs!), :198-204 (finish_aug_solve!), :161-180 (_kktmul!)
 * written as the scalar loops Julia's broadcasts compile to (same operations per element, same order). */
void okkt_set_aug_diagonal(int64_t n_tot, int64_t nlb, int64_t nub, const int64_t* ind_lb, const int64_t* ind_ub,
                           const double* reg, const double* l_lower, const double* l_diag, const double* u_lower,
                           const double* u_diag, double* pr_diag) {
    for (int64_t i = 0; i < n_tot; ++i) pr_diag[i] = reg[i];
    for (int64_t k = 0; k < nlb; ++k) pr_diag[ind_lb[k]] -= l_lower[k] / l_diag[k];
    for (int64_t k = 0; k < nub; ++k) pr_diag[ind_ub[k]] -= u_lower[k] / u_diag[k];
}

void okkt_reduce_rhs(int64_t nlb, int64_t nub, const int64_t* ind_lb, const int64_t* ind_ub, const double* l_diag,
                     const double* u_diag, double* xp, const double* dlb, const double* dub) {
    for (int64_t k = 0; k < nlb; ++k) xp[ind_lb[k]] -= dlb[k] / l_diag[k];
    for (int64_t k = 0; k < nub; ++k) xp[ind_ub[k]] -= dub[k] / u_diag[k];
}

void okkt_finish_aug_solve(int64_t nlb, int64_t nub, const int64_t* ind_lb, const int64_t* ind_ub, const double* l_lower,
                           const double* u_lower, const double* l_diag, const double* u_diag, const double* xp, double* dlb,
                           double* dub) {
    for (int64_t k = 0; k < nlb; ++k) dlb[k] = (-dlb[k] + l_lower[k] * xp[ind_lb[k]]) / l_diag[k];
    for (int64_t k = 0; k < nub; ++k) dub[k] = (dub[k] - u_lower[k] * xp[ind_ub[k]]) / u_diag[k];
}

void okkt_kktmul(int64_t n_tot, int64_t m, int64_t nlb, int64_t nub, const int64_t* ind_lb, const int64_t* ind_ub,
                 const double* reg, const double* du_diag, const double* l_lower, const double* u_lower, const double* l_diag,
                 const double* u_diag, double alpha, double beta, const double* x, double* w) {
    const double *xp = x, *xd = x + n_tot, *xlb = x + n_tot + m, *xub = x + n_tot + m + nlb;
    double *wp = w, *wd = w + n_tot, *wlb = w + n_tot + m, *wub = w + n_tot + m + nlb;
    for (int64_t i = 0; i < n_tot; ++i) wp[i] += alpha * reg[i] * xp[i];
    for (int64_t i = 0; i < m; ++i) wd[i] += alpha * du_diag[i] * xd[i];
    for (int64_t k = 0; k < nlb; ++k) wp[ind_lb[k]] -= alpha * xlb[k];
    for (int64_t k = 0; k < nub; ++k) wp[ind_ub[k]] += alpha * xub[k];
    for (int64_t k = 0; k < nlb; ++k) wlb[k] = beta * wlb[k] + alpha * (xp[ind_lb[k]] * l_lower[k] - xlb[k] * l_diag[k]);
    for (int64_t k = 0; k < nub; ++k) wub[k] = beta * wub[k] + alpha * (xp[ind_ub[k]] * u_lower[k] + xub[k] * u_diag[k]);
}
