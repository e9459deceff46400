// IPM reductions and the right-hand-side kernel around the KKT solve: the scalar quantities the
// filter line-search reads every iteration -- step lengths, barrier objective and its directional derivative, the three
// optimality errors, complementarity measures, scaling factors -- each as ONE single-pass kernel over device vectors.
// The reference computes them with scalar loops on the CPU (src/IPM/kernels.jl:263-388,675-695) and, on the GPU, with
// allocating mapreduce calls (lib/MadNLPGPU/src/IPM/kernels.jl).
//
// Determinism: a fixed grid of B2_RED_BLOCKS CTAs; every CTA reduces its grid-stride slice in a fixed order and writes one
// partial; the CTA that arrives last (atomic ticket) combines the partials in index order and applies the final scaling.
// min / max propagate NaN like Julia's `min` / `max`.
#include <algorithm>
#include <cmath>

#include "bounds.cuh"
#include "common.cuh"

using namespace b2;

namespace {

enum { R_SUM = 0, R_MIN = 1, R_MAX = 2 };

template <int KIND>
__device__ __forceinline__ double comb(double a, double b) {
    if (KIND == R_SUM) return a + b;
    if (a != a || b != b) return a + b;                 // NaN in, NaN out
    return (KIND == R_MIN) ? (a < b ? a : b) : (a > b ? a : b);
}

// F: struct with `__device__ double term(int64_t i) const` over i in [0, n), `double init` semantics via identity(),
// and `__device__ double finish(double r) const` applied once to the reduced value.
template <int KIND, class F>
__global__ void __launch_bounds__(256) k_reduce(int64_t n, F f, double identity, double* __restrict__ part, unsigned* ticket,
                                                double* __restrict__ out) {
    __shared__ double sm[8];
    __shared__ bool last;
    pdl_sync();
    double acc = identity;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc = comb<KIND>(acc, f.term(i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = comb<KIND>(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = sm[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) r = comb<KIND>(r, sm[w]);
        part[blockIdx.x] = r;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA: partials in index order (thread t owns partials t, t+256, ...; then the same tree as above)
    double r = identity;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) r = comb<KIND>(r, __ldcg(part + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = comb<KIND>(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = sm[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) t = comb<KIND>(t, sm[w]);
        out[0] = f.finish(t);
        *ticket = 0;                                    // ready for the next reduction on this stream
    }
}

template <int KIND, class F>
int run_reduce(b2_bounds* b, int64_t n, const F& f, double identity, double* out_d, void* stream, const char* who) {
    if (!b || !out_d || n < 0) { set_error(std::string(who) + ": invalid argument"); return B2_ERR_INVALID; }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, B2_RED_BLOCKS));
    cudaError_t e = launch_pdl(k_reduce<KIND, F>, dim3(grid), dim3(256), 0, as_stream(stream), n, f, identity, b->red_part.p, b->red_ticket.p, out_d);
    if (e != cudaSuccess) return cuda_fail(e, who, __FILE__, __LINE__);
    return B2_OK;
}

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

// ---- get_alpha_max (src/IPM/kernels.jl:356-371)
struct AlphaMax {
    const double *x, *xl, *xu, *dx; double tau;
    __device__ double term(int64_t i) const {
        const double d = dx[i];
        // one of the two reference terms is +Inf for every sign of d, so only the taken one is evaluated: ONE fp64
        // division per element, and only the bound on the side d points to is read
        if (d == 0.0 || d != d) return dinf();
        const double bnd = d < 0.0 ? xl[i] : xu[i];
        return (-x[i] + bnd) * tau / d;
    }
    __device__ double finish(double r) const { return r; }
};
// ---- get_alpha_z (:373-388): i < nlb over (zl_r, dzl), else over (zu_r, dzu)
struct AlphaZ {
    const int64_t *ind_lb, *ind_ub; int64_t nlb; const double *zl, *zu, *dzl, *dzu; double tau;
    __device__ double term(int64_t i) const {
        if (i < nlb) { const double d = dzl[i]; return d < 0.0 ? (-zl[ind_lb[i]]) * tau / d : dinf(); }
        const int64_t j = i - nlb; const double d = dzu[j];
        return d < 0.0 ? (-zu[ind_ub[j]]) * tau / d : dinf();
    }
    __device__ double finish(double r) const { return r; }
};
// ---- get_varphi (:263-283): obj_val - mu * sum log(slack), +Inf for a negative slack
struct Varphi {
    const int64_t *ind_lb, *ind_ub; int64_t nlb; const double *x, *xl, *xu; double mu, obj_val;
    __device__ double term(int64_t i) const {
        double s;
        if (i < nlb) { const int64_t k = ind_lb[i]; s = x[k] - xl[k]; } else { const int64_t k = ind_ub[i - nlb]; s = xu[k] - x[k]; }
        return s < 0.0 ? dinf() : -mu * log(s);
    }
    __device__ double finish(double r) const { return obj_val + r; }
};
// ---- get_varphi_d (:341-354)
struct VarphiD {
    const double *f, *x, *xl, *xu, *dx; double mu;
    __device__ double term(int64_t i) const { return (f[i] - mu / (x[i] - xl[i]) + mu / (xu[i] - x[i])) * dx[i]; }
    __device__ double finish(double r) const { return r; }
};
// ---- get_inf_du (:285-291)
struct InfDu {
    const double *f, *zl, *zu, *jacl; double sd;
    __device__ double term(int64_t i) const { return fabs(f[i] - zl[i] + zu[i] + jacl[i]); }
    __device__ double finish(double r) const { return r / sd; }
};
// ---- get_inf_compl (:293-303)
struct InfCompl {
    const int64_t *ind_lb, *ind_ub; int64_t nlb; const double *x, *xl, *xu, *zl, *zu; double mu, sc;
    __device__ double term(int64_t i) const {
        if (i < nlb) { const int64_t k = ind_lb[i]; return fabs((x[k] - xl[k]) * zl[k] - mu); }
        const int64_t k = ind_ub[i - nlb];
        return fabs((xu[k] - x[k]) * zu[k] - mu);
    }
    __device__ double finish(double r) const { return r / sc; }
};
// ---- get_average_complementarity (:305-314) / get_min_complementarity (:322-333)
struct ComplTerm {
    const int64_t *ind_lb, *ind_ub; int64_t nlb, ntot; const double *x, *xl, *xu, *zl, *zu; int average;
    __device__ double term(int64_t i) const {
        if (i < nlb) { const int64_t k = ind_lb[i]; return (x[k] - xl[k]) * zl[k]; }
        const int64_t k = ind_ub[i - nlb];
        return (xu[k] - x[k]) * zu[k];
    }
    __device__ double finish(double r) const { return average ? (ntot == 0 ? 0.0 : r / (double)ntot) : r; }
};
// ---- get_rel_search_norm (:675-681)
struct RelSearch {
    const double *x, *dx;
    __device__ double term(int64_t i) const { return fabs(dx[i]) / (1.0 + fabs(x[i])); }
    __device__ double finish(double r) const { return r; }
};
// ---- get_sd / get_sc (:684-695): max(s_max, (||l||_1 + ||zl_r||_1 + ||zu_r||_1) / max(1, count)) / s_max
struct ScaleSum {
    const int64_t *ind_lb, *ind_ub; int64_t m, nlb, count; const double *l, *zl, *zu; double s_max;
    __device__ double term(int64_t i) const {
        if (i < m) return fabs(l[i]);
        if (i < m + nlb) return fabs(zl[ind_lb[i - m]]);
        return fabs(zu[ind_ub[i - m - nlb]]);
    }
    __device__ double finish(double r) const {
        const double avg = r / (double)(count > 1 ? count : 1);
        return (s_max > avg ? s_max : avg) / s_max;
    }
};

// ---- set_aug_rhs! (:113-130): px = -f + zl - zu - jacl ; py = -c ; pzl = (xl_r - x_lr) zl_r + mu ; pzu = (xu_r - x_ur) zu_r - mu
__global__ void k_set_aug_rhs(int64_t n_tot, int64_t m, int64_t nlb, int64_t nub, const int64_t* __restrict__ ind_lb,
                              const int64_t* __restrict__ ind_ub, const double* __restrict__ x, const double* __restrict__ xl,
                              const double* __restrict__ xu, const double* __restrict__ f, const double* __restrict__ zl,
                              const double* __restrict__ zu, const double* __restrict__ jacl, const double* __restrict__ c, double mu,
                              double* __restrict__ p) {
    pdl_sync();
    const int64_t tot = n_tot + m + nlb + nub;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
        double v;
        if (t < n_tot) v = -f[t] + zl[t] - zu[t] - jacl[t];
        else if (t < n_tot + m) v = -c[t - n_tot];
        else if (t < n_tot + m + nlb) { const int64_t k = ind_lb[t - n_tot - m]; v = __dadd_rn(__dmul_rn(xl[k] - x[k], zl[k]), mu); }   // (no FMA contraction:
        else { const int64_t k = ind_ub[t - n_tot - m - nlb]; v = __dadd_rn(__dmul_rn(xu[k] - x[k], zu[k]), -mu); }                   //  bit-identical to the broadcast)
        p[t] = v;
    }
}

}  // namespace

#define B2_NEED(cond, who) do { if (!(cond)) { set_error(who ": invalid argument"); return B2_ERR_INVALID; } } while (0)

extern "C" {

int b2_get_alpha_max(b2_bounds* b, const double* x_d, const double* xl_d, const double* xu_d, const double* dx_d, double tau,
                     double* out_d, void* stream) {
    B2_NEED(b && (b->n_tot == 0 || (x_d && xl_d && xu_d && dx_d)), "b2_get_alpha_max");
    AlphaMax f{x_d, xl_d, xu_d, dx_d, tau};
    return run_reduce<R_MIN>(b, b->n_tot, f, 1.0, out_d, stream, "b2_get_alpha_max");
}

int b2_get_alpha_z(b2_bounds* b, const double* zl_d, const double* zu_d, const double* dzl_d, const double* dzu_d, double tau,
                   double* out_d, void* stream) {
    B2_NEED(b && (b->nlb == 0 || (zl_d && dzl_d)) && (b->nub == 0 || (zu_d && dzu_d)), "b2_get_alpha_z");
    AlphaZ f{b->ind_lb.p, b->ind_ub.p, b->nlb, zl_d, zu_d, dzl_d, dzu_d, tau};
    return run_reduce<R_MIN>(b, b->nlb + b->nub, f, 1.0, out_d, stream, "b2_get_alpha_z");
}

int b2_get_varphi(b2_bounds* b, double obj_val, const double* x_d, const double* xl_d, const double* xu_d, double mu, double* out_d,
                  void* stream) {
    B2_NEED(b && (b->nlb + b->nub == 0 || (x_d && xl_d && xu_d)), "b2_get_varphi");
    Varphi f{b->ind_lb.p, b->ind_ub.p, b->nlb, x_d, xl_d, xu_d, mu, obj_val};
    return run_reduce<R_SUM>(b, b->nlb + b->nub, f, 0.0, out_d, stream, "b2_get_varphi");
}

int b2_get_varphi_d(b2_bounds* b, const double* f_d, const double* x_d, const double* xl_d, const double* xu_d, const double* dx_d,
                    double mu, double* out_d, void* stream) {
    B2_NEED(b && (b->n_tot == 0 || (f_d && x_d && xl_d && xu_d && dx_d)), "b2_get_varphi_d");
    VarphiD f{f_d, x_d, xl_d, xu_d, dx_d, mu};
    return run_reduce<R_SUM>(b, b->n_tot, f, 0.0, out_d, stream, "b2_get_varphi_d");
}

int b2_get_inf_du(b2_bounds* b, const double* f_d, const double* zl_d, const double* zu_d, const double* jacl_d, double sd,
                  double* out_d, void* stream) {
    B2_NEED(b && (b->n_tot == 0 || (f_d && zl_d && zu_d && jacl_d)), "b2_get_inf_du");
    InfDu f{f_d, zl_d, zu_d, jacl_d, sd};
    return run_reduce<R_MAX>(b, b->n_tot, f, 0.0, out_d, stream, "b2_get_inf_du");
}

int b2_get_inf_compl(b2_bounds* b, const double* x_d, const double* xl_d, const double* xu_d, const double* zl_d, const double* zu_d,
                     double mu, double sc, double* out_d, void* stream) {
    B2_NEED(b && (b->nlb + b->nub == 0 || (x_d && xl_d && xu_d && zl_d && zu_d)), "b2_get_inf_compl");
    InfCompl f{b->ind_lb.p, b->ind_ub.p, b->nlb, x_d, xl_d, xu_d, zl_d, zu_d, mu, sc};
    return run_reduce<R_MAX>(b, b->nlb + b->nub, f, 0.0, out_d, stream, "b2_get_inf_compl");
}

int b2_get_average_complementarity(b2_bounds* b, const double* x_d, const double* xl_d, const double* xu_d, const double* zl_d,
                                   const double* zu_d, double* out_d, void* stream) {
    B2_NEED(b && (b->nlb + b->nub == 0 || (x_d && xl_d && xu_d && zl_d && zu_d)), "b2_get_average_complementarity");
    ComplTerm f{b->ind_lb.p, b->ind_ub.p, b->nlb, b->nlb + b->nub, x_d, xl_d, xu_d, zl_d, zu_d, 1};
    return run_reduce<R_SUM>(b, b->nlb + b->nub, f, 0.0, out_d, stream, "b2_get_average_complementarity");
}

int b2_get_min_complementarity(b2_bounds* b, const double* x_d, const double* xl_d, const double* xu_d, const double* zl_d,
                               const double* zu_d, double* out_d, void* stream) {
    B2_NEED(b && (b->nlb + b->nub == 0 || (x_d && xl_d && xu_d && zl_d && zu_d)), "b2_get_min_complementarity");
    ComplTerm f{b->ind_lb.p, b->ind_ub.p, b->nlb, b->nlb + b->nub, x_d, xl_d, xu_d, zl_d, zu_d, 0};
    return run_reduce<R_MIN>(b, b->nlb + b->nub, f, HUGE_VAL, out_d, stream, "b2_get_min_complementarity");
}

int b2_get_rel_search_norm(b2_bounds* b, int64_t n, const double* x_d, const double* dx_d, double* out_d, void* stream) {
    B2_NEED(b && n >= 0 && (n == 0 || (x_d && dx_d)), "b2_get_rel_search_norm");
    RelSearch f{x_d, dx_d};
    return run_reduce<R_MAX>(b, n, f, 0.0, out_d, stream, "b2_get_rel_search_norm");
}

int b2_get_sd(b2_bounds* b, int64_t m, const double* l_d, const double* zl_d, const double* zu_d, double s_max, double* out_d, void* stream) {
    B2_NEED(b && m >= 0 && (m == 0 || l_d) && (b->nlb == 0 || zl_d) && (b->nub == 0 || zu_d) && s_max > 0.0, "b2_get_sd");
    ScaleSum f{b->ind_lb.p, b->ind_ub.p, m, b->nlb, m + b->nlb + b->nub, l_d, zl_d, zu_d, s_max};
    return run_reduce<R_SUM>(b, m + b->nlb + b->nub, f, 0.0, out_d, stream, "b2_get_sd");
}

int b2_get_sc(b2_bounds* b, const double* zl_d, const double* zu_d, double s_max, double* out_d, void* stream) {
    B2_NEED(b && (b->nlb == 0 || zl_d) && (b->nub == 0 || zu_d) && s_max > 0.0, "b2_get_sc");
    ScaleSum f{b->ind_lb.p, b->ind_ub.p, 0, b->nlb, b->nlb + b->nub, nullptr, zl_d, zu_d, s_max};
    return run_reduce<R_SUM>(b, b->nlb + b->nub, f, 0.0, out_d, stream, "b2_get_sc");
}

int b2_set_aug_rhs(b2_bounds* b, int64_t m, const double* x_d, const double* xl_d, const double* xu_d, const double* f_d,
                   const double* zl_d, const double* zu_d, const double* jacl_d, const double* c_d, double mu, double* p_d, void* stream) {
    B2_NEED(b && m >= 0, "b2_set_aug_rhs");
    const int64_t tot = b->n_tot + m + b->nlb + b->nub;
    if (tot == 0) return B2_OK;
    B2_NEED(p_d && (b->n_tot == 0 || (x_d && xl_d && xu_d && f_d && zl_d && zu_d && jacl_d)) && (m == 0 || c_d), "b2_set_aug_rhs");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((tot + 255) / 256, 8 * sm_count()));
    cudaError_t e = launch_pdl(k_set_aug_rhs, dim3(grid), dim3(256), 0, as_stream(stream), b->n_tot, m, b->nlb, b->nub, b->ind_lb.p, b->ind_ub.p,
                               x_d, xl_d, xu_d, f_d, zl_d, zu_d, jacl_d, c_d, mu, p_d);
    if (e != cudaSuccess) return cuda_fail(e, "b2_set_aug_rhs", __FILE__, __LINE__);
    return B2_OK;
}

}  // extern "C"
