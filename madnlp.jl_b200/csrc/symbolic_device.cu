// Device-side symbolic phase: COO -> CSC pattern + map built with device sorts, the way the reference's GPU
// path does it (lib/MadNLPGPU/src/KKT/gpu_sparse.jl:260-302: sortperm of the (col,row) keys, unique, scatter of the map) instead of
// the host std::stable_sort of b2_coo_to_csc.  Output is IDENTICAL to b2_coo_to_csc (slots numbered by ascending (col,row) key,
// duplicates share a slot): tests/test_gpu_symbolic.py compares the two bit for bit.
//   keys = J*m + I  ->  cub radix sort (key, position)  ->  head flags + inclusive scan = slot id  ->  rowval / colptr / map
#include <cub/cub.cuh>

#include "common.cuh"

using namespace b2;

namespace {
__global__ void k_make_keys(int64_t nnz, int64_t m, const int32_t* __restrict__ I, const int32_t* __restrict__ J, int64_t* __restrict__ key,
                            int32_t* __restrict__ pos, int* __restrict__ bad, int32_t nrow, int32_t ncol) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = I[k], j = J[k];
        if (i < 0 || i >= nrow || j < 0 || j >= ncol) atomicExch(bad, 1);
        key[k] = (int64_t)j * m + i;
        pos[k] = (int32_t)k;
    }
}
__global__ void k_heads(int64_t nnz, const int64_t* __restrict__ key, int32_t* __restrict__ head) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        head[k] = (k == 0 || key[k] != key[k - 1]) ? 1 : 0;
}
// slot[k] = (inclusive scan of head)[k] - 1;  heads write rowval and bump their column's count; every entry writes its map
__global__ void k_emit(int64_t nnz, int64_t m, const int64_t* __restrict__ key, const int32_t* __restrict__ pos, const int32_t* __restrict__ head,
                       const int32_t* __restrict__ scan, int32_t* __restrict__ rowval, int32_t* __restrict__ colcnt, int64_t* __restrict__ map) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t slot = scan[k] - 1;
        if (head[k]) {
            rowval[slot] = (int32_t)(key[k] % m);
            atomicAdd(colcnt + (int32_t)(key[k] / m) + 1, 1);            // integer counts: order-independent
        }
        map[pos[k]] = slot;
    }
}
}  // namespace

// I_d, J_d: [nnz_coo] device, 0-based.  colptr_d [n+1], rowval_d [capacity nnz_coo], map_d [nnz_coo] device outputs.
// *nnz_csc (host) receives the number of distinct (row, col) positions; synchronises `stream` once for it.
extern "C" int b2_coo_to_csc_device(int32_t m, int32_t n, int64_t nnz_coo, const int32_t* I_d, const int32_t* J_d, int32_t* colptr_d,
                                    int32_t* rowval_d, int64_t* map_d, int64_t* nnz_csc, void* stream) {
    if (m < 0 || n < 0 || nnz_coo < 0 || nnz_coo > INT32_MAX || (nnz_coo && (!I_d || !J_d || !rowval_d || !map_d)) || !colptr_d) {
        set_error("b2_coo_to_csc_device: invalid argument");
        return B2_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("b2_coo_to_csc_device: no CUDA device"); return B2_ERR_NO_DEVICE; }
    cudaStream_t st = as_stream(stream);
    B2_CUDA(cudaMemsetAsync(colptr_d, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    if (nnz_coo == 0) { if (nnz_csc) *nnz_csc = 0; return B2_OK; }
    DevBuf<int64_t> key, key2;
    DevBuf<int32_t> pos, pos2, head, scan, bad;
    if (key.alloc(nnz_coo) != cudaSuccess || key2.alloc(nnz_coo) != cudaSuccess || pos.alloc(nnz_coo) != cudaSuccess ||
        pos2.alloc(nnz_coo) != cudaSuccess || head.alloc(nnz_coo) != cudaSuccess || scan.alloc(nnz_coo) != cudaSuccess || bad.alloc(1) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "b2_coo_to_csc_device alloc", __FILE__, __LINE__);
    B2_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int32_t), st));
    const int grid = (int)std::min<int64_t>((nnz_coo + 255) / 256, 8 * sm_count());
    k_make_keys<<<grid, 256, 0, st>>>(nnz_coo, (int64_t)m, I_d, J_d, key.p, pos.p, bad.p, m, n);
    // stable LSD radix sort over the significant key bits only
    int end_bit = 1;
    while (end_bit < 63 && ((int64_t)1 << end_bit) <= (int64_t)n * (int64_t)std::max(m, 1)) ++end_bit;
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, pos.p, pos2.p, (int)nnz_coo, 0, end_bit, st);
    cub::DeviceScan::InclusiveSum(nullptr, tmp2, head.p, scan.p, (int)nnz_coo, st);
    DevBuf<unsigned char> tmp;
    if (tmp.alloc(std::max(tmp_bytes, tmp2)) != cudaSuccess) return cuda_fail(cudaGetLastError(), "b2_coo_to_csc_device temp", __FILE__, __LINE__);
    tmp_bytes = tmp2 = tmp.bytes();
    B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, pos.p, pos2.p, (int)nnz_coo, 0, end_bit, st));
    k_heads<<<grid, 256, 0, st>>>(nnz_coo, key2.p, head.p);
    B2_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp2, head.p, scan.p, (int)nnz_coo, st));
    k_emit<<<grid, 256, 0, st>>>(nnz_coo, (int64_t)m, key2.p, pos2.p, head.p, scan.p, rowval_d, colptr_d, map_d);
    // colptr: counts were accumulated at colptr[j+1]; inclusive scan in place gives the pointers
    size_t tmp3 = tmp.bytes();
    {
        size_t need = 0;
        cub::DeviceScan::InclusiveSum(nullptr, need, colptr_d, colptr_d, n + 1, st);
        if (need > tmp3) { set_error("b2_coo_to_csc_device: temp storage"); return B2_ERR_CUDA; }
    }
    B2_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp3, colptr_d, colptr_d, n + 1, st));
    int32_t h_last = 0, h_bad = 0;
    B2_CUDA(cudaMemcpyAsync(&h_last, scan.p + (nnz_coo - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (h_bad) { set_error("b2_coo_to_csc_device: index out of range"); return B2_ERR_INVALID; }
    if (nnz_csc) *nnz_csc = h_last;
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// build_condensed_aug_symbolic on the device (src/KKT/Sparse/condensed.jl:201-301; the reference's GPU path sorts the same key list
// on the device, lib/MadNLPGPU/src/KKT/gpu_sparse.jl:100-130 + condensed.jl:251): enumerate diag | hess | Jt column pairs in the
// reference's order, STABLE radix sort by (col, row), unique -> slots; the per-slot source lists keep the enumeration order, so the
// plan -- and therefore every bit of the assembled matrix -- equals the host construction's (tests/test_gpu_symbolic.py).
// ---------------------------------------------------------------------------------------------------------------------
#include "condensed_plan.cuh"

namespace {
__global__ void k_pair_counts(int m, const int32_t* __restrict__ Jc, int32_t* __restrict__ cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int c = Jc[i + 1] - Jc[i];
        cnt[i] = c * (c + 1) / 2;
    }
}
__global__ void k_fill_diag_hess(int n, const int32_t* __restrict__ Hc, const int32_t* __restrict__ Hr, int64_t* __restrict__ key,
                                 int32_t* __restrict__ idx, int* __restrict__ bad) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        key[j] = (int64_t)j * n + j; idx[j] = j;
        for (int p = Hc[j]; p < Hc[j + 1]; ++p) {
            const int r = Hr[p];
            if (r < j || r >= n) atomicExch(bad, 1);
            key[n + p] = (int64_t)j * n + r; idx[n + p] = n + p;
        }
    }
}
__global__ void k_fill_pairs(int n, int m, int64_t base, const int32_t* __restrict__ Jc, const int32_t* __restrict__ Jr,
                             const int32_t* __restrict__ poff, int64_t* __restrict__ key, int32_t* __restrict__ idx, int4* __restrict__ rec,
                             int* __restrict__ bad) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        int q = poff[i];
        for (int j = Jc[i]; j < Jc[i + 1]; ++j)
            for (int k = j; k < Jc[i + 1]; ++k, ++q) {
                const int c1 = Jr[j], c2 = Jr[k];                       // (col, row) of the entry; rows of a Jt column are ascending
                if (c2 < c1 || c2 >= n || c1 < 0) atomicExch(bad, 1);
                key[base + q] = (int64_t)c1 * n + c2; idx[base + q] = (int32_t)(base + q);
                rec[q] = make_int4(i, j, k, 0);
            }
    }
}
// per sorted position: slot bookkeeping; marks triple entries
__global__ void k_cond_emit(int64_t tot, int n, int64_t n_hess_end, const int64_t* __restrict__ key, const int32_t* __restrict__ idx,
                            const int32_t* __restrict__ head, const int32_t* __restrict__ scan, int32_t* __restrict__ rowval,
                            int32_t* __restrict__ colcnt, int32_t* __restrict__ hsrc, int32_t* __restrict__ dsrc, int32_t* __restrict__ is_trip,
                            int* __restrict__ bad) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < tot; s += (int64_t)gridDim.x * blockDim.x) {
        const int32_t slot = scan[s] - 1, t = idx[s];
        if (head[s]) { rowval[slot] = (int32_t)(key[s] % n); atomicAdd(colcnt + (int32_t)(key[s] / n) + 1, 1); }
        int trip = 0;
        if (t < n) dsrc[slot] = t;
        else if (t < n_hess_end) { if (atomicExch(hsrc + slot, t - n) != -1) atomicExch(bad, 2); }      // duplicate entry in H
        else trip = 1;
        is_trip[s] = trip;
    }
}
__global__ void k_cond_trip(int64_t tot, int64_t base, const int32_t* __restrict__ idx, const int32_t* __restrict__ head, const int32_t* __restrict__ scan,
                            const int32_t* __restrict__ is_trip, const int32_t* __restrict__ texcl, const int4* __restrict__ rec,
                            int32_t* __restrict__ tptr, int4* __restrict__ trip) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < tot; s += (int64_t)gridDim.x * blockDim.x) {
        if (head[s]) tptr[scan[s] - 1] = texcl[s];
        if (is_trip[s]) trip[texcl[s]] = rec[idx[s] - base];
    }
}
}  // namespace

extern "C" int b2_condensed_symbolic_device(int32_t n, int32_t m, const int32_t* H_colptr_d, const int32_t* H_rowval_d,
                                            const int32_t* Jt_colptr_d, const int32_t* Jt_rowval_d, b2_condensed_plan** out,
                                            int64_t* nnz_aug, void* stream) {
    if (!out || n <= 0 || m < 0 || !H_colptr_d || !Jt_colptr_d) { set_error("b2_condensed_symbolic_device: invalid argument"); return B2_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("b2_condensed_symbolic_device: no CUDA device"); return B2_ERR_NO_DEVICE; }
    cudaStream_t st = as_stream(stream);
    const int grid_m = std::max(1, std::min((m + 255) / 256, 8 * sm_count())), grid_n = std::max(1, std::min((n + 255) / 256, 8 * sm_count()));
    DevBuf<int32_t> cnt, poff, bad;
    if (cnt.alloc(m + 1) != cudaSuccess || poff.alloc(m + 1) != cudaSuccess || bad.alloc(1) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "condensed symbolic alloc", __FILE__, __LINE__);
    B2_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int32_t), st));
    B2_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)(m + 1) * sizeof(int32_t), st));
    if (m > 0) k_pair_counts<<<grid_m, 256, 0, st>>>(m, Jt_colptr_d, cnt.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, poff.p, m + 1, st);
    DevBuf<unsigned char> tmp0;
    if (tmp0.alloc(std::max<size_t>(tb, 16)) != cudaSuccess) return cuda_fail(cudaGetLastError(), "condensed symbolic temp", __FILE__, __LINE__);
    B2_CUDA(cub::DeviceScan::ExclusiveSum(tmp0.p, tb, cnt.p, poff.p, m + 1, st));
    int32_t nnzH = 0, nj = 0;
    B2_CUDA(cudaMemcpyAsync(&nnzH, H_colptr_d + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&nj, poff.p + m, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    const int64_t base = (int64_t)n + nnzH, tot = base + nj;
    if (tot > INT32_MAX) { set_error("b2_condensed_symbolic_device: too many entries"); return B2_ERR_INVALID; }
    DevBuf<int64_t> key, key2;
    DevBuf<int32_t> idx, idx2, head, scan, is_trip, texcl, colcnt, rowval;
    DevBuf<int4> rec;
    auto* p = new b2_condensed_plan();
    p->n = n; p->m = m;
    const int grid_t = (int)std::max<int64_t>(1, std::min<int64_t>((tot + 255) / 256, 8 * sm_count()));
    if (key.alloc(tot) != cudaSuccess || key2.alloc(tot) != cudaSuccess || idx.alloc(tot) != cudaSuccess || idx2.alloc(tot) != cudaSuccess ||
        head.alloc(tot) != cudaSuccess || scan.alloc(tot) != cudaSuccess || is_trip.alloc(tot) != cudaSuccess || texcl.alloc(tot) != cudaSuccess ||
        colcnt.alloc(n + 1) != cudaSuccess || rowval.alloc(tot) != cudaSuccess || rec.alloc(std::max(nj, 1)) != cudaSuccess) {
        delete p;
        return cuda_fail(cudaGetLastError(), "condensed symbolic alloc", __FILE__, __LINE__);
    }
    k_fill_diag_hess<<<grid_n, 256, 0, st>>>(n, H_colptr_d, H_rowval_d, key.p, idx.p, bad.p);
    if (m > 0 && nj > 0) k_fill_pairs<<<grid_m, 256, 0, st>>>(n, m, base, Jt_colptr_d, Jt_rowval_d, poff.p, key.p, idx.p, rec.p, bad.p);
    int end_bit = 1;
    while (end_bit < 63 && ((int64_t)1 << end_bit) <= (int64_t)n * n) ++end_bit;
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, key.p, key2.p, idx.p, idx2.p, (int)tot, 0, end_bit, st);
    cub::DeviceScan::InclusiveSum(nullptr, t2, head.p, scan.p, (int)tot, st);
    cub::DeviceScan::InclusiveSum(nullptr, t3, colcnt.p, colcnt.p, n + 1, st);
    DevBuf<unsigned char> tmp;
    if (tmp.alloc(std::max(std::max(t1, t2), std::max(t3, (size_t)16))) != cudaSuccess) { delete p; return cuda_fail(cudaGetLastError(), "condensed symbolic temp", __FILE__, __LINE__); }
    size_t tsz = tmp.bytes();
    B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tsz, key.p, key2.p, idx.p, idx2.p, (int)tot, 0, end_bit, st));
    k_heads<<<grid_t, 256, 0, st>>>(tot, key2.p, head.p);
    tsz = tmp.bytes();
    B2_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tsz, head.p, scan.p, (int)tot, st));
    int32_t nslot = 0;
    B2_CUDA(cudaMemcpyAsync(&nslot, scan.p + (tot - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (p->hsrc.alloc(nslot) != cudaSuccess || p->dsrc.alloc(nslot) != cudaSuccess || p->tptr.alloc((size_t)nslot + 1) != cudaSuccess ||
        p->trip.alloc(std::max(nj, 1)) != cudaSuccess) { delete p; return cuda_fail(cudaGetLastError(), "condensed plan alloc", __FILE__, __LINE__); }
    B2_CUDA(cudaMemsetAsync(p->hsrc.p, 0xFF, (size_t)nslot * sizeof(int32_t), st));
    B2_CUDA(cudaMemsetAsync(p->dsrc.p, 0xFF, (size_t)nslot * sizeof(int32_t), st));
    B2_CUDA(cudaMemsetAsync(colcnt.p, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    B2_CUDA(cudaMemsetAsync(p->trip.p, 0, sizeof(int4), st));
    k_cond_emit<<<grid_t, 256, 0, st>>>(tot, n, base, key2.p, idx2.p, head.p, scan.p, rowval.p, colcnt.p, p->hsrc.p, p->dsrc.p, is_trip.p, bad.p);
    tsz = tmp.bytes();
    B2_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tsz, is_trip.p, texcl.p, (int)tot, st));
    k_cond_trip<<<grid_t, 256, 0, st>>>(tot, base, idx2.p, head.p, scan.p, is_trip.p, texcl.p, rec.p, p->tptr.p, p->trip.p);
    B2_CUDA(cudaMemcpyAsync(p->tptr.p + nslot, &nj, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    tsz = tmp.bytes();
    B2_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tsz, colcnt.p, colcnt.p, n + 1, st));
    p->colptr.resize(n + 1); p->rowval.resize(nslot);
    int32_t h_bad = 0;
    B2_CUDA(cudaMemcpyAsync(p->colptr.data(), colcnt.p, (size_t)(n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(p->rowval.data(), rowval.p, (size_t)nslot * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        delete p;
        set_error(h_bad == 2 ? "b2_condensed_symbolic_device: duplicate entry in H" : "b2_condensed_symbolic_device: entry above the diagonal (H must be lower, Jt rows sorted)");
        return B2_ERR_INVALID;
    }
    p->nnz_aug = nslot; p->n_dptr = n; p->n_hptr = nnzH; p->n_jptr = nj;
    *out = p;
    if (nnz_aug) *nnz_aug = nslot;
    return B2_OK;
}
