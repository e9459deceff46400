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
