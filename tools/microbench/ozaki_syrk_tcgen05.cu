// Micro-benchmark SURVEY.md section 7 asked for: the fp64 SYRK of BASELINE.json configs[1] (W = A' A, A = sqrt(D) .* J_ineq,
// K = ns = 2048 rows, n = 4096 columns) on the 5th-generation tensor cores.  tcgen05 has no fp64 kind, so fp64 is SLICED onto
// tcgen05.mma.kind::i8 (Ozaki scheme, int8 variant -- Ootomo/Ozaki/Yokota, "DGEMM on integer matrix multiplication unit", 2024):
//
//   1. per column m:  e_m = exponent of max_i |a_im|;  x = a_im * 2^-e_m in (-1, 1) is cut into S = 8 signed 7-bit digits
//      x = sum_s q_s 2^(-7(s+1))   (q_s int8, exact: 56 bits cover the fp64 mantissa of the column's largest entries)
//   2. G_d = sum_{s+t=d} Q_s' Q_t  for d = 0..S-1 : 36 EXACT int8 x int8 -> int32 GEMMs (|G_d| <= 8 * 2048 * 127^2 < 2^31),
//      the d-sums accumulate inside the tensor-core accumulator (8 accumulators of 64 columns = the SM's whole TMEM)
//   3. W(m,n) = 2^(e_m + e_n - 14) * sum_d 2^(-7d) G_d(m,n)   evaluated in fp64 (Horner) by the epilogue
//
// Kernel (one 128 x 64 output tile per CTA, lower triangle only, 192 threads):
//   warp 4 lane 0 : TMA producer  -- cp.async.bulk.tensor.3d (64-byte swizzle) of the 8 A-digit tiles and 8 B-digit tiles of a
//                                    64-deep K block into a 2-stage shared-memory ring, mbarrier complete_tx
//   warp 5 lane 0 : MMA issuer    -- 72 tcgen05.mma.cta_group::1.kind::i8 (M128 N64 K32) per K block, tcgen05.commit frees the stage
//   warps 0..3    : epilogue      -- tcgen05.ld of the 8 accumulators, Horner in fp64, scaling, coalesced stores
// Reported: max error relative to max|W| against cuBLAS DGEMM (parity gate 1e-13), time of the digit split and of the GEMM,
// int8 TOP/s, fp64-equivalent TFLOP/s, next to cuBLAS DGEMM/DSYRK on the same matrix.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ozaki_syrk_tcgen05 ozaki_syrk_tcgen05.cu -lcublas
#include <cublas_v2.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int S = 8;             // digits
constexpr int WB = 7;            // bits per digit
constexpr int BM = 128, BN = 64; // output tile
constexpr int BKB = 64;          // K bytes (= int8 elements) per pipeline stage: one 64-byte swizzle row
constexpr int STAGES = 2;
constexpr int A_TILE = BM * BKB, B_TILE = BN * BKB;                 // bytes of one digit tile
constexpr int STAGE_BYTES = S * (A_TILE + B_TILE);                  // 96 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;             // + alignment slack
constexpr int NTHREADS = 192;
constexpr uint32_t SPIN_MAX = 1u << 22;

// ------------------------------------------------------------------------------------------------ digit split
// one CTA per column m of A (K contiguous doubles): exponent, then S int8 digits per element; Q[s][m][i]
__global__ void __launch_bounds__(256) k_split(int K, int M, const double* __restrict__ A, int8_t* __restrict__ Q, int* __restrict__ expo) {
    const int m = blockIdx.x;
    const double* col = A + (size_t)m * K;
    __shared__ double red[256];
    double mx = 0.0;
    for (int i = threadIdx.x; i < K; i += 256) mx = fmax(mx, fabs(col[i]));
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
    mx = red[0];
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                    // mx = f * 2^e, f in [0.5, 1)  ->  |a| * 2^-e < 1
    if (threadIdx.x == 0) expo[m] = e;
    for (int i = threadIdx.x; i < K; i += 256) {
        double x = ldexp(col[i], -e);               // exact
#pragma unroll
        for (int s = 0; s < S; ++s) {
            x *= (double)(1 << WB);                 // exact
            const double q = trunc(x);              // |q| <= 127
            Q[((size_t)s * M + m) * K + i] = (int8_t)(int)q;
            x -= q;                                 // exact
        }
    }
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {     // bounded: never hang the device
    uint32_t done = 0;
    for (uint32_t it = 0; it < SPIN_MAX; ++it) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major operand tile, 64-byte swizzle: rows of 64 bytes, 8-row groups 512 bytes apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                         // descriptor version
    d |= (uint64_t)4 << 61;                         // SWIZZLE_64B
    return d;
}
// instruction descriptor, kind::i8: D = S32, A = B = signed int8, both K-major, M = 128, N = 64
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// ------------------------------------------------------------------------------------------------ the GEMM
// tile list: tiles[t] = (bm, bn) with bn*BN < (bm+1)*BM  (touches the lower triangle)
__global__ void __launch_bounds__(NTHREADS, 1) k_ozaki_syrk(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                            int K, int M, const int2* __restrict__ tiles, const int* __restrict__ expo,
                                                            double* __restrict__ C, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
    __shared__ uint32_t tmem_base_sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bm = tiles[blockIdx.x].x, bn = tiles[blockIdx.x].y;
    const int nkb = K / BKB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                   // the whole tensor memory of the SM: 8 accumulators x 64 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sm)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sm;

    if (warp == 4 && lane == 0) {
        // ---------------- TMA producer
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % STAGES;
            if (kb >= STAGES && !mbar_wait(&empty_bar[st], ((kb / STAGES) - 1) & 1, err)) break;
            uint8_t* sa = smem + (size_t)st * STAGE_BYTES;
            uint8_t* sb = sa + S * A_TILE;
            mbar_expect_tx(&full_bar[st], STAGE_BYTES);
            tma_load_3d(sa, &mapA, &full_bar[st], kb * BKB, bm * BM, 0);      // box (64 B of K, 128 rows, 8 digits)
            tma_load_3d(sb, &mapB, &full_bar[st], kb * BKB, bn * BN, 0);      // box (64 B of K,  64 rows, 8 digits)
        }
    } else if (warp == 5 && lane == 0) {
        // ---------------- MMA issuer
        uint32_t started = 0;                           // bit d: accumulator d has been written once
        bool ok = true;
        for (int kb = 0; kb < nkb && ok; ++kb) {
            const int st = kb % STAGES;
            ok = mbar_wait(&full_bar[st], (kb / STAGES) & 1, err);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
            const uint32_t sb = sa + S * A_TILE;
#pragma unroll 1
            for (int t = 0; t < S; ++t) {
#pragma unroll 1
                for (int s = 0; s + t < S; ++s) {
                    const int d = s + t;
                    const uint64_t ad = umma_desc_k_sw64(sa + s * A_TILE), bd = umma_desc_k_sw64(sb + t * B_TILE);
#pragma unroll
                    for (int k2 = 0; k2 < BKB / 32; ++k2) {             // UMMA_K = 32 bytes: advance the start address inside the swizzle row
                        umma_i8(tmem_base + d * BN, ad + (uint64_t)(k2 * 2), bd + (uint64_t)(k2 * 2), (started >> d) & 1u);
                        started |= 1u << d;
                    }
                }
            }
            umma_commit(&empty_bar[st]);                // the stage may be refilled once these MMAs have read it
        }
        umma_commit(&accum_bar);                         // all accumulators final
    } else if (warp < 4) {
        // ---------------- epilogue: warp w owns TMEM lanes 32w .. 32w+31 = tile rows
        if (mbar_wait(&accum_bar, 0, err)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = bm * BM + warp * 32 + lane;
            const int em = expo[m];
            const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {      // 32 columns at a time (register budget)
                double h[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) h[j] = 0.0;
#pragma unroll 1
                for (int d = S - 1; d >= 0; --d) {
                    uint32_t r[32];
                    tmem_ld32(lane_base + d * BN + half * 32, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) h[j] = fma(h[j], 1.0 / (1 << WB), (double)(int)r[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = bn * BN + half * 32 + j;
                    if (n < M && m < M) C[(size_t)n * M + m] = ldexp(h[j], em + expo[n] - 2 * WB);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ host
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

static CUtensorMap make_map(PFN_cuTensorMapEncodeTiled_v12000 enc, const int8_t* Q, int K, int M, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)K * M};        // bytes, dims 1 and 2
    cuuint32_t box[3] = {(cuuint32_t)BKB, (cuuint32_t)box_rows, (cuuint32_t)S};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)Q, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"error\": \"cuTensorMapEncodeTiled failed: %d\"}\n", (int)r); exit(1); }
    return m;
}

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 2048, M = argc > 2 ? atoi(argv[2]) : 4096;
    const int reps = argc > 3 ? atoi(argv[3]) : 10;
    if (K % BKB || M % BM) { printf("{\"error\": \"K must be a multiple of %d and M of %d\"}\n", BKB, BM); return 1; }
    // A = sqrt(D) .* J:  J ~ N(0,1)/sqrt(M),  D log-uniform in [1e-8, 1e8] per ROW (the contraction index) -- SURVEY 8d C2
    std::vector<double> hA((size_t)K * M);
    {
        std::mt19937_64 rng(1);
        std::normal_distribution<double> nd(0.0, 1.0);
        std::uniform_real_distribution<double> ud(-8.0, 8.0);
        std::vector<double> sd(K);
        for (int i = 0; i < K; ++i) sd[i] = sqrt(pow(10.0, ud(rng)));
        for (int m = 0; m < M; ++m)
            for (int i = 0; i < K; ++i) hA[(size_t)m * K + i] = sd[i] * nd(rng) / sqrt((double)M);
    }
    double *dA, *dC, *dRef;
    int8_t* dQ;
    int *dE, *dErr;
    CK(cudaMalloc(&dA, sizeof(double) * K * M));
    CK(cudaMalloc(&dC, sizeof(double) * M * M));
    CK(cudaMalloc(&dRef, sizeof(double) * M * M));
    CK(cudaMalloc(&dQ, (size_t)S * M * K));
    CK(cudaMalloc(&dE, sizeof(int) * M));
    CK(cudaMalloc(&dErr, sizeof(int)));
    CK(cudaMemset(dErr, 0, sizeof(int)));
    CK(cudaMemset(dC, 0, sizeof(double) * M * M));
    CK(cudaMemcpy(dA, hA.data(), sizeof(double) * K * M, cudaMemcpyHostToDevice));
    std::vector<int2> tiles;
    for (int bm = 0; bm < M / BM; ++bm)
        for (int bn = 0; bn * BN < (bm + 1) * BM; ++bn) tiles.push_back(make_int2(bm, bn));
    int2* dT;
    CK(cudaMalloc(&dT, sizeof(int2) * tiles.size()));
    CK(cudaMemcpy(dT, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice));
    auto enc = get_encode();
    CUtensorMap mapA = make_map(enc, dQ, K, M, BM), mapB = make_map(enc, dQ, K, M, BN);
    CK(cudaFuncSetAttribute(k_ozaki_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));

    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float ms_split = 1e30f, ms_gemm = 1e30f;
    for (int r = 0; r < reps + 2; ++r) {
        CK(cudaEventRecord(e0));
        k_split<<<M, 256>>>(K, M, dA, dQ, dE);
        CK(cudaEventRecord(e1));
        k_ozaki_syrk<<<(int)tiles.size(), NTHREADS, SMEM_BYTES>>>(mapA, mapB, K, M, dT, dE, dC, dErr);
        CK(cudaEventRecord(e2));
        CK(cudaEventSynchronize(e2));
        CK(cudaGetLastError());
        float a, b;
        CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
        if (r >= 2) { ms_split = fminf(ms_split, a); ms_gemm = fminf(ms_gemm, b); }
    }
    int herr = 0;
    CK(cudaMemcpy(&herr, dErr, sizeof(int), cudaMemcpyDeviceToHost));
    // reference + library bars: cuBLAS DGEMM (what the reference calls, Dense/condensed.jl:171) and DSYRK
    cublasHandle_t hb;
    cublasCreate(&hb);
    const double one = 1.0, zero = 0.0;
    float ms_dgemm = 1e30f, ms_dsyrk = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        cublasDgemm(hb, CUBLAS_OP_T, CUBLAS_OP_N, M, M, K, &one, dA, K, dA, K, &zero, dRef, M);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float a; CK(cudaEventElapsedTime(&a, e0, e1));
        if (r) ms_dgemm = fminf(ms_dgemm, a);
    }
    {
        double* dTmp;
        CK(cudaMalloc(&dTmp, sizeof(double) * M * M));
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0));
            cublasDsyrk(hb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, M, K, &one, dA, K, &zero, dTmp, M);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float a; CK(cudaEventElapsedTime(&a, e0, e1));
            if (r) ms_dsyrk = fminf(ms_dsyrk, a);
        }
        cudaFree(dTmp);
    }
    std::vector<double> hC((size_t)M * M), hR((size_t)M * M);
    CK(cudaMemcpy(hC.data(), dC, sizeof(double) * M * M, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hR.data(), dRef, sizeof(double) * M * M, cudaMemcpyDeviceToHost));
    double maxabs = 0.0, maxerr = 0.0;
    for (int n = 0; n < M; ++n)
        for (int m = n; m < M; ++m) {
            const double r = hR[(size_t)n * M + m];
            maxabs = fmax(maxabs, fabs(r));
            maxerr = fmax(maxerr, fabs(hC[(size_t)n * M + m] - r));
        }
    // exactness of the integer part on a sample: recompute W(m,n) from the fp64 definition in long double
    double maxerr_ld = 0.0;
    for (int q = 0; q < 64; ++q) {
        const int m = (q * 977 + 13) % M, n = (q * 131) % (m + 1);
        long double acc = 0.0L;
        for (int i = 0; i < K; ++i) acc += (long double)hA[(size_t)m * K + i] * (long double)hA[(size_t)n * K + i];
        maxerr_ld = fmax(maxerr_ld, fabs((double)(acc - (long double)hC[(size_t)n * M + m])));
    }
    const double flop = (double)M * (M + 1) * K;                                  // SYRK, lower triangle
    const double tile_ops = (double)tiles.size() * BM * BN * (double)K * 2.0 * (S * (S + 1) / 2);
    printf("{\"bench\": \"ozaki_int8_syrk_tcgen05\", \"K\": %d, \"n\": %d, \"digits\": %d, \"bits_per_digit\": %d, \"int8_gemms\": %d, "
           "\"tile\": \"%dx%d, K block %d, %d stages\", \"tiles\": %d, \"err_flag\": %d, "
           "\"max_abs_err_vs_cublas_dgemm\": %.3e, \"max_abs_W\": %.3e, \"rel_err\": %.3e, \"rel_err_vs_long_double_sample\": %.3e, \"parity_gate_1e-13\": %s, "
           "\"ms_split\": %.4f, \"ms_gemm\": %.4f, \"ms_total\": %.4f, \"int8_tops\": %.1f, \"fp64_equiv_tflops\": %.2f, "
           "\"cublas_dgemm_ms\": %.4f, \"cublas_dsyrk_ms\": %.4f, \"cublas_dsyrk_tflops\": %.2f}\n",
           K, M, S, WB, S * (S + 1) / 2, BM, BN, BKB, STAGES, (int)tiles.size(), herr, maxerr, maxabs, maxerr / maxabs, maxerr_ld / maxabs,
           (maxerr / maxabs <= 1e-13 && herr == 0) ? "true" : "false", ms_split, ms_gemm, ms_split + ms_gemm, tile_ops / (ms_gemm * 1e-3) / 1e12,
           flop / ((ms_split + ms_gemm) * 1e-3) / 1e12, ms_dgemm, ms_dsyrk, flop / (ms_dsyrk * 1e-3) / 1e12);
    return 0;
}
