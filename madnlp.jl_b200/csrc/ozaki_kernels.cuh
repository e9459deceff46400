// Ozaki-scheme fp64 SYRK on the 5th-generation tensor cores: W = A' A with A (K x n) cut into S = 8 signed 7-bit digits per
// entry, the 36 digit-pair products run as EXACT int8 GEMMs on tcgen05.mma.kind::i8 (accumulators in TMEM), operands staged by
// TMA (cp.async.bulk.tensor, 64-byte swizzle), fp64 reconstruction in the epilogue.  See tools/microbench/ozaki_syrk_tcgen05.cu
// for the stand-alone measurement and DESIGN.md section 3 for the error analysis; used by b2d_condensed_assemble_ozaki
// (build_kkt!(::DenseCondensedKKTSystem), src/KKT/Dense/condensed.jl:157-186, in place of cuBLAS mul!(W, J', J)).
//
//   1. per column m:  e_m = exponent of max_i |a_im|;  x = a_im * 2^-e_m in (-1, 1) is cut into S = 8 signed 7-bit digits
//      x = sum_s q_s 2^(-7(s+1))   (q_s int8, exact: 56 bits cover the fp64 mantissa of the column's largest entries)
//   2. G_d = sum_{s+t=d} Q_s' Q_t  for d = 0..S-1 : 36 exact int8 x int8 -> int32 GEMMs (|G_d| <= 8 * K * 127^2 < 2^31 for
//      K <= 16384), the d-sums accumulate inside the tensor-core accumulators (8 accumulators of 64 columns = the SM's whole TMEM)
//   3. W(m,n) = 2^(e_m + e_n - 14) * sum_d 2^(-7d) G_d(m,n)   evaluated in fp64 (Horner) by the epilogue
//
// Kernel (one 128 x 64 output tile per CTA, 192 threads):
//   warp 4 lane 0 : TMA producer  -- the 8 A-digit tiles and 8 B-digit tiles of a 64-deep K block into a 2-stage ring
//   warp 5 lane 0 : MMA issuer    -- 72 tcgen05.mma.cta_group::1.kind::i8 (M128 N64 K32) per K block, tcgen05.commit frees the stage
//   warps 0..3    : epilogue      -- tcgen05.ld of the 8 accumulators, Horner in fp64, scaling, coalesced stores
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace ozk {

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
// one CTA per column m of the operand A (K x M): a(i, m) = (scale ? sqrt(scale[i]) : 1) * src[m*lds + (rows ? rows[i] : i)];
// exponent of the column, then S int8 digits per element into Q[s][m][i] (row length Kpad, rows beyond K stay zero)
__global__ void __launch_bounds__(256) k_ozaki_split(int K, int Kpad, int Mpad, const double* __restrict__ src, int64_t lds,
                                                     const int64_t* __restrict__ rows, const double* __restrict__ scale,
                                                     int8_t* __restrict__ Q, int* __restrict__ expo) {
    const int m = blockIdx.x;
    const double* col = src + (size_t)m * lds;
    __shared__ double red[256];
    double mx = 0.0;
    for (int i = threadIdx.x; i < K; i += 256) {
        const double a = col[rows ? rows[i] : i] * (scale ? sqrt(scale[i]) : 1.0);
        mx = fmax(mx, fabs(a));
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
    mx = red[0];
    int e = 0;
    if (mx > 0.0 && mx < 1e300) frexp(mx, &e);      // mx = f * 2^e, f in [0.5, 1)  ->  |a| * 2^-e < 1   (inf/nan columns: digits 0)
    if (threadIdx.x == 0) expo[m] = e;
    for (int i = threadIdx.x; i < K; i += 256) {
        double x = ldexp(col[rows ? rows[i] : i] * (scale ? sqrt(scale[i]) : 1.0), -e);    // exact scaling
        if (!(fabs(x) < 1.0)) x = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            x *= (double)(1 << WB);                 // exact
            const double q = trunc(x);              // |q| <= 127
            Q[((size_t)s * Mpad + m) * Kpad + i] = (int8_t)(int)q;
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
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

// 2^e for |e| <= 1022 (exponent field only)
__device__ __forceinline__ double pow2i(int e) { return __longlong_as_double((long long)(min(max(e, -1022), 1023) + 1023) << 52); }

// ------------------------------------------------------------------------------------------------ the GEMM
// tile list: tiles[t] = (bm, bn) with bn*BN < (bm+1)*BM  (touches the lower triangle)
// epilogue: C(m, n) = W(m, n) [+ hess(m, n)] [+ pr(m) on the diagonal] for m, n < nvalid (and m >= n when lower_only); ld = ldc / ldh
__global__ void __launch_bounds__(NTHREADS, 1) k_ozaki_syrk(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                            int K, int nvalid, const int2* __restrict__ tiles, const int* __restrict__ expo,
                                                            double* __restrict__ C, int64_t ldc, const double* __restrict__ hess, int64_t ldh,
                                                            const double* __restrict__ pr, int lower_only, int* err) {
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
    } else if (warp == 5) {
        // ---------------- MMA issuer: the WHOLE warp runs the (warp-uniform) loop so that descriptors and predicates live in uniform
        // registers; one elected lane issues.  The 72 MMAs of a K block are straight-line code: their operand descriptors differ from
        // the stage's base descriptors by compile-time constants (digit tile offset + 32-byte K step), the accumulator by d * 64
        // columns -- a single thread cannot afford ~100 cycles of address arithmetic per 32-cycle MMA (measured: v1 of this kernel).
        const bool leader = elect_one();
        bool ok = true;
        for (int kb = 0; kb < nkb && ok; ++kb) {
            const int st = kb % STAGES;
            ok = mbar_wait(&full_bar[st], (kb / STAGES) & 1, err);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
                const uint64_t ad0 = umma_desc_k_sw64(sa), bd0 = umma_desc_k_sw64(sa + S * A_TILE);
                const uint32_t acc_any = kb > 0 ? 1u : 0u;
#pragma unroll
                for (int t = 0; t < S; ++t) {
#pragma unroll
                    for (int s = 0; s < S - t; ++s) {
#pragma unroll
                        for (int k2 = 0; k2 < BKB / 32; ++k2) {
                            // first write of accumulator d = s + t: K block 0, t == 0, k2 == 0 (t is the outer loop, so every d is first met at t = 0)
                            const uint32_t acc = (t > 0 || k2 > 0) ? 1u : acc_any;
                            umma_i8(tmem_base + (uint32_t)((s + t) * BN), ad0 + (uint64_t)((s * A_TILE + k2 * 32) >> 4),
                                    bd0 + (uint64_t)((t * B_TILE + k2 * 32) >> 4), acc);
                        }
                    }
                }
                umma_commit(&empty_bar[st]);            // the stage may be refilled once these MMAs have read it
            }
            __syncwarp();
        }
        if (leader) umma_commit(&accum_bar);             // all accumulators final
        __syncwarp();
    } else if (warp < 4) {
        // ---------------- epilogue: warp w owns TMEM lanes 32w .. 32w+31 = tile rows
        if (mbar_wait(&accum_bar, 0, err)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = bm * BM + warp * 32 + lane;
            const int em = expo[m];
            const double pm = pow2i(em - 2 * WB);                 // 2^(e_m - 14): exact scaling factors, applied as two
                                                                       // multiplications (cheaper than one ldexp per entry)
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
                    if (n < nvalid && m < nvalid && (!lower_only || m >= n)) {
                        double v = (h[j] * pm) * pow2i(expo[n]);
                        if (hess) v += hess[(size_t)n * ldh + m];
                        if (pr && m == n) v += pr[m];
                        C[(size_t)n * ldc + m] = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}


}  // namespace ozk
