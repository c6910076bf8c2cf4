// gram_umma.cuh -- the Gram-matrix product  out[c][i] = sum_j G_ij X[c][j],  G_ij = 2^-(|a_i - a_j|^2),  on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).  It replaces the CUDA-core kernel lr_gram_apply_kernel
// (lowrank.cuh) in the randomised range finder of the low-rank non-rigid path; the op it stands for in the reference is the dense
// product with G = rbf_kernel(Y, Y, beta) of probreg/transformation.py:91-102 + probreg/cpd.py:296-297 (G is never stored here).
//
// Shape of the problem: an M x M x K GEMM (M = points, K = rank <= 256) whose A operand does not exist in memory -- it is
// generated tile by tile by the CUDA cores (3 sub, 3 fma-ish, one MUFU.EX2 per element) straight into the shared-memory layout
// tcgen05.mma reads.  One generated A tile is multiplied into ALL K columns (the old kernel regenerated G for every 16 columns).
//
// Precision: a single TF32 product is ~800x too coarse for the moved points (profiles/r1_tf32_split_probe.txt); the 3-term split
//     G X  ~=  G_hi X_hi + G_lo X_hi + G_hi X_lo,        v_hi = v with the low 13 mantissa bits cleared,  v_lo = v - v_hi
// is indistinguishable from the FP32 CUDA-core product.  Accumulation is FP32 in TMEM over one j-chunk (GU_CHUNK points), the
// chunk partials are summed in FP64 in a fixed order by gu_reduce_kernel => bit-reproducible, independent of the row sharding.
//
// CTA = 256 rows of G (two M = 128 MMAs share every B tile: halves the L2 traffic of X, which at M = 128 per CTA would sit at the
// ~12 TB/s L2 limit), persistent over work units {row tile, j-chunk}.  Warp roles (14 warps):
//     warp 0      TMA producer: X_hi / X_lo tiles [N x 16] -> shared memory, 64-byte swizzle (cp.async.bulk.tensor, SASS UTMALDG)
//     warp 1      TMEM allocation; one lane issues the MMAs (12 per 16-point stage) and the commits
//     warps 2-5   epilogue: tcgen05.ld the 2 x [128 x N] FP32 accumulators, store the chunk partial (coalesced along i)
//     warps 6-13  generators: thread r owns row r of the tile, writes G_hi / G_lo rows in the swizzled K-major layout
// Three 64 KB stages {A_hi, A_lo, B_hi, B_lo}; mbarriers full_a (8 generator warps), full_b (TMA bytes), empty (tcgen05.commit),
// acc_full (commit after a unit's last MMA), acc_empty (4 epilogue warps).
//
// Shared-memory operand layout (both operands K-major, SWIZZLE_64B; encodings as in cute/arch/mma_sm100_desc.hpp):
//     row r of a tile starts at byte 64 r (8-row groups 512 B apart = the descriptor's stride byte offset); within the row the
//     16-byte chunk c (4 floats) sits at chunk position c ^ ((r >> 1) & 3).  One MMA consumes K = 8 floats = 32 bytes of every
//     row; the second K-step of a stage is reached by advancing the descriptor start address by 32 bytes.
#pragma once
#ifndef CPD_HOST_EMU
#include <cuda.h>          // CUtensorMap (type only: the encode entry point is fetched from the driver at run time)
#include "kernels.cuh"

namespace cpd {

constexpr int GU_ROWS = 256;           // rows of G per CTA (2 x UMMA M = 128)
constexpr int GU_KS = 16;              // j-points per pipeline stage (64-byte rows)
constexpr int GU_NMAX = 256;           // most columns (UMMA N) of one launch
constexpr int GU_STAGES = 3;
constexpr int GU_A_BYTES = GU_ROWS * GU_KS * 4;        // 16 KB: one of A_hi / A_lo
constexpr int GU_B_BYTES = GU_NMAX * GU_KS * 4;        // 16 KB: one of B_hi / B_lo (N x 64 bytes used)
constexpr int GU_STAGE_BYTES = 2 * GU_A_BYTES + 2 * GU_B_BYTES;
constexpr int GU_THREADS = 14 * 32;
constexpr int GU_SMEM = GU_STAGES * GU_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers, TMEM base */;
constexpr uint32_t GU_TF32_MASK = 0xffffe000u;

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gu_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait.  With -DGU_DEBUG_WAIT (tools/umma_probe.cu) a wait gives up after ~1 s, records which one it was in
// gu_timeout_code and lets every later wait fall through, so a protocol error ends the kernel with a diagnosis instead of a hang.
#ifdef GU_DEBUG_WAIT
__device__ int gu_timeout_code = 0;
__device__ __forceinline__ void gu_wait(uint64_t* bar, uint32_t parity, int code) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (*(volatile int*)&gu_timeout_code != 0) return;
        if (clock64() - t0 > 2000000000ll) { atomicCAS(&gu_timeout_code, 0, code); return; }
    }
}
#else
__device__ __forceinline__ void gu_wait(uint64_t* bar, uint32_t parity, int) { mbar_wait(bar, parity); }
#endif
// a wait that is not on the critical path (the epilogue warps wait a whole work unit for acc_full): poll with a pause, so that the
// waiting warps do not take issue slots from the generators (their spin loops were 30 % of all issued instructions)
__device__ __forceinline__ void gu_wait_relaxed(uint64_t* bar, uint32_t parity, int code) {
#ifdef GU_DEBUG_WAIT
    gu_wait(bar, parity, code);
#else
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(400);
    }
#endif
}
// one lane of a converged warp (elect.sync).  The MMA-issuing warp runs its loop with all 32 lanes and elects only around the
// tcgen05 instructions: the descriptor arithmetic then stays warp-uniform and compiles to the uniform datapath.  Inside an
// `if (lane == 0)` branch the same values are per-thread registers and every tcgen05.mma needs ~8 R2UR transfers on one thread --
// the issue loop, not the tensor pipe, set the pace (tensor pipe 37 % active, profiles/r2_ncu_gi_gram_v5.txt).
__device__ __forceinline__ bool gu_elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "     elect.sync %%rx|%%px, %2;\n"
        "@%%px mov.s32 %1, 1;\n"
        "     mov.s32 %0, %%rx;\n"
        "}\n"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFF));
    return pred != 0;
}
__device__ __forceinline__ void gu_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void gu_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gu_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gu_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void gu_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of every tcgen05.mma issued so far by this thread -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void gu_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, one K-step (8 TF32 values per row)
__device__ __forceinline__ void gu_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void gu_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// 16 consecutive FP32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void gu_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}

// ---- descriptors --------------------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_64B (cute SmemDescriptor: start >> 4 in [0,14), leading byte offset >> 4 in
// [16,30) (= 1, unused for swizzled K-major), stride byte offset >> 4 in [32,46) (8 rows x 64 B = 512), version 1 in [46,48),
// layout type in [61,64): 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t gu_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// instruction descriptor (cute InstrDescriptor): D = F32 (1 at [4,6)), A = B = TF32 (2 at [7,10) and [10,13)), both K-major
// (0 at 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ __forceinline__ uint32_t gu_instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of the 16-byte chunk c (0..3) of row r inside a tile with 64-byte rows, 64-byte swizzle
__device__ __forceinline__ uint32_t gu_row_chunk_offset(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// ---- operand preparation: X (FP64, [rank][ld]) -> TF32 hi / lo planes ([2 n16][ldx] floats: rows 0..n16 hi, n16..2 n16 lo) -------
__global__ void __launch_bounds__(THREADS)
gu_split_kernel(const double* __restrict__ X, long long m, long long ld, int rank, int n16, long long ldx, float* __restrict__ planes) {
    const long long j = (long long)blockIdx.x * THREADS + threadIdx.x;
    const int c = blockIdx.y;
    if (j < ldx && c < n16) {
        float hi = 0.0f, lo = 0.0f;
        if (j < m && c < rank) {
            const double x = X[(long long)c * ld + j];
            hi = __uint_as_float(__float_as_uint((float)x) & GU_TF32_MASK);
            lo = (float)(x - (double)hi);
        }
        planes[(long long)c * ldx + j] = hi;
        planes[(long long)(n16 + c) * ldx + j] = lo;
    }
}

// out[c][i_begin + ii] = sum over the j-chunks of part[q][c][ii], in chunk order, FP64.  Four rows per thread (one 16-byte load
// per chunk, four independent sums): the kernel is a pure stream over the partials (nq x n16 x ldp floats, read once).
__global__ void __launch_bounds__(THREADS)
gu_reduce_kernel(const float* __restrict__ part, int nq, int n16, long long ldp, int rank, long long rows, long long i_begin, long long ld,
                 double* __restrict__ out) {
    const long long ii = ((long long)blockIdx.x * THREADS + threadIdx.x) * 4;      // ldp is a multiple of 256: 16-byte aligned
    const int c = blockIdx.y;
    if (ii < rows && c < rank) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const float* p = part + (long long)c * ldp + ii;
        const long long step = (long long)n16 * ldp;
#pragma unroll 4
        for (int q = 0; q < nq; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + (long long)q * step));
            s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
        }
        double* o = out + (long long)c * ld + i_begin + ii;
        o[0] = s0;
        if (ii + 1 < rows) o[1] = s1;
        if (ii + 2 < rows) o[2] = s2;
        if (ii + 3 < rows) o[3] = s3;
    }
}

// first-use self-check: the largest |a - b| and |b| over rows [0, rows) of `cols` columns  ->  res[0], res[1] (one CTA)
__global__ void __launch_bounds__(THREADS)
gu_compare_kernel(const double* __restrict__ a, const double* __restrict__ b, long long ld, long long i_begin, long long rows, int cols,
                  double* __restrict__ res) {
    __shared__ double sd[THREADS], sv[THREADS];
    double d = 0.0, v = 0.0;
    for (int c = 0; c < cols; ++c)
        for (long long i = threadIdx.x; i < rows; i += THREADS) {
            const double x = a[(long long)c * ld + i_begin + i], y = b[(long long)c * ld + i_begin + i];
            d = fmax(d, fabs(x - y)); v = fmax(v, fabs(y));
            if (!(x == x)) d = 1e300;                      // NaN in the tensor-core result
        }
    sd[threadIdx.x] = d; sv[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < THREADS; ++t) { d = fmax(d, sd[t]); v = fmax(v, sv[t]); }
        res[0] = d; res[1] = v;
    }
}

// ---- the product ----------------------------------------------------------------------------------------------------------------
// pts: float4 {a, 0} per point in the scaled frame of lr_pack_kernel, padded to jpad with far-away records (G == 0);
// rows [i_begin, i_end) of G, all jpad columns, in chunks of `chunk` points (a multiple of GU_KS);
// part[q][c][ii]: FP32 partial of chunk q, column c < n16, row ii = i - i_begin (ldp >= ntiles * GU_ROWS).
__global__ void __launch_bounds__(GU_THREADS, 1)
gu_gram_kernel(const __grid_constant__ CUtensorMap xmap, const float4* __restrict__ pts, long long jpad, int chunk, long long i_begin,
               long long i_end, int n16, float* __restrict__ part, long long ldp) {
    extern __shared__ __align__(1024) unsigned char gu_smem_raw[];
    // 1024-byte alignment (declared on the array, honoured for dynamic shared memory) keeps the swizzle pattern of every tile
    // anchored the way the TMA unit and the MMA unit both expect.  NOT aligned by pointer arithmetic: that launders the address
    // space and every access through the pointer becomes a generic LD/ST instead of LDS/STS (measured: the j-point loads of the
    // int8 kernel showed up as LD.E.64 with long-scoreboard stalls, profiles/r2_ncu_gi_gram_v3_source.txt)
    unsigned char* smem = gu_smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GU_STAGES * GU_STAGE_BYTES);
    uint64_t* full_a = bars;                    // [GU_STAGES]
    uint64_t* full_b = bars + GU_STAGES;        // [GU_STAGES]
    uint64_t* empty = bars + 2 * GU_STAGES;     // [GU_STAGES]
    uint64_t* acc_full = bars + 3 * GU_STAGES;
    uint64_t* acc_empty = bars + 3 * GU_STAGES + 1;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 3 * GU_STAGES + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rows = i_end - i_begin;
    const int ntiles = (int)((rows + GU_ROWS - 1) / GU_ROWS);
    const int nq = (int)((jpad + chunk - 1) / chunk);
    const long long nunits = (long long)ntiles * nq;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GU_STAGES; ++s) { mbar_init(&full_a[s], 8); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        mbar_fence_init();
    }
    if (warp == 1) gu_tmem_alloc(tmem_base_slot, 512);
    gu_tc_fence_before();
    __syncthreads();
    gu_tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            uint32_t stage = 0, phase = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
                const int q = (int)(u / ntiles);
                const long long j0 = (long long)q * chunk;
                const int nst = (int)((min((long long)chunk, jpad - j0)) / GU_KS);
                for (int kb = 0; kb < nst; ++kb) {
                    gu_wait(&empty[stage], phase ^ 1, 1);
                    unsigned char* sb = smem + stage * GU_STAGE_BYTES + 2 * GU_A_BYTES;
                    mbar_expect_tx(&full_b[stage], (uint32_t)(2 * n16 * GU_KS * 4));
                    gu_tma_load_2d(sb, &xmap, (int)(j0 + (long long)kb * GU_KS), 0, &full_b[stage]);
                    gu_tma_load_2d(sb + GU_B_BYTES, &xmap, (int)(j0 + (long long)kb * GU_KS), n16, &full_b[stage]);
                    if (++stage == GU_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the pipeline, one elected lane issues =====
        {
            const uint32_t idesc = gu_instr_desc(n16);
            uint32_t stage = 0, phase = 0, acc_phase = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
                const int q = (int)(u / ntiles);
                const long long j0 = (long long)q * chunk;
                const int nst = (int)((min((long long)chunk, jpad - j0)) / GU_KS);
                gu_wait(acc_empty, acc_phase ^ 1, 2);          // the epilogue has drained the previous unit's accumulators
                gu_tc_fence_after();
                for (int kb = 0; kb < nst; ++kb) {
                    gu_wait(&full_a[stage], phase, 3);
                    gu_wait(&full_b[stage], phase, 4);
                    gu_tc_fence_after();
                    const uint32_t sa = smem_u32(smem) + stage * GU_STAGE_BYTES;
                    const uint32_t sb = sa + 2 * GU_A_BYTES;
                    if (gu_elect_one()) {
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint32_t d = tmem_base + (uint32_t)(half * n16);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const uint64_t a_hi = gu_smem_desc(sa + half * (128 * 64) + kk * 32);
                                const uint64_t a_lo = gu_smem_desc(sa + GU_A_BYTES + half * (128 * 64) + kk * 32);
                                const uint64_t b_hi = gu_smem_desc(sb + kk * 32);
                                const uint64_t b_lo = gu_smem_desc(sb + GU_B_BYTES + kk * 32);
                                gu_mma_tf32(d, a_lo, b_hi, idesc, (kb | kk) != 0 ? 1u : 0u);     // small terms first
                                gu_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                                gu_mma_tf32(d, a_hi, b_hi, idesc, 1u);
                            }
                        }
                        gu_commit(&empty[stage]);                     // frees the stage once these MMAs have read it
                        if (kb == nst - 1) gu_commit(acc_full);       // ... and tells the epilogue the unit is complete
                    }
                    __syncwarp();
                    if (++stage == GU_STAGES) { stage = 0; phase ^= 1; }
                }
                acc_phase ^= 1;
            }
        }
    } else if (warp < 6) {
        // ===== epilogue: TMEM -> chunk partial =====
        const int quarter = warp & 3;                             // the TMEM lanes this warp may read: 32 quarter .. +31
        uint32_t acc_phase = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            const int q = (int)(u / ntiles), t = (int)(u % ntiles);
            gu_wait_relaxed(acc_full, acc_phase, 5);
            gu_tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const long long ii = (long long)t * GU_ROWS + half * 128 + quarter * 32 + lane;
                float* dst = part + (long long)q * n16 * ldp + ii;
#pragma unroll 1
                for (int c0 = 0; c0 < n16; c0 += 16) {
                    float v[16];
                    gu_tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * n16 + c0), v);
#pragma unroll
                    for (int k = 0; k < 16; ++k) dst[(long long)(c0 + k) * ldp] = v[k];
                }
            }
            gu_tc_fence_before();
            __syncwarp();
            if (lane == 0) gu_mbar_arrive(acc_empty);
            acc_phase ^= 1;
        }
    } else {
        // ===== generators: thread r writes row r of G_hi / G_lo =====
        const int r = threadIdx.x - 6 * 32;
        uint32_t stage = 0, phase = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            const int q = (int)(u / ntiles), t = (int)(u % ntiles);
            const long long j0 = (long long)q * chunk;
            const int nst = (int)((min((long long)chunk, jpad - j0)) / GU_KS);
            long long i = i_begin + (long long)t * GU_ROWS + r;
            if (i >= i_end) i = i_end - 1;                        // rows past the end: computed, never read back
            const float4 a = pts[i];
            const float4* bj = pts + j0;
            for (int kb = 0; kb < nst; ++kb) {
                float e[GU_KS];
#pragma unroll
                for (int jj = 0; jj < GU_KS; ++jj) {
                    const float4 b = __ldg(bj + kb * GU_KS + jj);
                    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
                    e[jj] = ex2(-fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                }
                gu_wait(&empty[stage], phase ^ 1, 6);
                unsigned char* sa = smem + stage * GU_STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(e[4 * c]) & GU_TF32_MASK);
                    hi.y = __uint_as_float(__float_as_uint(e[4 * c + 1]) & GU_TF32_MASK);
                    hi.z = __uint_as_float(__float_as_uint(e[4 * c + 2]) & GU_TF32_MASK);
                    hi.w = __uint_as_float(__float_as_uint(e[4 * c + 3]) & GU_TF32_MASK);
                    lo.x = e[4 * c] - hi.x; lo.y = e[4 * c + 1] - hi.y; lo.z = e[4 * c + 2] - hi.z; lo.w = e[4 * c + 3] - hi.w;
                    const uint32_t off = gu_row_chunk_offset(r, c);
                    *reinterpret_cast<float4*>(sa + off) = hi;
                    *reinterpret_cast<float4*>(sa + GU_A_BYTES + off) = lo;
                }
                gu_fence_async_smem();                            // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) gu_mbar_arrive(&full_a[stage]);
                if (++stage == GU_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    gu_tc_fence_before();
    __syncthreads();
    if (warp == 1) gu_tmem_dealloc(tmem_base, 512);
}

// ---- layout probe (tools/umma_probe.cu, first-use self-check): D[128][n16] = A[128][16] * B[n16][16]^T through exactly the
// descriptor / swizzle / TMA / TMEM conventions of the kernel above, one CTA of 128 threads, plain FP32 inputs (TF32-truncated
// by the tensor core).  bmap: tensor map over B ([n16][16] floats), box {16, n16}, 64-byte swizzle.
__global__ void __launch_bounds__(128, 1)
gu_layout_probe_kernel(const __grid_constant__ CUtensorMap bmap, const float* __restrict__ A, int n16, float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char gu_smem_raw[];
    unsigned char* smem = gu_smem_raw;
    unsigned char* sa = smem;                              // 128 rows x 64 B
    unsigned char* sb = smem + 8192;                       // n16 rows x 64 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8192 + GU_B_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, r = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (warp == 0) gu_tmem_alloc(slot, 256);
    gu_tc_fence_before();
    __syncthreads();
    gu_tc_fence_after();
    const uint32_t tmem_base = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], (uint32_t)(n16 * 64));
        gu_tma_load_2d(sb, &bmap, 0, 0, &bars[0]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(sa + gu_row_chunk_offset(r, c)) = *reinterpret_cast<const float4*>(A + r * 16 + 4 * c);
    gu_fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        gu_wait(&bars[0], 0, 7);
        gu_tc_fence_after();
        const uint32_t idesc = gu_instr_desc(n16);
        for (int kk = 0; kk < 2; ++kk)
            gu_mma_tf32(tmem_base, gu_smem_desc(smem_u32(sa) + kk * 32), gu_smem_desc(smem_u32(sb) + kk * 32), idesc, kk);
        gu_commit(&bars[1]);
    }
    gu_wait(&bars[1], 0, 8);
    gu_tc_fence_after();
    for (int c0 = 0; c0 < n16; c0 += 16) {
        float v[16];
        gu_tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int k = 0; k < 16; ++k) D[r * n16 + c0 + k] = v[k];
    }
    gu_tc_fence_before();
    __syncthreads();
    if (warp == 0) gu_tmem_dealloc(tmem_base, 256);
}


// ---- host: tensor map over the hi / lo planes (cuTensorMapEncodeTiled is fetched from the driver: no link against libcuda) ------
typedef CUresult (*gu_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// 2-D FP32 tensor [rows][ld] (ld floats per row, contiguous), box = {GU_KS floats, box_rows}, 64-byte swizzle.  0 on success.
inline int gu_make_map(CUtensorMap* map, const float* base, long long ld, long long rows, int box_rows) {
    static gu_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
        encode = reinterpret_cast<gu_encode_fn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)GU_KS, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace cpd
#endif  // CPD_HOST_EMU
