// Small-GEMM kernel on tcgen05 with in-kernel operand preparation ("manual fill"):
//
//   C[m,n] = sum_k A(m,k) * B(n,k)       A, B: fp32 functors (gathers, transposes, on-the-fly elementwise math)
//
// The producers evaluate the functors 8 consecutive k at a time, split every value into bf16 hi + lo
// (hi = rn(v), lo = rn(v - hi)) and store both into shared memory in the UMMA canonical K-major SWIZZLE_128B layout
// (row r of a 64-element k-block at r*128 B, its 16-byte chunk c at position c ^ (r & 7)) -- what TMA would have
// written had the operand existed in HBM as bf16.  One thread then issues the three bf16 passes
// A_lo*B_hi + A_hi*B_lo + A_hi*B_hi into one fp32 TMEM accumulator (fp32-grade product, see pph_similarity_tc.cu).
// This removes every "convert / transpose / gather into a staging buffer" pre-pass the add-on layer GEMMs would
// otherwise need in front of a TMA-fed kernel.
//
// CTA = 128 output rows (UMMA M = 128) x all N columns in tiles of BN <= 256, one k-split (blockIdx.y); 256 threads:
// all 8 warps fill, thread 0 issues MMAs, all 8 warps drain TMEM (warp w: lane quarter w & 3, column half w >> 2).
// Two shared-memory stages (fill of k-block i+1 overlaps the MMAs of k-block i).
#pragma once

#include "pph_common.cuh"
#include "pph_tc_ptx.cuh"

namespace pph {

constexpr int kTgThreads = 256;
constexpr int kTgBM = 128;
constexpr int kTgBK = 64;
constexpr int kTgMaxBN = 256;

__host__ __device__ inline size_t tcgemm_stage_bytes(int BN) { return (size_t)2 * (kTgBM + BN) * kTgBK * 2; }
inline size_t tcgemm_smem_bytes(int BN) { return 1024 + 2 * tcgemm_stage_bytes(BN) + 128 * 8 * 4 + 256; }

__device__ __forceinline__ void tg_store_split8(uint8_t* hi_tile, uint8_t* lo_tile, int row, int c, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint16_t h0 = bf16_bits(v[2 * i]), h1 = bf16_bits(v[2 * i + 1]);
        const uint16_t l0 = bf16_bits(v[2 * i] - bf16_to_float(h0)), l1 = bf16_bits(v[2 * i + 1] - bf16_to_float(h1));
        h[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        l[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    const int off = row * 128 + ((c ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Operand functor contract:
//   static constexpr bool kContigK;                              // memory-contiguous along k (else along rows)
//   __device__ void load8(int row, int k0, float (&v)[8]) const; // values (row, k0..k0+7); zeros outside the operand
// Epilogue functor contract:
//   struct State;  __device__ void init(State&) const;
//   __device__ void chunk(State&, int m, int n0, const uint32_t (&acc)[32]) const;   // columns n0..n0+31 of row m
//   __device__ void finish(State&, int m, int row_local, int half, float* scratch) const;  // after all column tiles
template <class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(kTgThreads, 1)
tcgemm_kernel(int M, int N, int Kd, int BN, int k_per_split, AOp a_op, BOp b_op, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const size_t stage_bytes = tcgemm_stage_bytes(BN);
    const int a_bytes = kTgBM * kTgBK * 2, b_bytes = BN * kTgBK * 2;
    float* scratch = reinterpret_cast<float*>(smem + 2 * stage_bytes);            // [128][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes + 128 * 8 * 4);
    uint64_t* empty = bars;            // [2]
    uint64_t* acc_done = bars + 2;     // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kTgBM;
    const int kz0 = blockIdx.y * k_per_split, kz1 = min(Kd, kz0 + k_per_split);

    if (tid == 0) {
        ptx::mbar_init(&empty[0], 1);
        ptx::mbar_init(&empty[1], 1);
        ptx::mbar_init(acc_done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_ptr, 256);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t idesc = ptx::umma_idesc_bf16(kTgBM, BN);

    typename Epi::State st;
    epi.init(st);
    const int quarter = warp & 3, half = warp >> 2;
    const int m_row = m0 + quarter * 32 + lane;
    int fill = 0;                                   // k-blocks filled so far (stage = fill & 1)
    int ntile = 0;
    for (int n0 = 0; n0 < N; n0 += BN, ++ntile) {
        for (int kb = kz0; kb < kz1; kb += kTgBK, ++fill) {
            const int stage = fill & 1, use = fill >> 1;
            if (use > 0) ptx::mbar_wait(&empty[stage], (uint32_t)(use - 1) & 1u);   // MMAs that read this stage retired
            uint8_t* sA_hi = smem + stage * stage_bytes;
            uint8_t* sA_lo = sA_hi + a_bytes;
            uint8_t* sB_hi = sA_lo + a_bytes;
            uint8_t* sB_lo = sB_hi + b_bytes;
            // ---- fill A: 128 rows x 8 chunks
#pragma unroll 2
            for (int g = tid; g < kTgBM * 8; g += kTgThreads) {
                const int row = AOp::kContigK ? (g >> 3) : (g & (kTgBM - 1));
                const int c = AOp::kContigK ? (g & 7) : (g >> 7);
                float v[8];
                a_op.load8(m0 + row, kb + c * 8, v);
                tg_store_split8(sA_hi, sA_lo, row, c, v);
            }
            // ---- fill B: BN rows x 8 chunks
#pragma unroll 2
            for (int g = tid; g < BN * 8; g += kTgThreads) {
                const int row = BOp::kContigK ? (g >> 3) : (g % BN);
                const int c = BOp::kContigK ? (g & 7) : (g / BN);
                float v[8];
                b_op.load8(n0 + row, kb + c * 8, v);
                tg_store_split8(sB_hi, sB_lo, row, c, v);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                ptx::tc_fence_after();
                const uint32_t a_hi = ptx::smem_u32(sA_hi), a_lo = ptx::smem_u32(sA_lo);
                const uint32_t b_hi = ptx::smem_u32(sB_hi), b_lo = ptx::smem_u32(sB_lo);
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t a = term == 0 ? a_lo : a_hi, b = term == 1 ? b_lo : b_hi;
#pragma unroll
                    for (int kk = 0; kk < kTgBK / 16; ++kk)
                        ptx::mma_bf16_ss(tmem_base, ptx::umma_desc_k_sw128(a + kk * 32), ptx::umma_desc_k_sw128(b + kk * 32),
                                         idesc, (uint32_t)(kb != kz0 || term != 0 || kk != 0));
                }
                ptx::mma_commit(&empty[stage]);
                if (kb + kTgBK >= kz1) ptx::mma_commit(acc_done);
            }
        }
        // ---- epilogue of this column tile
        ptx::mbar_wait(acc_done, (uint32_t)ntile & 1u);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int ncols = min(BN, N - n0);
        for (int c0 = half * 32; c0 < ncols; c0 += 64) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(taddr + c0, v);
            ptx::tmem_ld_wait(v);
            if (m_row < M) epi.chunk(st, m_row, n0 + c0, v);
        }
        ptx::tc_fence_before();
        __syncthreads();                           // every TMEM read done before the next tile's MMAs overwrite
        ptx::tc_fence_after();
    }
    epi.finish(st, m_row, quarter * 32 + lane, half, scratch, m_row < M);
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 256);
    }
}

template <class AOp, class BOp, class Epi>
inline int launch_tcgemm(int M, int N, int Kd, int BN, int splits, AOp a, BOp b, Epi e, cudaStream_t st,
                         const char* what) {
    auto kern = tcgemm_kernel<AOp, BOp, Epi>;
    const size_t smem = tcgemm_smem_bytes(BN);
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)tcgemm_smem_bytes(kTgMaxBN));
        if (err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(err)); return (int)err; }
        configured = true;
    }
    if (splits < 1) splits = 1;
    int k_per_split = ceil_div(ceil_div(Kd, splits), kTgBK) * kTgBK;
    splits = ceil_div(Kd, k_per_split);
    dim3 grid(ceil_div(M, kTgBM), splits);
    kern<<<grid, kTgThreads, smem, st>>>(M, N, Kd, BN, k_per_split, a, b, e);
    return launch_status(what);
}

// column-tile width for an N-wide output: a multiple of 16, <= 256, balanced over ceil(N/256) tiles
inline int tcgemm_pick_bn(int N) {
    const int tiles = ceil_div(N, kTgMaxBN);
    return ceil_div(ceil_div(N, tiles), 16) * 16;
}

}  // namespace pph
