// (a8, part 2) Backward of max-pool + log-similarity + squared-L2 distance in its argmin-routed (sparse) form,
// SURVEY.md 8(d)(iv).  Autograd of protopformer.py:201-247 would mirror ~10 passes over the (B,P,K) map; here the
// upstream gradient is one scalar g[b,p] placed at token argmin[b,p]:
//   dPl[p,:]   = 2 (Pl[p,:] sum_b g[b,p] - sum_b g[b,p] Zs[b,argmin[b,p],:])          proto_grad_kernel (gather of token rows)
//   dZs[b,k,:] = 2 (Zs[b,k,:] sum_{p in bin(b,k)} g[b,p] - sum_{p in bin(b,k)} g[b,p] Pl[p,:])   token_grad_kernel
//   dPg, dZc   : same with one token per image (dense; dZc through the generic GEMM)
// Byte-bound: every (b,p) pair moves one D-float row (L2-resident operands); rows are read as coalesced 128B lines.
#include "pph_common.cuh"
#include "pph_sgemm.cuh"

namespace pph {

constexpr int kBwdMaxDV = 16;   // D <= 512

// ---- prototype gradients: one warp per prototype row (local rows first, then global rows) ----------------------
__global__ void __launch_bounds__(256)
proto_grad_kernel(const float* __restrict__ g_l, const float* __restrict__ g_g, const int32_t* __restrict__ argmin_l,
                  const float* __restrict__ Zs, const float* __restrict__ Zc, const float* __restrict__ Pl,
                  const float* __restrict__ Pgl, int B, int K, int D, int P, int Pg,
                  float* __restrict__ dPl, float* __restrict__ dPg) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= P + Pg) return;
    const bool global = row >= P;
    const int p = global ? row - P : row;
    const int np = global ? Pg : P;
    const float* g = global ? g_g : g_l;
    float acc[kBwdMaxDV];
#pragma unroll
    for (int i = 0; i < kBwdMaxDV; ++i) acc[i] = 0.f;
    float gsum = 0.f;
    for (int b0 = 0; b0 < B; b0 += 32) {
        const int bl = b0 + lane;
        const float gv = bl < B ? __ldg(g + (size_t)bl * np + p) : 0.f;
        const int av = (!global && bl < B) ? __ldg(argmin_l + (size_t)bl * P + p) : 0;
        const int cnt = min(32, B - b0);
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const float gg = __shfl_sync(0xffffffffu, gv, t);
            const int aa = __shfl_sync(0xffffffffu, av, t);
            const float* zr = global ? Zc + (size_t)(b0 + t) * D : Zs + ((size_t)(b0 + t) * K + aa) * D;
            gsum += gg;
#pragma unroll
            for (int i = 0; i < kBwdMaxDV; ++i)
                if (i * 32 + lane < D) acc[i] = fmaf(gg, __ldg(zr + i * 32 + lane), acc[i]);
        }
    }
    const float* pr = (global ? Pgl : Pl) + (size_t)p * D;
    float* out = (global ? dPg : dPl) + (size_t)p * D;
#pragma unroll
    for (int i = 0; i < kBwdMaxDV; ++i)
        if (i * 32 + lane < D) out[i * 32 + lane] = 2.0f * (__ldg(pr + i * 32 + lane) * gsum - acc[i]);
}

// ---- token gradients: CTA = (image, token range); prototypes binned by their argmin token in shared memory ------
constexpr int kTokThreads = 256;

__global__ void __launch_bounds__(kTokThreads)
token_grad_kernel(const float* __restrict__ g_l, const int32_t* __restrict__ argmin_l, const float* __restrict__ Zs,
                  const float* __restrict__ Pl, int K, int D, int P, int tok_per_cta, float* __restrict__ dZs) {
    extern __shared__ int smi[];
    int* start = smi;              // [K+1] bin offsets
    int* cursor = start + K + 1;   // [K]
    int* list = cursor + K;        // [P] prototype ids grouped by token
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k_begin = blockIdx.y * tok_per_cta, k_end = min(K, k_begin + tok_per_cta);
    const int32_t* am = argmin_l + (size_t)b * P;
    for (int k = tid; k <= K; k += kTokThreads) start[k] = 0;
    __syncthreads();
    for (int p = tid; p < P; p += kTokThreads) {
        const int a = __ldg(am + p);
        if (a >= k_begin && a < k_end) atomicAdd(&start[a + 1], 1);
    }
    __syncthreads();
    if (warp == 0) {               // inclusive scan of the counts -> bin offsets
        int carry = 0;
        for (int k0 = 0; k0 <= K; k0 += 32) {
            const int k = k0 + lane;
            int v = k <= K ? start[k] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            if (k <= K) start[k] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += kTokThreads) cursor[k] = start[k];
    __syncthreads();
    for (int p = tid; p < P; p += kTokThreads) {
        const int a = __ldg(am + p);
        if (a >= k_begin && a < k_end) list[atomicAdd(&cursor[a], 1)] = p;
    }
    __syncthreads();
    const float* gb = g_l + (size_t)b * P;
    for (int k = k_begin + warp; k < k_end; k += kTokThreads / 32) {
        float acc[kBwdMaxDV];
#pragma unroll
        for (int i = 0; i < kBwdMaxDV; ++i) acc[i] = 0.f;
        float gsum = 0.f;
        const int e0 = start[k], e1 = start[k + 1];
#pragma unroll 4
        for (int e = e0; e < e1; ++e) {
            const int p = list[e];
            const float gg = __ldg(gb + p);
            const float* pr = Pl + (size_t)p * D;
            gsum += gg;
#pragma unroll
            for (int i = 0; i < kBwdMaxDV; ++i)
                if (i * 32 + lane < D) acc[i] = fmaf(gg, __ldg(pr + i * 32 + lane), acc[i]);
        }
        const float* zr = Zs + ((size_t)b * K + k) * D;
        float* out = dZs + ((size_t)b * K + k) * D;
#pragma unroll
        for (int i = 0; i < kBwdMaxDV; ++i)
            if (i * 32 + lane < D) out[i * 32 + lane] = 2.0f * (__ldg(zr + i * 32 + lane) * gsum - acc[i]);
    }
}

struct ClsGradEpi {   // dZc[b,d] += 2 (Zc[b,d] * rowsum_part - acc_part)     (split over prototypes)
    const float* Zc;
    float* dZc;
    int D;
    __device__ __forceinline__ void operator()(int b, int d, float acc, float rs) const {
        const size_t o = (size_t)b * D + d;
        atomicAdd(dZc + o, 2.0f * (__ldg(Zc + o) * rs - acc));
    }
};

}  // namespace pph

extern "C" int pph_similarity_bwd(const float* g_l, const float* g_g, const int32_t* argmin_l,
                                  const float* Zs, const float* Zc, const float* Pl, const float* Pgl,
                                  int B, int K, int D, int P, int Pg,
                                  float* dZs, float* dZc, float* dPl, float* dPg, pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(g_l && argmin_l && Zs && Pl && dZs && dPl, PPH_EINVAL, "pph_similarity_bwd: null local pointer");
    PPH_REQUIRE(Pg == 0 || (g_g && Zc && Pgl && dZc && dPg), PPH_EINVAL, "pph_similarity_bwd: null global pointer");
    PPH_REQUIRE(B >= 0 && K >= 1 && D >= 1 && D <= 32 * kBwdMaxDV && P >= 1 && Pg >= 0, PPH_EINVAL,
                "pph_similarity_bwd: bad dims B=%d K=%d D=%d P=%d Pg=%d", B, K, D, P, Pg);
    cudaStream_t st = as_stream(stream);
    if (B == 0) {
        cudaMemsetAsync(dPl, 0, sizeof(float) * (size_t)P * D, st);
        if (Pg > 0) cudaMemsetAsync(dPg, 0, sizeof(float) * (size_t)Pg * D, st);
        return launch_status("pph_similarity_bwd(empty)");
    }
    proto_grad_kernel<<<ceil_div(P + Pg, 8), 256, 0, st>>>(g_l, g_g, argmin_l, Zs, Zc, Pl, Pgl, B, K, D, P, Pg, dPl,
                                                           dPg);
    int rc = launch_status("pph_similarity_bwd(proto)");
    if (rc) return rc;
    {
        // enough (image, token range) CTAs for ~2 per SM
        int parts = ceil_div(2 * 148, B);
        if (parts > K) parts = K;
        if (parts < 1) parts = 1;
        const int tok_per_cta = ceil_div(K, parts);
        parts = ceil_div(K, tok_per_cta);
        const size_t smem = sizeof(int) * ((size_t)2 * K + 1 + P);
        PPH_REQUIRE(smem <= 200 * 1024, PPH_EUNSUP, "pph_similarity_bwd: P=%d too large for the bin list", P);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(token_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("pph_similarity_bwd: %s", cudaGetErrorString(e)); return (int)e; }
        }
        token_grad_kernel<<<dim3(B, parts), kTokThreads, smem, st>>>(g_l, argmin_l, Zs, Pl, K, D, P, tok_per_cta, dZs);
        rc = launch_status("pph_similarity_bwd(token)");
        if (rc) return rc;
    }
    if (Pg > 0) {
        cudaError_t e = cudaMemsetAsync(dZc, 0, sizeof(float) * (size_t)B * D, st);
        if (e != cudaSuccess) { set_error("pph_similarity_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
        StridedOp<true> a{g_g, B, Pg, 1};
        StridedOp<false> bop{Pgl, D, 1, D};       // (row = d, k = p) -> Pg[p*D + d]
        ClsGradEpi epi{Zc, dZc, D};
        const int tiles = ceil_div(B, kGemmBM) * ceil_div(D, kGemmBN);
        launch_sgemm<true>(B, D, Pg, ceil_div(148, tiles), a, bop, epi, st);
        rc = launch_status("pph_similarity_bwd(cls)");
        if (rc) return rc;
    }
    return 0;
}
