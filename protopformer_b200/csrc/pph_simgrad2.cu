// (a8, part 2) second implementation of the argmin-routed backward of max-pool + similarity + squared-L2 distance
// (autograd of protopformer.py:201-247; restatement SURVEY.md 8(d)(iv)):
//   dPl[p,:]   = 2 (Pl[p,:] sum_b g[b,p] - sum_b g[b,p] Zs[b,argmin[b,p],:])      dPg[p,:] likewise with Zc[b,:]
//   dZs[b,k,:] = 2 (Zs[b,k,:] sum_{p in bin(b,k)} g[b,p] - sum_{p in bin(b,k)} g[b,p] Pl[p,:])
//   dZc[b,:]   = 2 (Zc[b,:] sum_p g_g[b,p] - sum_p g_g[b,p] Pg[p,:])
// The round-1 kernel (pph_similarity_bwd.cu) gathers one D-float row per (image, prototype) pair from L2, twice:
// 197 MB of L2 traffic for 16 MB of operands at the CUB shape, 33 us.  Here the feature axis is cut into 16-float
// slices and the operand that is gathered FROM is staged in shared memory once per CTA, so the gathers are
// shared-memory reads (64 B rows) and L2 only sees coalesced 64-byte segments:
//   kind A (slice, image group)      Pl[:, slice] resident (P x 64 B); per image the bin-sorted (p, g) list is staged,
//                                    a 4-lane group owns one bin and walks it: dZs rows
//   kind B (slice, prototype tile)   Zs/Zc[:, :, slice] of 16 images at a time (double-buffered cp.async); a 4-lane
//                                    group owns 2 prototypes and walks pairT[p][b] = (g, argmin): dPl and dPg rows
//                                    (global prototypes route to the CLS row K of every image)
//   kind C (slice, image group)      Pg[:, slice] resident; one warp per image, dense: dZc rows
// Every output element has one writer and a fixed summation order -> bit-reproducible, no atomics, no partials.
// The PPC-loss gradients computed by pph_head_mid are added while the rows are written (dZs_ppc elementwise, the
// per-image prototype rows dP_img summed over the images of the prototype's class in image order).
// Shared-memory bandwidth is what bounds it: every pair moves 64 B per slice through LDS (197 MB per launch at the
// CUB shape = 10.4 k cycles at 128 B/clk/SM over 148 SMs).
#include "pph_common.cuh"
#include "pph_step2.cuh"

namespace pph {

constexpr int kG2Threads = 512;
constexpr int kG2Warps = kG2Threads / 32;
constexpr int kG2DS = 16;            // floats per slice row
constexpr int kG2PT = 256;           // prototypes per kind-B CTA (16 warps x 8 lane groups x 2)
constexpr int kG2MaxIC = 16;         // images per kind-B staging chunk (fewer when K is large)

struct Grad2Args {
    int B, Bp, K, D, P, Pg, m;
    int NS;                       // D / 16 slices
    int IA, nIG;                  // kind A: images per CTA, image groups
    int nPT;                      // kind B: prototype tiles over [0, P + Pg)
    int IGC, nIGC;                // kind C: images per CTA, image groups
    int IC;                       // kind B: images per staging chunk (even)
    int nA, nB, nC;
    const float *g_l, *g_g;
    const float2* pairT;
    const int32_t *bin_start, *bin_list;
    const float *Zs, *Zc, *Pl, *Pgl;
    const float *add_dZs, *dP_img;
    const int64_t* labels;
    float *dZs, *dZc, *dPl, *dPg;
};

__device__ __forceinline__ float4 fma4(float s, float4 v, float4 a) {
    a.x = fmaf(s, v.x, a.x); a.y = fmaf(s, v.y, a.y); a.z = fmaf(s, v.z, a.z); a.w = fmaf(s, v.w, a.w);
    return a;
}

// stage rows [0, R) of a [R][D] matrix, columns [c0, c0 + 16), into dst[R][16] with 16-byte cp.async copies
__device__ __forceinline__ void stage_slice(float* dst, const float* __restrict__ src, int R, int D, int c0) {
    for (int i = threadIdx.x; i < R * 4; i += kG2Threads) {
        const int row = i >> 2, q = i & 3;
        cp_async16(dst + (size_t)row * kG2DS + q * 4, src + (size_t)row * D + c0 + q * 4);
    }
}

// ---- kind A: token-gradient rows -----------------------------------------------------------------------------------
__device__ __forceinline__ void grad2_tokens(const Grad2Args& a, int vb, float* sm) {
    const int sa = vb % a.NS, ig = vb / a.NS;
    const int K = a.K, P = a.P, D = a.D;
    float* Psl = sm;                                              // [P][16]
    float2* sp = reinterpret_cast<float2*>(Psl + (size_t)P * kG2DS);   // [2][P]  (p bits, g) in bin order
    int* bst = reinterpret_cast<int*>(sp + 2 * (size_t)P);       // [2][K+1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = lane >> 2, dq = lane & 3;
    const int c0 = sa * kG2DS;
    const int bA = ig * a.IA, bE = min(a.B, bA + a.IA);
    stage_slice(Psl, a.Pl, P, D, c0);
    cp_async_commit();
    auto stage_image = [&](int b, int buf) {
        float2* spb = sp + (size_t)buf * P;
        const int32_t* list = a.bin_list + (size_t)b * P;
        const float* gb = a.g_l + (size_t)b * P;
        for (int e = tid; e < P; e += kG2Threads) {
            const int p = __ldg(list + e);
            spb[e] = make_float2(__int_as_float(p), __ldg(gb + p));
        }
        for (int k = tid; k <= K; k += kG2Threads) bst[buf * (K + 1) + k] = __ldg(a.bin_start + (size_t)b * (K + 1) + k);
    };
    if (bA < bE) stage_image(bA, 0);
    cp_async_wait<0>();
    __syncthreads();
    for (int b = bA; b < bE; ++b) {
        const int buf = (b - bA) & 1;
        if (b + 1 < bE) stage_image(b + 1, buf ^ 1);              // next image's list lands while this one is walked
        const float2* spb = sp + (size_t)buf * P;
        const int* bs = bst + buf * (K + 1);
        for (int k0 = 0; k0 < K; k0 += kG2Warps * 8) {
            const int k = k0 + warp * 8 + grp;
            int e0 = 0, e1 = 0;
            if (k < K) { e0 = bs[k]; e1 = bs[k + 1]; }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float gs = 0.f;
            float2 pg = e0 < e1 ? spb[e0] : make_float2(0.f, 0.f);
            for (int e = e0; e < e1; ++e) {
                const float2 nxt = e + 1 < e1 ? spb[e + 1] : pg;
                const float4 v = *reinterpret_cast<const float4*>(Psl + (size_t)__float_as_int(pg.x) * kG2DS + dq * 4);
                acc = fma4(pg.y, v, acc);
                gs += pg.y;
                pg = nxt;
            }
            if (k < K) {
                const size_t o = ((size_t)b * K + k) * D + c0 + dq * 4;
                const float4 z = __ldg(reinterpret_cast<const float4*>(a.Zs + o));
                float4 r;
                r.x = 2.0f * (z.x * gs - acc.x); r.y = 2.0f * (z.y * gs - acc.y);
                r.z = 2.0f * (z.z * gs - acc.z); r.w = 2.0f * (z.w * gs - acc.w);
                if (a.add_dZs) {
                    const float4 ad = __ldg(reinterpret_cast<const float4*>(a.add_dZs + o));
                    r.x += ad.x; r.y += ad.y; r.z += ad.z; r.w += ad.w;
                }
                *reinterpret_cast<float4*>(a.dZs + o) = r;
            }
        }
        __syncthreads();
    }
}

// ---- kind B: prototype-gradient rows (local and global) -------------------------------------------------------------
__device__ __forceinline__ void grad2_protos(const Grad2Args& a, int vb, float* sm) {
    const int sb = vb % a.NS, pt = vb / a.NS;
    const int K = a.K, D = a.D, B = a.B, P = a.P, R = K + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = lane >> 2, dq = lane & 3;
    const int c0 = sb * kG2DS;
    const int IC = a.IC;
    const size_t buf_floats = (size_t)IC * R * kG2DS;
    float* Zsl = sm;                                             // [2][IC][K+1][16]
    int pj[2];
    bool ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        pj[j] = pt * kG2PT + j * (kG2PT / 2) + warp * 8 + grp;
        ok[j] = pj[j] < P + a.Pg;
    }
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    float gs[2] = {0.f, 0.f};
    auto stage_chunk = [&](int ch, int buf) {
        const int b0 = ch * IC, n = min(IC, B - b0);
        float* dst = Zsl + buf * buf_floats;
        for (int i = tid; i < n * R * 4; i += kG2Threads) {
            const int row = i >> 2, q = i & 3;
            const int bi = row / R, kk = row - bi * R;
            const float* src = kk < K ? a.Zs + ((size_t)(b0 + bi) * K + kk) * D : a.Zc + (size_t)(b0 + bi) * D;
            cp_async16(dst + (size_t)row * kG2DS + q * 4, src + c0 + q * 4);
        }
        cp_async_commit();
    };
    const int nch = (B + IC - 1) / IC;
    stage_chunk(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nch) {
            stage_chunk(ch + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int b0 = ch * IC, n = min(IC, B - b0);
        const float* zb = Zsl + buf * buf_floats + dq * 4;
#pragma unroll 2
        for (int bi = 0; bi < n; bi += 2) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (!ok[j]) continue;
                const float4 pr = __ldcg(reinterpret_cast<const float4*>(a.pairT + (size_t)pj[j] * a.Bp + b0 + bi));
                const float4 v0 = *reinterpret_cast<const float4*>(zb + ((size_t)bi * R + __float_as_int(pr.y)) * kG2DS);
                acc[j] = fma4(pr.x, v0, acc[j]);
                gs[j] += pr.x;
                if (bi + 1 < n) {
                    const float4 v1 = *reinterpret_cast<const float4*>(zb + ((size_t)(bi + 1) * R + __float_as_int(pr.w)) * kG2DS);
                    acc[j] = fma4(pr.z, v1, acc[j]);
                    gs[j] += pr.z;
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (!ok[j]) continue;
        const int p = pj[j];
        const bool glob = p >= P;
        const size_t o = (size_t)(glob ? p - P : p) * D + c0 + dq * 4;
        const float4 base = __ldg(reinterpret_cast<const float4*>((glob ? a.Pgl : a.Pl) + o));
        float4 r;
        r.x = 2.0f * (base.x * gs[j] - acc[j].x); r.y = 2.0f * (base.y * gs[j] - acc[j].y);
        r.z = 2.0f * (base.z * gs[j] - acc[j].z); r.w = 2.0f * (base.w * gs[j] - acc[j].w);
        if (!glob && a.dP_img) {       // PPC rows of the images labelled with this prototype's class, in image order
            const int cls = p / a.m, jj = p - cls * a.m;
            for (int b = 0; b < B; ++b) {
                long y = __ldg(a.labels + b);
                if (y < 0) y = 0;
                if (y * a.m + a.m > P) y = P / a.m - 1;
                if ((int)y == cls) {
                    const float4 ad = __ldcg(reinterpret_cast<const float4*>(a.dP_img + ((size_t)b * a.m + jj) * D + c0 + dq * 4));
                    r.x += ad.x; r.y += ad.y; r.z += ad.z; r.w += ad.w;
                }
            }
        }
        *reinterpret_cast<float4*>((glob ? a.dPg : a.dPl) + o) = r;
    }
}

// ---- kind C: CLS-token gradient rows (dense) -----------------------------------------------------------------------
__device__ __forceinline__ void grad2_cls(const Grad2Args& a, int vb, float* sm) {
    const int sc = vb % a.NS, ig = vb / a.NS;
    const int D = a.D, Pg = a.Pg;
    float* Gsl = sm;                                              // [Pg][16]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = lane >> 2, dq = lane & 3;
    const int c0 = sc * kG2DS;
    stage_slice(Gsl, a.Pgl, Pg, D, c0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int bA = ig * a.IGC, bE = min(a.B, bA + a.IGC);
    for (int b = bA + warp; b < bE; b += kG2Warps) {
        const float* gb = a.g_g + (size_t)b * Pg;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float gs = 0.f;
#pragma unroll 4
        for (int p = grp; p < Pg; p += 8) {
            const float g = __ldcg(gb + p);
            const float4 v = *reinterpret_cast<const float4*>(Gsl + (size_t)p * kG2DS + dq * 4);
            acc = fma4(g, v, acc);
            gs += g;
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {          // fixed-order tree over the 8 lane groups
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
            gs += __shfl_xor_sync(0xffffffffu, gs, o);
        }
        if (grp == 0) {
            const size_t o = (size_t)b * D + c0 + dq * 4;
            const float4 z = __ldg(reinterpret_cast<const float4*>(a.Zc + o));
            float4 r;
            r.x = 2.0f * (z.x * gs - acc.x); r.y = 2.0f * (z.y * gs - acc.y);
            r.z = 2.0f * (z.z * gs - acc.z); r.w = 2.0f * (z.w * gs - acc.w);
            *reinterpret_cast<float4*>(a.dZc + o) = r;
        }
    }
}

__global__ void __launch_bounds__(kG2Threads, 1)
sim_grads2_kernel(const Grad2Args a) {
    pdl_sync();
    extern __shared__ __align__(16) float sm_g2[];
    int vb = blockIdx.x;
    if (vb < a.nA) { grad2_tokens(a, vb, sm_g2); return; }
    vb -= a.nA;
    if (vb < a.nB) { grad2_protos(a, vb, sm_g2); return; }
    vb -= a.nB;
    grad2_cls(a, vb, sm_g2);
}

static int grad2_ic(int K) {
    int ic = (int)((size_t)190 * 1024 / (2 * (size_t)(K + 1) * kG2DS * sizeof(float))) & ~1;
    return ic > kG2MaxIC ? kG2MaxIC : ic;
}

static size_t grad2_smem(int K, int P, int Pg) {
    const size_t sa = sizeof(float) * (size_t)P * kG2DS + sizeof(float2) * 2 * (size_t)P + sizeof(int) * 2 * (size_t)(K + 1);
    const size_t sb = sizeof(float) * 2 * (size_t)grad2_ic(K) * (K + 1) * kG2DS;
    const size_t sc = sizeof(float) * (size_t)Pg * kG2DS;
    size_t s = sa > sb ? sa : sb;
    return s > sc ? s : sc;
}

}  // namespace pph

extern "C" int pph_similarity_bwd2_supported(int B, int K, int D, int P, int Pg) {
    using namespace pph;
    if (B < 1 || K < 1 || P < 1 || Pg < 0 || D < kG2DS || D % kG2DS != 0) return 0;
    return (grad2_ic(K) >= 2 && grad2_smem(K, P, Pg) <= 220 * 1024) ? 1 : 0;
}

extern "C" int pph_similarity_bwd2_ws_bytes(int B, int K, int P, long long* bytes) {
    using namespace pph;
    PPH_REQUIRE(bytes && B >= 1 && K >= 1 && P >= 1, PPH_EINVAL, "pph_similarity_bwd2_ws_bytes: bad args");
    *bytes = (long long)carve_bins(nullptr, B, K, P).bytes;
    return 0;
}

extern "C" int pph_similarity_bwd2(const float* g_l, const float* g_g, const float* pairT, const void* bwd_workspace,
                                   const float* Zs, const float* Zc, const float* Pl, const float* Pgl,
                                   int B, int K, int D, int P, int Pg, int m,
                                   const float* add_dZs, const float* dP_img, const int64_t* labels,
                                   float* dZs, float* dZc, float* dPl, float* dPg, pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(g_l && pairT && bwd_workspace && Zs && Pl && dZs && dPl, PPH_EINVAL, "pph_similarity_bwd2: null pointer");
    PPH_REQUIRE(Pg == 0 || (g_g && Zc && Pgl && dZc && dPg), PPH_EINVAL, "pph_similarity_bwd2: null global pointer");
    PPH_REQUIRE(!dP_img || (labels && m >= 1), PPH_EINVAL, "pph_similarity_bwd2: dP_img needs labels and m");
    PPH_REQUIRE(pph_similarity_bwd2_supported(B, K, D, P, Pg), PPH_EUNSUP,
                "pph_similarity_bwd2: shape B=%d K=%d D=%d P=%d Pg=%d (D %% 16, shared memory)", B, K, D, P, Pg);
    const Step2Bins bins = carve_bins(const_cast<void*>(bwd_workspace), B, K, P);
    Grad2Args a;
    a.B = B; a.Bp = ceil_div(B, 64) * 64; a.K = K; a.D = D; a.P = P; a.Pg = Pg; a.m = m > 0 ? m : 1;
    a.NS = D / kG2DS;
    int sms = pph_sm_count();
    if (sms <= 0) sms = 148;
    // kind A: about one CTA per SM-slot third; at least 4 images per CTA so the resident P slice is amortised
    a.IA = ceil_div(B * a.NS, sms);
    if (a.IA < 4) a.IA = 4;
    if (a.IA > B) a.IA = B;
    a.nIG = ceil_div(B, a.IA);
    a.nPT = ceil_div(P + Pg, kG2PT);
    a.IC = grad2_ic(K);
    a.IGC = 32;
    a.nIGC = Pg > 0 ? ceil_div(B, a.IGC) : 0;
    a.nA = a.NS * a.nIG;
    a.nB = a.NS * a.nPT;
    a.nC = a.NS * a.nIGC;
    a.g_l = g_l; a.g_g = g_g; a.pairT = reinterpret_cast<const float2*>(pairT);
    a.bin_start = bins.bin_start; a.bin_list = bins.bin_list;
    a.Zs = Zs; a.Zc = Zc; a.Pl = Pl; a.Pgl = Pgl; a.add_dZs = add_dZs; a.dP_img = dP_img; a.labels = labels;
    a.dZs = dZs; a.dZc = dZc; a.dPl = dPl; a.dPg = dPg;
    const size_t smem = grad2_smem(K, P, Pg);
    cudaError_t e = cudaFuncSetAttribute(sim_grads2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("pph_similarity_bwd2: %s", cudaGetErrorString(e)); return (int)e; }
    launch_k(sim_grads2_kernel, dim3(a.nA + a.nB + a.nC), dim3(kG2Threads), smem, as_stream(stream), a);
    return launch_status("pph_similarity_bwd2");
}
