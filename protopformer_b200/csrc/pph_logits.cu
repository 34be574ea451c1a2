// (a6) prototype-to-class last layers + global/local combine (protopformer.py:297-300, 314-316) and the first
// step of their backward, collapsed with the similarity derivative into one scalar per (image, prototype).
// The last layers are frozen in the reference (protopformer.py:130-131) -> no weight gradient is produced.
// FP32 FMA on the CUDA cores through the generic tile GEMM; local and global branches share one launch.
#include "pph_common.cuh"
#include "pph_sgemm.cuh"

namespace pph {

// ---- forward ---------------------------------------------------------------------------------------------------
// contraction index k runs over [0,Plpad) = local prototypes (zero padded to a split boundary), then the global ones
struct ActConcatOp {
    static constexpr bool kContigK = true;
    const float *l, *g;
    int rows, Kl, Klpad, Kg;
    __device__ __forceinline__ float operator()(int row, int k) const {
        if (row >= rows) return 0.f;
        if (k < Klpad) return k < Kl ? __ldg(l + (size_t)row * Kl + k) : 0.f;
        const int kk = k - Klpad;
        return kk < Kg ? __ldg(g + (size_t)row * Kg + kk) : 0.f;
    }
};

struct LogitsEpi {
    float *logits, *logits_g, *logits_l;
    int C, Klpad, k_per_split;
    float gc;
    __device__ __forceinline__ void operator()(int b, int c, float acc, float) const {
        const bool global = (int)blockIdx.z * k_per_split >= Klpad;
        const size_t o = (size_t)b * C + c;
        if (global) {
            atomicAdd(logits_g + o, acc);
            atomicAdd(logits + o, gc * acc);
        } else {
            atomicAdd(logits_l + o, acc);
            atomicAdd(logits + o, (1.0f - gc) * acc);
        }
    }
};

__global__ void zero3_kernel(float* a, float* b, float* c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = 0.f; b[i] = 0.f; c[i] = 0.f; }
}

// ---- backward --------------------------------------------------------------------------------------------------
// output column n runs over [0,Plpad) = local prototypes (padded to the tile width), then the global ones
struct UpstreamOp {     // (row = b, k = c): upstream gradient of the branch this CTA's columns belong to
    static constexpr bool kContigK = true;
    const float *dlogits, *dlogits_g, *dlogits_l;
    int rows, C, Plpad;
    float gc;
    __device__ __forceinline__ float operator()(int b, int c) const {
        if (b >= rows || c >= C) return 0.f;
        const bool global = (int)blockIdx.x * kGemmBN >= Plpad;
        const size_t o = (size_t)b * C + c;
        float v = (global ? gc : 1.0f - gc) * __ldg(dlogits + o);
        const float* extra = global ? dlogits_g : dlogits_l;
        if (extra) v += __ldg(extra + o);
        return v;
    }
};

struct LastLayerTOp {   // (row = n, k = c) -> W[c, p]
    static constexpr bool kContigK = false;
    const float *Wl, *Wg;
    int P, Plpad, Pg, C;
    __device__ __forceinline__ float operator()(int n, int c) const {
        if (c >= C) return 0.f;
        if (n < Plpad) return n < P ? __ldg(Wl + (size_t)c * P + n) : 0.f;
        const int p = n - Plpad;
        return p < Pg ? __ldg(Wg + (size_t)c * Pg + p) : 0.f;
    }
};

struct RouteEpi {
    const float *dmin_l, *dmin_g;
    float *g_l, *g_g;
    int P, Plpad, Pg, act_fn;
    float eps;
    __device__ __forceinline__ void operator()(int b, int n, float acc, float) const {
        if (n < Plpad) {
            if (n < P) {
                const size_t o = (size_t)b * P + n;
                g_l[o] = acc * dact_of_dist(__ldg(dmin_l + o), act_fn, eps);
            }
        } else {
            const int p = n - Plpad;
            if (p < Pg) {
                const size_t o = (size_t)b * Pg + p;
                g_g[o] = acc * dact_of_dist(__ldg(dmin_g + o), act_fn, eps);
            }
        }
    }
};

}  // namespace pph

extern "C" int pph_logits_fwd(const float* act_l, const float* act_g, const float* Wl, const float* Wg,
                              int B, int P, int Pg, int C, float global_coe,
                              float* logits, float* logits_g, float* logits_l, pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(act_l && Wl && logits && logits_g && logits_l && (Pg == 0 || (act_g && Wg)), PPH_EINVAL,
                "pph_logits_fwd: null pointer");
    PPH_REQUIRE(B >= 0 && P >= 1 && Pg >= 0 && C >= 1, PPH_EINVAL, "pph_logits_fwd: bad dims");
    if (B == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const int n = B * C;
    zero3_kernel<<<ceil_div(n, 256), 256, 0, st>>>(logits, logits_g, logits_l, n);
    // split the contraction so ~2 waves of CTAs exist; local part padded to a split boundary
    const int tiles = ceil_div(B, kGemmBM) * ceil_div(C, kGemmBN);
    int want = ceil_div(2 * 148, tiles);
    int k_per_split = ceil_div(ceil_div(P + Pg, want), kGemmBK) * kGemmBK;
    if (k_per_split < 4 * kGemmBK) k_per_split = 4 * kGemmBK;
    const int Klpad = ceil_div(P, k_per_split) * k_per_split;
    const int Ktot = Klpad + Pg;
    ActConcatOp a{act_l, act_g, B, P, Klpad, Pg};
    ActConcatOp w{Wl, Wg, C, P, Klpad, Pg};
    LogitsEpi epi{logits, logits_g, logits_l, C, Klpad, k_per_split, global_coe};
    dim3 grid(ceil_div(C, kGemmBN), ceil_div(B, kGemmBM), ceil_div(Ktot, k_per_split));
    sgemm_kernel<false><<<grid, kGemmThreads, 0, st>>>(B, C, Ktot, k_per_split, a, w, epi);
    return launch_status("pph_logits_fwd");
}

extern "C" int pph_logits_bwd(const float* dlogits, const float* dlogits_g, const float* dlogits_l,
                              const float* Wl, const float* Wg, const float* dmin_l, const float* dmin_g,
                              int B, int P, int Pg, int C, float global_coe, int act_fn, float eps,
                              float* g_l, float* g_g, pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(dlogits && Wl && dmin_l && g_l && (Pg == 0 || (Wg && dmin_g && g_g)), PPH_EINVAL,
                "pph_logits_bwd: null pointer");
    PPH_REQUIRE(B >= 0 && P >= 1 && Pg >= 0 && C >= 1, PPH_EINVAL, "pph_logits_bwd: bad dims");
    if (B == 0) return 0;
    const int Plpad = ceil_div(P, kGemmBN) * kGemmBN;
    UpstreamOp a{dlogits, dlogits_g, dlogits_l, B, C, Plpad, global_coe};
    LastLayerTOp w{Wl, Wg, P, Plpad, Pg, C};
    RouteEpi epi{dmin_l, dmin_g, g_l, g_g, P, Plpad, Pg, act_fn, eps};
    launch_sgemm<false>(B, Plpad + Pg, C, 1, a, w, epi, as_stream(stream));
    return launch_status("pph_logits_bwd");
}
