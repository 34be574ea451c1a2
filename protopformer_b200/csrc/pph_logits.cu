// (a6) prototype-to-class last layers + global/local combine (protopformer.py:297-300, 314-316) and the first
// step of their backward, collapsed with the similarity derivative into one scalar per (image, prototype).
// The last layers are frozen in the reference (protopformer.py:130-131) -> no weight gradient is produced.
//
// Both are skinny FP32 contractions (B x C x (P+Pg) with B = 64): too small for a tensor-core tile grid, so they are
// laid out for L2 traffic and determinism instead:
//   forward : CTA = 8 images x 8 classes, the 256 threads split the prototype axis (coalesced row reads), 64
//             register accumulators per thread, one block reduction -> every logit has ONE writer (deterministic).
//   backward: CTA = 64 images x 32 prototypes, contraction over the C classes from shared memory, 2x4 register tile.
#include "pph_common.cuh"

namespace pph {

constexpr int kLogThreads = 256;
constexpr int kLogTB = 8, kLogTC = 8;

__global__ void __launch_bounds__(kLogThreads)
logits_fwd_kernel(const float* __restrict__ act_l, const float* __restrict__ act_g, const float* __restrict__ Wl,
                  const float* __restrict__ Wg, int B, int P, int Pg, int C, float gc,
                  float* __restrict__ logits, float* __restrict__ logits_g, float* __restrict__ logits_l) {
    __shared__ float red[kLogThreads / 32][kLogTB * kLogTC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.y * kLogTB, c0 = blockIdx.x * kLogTC;
    float res[2] = {0.f, 0.f};      // this thread's (b,c) output: local, global   (thread tid < 64 owns output tid)
    for (int branch = 0; branch < 2; ++branch) {
        const float* A = branch ? act_g : act_l;
        const float* W = branch ? Wg : Wl;
        const int Kd = branch ? Pg : P;
        float acc[kLogTB][kLogTC];
#pragma unroll
        for (int i = 0; i < kLogTB; ++i)
#pragma unroll
            for (int j = 0; j < kLogTC; ++j) acc[i][j] = 0.f;
        for (int p = tid; p < Kd; p += kLogThreads) {
            float a[kLogTB], w[kLogTC];
#pragma unroll
            for (int i = 0; i < kLogTB; ++i) a[i] = (b0 + i < B) ? __ldg(A + (size_t)(b0 + i) * Kd + p) : 0.f;
#pragma unroll
            for (int j = 0; j < kLogTC; ++j) w[j] = (c0 + j < C) ? __ldg(W + (size_t)(c0 + j) * Kd + p) : 0.f;
#pragma unroll
            for (int i = 0; i < kLogTB; ++i)
#pragma unroll
                for (int j = 0; j < kLogTC; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        // block reduction of the 64 accumulators: warp shuffles, then a fixed-order sum over the 8 warps
#pragma unroll
        for (int i = 0; i < kLogTB; ++i)
#pragma unroll
            for (int j = 0; j < kLogTC; ++j) {
                const float v = warp_sum(acc[i][j]);
                if (lane == 0) red[warp][i * kLogTC + j] = v;
            }
        __syncthreads();
        if (tid < kLogTB * kLogTC) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kLogThreads / 32; ++w) s += red[w][tid];
            res[branch] = s;
        }
        __syncthreads();
    }
    if (tid < kLogTB * kLogTC) {
        const int b = b0 + tid / kLogTC, c = c0 + tid % kLogTC;
        if (b < B && c < C) {
            const size_t o = (size_t)b * C + c;
            logits_l[o] = res[0];
            logits_g[o] = res[1];
            logits[o] = gc * res[1] + (1.0f - gc) * res[0];
        }
    }
}

// g[b,p] = (coef * sum_c dlogits[b,c] W[c,p] + sum_c extra[b,c] W[c,p]) * act'(dmin[b,p]) * [dmin > 0]
constexpr int kLbTB = 64, kLbTP = 32, kLbCC = 64;   // images, prototypes per CTA; class chunk staged in smem

__global__ void __launch_bounds__(kLogThreads)
logits_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ dlogits_g,
                  const float* __restrict__ dlogits_l, const float* __restrict__ Wl, const float* __restrict__ Wg,
                  const float* __restrict__ dmin_l, const float* __restrict__ dmin_g,
                  int B, int P, int Pg, int C, int tiles_l, float gc, int act_fn, float eps,
                  float* __restrict__ g_l, float* __restrict__ g_g) {
    __shared__ __align__(16) float up[kLbCC][kLbTB + 2];     // upstream gradient chunk, transposed [c][b]
    __shared__ __align__(16) float wt[kLbCC][kLbTP];         // last-layer chunk [c][p]
    const int tid = threadIdx.x;
    const bool global = (int)blockIdx.x >= tiles_l;
    const int p0 = (global ? blockIdx.x - tiles_l : blockIdx.x) * kLbTP;
    const int b0 = blockIdx.y * kLbTB;
    const int np = global ? Pg : P;
    const float* W = global ? Wg : Wl;
    const float* extra = global ? dlogits_g : dlogits_l;
    const float coef = global ? gc : 1.0f - gc;
    const int tb = tid & 31, tp = tid >> 5;      // images 2*tb, 2*tb+1; prototypes 4*tp .. 4*tp+3
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    // the next class chunk is fetched into registers while the current one is contracted from shared memory
    constexpr int NU = kLbCC * kLbTB / kLogThreads, NW = kLbCC * kLbTP / kLogThreads;     // 16, 8
    float ru[NU], rw[NW];
    auto fetch = [&](int cc) {
#pragma unroll
        for (int q = 0; q < NU; ++q) {
            const int i = tid + q * kLogThreads;
            const int b = i / kLbCC, c = i - b * kLbCC;       // consecutive threads: consecutive classes (coalesced)
            float v = 0.f;
            if (b0 + b < B && cc + c < C) {
                const size_t o = (size_t)(b0 + b) * C + cc + c;
                v = coef * __ldg(dlogits + o);
                if (extra) v += __ldg(extra + o);
            }
            ru[q] = v;
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            const int i = tid + q * kLogThreads;
            const int c = i / kLbTP, p = i - c * kLbTP;
            rw[q] = (cc + c < C && p0 + p < np) ? __ldg(W + (size_t)(cc + c) * np + p0 + p) : 0.f;
        }
    };
    fetch(0);
    for (int cc = 0; cc < C; cc += kLbCC) {
#pragma unroll
        for (int q = 0; q < NU; ++q) {
            const int i = tid + q * kLogThreads;
            const int b = i / kLbCC, c = i - b * kLbCC;
            up[c][b] = ru[q];
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            const int i = tid + q * kLogThreads;
            const int c = i / kLbTP, p = i - c * kLbTP;
            wt[c][p] = rw[q];
        }
        __syncthreads();
        if (cc + kLbCC < C) fetch(cc + kLbCC);
#pragma unroll 8
        for (int c = 0; c < kLbCC; ++c) {
            const float2 u = *reinterpret_cast<const float2*>(&up[c][2 * tb]);
            const float4 w = *reinterpret_cast<const float4*>(&wt[c][4 * tp]);
            acc[0][0] = fmaf(u.x, w.x, acc[0][0]); acc[0][1] = fmaf(u.x, w.y, acc[0][1]);
            acc[0][2] = fmaf(u.x, w.z, acc[0][2]); acc[0][3] = fmaf(u.x, w.w, acc[0][3]);
            acc[1][0] = fmaf(u.y, w.x, acc[1][0]); acc[1][1] = fmaf(u.y, w.y, acc[1][1]);
            acc[1][2] = fmaf(u.y, w.z, acc[1][2]); acc[1][3] = fmaf(u.y, w.w, acc[1][3]);
        }
        __syncthreads();
    }
    const float* dmin = global ? dmin_g : dmin_l;
    float* g = global ? g_g : g_l;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int b = b0 + 2 * tb + i;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = p0 + 4 * tp + j;
            if (p < np) {
                const size_t o = (size_t)b * np + p;
                g[o] = acc[i][j] * dact_of_dist(__ldg(dmin + o), act_fn, eps);
            }
        }
    }
}

}  // namespace pph

extern "C" int pph_logits_fwd(const float* act_l, const float* act_g, const float* Wl, const float* Wg,
                              int B, int P, int Pg, int C, float global_coe,
                              float* logits, float* logits_g, float* logits_l, pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(act_l && Wl && logits && logits_g && logits_l && (Pg == 0 || (act_g && Wg)), PPH_EINVAL,
                "pph_logits_fwd: null pointer");
    PPH_REQUIRE(B >= 0 && P >= 1 && Pg >= 0 && C >= 1, PPH_EINVAL, "pph_logits_fwd: bad dims");
    if (B == 0) return 0;
    dim3 grid(ceil_div(C, kLogTC), ceil_div(B, kLogTB));
    logits_fwd_kernel<<<grid, kLogThreads, 0, as_stream(stream)>>>(act_l, act_g, Wl, Wg, B, P, Pg, C, global_coe,
                                                                  logits, logits_g, logits_l);
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
    const int tiles_l = ceil_div(P, kLbTP), tiles_g = Pg > 0 ? ceil_div(Pg, kLbTP) : 0;
    dim3 grid(tiles_l + tiles_g, ceil_div(B, kLbTB));
    logits_bwd_kernel<<<grid, kLogThreads, 0, as_stream(stream)>>>(dlogits, dlogits_g, dlogits_l, Wl, Wg, dmin_l,
                                                                  dmin_g, B, P, Pg, C, tiles_l, global_coe, act_fn,
                                                                  eps, g_l, g_g);
    return launch_status("pph_logits_bwd");
}
