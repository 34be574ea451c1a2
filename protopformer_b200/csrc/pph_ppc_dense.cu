// PPC loss on a DENSE activation map (protopformer.py:259-288 as written: `get_PPC_loss(total_proto_act, ...)` accepts any
// (B,P,h,w) tensor).  The fused path never materialises that map (pph_ppc_fwd recomputes the label-class slice from the
// token features); this pair of kernels serves callers that hold the map as a tensor -- same arithmetic, the slice is
// gathered instead of recomputed, and the gradient goes back to the map.
//
//   weights of prototype j of image b : w_k = act[b, label_b*m + j, k] on grid cell pos_k = (idx_k / side, idx_k % side)
//   S = sum w,  mu = sum w pos / S,  var_d = N / (S (N-1)) * sum w (pos_d - mu_d)^2              (:250-257, zero weight elsewhere)
//   L_cov = mean_{b,j} relu((var_x + var_y)/2 - t_cov),  L_mean = mean_{b,j,j'} relu(t_mean - |mu_j - mu_j'|) [j != j']
//
// One CTA per image, one warp per label-class prototype (m <= 32); fixed-order sums -> bit-reproducible.
#include "pph_common.cuh"

namespace pph {
namespace {

constexpr int kPdStats = 8;      // S, mu_x, mu_y, V_x, V_y, cov-active flag, unused, unused

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void ppc_dense_fwd_kernel(const float* __restrict__ act, const int* __restrict__ idx32,
                                     const long long* __restrict__ labels, int B, int P, int K, int m, int N, int side,
                                     float cov_thresh, float mean_thresh, float* __restrict__ stats,
                                     float* __restrict__ partial, unsigned* __restrict__ counter, float* __restrict__ losses) {
    extern __shared__ float sm[];                 // m * 2 means, m cov terms, m mean-term row sums
    float* mu = sm;
    float* covt = sm + 2 * m;
    float* meant = covt + m;
    const int b = blockIdx.x, j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long lab = labels[b];
    lab = lab < 0 ? 0 : (lab >= P / m ? P / m - 1 : lab);
    const float* w = act + ((size_t)b * P + lab * m + j) * K;
    const int* ix = idx32 + (size_t)b * K;
    float S = 0.f, sx = 0.f, sy = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float v = w[k];
        const int n = ix[k];
        S += v;
        sx += v * (float)(n / side);
        sy += v * (float)(n % side);
    }
    S = warp_sum(S); sx = warp_sum(sx); sy = warp_sum(sy);
    const float mx = sx / S, my = sy / S;
    float vx = 0.f, vy = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float v = w[k];
        const int n = ix[k];
        const float dx = (float)(n / side) - mx, dy = (float)(n % side) - my;
        vx += v * dx * dx;
        vy += v * dy * dy;
    }
    vx = warp_sum(vx); vy = warp_sum(vy);
    const float c = (float)N / (float)(N - 1);
    const float cov = 0.5f * (c * vx / S + c * vy / S) - cov_thresh;
    if (lane == 0) {
        float* st = stats + ((size_t)b * m + j) * kPdStats;
        st[0] = S; st[1] = mx; st[2] = my; st[3] = vx; st[4] = vy; st[5] = cov > 0.f ? 1.f : 0.f;
        mu[2 * j] = mx; mu[2 * j + 1] = my;
        covt[j] = cov > 0.f ? cov : (cov != cov ? cov : 0.f);          // relu that propagates NaN like torch
    }
    __syncthreads();
    float t = 0.f;
    for (int q = lane; q < m; q += 32)
        if (q != j) {
            const float dx = mx - mu[2 * q], dy = my - mu[2 * q + 1];
            const float r = mean_thresh - sqrtf(dx * dx + dy * dy);
            t += r > 0.f ? r : (r != r ? r : 0.f);
        }
    t = warp_sum(t);
    if (lane == 0) meant[j] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, d = 0.f;
        for (int q = 0; q < m; ++q) { a += covt[q]; d += meant[q]; }
        partial[2 * b] = a;
        partial[2 * b + 1] = d;
        __threadfence();
        if (atomicAdd(counter, 1u) == (unsigned)(B - 1)) {           // last image: ordered final sum
            __threadfence();
            float ca = 0.f, da = 0.f;
            for (int i = 0; i < B; ++i) {
                ca += __ldcg(partial + 2 * i);
                da += __ldcg(partial + 2 * i + 1);
            }
            losses[0] = ca / ((float)B * (float)m);
            losses[1] = da / ((float)B * (float)m * (float)m);
            *counter = 0u;
        }
    }
}

__global__ void ppc_dense_bwd_kernel(const float* __restrict__ act, const int* __restrict__ idx32,
                                     const long long* __restrict__ labels, const float* __restrict__ stats,
                                     const float* __restrict__ g_cov, const float* __restrict__ g_mean, int B, int P, int K, int m,
                                     int N, int side, float mean_thresh, float* __restrict__ dact) {
    const int b = blockIdx.x, j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long lab = labels[b];
    lab = lab < 0 ? 0 : (lab >= P / m ? P / m - 1 : lab);
    // every other row of this image's map has zero gradient
    {
        float* base = dact + (size_t)b * P * K;
        const size_t lo = (size_t)lab * m * K, hi = lo + (size_t)m * K, tot = (size_t)P * K;
        for (size_t i = threadIdx.x; i < tot; i += blockDim.x)
            if (i < lo || i >= hi) base[i] = 0.f;
    }
    const float* st = stats + ((size_t)b * m + j) * kPdStats;
    const float S = st[0], mx = st[1], my = st[2], vx = st[3], vy = st[4], on = st[5];
    const float gc = g_cov[0] / ((float)B * (float)m) * on * 0.5f * ((float)N / (float)(N - 1)) / S;
    // dL_mean / dmu_j: every unordered pair appears twice in the (m x m) mean
    float gx = 0.f, gy = 0.f;
    for (int q = 0; q < m; ++q)
        if (q != j) {
            const float* sq = stats + ((size_t)b * m + q) * kPdStats;
            const float dx = mx - sq[1], dy = my - sq[2];
            const float dist = sqrtf(dx * dx + dy * dy);
            if (mean_thresh - dist > 0.f && dist > 0.f) {
                gx -= dx / dist;
                gy -= dy / dist;
            }
        }
    const float gm = 2.f * g_mean[0] / ((float)B * (float)m * (float)m) / S;
    gx *= gm; gy *= gm;
    const float* w = act + ((size_t)b * P + lab * m + j) * K;
    float* dw = dact + ((size_t)b * P + lab * m + j) * K;
    const int* ix = idx32 + (size_t)b * K;
    (void)w;
    for (int k = lane; k < K; k += 32) {
        const int n = ix[k];
        const float dx = (float)(n / side) - mx, dy = (float)(n % side) - my;
        dw[k] = gc * ((dx * dx - vx / S) + (dy * dy - vy / S)) + gx * dx + gy * dy;
    }
}

}  // namespace
}  // namespace pph

using namespace pph;

static int ppc_dense_check(const char* who, int B, int P, int K, int m, int N) {
    PPH_REQUIRE(B >= 0 && K >= 1 && m >= 1 && m <= 32 && P >= m && P % m == 0 && N >= K && N >= 2, PPH_EINVAL,
                "%s: bad dims B=%d P=%d K=%d m=%d N=%d (1 <= m <= 32, P %% m == 0, K <= N)", who, B, P, K, m, N);
    int side = 1;
    while (side * side < N) ++side;
    PPH_REQUIRE(side * side == N, PPH_EINVAL, "%s: N=%d is not a perfect square", who, N);
    return side;
}

extern "C" int pph_ppc_dense_fwd(const float* act, const int32_t* idx32, const int64_t* labels, int B, int P, int K, int m, int N,
                                 float cov_thresh, float mean_thresh, float* stats, float* partial, unsigned int* counter,
                                 float* losses, pph_stream_t stream) {
    PPH_REQUIRE(act && idx32 && labels && stats && partial && counter && losses, PPH_EINVAL, "pph_ppc_dense_fwd: null pointer");
    const int side = ppc_dense_check("pph_ppc_dense_fwd", B, P, K, m, N);
    if (side < 0) return side;
    if (B == 0) return 0;
    ppc_dense_fwd_kernel<<<B, 32 * m, sizeof(float) * 4 * m, as_stream(stream)>>>(
        act, idx32, reinterpret_cast<const long long*>(labels), B, P, K, m, N, side, cov_thresh, mean_thresh, stats, partial,
        counter, losses);
    return launch_status("pph_ppc_dense_fwd");
}

extern "C" int pph_ppc_dense_bwd(const float* act, const int32_t* idx32, const int64_t* labels, const float* stats,
                                 const float* g_cov, const float* g_mean, int B, int P, int K, int m, int N, float mean_thresh,
                                 float* dact, pph_stream_t stream) {
    PPH_REQUIRE(act && idx32 && labels && stats && g_cov && g_mean && dact, PPH_EINVAL, "pph_ppc_dense_bwd: null pointer");
    const int side = ppc_dense_check("pph_ppc_dense_bwd", B, P, K, m, N);
    if (side < 0) return side;
    if (B == 0) return 0;
    ppc_dense_bwd_kernel<<<B, 32 * m, 0, as_stream(stream)>>>(act, idx32, reinterpret_cast<const long long*>(labels), stats, g_cov,
                                                               g_mean, B, P, K, m, N, side, mean_thresh, dact);
    return launch_status("pph_ppc_dense_bwd");
}
