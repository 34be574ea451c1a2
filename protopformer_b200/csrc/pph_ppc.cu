// (a7) PPC (prototypical part concentration) loss, forward and backward.  Replaces PPNet.get_PPC_loss + batch_cov
// (protopformer.py:249-288): ~25 ATen launches, a third topk+sort, a (B*m,196,2) grid repeat and a bmm of
// B*m*196 2x1*1x2 outer products collapse to one CTA per image.
//
// Per image b with label y, for the m prototypes p_j = y*m + j (restatement verified in SURVEY.md 8(d)(iii)):
//   d[j,k]  = relu(z2[b,k] + (p2[p_j] - 2 <Z[b,k], P[p_j]>))      recomputed in fp32 (the reference gathers it)
//   w[j,k]  = act(d[j,k]);  pos_k = (idx[b,k] / side, idx[b,k] % side);  S_j = sum_k w
//   mu_j    = sum_k w pos_k / S_j;   V_j = sum_k w (pos_k - mu_j)^2 / S_j   (per coordinate);  var_j = V_j N/(N-1)
//   L_cov   = mean_{b,j} relu((var_r + var_c)/2 - cov_thresh)
//   L_mean  = mean_{b,i,j} relu(mean_thresh - |mu_i - mu_j|) [i != j]     (diagonal zeros stay in the denominator)
// HBM traffic per image: K*D*4 (token features, L2-resident after the add-on kernel) + m*D*4 + K*4 in,
// m*K*4 + m*32 out -- a latency-bound kernel; loads are coalesced 128B rows.
#include <math.h>

#include "pph_common.cuh"

namespace pph {

constexpr int kPpcThreads = 256;
constexpr int kPpcMaxDV = 16;     // D <= 512

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kPpcThreads / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(kPpcThreads)
ppc_fwd_kernel(const float* __restrict__ Zs, const float* __restrict__ z2s, const float* __restrict__ Pl,
               const float* __restrict__ p2l, const int32_t* __restrict__ idx, const int64_t* __restrict__ labels,
               int B, int K, int D, int P, int m, int N, int side, int act_fn, float eps,
               float cov_thresh, float mean_thresh,
               float* __restrict__ dslice, float* __restrict__ stats, float* partial, unsigned int* counter,
               float* __restrict__ losses) {
    extern __shared__ float sm[];
    float* Prow = sm;                 // [m][D]
    float* dsl = Prow + m * D;        // [m][K]
    float* mu = dsl + m * K;          // [m][2]
    __shared__ float red[kPpcThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kPpcThreads / 32;
    long y = labels[b];
    if (y < 0) y = 0;
    if (y * m + m > P) y = P / m - 1;                 // out-of-range labels are clamped (the reference would raise)
    const int prow0 = (int)y * m;
    for (int i = tid; i < m * D; i += kPpcThreads) Prow[i] = __ldg(Pl + (size_t)prow0 * D + i);
    __syncthreads();

    // distances of the label-class prototypes to every selected token: one warp per token
    for (int k = warp; k < K; k += nwarp) {
        const float* zr = Zs + ((size_t)b * K + k) * D;
        float z[kPpcMaxDV];
#pragma unroll
        for (int i = 0; i < kPpcMaxDV; ++i) z[i] = (i * 32 + lane < D) ? __ldg(zr + i * 32 + lane) : 0.f;
        const float zz = __ldg(z2s + (size_t)b * K + k);
        for (int j = 0; j < m; ++j) {
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < kPpcMaxDV; ++i)
                if (i * 32 < D) dot = fmaf(z[i], (i * 32 + lane < D) ? Prow[j * D + i * 32 + lane] : 0.f, dot);
            dot = warp_sum(dot);
            if (lane == 0) {
                const float d = fmaxf(zz + fmaf(-2.0f, dot, __ldg(p2l + prow0 + j)), 0.0f);
                dsl[j * K + k] = d;
                dslice[((size_t)b * m + j) * K + k] = d;
            }
        }
    }
    __syncthreads();

    // weighted mean / variance of the grid positions: one warp per prototype
    float cov_sum = 0.f;    // meaningful on lane 0 of each warp
    for (int j = warp; j < m; j += nwarp) {
        float S = 0.f, Sr = 0.f, Sc = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float w = act_of_dist(dsl[j * K + k], act_fn, eps);
            const int n = __ldg(idx + (size_t)b * K + k);
            S += w;
            Sr = fmaf(w, (float)(n / side), Sr);
            Sc = fmaf(w, (float)(n % side), Sc);
        }
        S = warp_sum(S); Sr = warp_sum(Sr); Sc = warp_sum(Sc);
        const float mr = Sr / S, mc = Sc / S;
        float Vr = 0.f, Vc = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float w = act_of_dist(dsl[j * K + k], act_fn, eps);
            const int n = __ldg(idx + (size_t)b * K + k);
            const float dr = (float)(n / side) - mr, dc = (float)(n % side) - mc;
            Vr = fmaf(w, dr * dr, Vr);
            Vc = fmaf(w, dc * dc, Vc);
        }
        Vr = warp_sum(Vr) / S;
        Vc = warp_sum(Vc) / S;
        const float scale = (float)N / (float)(N - 1);
        const float pre = (Vr * scale + Vc * scale) * 0.5f - cov_thresh;
        if (lane == 0) {
            mu[2 * j] = mr; mu[2 * j + 1] = mc;
            float* st = stats + ((size_t)b * m + j) * 8;
            st[0] = S; st[1] = mr; st[2] = mc; st[3] = Vr; st[4] = Vc; st[5] = pre; st[6] = 0.f; st[7] = 0.f;
            cov_sum += fmaxf(pre, 0.0f);
        }
    }
    const float cov_img = block_sum_256(lane == 0 ? cov_sum : 0.f, red);   // includes the barrier that publishes mu[]

    float mean_sum = 0.f;
    for (int t = tid; t < m * m; t += kPpcThreads) {
        const int i = t / m, j = t - i * m;
        if (i != j) {
            const float dr = mu[2 * i] - mu[2 * j], dc = mu[2 * i + 1] - mu[2 * j + 1];
            mean_sum += fmaxf(mean_thresh - sqrtf(dr * dr + dc * dc), 0.0f);
        }
    }
    const float mean_img = block_sum_256(mean_sum, red);

    // deterministic final sum: the last CTA to finish adds the per-image partials in image order
    if (tid == 0) {
        partial[2 * b] = cov_img;
        partial[2 * b + 1] = mean_img;
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        if (ticket == (unsigned int)(B - 1)) {
            __threadfence();
            float c = 0.f, s = 0.f;
            for (int i = 0; i < B; ++i) {
                c += __ldcg(partial + 2 * i);
                s += __ldcg(partial + 2 * i + 1);
            }
            losses[0] = c / ((float)B * (float)m);
            losses[1] = s / ((float)B * (float)m * (float)m);
            *counter = 0u;          // ready for the next launch / graph replay
        }
    }
}

__global__ void __launch_bounds__(kPpcThreads)
ppc_bwd_kernel(const float* __restrict__ Zs, const float* __restrict__ Pl, const int32_t* __restrict__ idx,
               const int64_t* __restrict__ labels, const float* __restrict__ dslice, const float* __restrict__ stats,
               const float* __restrict__ g_losses, int B, int K, int D, int P, int m, int N, int side,
               int act_fn, float eps, float mean_thresh, float* __restrict__ dZs, float* __restrict__ dP) {
    extern __shared__ float sm[];
    float* Prow = sm;                 // [m][D]
    float* dd = Prow + m * D;         // [m][K]   d loss / d distance
    float* st = dd + m * K;           // [m][8]   stats, then [6],[7] <- d loss / d mu
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kPpcThreads / 32;
    long y = labels[b];
    if (y < 0) y = 0;
    if (y * m + m > P) y = P / m - 1;
    const int prow0 = (int)y * m;
    for (int i = tid; i < m * D; i += kPpcThreads) Prow[i] = __ldg(Pl + (size_t)prow0 * D + i);
    for (int i = tid; i < m * 8; i += kPpcThreads) st[i] = __ldg(stats + (size_t)b * m * 8 + i);
    __syncthreads();
    const float g_cov = __ldg(g_losses) / ((float)B * (float)m);
    const float g_mean = __ldg(g_losses + 1) / ((float)B * (float)m * (float)m);

    // d loss / d mu_i from the pairwise term (both (i,j) and (j,i) depend on mu_i)
    if (tid < m) {
        const int i = tid;
        float gr = 0.f, gcn = 0.f;
        for (int j = 0; j < m; ++j) {
            if (j == i) continue;
            const float dr = st[i * 8 + 1] - st[j * 8 + 1], dc = st[i * 8 + 2] - st[j * 8 + 2];
            const float dist = sqrtf(dr * dr + dc * dc);
            if (mean_thresh - dist > 0.0f && dist > 0.0f) {
                gr -= 2.0f * g_mean * dr / dist;
                gcn -= 2.0f * g_mean * dc / dist;
            }
        }
        st[i * 8 + 6] = gr;
        st[i * 8 + 7] = gcn;
    }
    __syncthreads();
    const float scale = (float)N / (float)(N - 1);
    for (int t = tid; t < m * K; t += kPpcThreads) {
        const int j = t / K, k = t - j * K;
        const float S = st[j * 8], mr = st[j * 8 + 1], mc = st[j * 8 + 2], Vr = st[j * 8 + 3], Vc = st[j * 8 + 4];
        const float dV = st[j * 8 + 5] > 0.0f ? 0.5f * g_cov * scale : 0.0f;
        const int n = __ldg(idx + (size_t)b * K + k);
        const float dr = (float)(n / side) - mr, dc = (float)(n % side) - mc;
        const float dw = (dV * ((dr * dr - Vr) + (dc * dc - Vc)) + st[j * 8 + 6] * dr + st[j * 8 + 7] * dc) / S;
        const float d = __ldg(dslice + ((size_t)b * m + j) * K + k);
        dd[t] = dw * dact_of_dist(d, act_fn, eps);
    }
    __syncthreads();

    // token gradient rows: one warp per token, dZ[b,k,:] = sum_j dd[j,k] * 2 (Z[b,k,:] - P_j)
    for (int k = warp; k < K; k += nwarp) {
        const float* zr = Zs + ((size_t)b * K + k) * D;
        float z[kPpcMaxDV], acc[kPpcMaxDV];
#pragma unroll
        for (int i = 0; i < kPpcMaxDV; ++i) {
            z[i] = (i * 32 + lane < D) ? __ldg(zr + i * 32 + lane) : 0.f;
            acc[i] = 0.f;
        }
        for (int j = 0; j < m; ++j) {
            const float c2 = 2.0f * dd[j * K + k];
#pragma unroll
            for (int i = 0; i < kPpcMaxDV; ++i)
                if (i * 32 + lane < D) acc[i] = fmaf(c2, z[i] - Prow[j * D + i * 32 + lane], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < kPpcMaxDV; ++i)
            if (i * 32 + lane < D) dZs[((size_t)b * K + k) * D + i * 32 + lane] = acc[i];
    }
    // prototype gradient rows: one warp per label-class prototype, dP_j += sum_k dd[j,k] * 2 (P_j - Z[b,k,:])
    for (int j = warp; j < m; j += nwarp) {
        float acc[kPpcMaxDV];
#pragma unroll
        for (int i = 0; i < kPpcMaxDV; ++i) acc[i] = 0.f;
        for (int k = 0; k < K; ++k) {
            const float c2 = 2.0f * dd[j * K + k];
            const float* zr = Zs + ((size_t)b * K + k) * D;
#pragma unroll
            for (int i = 0; i < kPpcMaxDV; ++i)
                if (i * 32 + lane < D) acc[i] = fmaf(c2, Prow[j * D + i * 32 + lane] - __ldg(zr + i * 32 + lane), acc[i]);
        }
#pragma unroll
        for (int i = 0; i < kPpcMaxDV; ++i)
            if (i * 32 + lane < D) atomicAdd(dP + (size_t)(prow0 + j) * D + i * 32 + lane, acc[i]);
    }
}

static int ppc_check(int B, int K, int D, int P, int m, int N, int* side) {
    PPH_REQUIRE(B >= 0 && K >= 1 && D >= 1 && D <= 32 * kPpcMaxDV && m >= 1 && P >= m && N >= 2, PPH_EINVAL,
                "pph_ppc: bad dims B=%d K=%d D=%d P=%d m=%d N=%d", B, K, D, P, m, N);
    int s = (int)lrint(sqrt((double)N));
    PPH_REQUIRE(s * s == N, PPH_EINVAL, "pph_ppc: N=%d is not a perfect square", N);
    *side = s;
    return 0;
}

}  // namespace pph

extern "C" int pph_ppc_fwd(const float* Zs, const float* z2s, const float* Pl, const float* p2l,
                           const int32_t* idx32, const int64_t* labels,
                           int B, int K, int D, int P, int m, int N, int act_fn, float eps,
                           float cov_thresh, float mean_thresh,
                           float* dslice, float* stats, float* partial, uint32_t* counter, float* losses,
                           pph_stream_t stream) {
    using namespace pph;
    PPH_REQUIRE(Zs && z2s && Pl && p2l && idx32 && labels && dslice && stats && partial && counter && losses,
                PPH_EINVAL, "pph_ppc_fwd: null pointer");
    int side = 0, rc = ppc_check(B, K, D, P, m, N, &side);
    if (rc) return rc;
    if (B == 0) return 0;
    const size_t smem = sizeof(float) * ((size_t)m * D + (size_t)m * K + 2 * (size_t)m);
    PPH_REQUIRE(smem <= 200 * 1024, PPH_EUNSUP, "pph_ppc_fwd: m*D too large for shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(ppc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("pph_ppc_fwd: %s", cudaGetErrorString(e)); return (int)e; }
    }
    ppc_fwd_kernel<<<B, kPpcThreads, smem, as_stream(stream)>>>(Zs, z2s, Pl, p2l, idx32, labels, B, K, D, P, m, N,
                                                               side, act_fn, eps, cov_thresh, mean_thresh, dslice,
                                                               stats, partial, counter, losses);
    return launch_status("pph_ppc_fwd");
}

extern "C" int pph_ppc_bwd(const float* Zs, const float* Pl, const int32_t* idx32, const int64_t* labels,
                           const float* dslice, const float* stats, const float* g_losses,
                           int B, int K, int D, int P, int m, int N, int act_fn, float eps,
                           float cov_thresh, float mean_thresh, float* dZs, float* dP, pph_stream_t stream) {
    using namespace pph;
    (void)cov_thresh;
    PPH_REQUIRE(Zs && Pl && idx32 && labels && dslice && stats && g_losses && dZs && dP, PPH_EINVAL,
                "pph_ppc_bwd: null pointer");
    int side = 0, rc = ppc_check(B, K, D, P, m, N, &side);
    if (rc) return rc;
    if (B == 0) return 0;
    const size_t smem = sizeof(float) * ((size_t)m * D + (size_t)m * K + 8 * (size_t)m);
    PPH_REQUIRE(smem <= 200 * 1024, PPH_EUNSUP, "pph_ppc_bwd: m*D too large for shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(ppc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("pph_ppc_bwd: %s", cudaGetErrorString(e)); return (int)e; }
    }
    ppc_bwd_kernel<<<B, kPpcThreads, smem, as_stream(stream)>>>(Zs, Pl, idx32, labels, dslice, stats, g_losses, B, K,
                                                               D, P, m, N, side, act_fn, eps, mean_thresh, dZs, dP);
    return launch_status("pph_ppc_bwd");
}
