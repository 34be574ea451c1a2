// Token bins of the argmin-routed backward (shared by pph_similarity_bwd.cu and the fused mid-step kernel).
#pragma once

#include "pph_common.cuh"

namespace pph {

// ---- binning ---------------------------------------------------------------------------------------------------
// Stable counting sort of an image's prototypes by argmin token.  8 warps own 8 contiguous prototype ranges:
// pass A builds per-warp histograms (match.any groups equal tokens inside a 32-prototype step), an exclusive scan
// over (token, warp) turns them into per-warp write cursors, pass B replays the same steps and places every
// prototype -> inside a bin prototypes are in ascending order, independent of scheduling (deterministic sums).
// Also emits, per token, the first work item of the bin when bins are cut into chunks of kBinChunk entries, and (when
// item_desc is given) one 16-byte descriptor per work item -- (token, first entry, end entry, chunks of the bin | chunk << 16),
// token = -1 for the unused tail of the image's item range -- so that the gradient kernel starts a work item with ONE load
// instead of a search over item_start plus four bound loads (a chain of dependent L2 latencies in front of every gather).
constexpr int kBinChunk = 32;
__host__ __device__ inline int bin_items_per_image(int K, int P) { return K + (P + kBinChunk - 1) / kBinChunk; }

// One CTA of 256 threads per image b; smi: 8*K + 2*(K+1) ints of shared memory.
template <bool kCoherentKeys = false>
__device__ __forceinline__ void
bin_tokens_body(int b, const int32_t* __restrict__ argmin_l, int K, int P, int32_t* __restrict__ bin_start,
                int32_t* __restrict__ item_start, int32_t* __restrict__ bin_list, int* smi, int4* __restrict__ item_desc = nullptr) {
    int* hist = smi;                      // [8][K]   per-warp histogram, then per-warp cursor
    int* tot = smi + 8 * K;               // [K+1]    bin offsets
    int* itm = tot + K + 1;               // [K+1]    work-item offsets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* am = argmin_l + (size_t)b * P;
    const int per = ((P + 7) / 8 + 31) & ~31;            // prototypes per warp, multiple of 32
    const int pa = warp * per, pb = min(P, pa + per);
    for (int i = tid; i < 8 * K; i += 256) hist[i] = 0;
    __syncthreads();
    for (int p0 = pa; p0 < pb; p0 += 32) {               // pass A
        const int p = p0 + lane;
        const bool valid = p < pb;
        const int a = valid ? min(max(kCoherentKeys ? __ldcg(am + p) : __ldg(am + p), 0), K - 1) : -1 - lane;   // clamped: a bad index must not leave the bins
        const unsigned peers = __match_any_sync(0xffffffffu, a);
        if (valid && lane == __ffs(peers) - 1) hist[warp * K + a] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {                                      // scan over tokens of the 8-warp totals
        int carry = 0, icarry = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            int n = 0;
            if (k < K)
                for (int w = 0; w < 8; ++w) n += hist[w * K + k];
            int v = n, c = (n + kBinChunk - 1) / kBinChunk;
            if (k < K && c == 0) c = 1;                   // empty bins still own one item (they write zeros)
            int ci = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                const int ui = __shfl_up_sync(0xffffffffu, ci, o);
                if (lane >= o) { v += u; ci += ui; }
            }
            if (k < K) { tot[k] = carry + v - n; itm[k] = icarry + ci - c; }
            carry += __shfl_sync(0xffffffffu, v, 31);
            icarry += __shfl_sync(0xffffffffu, ci, 31);
        }
        if (lane == 0) { tot[K] = carry; itm[K] = icarry; }
    }
    __syncthreads();
    for (int k = tid; k < K; k += 256) {                  // per-warp cursors: bin offset + counts of earlier warps
        int run = tot[k];
        for (int w = 0; w < 8; ++w) {
            const int n = hist[w * K + k];
            hist[w * K + k] = run;
            run += n;
        }
    }
    for (int k = tid; k <= K; k += 256) {
        bin_start[(size_t)b * (K + 1) + k] = tot[k];
        item_start[(size_t)b * (K + 1) + k] = itm[k];
    }
    if (item_desc) {
        const int ipi = bin_items_per_image(K, P);
        int4* dsc = item_desc + (size_t)b * ipi;
        for (int k = tid; k < K; k += 256) {
            const int e0 = tot[k], e1 = tot[k + 1], i0 = itm[k], nch = itm[k + 1] - i0;
            for (int c = 0; c < nch; ++c)
                dsc[i0 + c] = make_int4(k, e0 + c * kBinChunk, min(e1, e0 + (c + 1) * kBinChunk), nch | (c << 16));
        }
        for (int i = itm[K] + tid; i < ipi; i += 256) dsc[i] = make_int4(-1, 0, 0, 0);
    }
    __syncthreads();
    int32_t* list = bin_list + (size_t)b * P;
    for (int p0 = pa; p0 < pb; p0 += 32) {               // pass B
        const int p = p0 + lane;
        const bool valid = p < pb;
        const int a = valid ? min(max(kCoherentKeys ? __ldcg(am + p) : __ldg(am + p), 0), K - 1) : -1 - lane;   // clamped: a bad index must not leave the bins
        const unsigned peers = __match_any_sync(0xffffffffu, a);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (valid && lane == leader) {
            base = hist[warp * K + a];
            hist[warp * K + a] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) list[base + rank] = p;
        __syncwarp();
    }
}


inline size_t bin_tokens_smem_bytes(int K) { return sizeof(int) * ((size_t)8 * K + 2 * (size_t)(K + 1)); }

}  // namespace pph
