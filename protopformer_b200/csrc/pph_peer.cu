// Gradient exchange of the data-parallel head over NVLink / NVSwitch peer memory: one kernel, in place.
//
// The reference leaves this to DistributedDataParallel (main.py:369-371: NCCL bucket all-reduce overlapped with the
// backward).  Here every rank's flat gradient buffer is a peer-mapped ("symmetric") allocation, and the averaging
// all-reduce is one launch on the step's own stream / graph branch:
//
//   barrier 1   CTA c of rank r tells CTA c of every peer "my gradients are final" (a flag in the peer's memory,
//               st.release.sys) and waits for theirs (ld.acquire.sys)
//   reduce      rank r owns chunk r of [lo, lo + n): it sums chunk r of every rank's buffer in rank order (so the
//               result does not depend on timing), scales by 1/world and writes the average into chunk r of EVERY
//               rank's buffer.  With a multicast mapping (NVLS) the sum is one multimem.ld_reduce executed by the
//               switch and the write one multimem.st: each GPU then moves n/world floats each way instead of
//               (world-1) n/world.
//   barrier 2   "my chunk is written everywhere"; when a rank's kernel ends all chunks of its buffer are final.
//
// Chunk r of a rank's buffer is read only by rank r, and overwritten only by rank r after it has read it, so the
// operation is in place without staging.  Flags are monotonically increasing epochs kept next to the buffer (two per
// call), so graph replays need no reset.  Every rank must issue the same sequence of calls per `slot`.
#include "pph_common.cuh"

namespace pph {
namespace {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxCtas = 64;
constexpr int kPeerSlots = 4;
constexpr int kPeerThreads = 512;

struct PeerArgs {
    float* buf[kPeerMaxWorld];          // this buffer as mapped on every rank (buf[rank] = the local one)
    float* mc;                          // multicast mapping of the same buffer or nullptr
    long long flag_off;                 // byte offset of the flag block inside each buffer
    long long lo, n;                    // floats, multiples of 4
    int rank, world, slot;
    float inv;
};

// flag block: [slot][cta][src rank] arrival epochs written by the peers, then [slot][cta] this rank's own epoch
__device__ __forceinline__ unsigned* flag_ptr(float* base, long long off, int slot, int cta, int src) {
    return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(base) + off) +
           ((size_t)slot * kPeerMaxCtas + cta) * kPeerMaxWorld + src;
}
__device__ __forceinline__ unsigned* epoch_ptr(float* base, long long off, int slot, int cta) {
    return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(base) + off) +
           (size_t)kPeerSlots * kPeerMaxCtas * kPeerMaxWorld + (size_t)slot * kPeerMaxCtas + cta;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 mc_ld_reduce4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st4(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all `world` peers' CTA `cta` reach this point (bounded spin: a missing peer traps instead of hanging the GPU).
// kRelease = false: the arrival only announces data that EARLIER kernels of this rank left in its own memory (visible
// to peer loads since the kernel boundary), so the flag is a relaxed store -- a system-scope fence over NVLink costs
// microseconds; kRelease = true: this CTA's stores into the peers' memory are ordered before the flag.
template <bool kRelease>
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int cta, unsigned target) {
    const int t = threadIdx.x;
    __syncthreads();
    if (t < a.world) {
        unsigned* theirs = flag_ptr(a.buf[t], a.flag_off, a.slot, cta, a.rank);
        if constexpr (kRelease) st_release_sys(theirs, target);
        else st_relaxed_sys(theirs, target);
        const unsigned* mine = flag_ptr(a.buf[a.rank], a.flag_off, a.slot, cta, t);
        long long spins = 0;
        while ((int)(ld_acquire_sys(mine) - target) < 0) {
            if (++spins > (1ll << 28)) __trap();
        }
    }
    __syncthreads();
}

template <bool kMulticast, int kWorld, int kBatch>
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs a) {
    const int cta = blockIdx.x, nc = gridDim.x;
    unsigned* ep = epoch_ptr(a.buf[a.rank], a.flag_off, a.slot, cta);
    const unsigned e = *reinterpret_cast<volatile unsigned*>(ep);
    peer_barrier<false>(a, cta, e + 1);
    const long long n4 = a.n >> 2;
    const long long per = (n4 + a.world - 1) / a.world;
    const long long c0 = a.rank * per, c1 = min(c0 + per, n4);
    // NVLink round trips are microseconds: every thread issues a whole batch of loads before it touches the first result,
    // so that the chunk is (nearly) one round trip deep instead of one per element
    const long long stride = (long long)nc * kPeerThreads;
    for (long long b = c0 + (long long)cta * kPeerThreads + threadIdx.x; b < c1; b += kBatch * stride) {
        float4 s[kBatch];
        if constexpr (kMulticast) {
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
                if (b + k * stride < c1) s[k] = mc_ld_reduce4(a.mc + a.lo + ((b + k * stride) << 2));
        } else {
            float4 v[kBatch][kWorld];
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
#pragma unroll
                for (int p = 0; p < kWorld; ++p)
                    if (p < a.world && b + k * stride < c1) v[k][p] = ld_peer4(a.buf[p] + a.lo + ((b + k * stride) << 2));
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                s[k] = v[k][0];
#pragma unroll
                for (int p = 1; p < kWorld; ++p)
                    if (p < a.world) { s[k].x += v[k][p].x; s[k].y += v[k][p].y; s[k].z += v[k][p].z; s[k].w += v[k][p].w; }
            }
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            if (b + k * stride >= c1) continue;
            const long long off = a.lo + ((b + k * stride) << 2);
            float4 r = s[k];
            r.x *= a.inv; r.y *= a.inv; r.z *= a.inv; r.w *= a.inv;
            if constexpr (kMulticast) {
                mc_st4(a.mc + off, r);
            } else {
#pragma unroll
                for (int p = 0; p < kWorld; ++p)
                    if (p < a.world) *reinterpret_cast<float4*>(a.buf[p] + off) = r;
            }
        }
    }
    peer_barrier<true>(a, cta, e + 2);
    if (threadIdx.x == 0) *ep = e + 2;
}

}  // namespace
}  // namespace pph

using namespace pph;

extern "C" int pph_peer_flag_bytes(long long* bytes) {
    PPH_REQUIRE(bytes, PPH_EINVAL, "pph_peer_flag_bytes: null argument");
    *bytes = (long long)sizeof(unsigned) * ((size_t)kPeerSlots * kPeerMaxCtas * kPeerMaxWorld + (size_t)kPeerSlots * kPeerMaxCtas);
    return 0;
}

extern "C" int pph_peer_allreduce(const unsigned long long* buf_ptrs, unsigned long long multicast_ptr,
                                  long long flag_offset_bytes, int rank, int world, long long lo, long long n,
                                  int n_ctas, int slot, pph_stream_t stream) {
    PPH_REQUIRE(buf_ptrs && world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, PPH_EINVAL,
                "pph_peer_allreduce: world in [1,16], rank in [0,world), buffer table required");
    PPH_REQUIRE(n_ctas >= 1 && n_ctas <= kPeerMaxCtas && slot >= 0 && slot < kPeerSlots, PPH_EINVAL,
                "pph_peer_allreduce: n_ctas in [1,64], slot in [0,4)");
    PPH_REQUIRE(lo >= 0 && n >= 0 && !(lo & 3) && !(n & 3) && !(flag_offset_bytes & 15) && flag_offset_bytes >= (lo + n) * 4,
                PPH_EINVAL, "pph_peer_allreduce: lo and n must be multiples of 4 floats; the flag block lies behind the data");
    if (n == 0 || world == 1) return 0;
    PeerArgs a{};
    for (int p = 0; p < world; ++p) {
        PPH_REQUIRE(buf_ptrs[p], PPH_EINVAL, "pph_peer_allreduce: null peer buffer");
        a.buf[p] = reinterpret_cast<float*>(buf_ptrs[p]);
    }
    a.mc = reinterpret_cast<float*>(multicast_ptr);
    a.flag_off = flag_offset_bytes;
    a.lo = lo; a.n = n; a.rank = rank; a.world = world; a.slot = slot;
    a.inv = 1.0f / (float)world;
    cudaStream_t st = as_stream(stream);
    // same (maximum) shared-memory carve-out as the tcgen05 kernels this launch runs beside: an SM never has to drain to
    // switch its L1 / shared split before it can take their CTAs
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        kern<<<n_ctas, kPeerThreads, 0, st>>>(a);
    };
    if (a.mc) go(peer_allreduce_kernel<true, 1, 8>);
    else if (world <= 2) go(peer_allreduce_kernel<false, 2, 8>);
    else if (world <= 4) go(peer_allreduce_kernel<false, 4, 4>);
    else if (world <= 8) go(peer_allreduce_kernel<false, 8, 2>);
    else go(peer_allreduce_kernel<false, 16, 1>);
    return launch_status("pph_peer_allreduce");
}
