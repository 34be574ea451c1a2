// Shared helpers for the prototype-head kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/protohead.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "protohead_b200 is written for sm_100a only"
#endif

namespace pph {

// ---------------------------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define PPH_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            ::pph::set_error(__VA_ARGS__);      \
            return (code);                      \
        }                                       \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline cudaStream_t as_stream(pph_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (env PPH_PDL=1, read once): every kernel of this library is launched with the
// programmatic-stream-serialization attribute, so the next kernel's CTAs are scheduled and run their prologue while
// the previous kernel drains; each kernel executes pdl_sync() before its first global-memory access, which blocks
// until every prerequisite grid has completed and flushed (so data hazards are exactly those of plain stream order).
// Under stream capture the attribute becomes a programmatic edge of the CUDA graph.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // status is read by launch_status()
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
// First statement of every kernel (see launch_k): let the dependent grid start its prologue, then wait for the
// prerequisite grids.  Both are no-ops for a launch without the programmatic attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
    pdl_launch_dependents();
    pdl_wait();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// activation of a distance, protopformer.py:228-234 (precise log/div: parity is 1e-4 on a cancelling quantity)
__device__ __forceinline__ float act_of_dist(float d, int act_fn, float eps) {
    return act_fn == PPH_ACT_LOG ? logf((d + 1.0f) / (d + eps)) : -d;
}

// d act / d d, including the relu mask of protopformer.py:216 (relu'(0) = 0 as in torch)
__device__ __forceinline__ float dact_of_dist(float d, int act_fn, float eps) {
    if (!(d > 0.0f)) return 0.0f;
    return act_fn == PPH_ACT_LOG ? (1.0f / (d + 1.0f) - 1.0f / (d + eps)) : -1.0f;
}

__device__ __forceinline__ uint16_t bf16_bits(float x) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float bf16_to_float(uint16_t b) {
    return __uint_as_float(((uint32_t)b) << 16);
}

}  // namespace pph
