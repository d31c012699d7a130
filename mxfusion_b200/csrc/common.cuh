// Shared device/host helpers for libmxf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdlib>
#include "../../include/mxf_b200.h"

namespace mxf {

extern std::atomic<uint64_t> g_launches;   // statistics only; never read by a kernel

inline int after_launch(int n = 1) {
    g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? MXF_OK : (int)e;
}

constexpr int kNumSMs = 148;   // B200

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The step is a long chain of short dependent kernels (blocked potrf / trsm).  A kernel that (a) calls
// pdl_launch_dependents() on entry and (b) calls pdl_wait() before its first global-memory access can be launched with
// launch_pdl(): its grid is scheduled while the predecessor drains, so launch latency and the prologue (mbarrier init,
// TMEM allocation) leave the critical path.  pdl_wait() returns only when the predecessor grid has completed and its
// writes are visible, so there is no hazard; a kernel that does not call pdl_wait() must NOT be launched this way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
    static int on = [] {
        const char* e = getenv("MXF_PDL");      // measured on B200: no gain under CUDA-graph replay -> opt-in
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
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
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in thread 0 (and in all threads of warp 0).
// `red` must hold >= 32 elements of T in shared memory.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();             // protect `red` from a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
    if (wid == 0) v = warp_sum(v);
    return v;
}

#define MXF_DISPATCH_DTYPE(dtype, ...)                          \
    do {                                                         \
        if ((dtype) == MXF_F32) { using T = float; __VA_ARGS__; } \
        else if ((dtype) == MXF_F64) { using T = double; __VA_ARGS__; } \
        else return MXF_EDTYPE;                                  \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace mxf
