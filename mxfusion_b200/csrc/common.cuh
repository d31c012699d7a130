// Shared device/host helpers for libmxf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/mxf_b200.h"

namespace mxf {

extern std::atomic<uint64_t> g_launches;   // statistics only; never read by a kernel

inline int after_launch(int n = 1) {
    g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? MXF_OK : (int)e;
}

constexpr int kNumSMs = 148;   // B200

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
