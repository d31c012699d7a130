// Micro-benchmark (B200): FP32 FMA issue rate -- scalar FFMA (3 distinct registers), FFMA2 (fma.rn.f32x2), per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 scripts/ubench/ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&d), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    d = *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float x, float y) {
    float acc[16];
    float2 acc2[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float a0 = x, a1 = y, a2 = x + y, a3 = x - y;
    float2 p0 = make_float2(x, y), p1 = make_float2(y, x);
    for (int t = 0; t < iters; ++t) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], (i & 1) ? a0 : a1, (i & 2) ? a2 : a3);     // 16 FFMA
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(acc2[i], (i & 1) ? p0 : p1, acc2[(i + 1) & 7]);            // 8 FFMA2 = 16 FMA
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc2[i].x + acc2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
    float* out;
    cudaMalloc(&out, 148 * 4 * 512 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int threads : {128, 256, 512}) {
        k<MODE><<<148 * 2, threads>>>(out, 100, 1.0001f, 0.9999f);
        cudaEventRecord(a);
        k<MODE><<<148 * 2, threads>>>(out, iters, 1.0001f, 0.9999f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double fma = 148.0 * 2 * threads * (double)iters * 16;
        printf("%s threads/CTA=%d (2 CTA/SM): %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM at 1.9 GHz\n", name, threads, ms,
               2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
    }
    cudaFree(out);
}

int main() {
    run<0>("FFMA ");
    run<1>("FFMA2");
    return 0;
}
