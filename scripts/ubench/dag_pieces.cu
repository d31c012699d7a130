// Micro-benchmark (B200, one CTA of 256 threads): cycles of the building blocks of csrc/chol_dag.cu.
#include <cstdio>
#include <vector>
#include "../../mxfusion_b200/csrc/chol_dag.cu"

namespace mxf { std::atomic<uint64_t> g_launches{0}; }
using namespace mxf;

__global__ void __launch_bounds__(256, 2) pieces(const float* A, float* out, long long* cyc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufX = reinterpret_cast<float*>(smem_raw);
    float* bufY = bufX + TILE_F;
    float* Dp = bufY + TILE_F;
    float* Wp = Dp + B * LDP;
    float* Tm = Wp + B * LDP;
    const Map m;
    for (int e = threadIdx.x; e < TILE_F; e += 256) { bufX[e] = A[e]; bufY[e] = A[(e * 7) % TILE_F]; }
    for (int e = threadIdx.x; e < 64 * 64; e += 256) Dp[(e >> 6) * LDP + (e & 63)] = A[e];
    __syncthreads();
    float acc[4][4] = {};
    for (int rep = 0; rep < 3; ++rep) {
        long long t0 = clock64();
        mma_nt<false>(bufX, bufY, acc, m);
        __syncthreads();
        long long t1 = clock64();
        mma_nn(bufX, bufY, acc, m);
        __syncthreads();
        long long t2 = clock64();
        small_mm32<true, false>(Tm, Dp + 32 * LDP, Wp, 1.f);
        __syncthreads();
        long long t3 = clock64();
        for (int e = threadIdx.x; e < 64 * 64; e += 256) Dp[(e >> 6) * LDP + (e & 63)] = A[e];
        __syncthreads();
        long long t4 = clock64();
        if (threadIdx.x < 32) warp_chol_panel32<true>(Dp, 0, threadIdx.x);
        __syncthreads();
        long long t5 = clock64();
        for (int e = threadIdx.x; e < 64 * 64; e += 256) Dp[(e >> 6) * LDP + (e & 63)] = A[e];
        __syncthreads();
        long long t6 = clock64();
        diag_block_64<true>(Dp, Wp, Tm);
        __syncthreads();
        long long t7 = clock64();
        if (threadIdx.x < 32) warp_inv32(Dp, Wp, 0, threadIdx.x);
        __syncthreads();
        long long t8 = clock64();
        if (threadIdx.x < 32) warp_chol_panel32<false>(Dp, 32, threadIdx.x);
        __syncthreads();
        long long t9 = clock64();
        if (threadIdx.x == 0) {
            long long* c = cyc + rep * 8;
            c[0] = t1 - t0; c[1] = t2 - t1; c[2] = t3 - t2; c[3] = t5 - t4; c[4] = t7 - t6; c[5] = t8 - t7; c[6] = t9 - t8;
        }
    }
    store_acc(out, 64, acc, 64, 64, true, m);
}

int main() {
    std::vector<float> A(64 * 64);
    for (int i = 0; i < 64; ++i)
        for (int j = 0; j < 64; ++j) A[i * 64 + j] = expf(-0.02f * (i - j) * (i - j)) + (i == j ? 0.1f : 0.f);
    float *dA, *dO;
    long long* dC;
    cudaMalloc(&dA, 64 * 64 * 4);
    cudaMalloc(&dO, 64 * 64 * 4);
    cudaMalloc(&dC, 24 * 8);
    cudaMemcpy(dA, A.data(), 64 * 64 * 4, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * (2 * TILE_F + 2 * B * LDP + 32 * LDP) + 64;
    cudaFuncSetAttribute(pieces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pieces<<<1, 256, smem>>>(dA, dO, dC);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[24];
    cudaMemcpy(c, dC, 24 * 8, cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(e));
    for (int rep = 0; rep < 3; ++rep)
        printf("rep %d: mma_nt %lld  mma_nn %lld  small_mm32 %lld  panel32<two> %lld  diag_block_64 %lld  inv32 %lld  panel32<one> %lld\n", rep, c[rep * 8],
               c[rep * 8 + 1], c[rep * 8 + 2], c[rep * 8 + 3], c[rep * 8 + 4], c[rep * 8 + 5], c[rep * 8 + 6]);
    return 0;
}
