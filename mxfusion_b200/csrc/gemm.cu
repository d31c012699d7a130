// Batched row-major GEMM on the FP32/FP64 FMA pipes:  C = alpha * op(A) op(B) + beta * C.
//
// Replaces mx.nd.linalg.gemm2 / syrk / trmm call sites of the reference
// (modules/gp_modules/svgp_regression.py:76,82,89,90; gp/kernels/stationary.py:94,102 are folded into
// kbuild.cu instead) and is the update engine of the blocked potrf / trsm in potrf.cu / trsm.cu.
// This is the exact-precision path (f32 FMA, f64 FMA); the f32 tensor-core path (3xTF32 on tcgen05)
// lives in gemm_tc.cu and is selected by mxf_gemm for large f32 problems when enabled.
//
// Tiling: CTA tile BM x BN, K-step BK, 256 threads, each thread a (2*H) x (2*H) micro-tile split
// into four H x H quadrants (rows ty*H and BM/2+ty*H; columns tx*H and BN/2+tx*H) so that the
// shared-memory reads are 16-byte and bank-conflict free.  Global tiles are prefetched into
// registers while the current tile is multiplied (software double buffering).
#include <algorithm>
#include "common.cuh"

namespace mxf {

template <typename T, int BM, int BN, int BK>
struct GemmCfg {
    static constexpr int THREADS = 256;
    static constexpr int H = BM / 32;              // quadrant edge: 4 (f32, 128) or 2 (f64, 64)
    static constexpr int PAD = 4;
    static constexpr int LDS_A = BM + PAD;
    static constexpr int LDS_B = BN + PAD;
    static constexpr int EA = BM * BK / THREADS;   // elements of the A tile per thread
    static constexpr int EB = BN * BK / THREADS;
};

// Load this thread's slice of a (MN x BK) operand tile into registers.
// kcontig: the operand is stored with K contiguous (A not transposed / B transposed).
template <typename T, int BMN, int BK, int E>
__device__ __forceinline__ void load_tile(T (&reg)[E], const T* __restrict__ P, int64_t ld, bool kcontig,
                                          int mn0, int k0, int mn_lim, int k_lim) {
    const int t = threadIdx.x;
    if (kcontig) {
        // thread -> (row = t / (BK/E), kk = (t % (BK/E)) * E .. +E)
        constexpr int TPR = BK / E;
        const int r = t / TPR, kk = (t % TPR) * E;
        const int mn = mn0 + r;
        const T* src = P + (int64_t)mn * ld + k0 + kk;
#pragma unroll
        for (int e = 0; e < E; ++e) reg[e] = (mn < mn_lim && k0 + kk + e < k_lim) ? src[e] : T(0);
    } else {
        // thread -> (kk = t / (BMN/E), mn = (t % (BMN/E)) * E .. +E)
        constexpr int TPK = BMN / E;
        const int kk = t / TPK, c = (t % TPK) * E;
        const T* src = P + (int64_t)(k0 + kk) * ld + mn0 + c;
#pragma unroll
        for (int e = 0; e < E; ++e) reg[e] = (k0 + kk < k_lim && mn0 + c + e < mn_lim) ? src[e] : T(0);
    }
}

template <typename T, int BMN, int BK, int E, int LDS>
__device__ __forceinline__ void store_tile(const T (&reg)[E], T* __restrict__ sm, bool kcontig) {
    const int t = threadIdx.x;
    if (kcontig) {
        constexpr int TPR = BK / E;
        const int r = t / TPR, kk = (t % TPR) * E;
#pragma unroll
        for (int e = 0; e < E; ++e) sm[(kk + e) * LDS + r] = reg[e];
    } else {
        constexpr int TPK = BMN / E;
        const int kk = t / TPK, c = (t % TPK) * E;
#pragma unroll
        for (int e = 0; e < E; ++e) sm[kk * LDS + c + e] = reg[e];
    }
}

template <typename T, int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
gemm_kernel(int transA, int transB, int m, int n, int k, T alpha, const T* __restrict__ A, int64_t lda,
            int64_t sA, const T* __restrict__ B, int64_t ldb, int64_t sB, T beta, T* __restrict__ C,
            int64_t ldc, int64_t sC, int tri) {
    using Cfg = GemmCfg<T, BM, BN, BK>;
    constexpr int H = Cfg::H;
    __shared__ __align__(16) T As[2][BK * Cfg::LDS_A];
    __shared__ __align__(16) T Bs[2][BK * Cfg::LDS_B];

    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    if ((tri & 1) && j0 > i0 + BM - 1) return;      // bits 2/4 (triangular A) are hints: the full product is the same
    const int s = blockIdx.z;
    A += (int64_t)s * sA;
    B += (int64_t)s * sB;
    C += (int64_t)s * sC;

    const bool a_kc = (transA == 0), b_kc = (transB != 0);
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;

    T acc[2 * H][2 * H];
#pragma unroll
    for (int r = 0; r < 2 * H; ++r)
#pragma unroll
        for (int c = 0; c < 2 * H; ++c) acc[r][c] = T(0);

    T ra[Cfg::EA], rb[Cfg::EB];
    load_tile<T, BM, BK, Cfg::EA>(ra, A, lda, a_kc, i0, 0, m, k);
    load_tile<T, BN, BK, Cfg::EB>(rb, B, ldb, b_kc, j0, 0, n, k);
    store_tile<T, BM, BK, Cfg::EA, Cfg::LDS_A>(ra, As[0], a_kc);
    store_tile<T, BN, BK, Cfg::EB, Cfg::LDS_B>(rb, Bs[0], b_kc);
    __syncthreads();

    int buf = 0;
    for (int k0 = 0; k0 < k; k0 += BK) {
        const bool more = k0 + BK < k;
        if (more) {
            load_tile<T, BM, BK, Cfg::EA>(ra, A, lda, a_kc, i0, k0 + BK, m, k);
            load_tile<T, BN, BK, Cfg::EB>(rb, B, ldb, b_kc, j0, k0 + BK, n, k);
        }
        const T* as = As[buf];
        const T* bs = Bs[buf];
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[2 * H], b[2 * H];
#pragma unroll
            for (int e = 0; e < H; ++e) {
                a[e] = as[kk * Cfg::LDS_A + ty * H + e];
                a[H + e] = as[kk * Cfg::LDS_A + BM / 2 + ty * H + e];
                b[e] = bs[kk * Cfg::LDS_B + tx * H + e];
                b[H + e] = bs[kk * Cfg::LDS_B + BN / 2 + tx * H + e];
            }
#pragma unroll
            for (int r = 0; r < 2 * H; ++r)
#pragma unroll
                for (int c = 0; c < 2 * H; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        if (more) {
            store_tile<T, BM, BK, Cfg::EA, Cfg::LDS_A>(ra, As[buf ^ 1], a_kc);
            store_tile<T, BN, BK, Cfg::EB, Cfg::LDS_B>(rb, Bs[buf ^ 1], b_kc);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int r = 0; r < 2 * H; ++r) {
        const int i = i0 + (r < H ? ty * H + r : BM / 2 + ty * H + (r - H));
        if (i >= m) continue;
#pragma unroll
        for (int c = 0; c < 2 * H; ++c) {
            const int j = j0 + (c < H ? tx * H + c : BN / 2 + tx * H + (c - H));
            if (j >= n) continue;
            T* dst = C + (int64_t)i * ldc + j;
            T v = alpha * acc[r][c];
            if (beta != T(0)) v = fma(beta, *dst, v);
            *dst = v;
        }
    }
}

template <typename T>
int gemm_simt(int transA, int transB, int m, int n, int k, double alpha, const T* A, int64_t lda, int64_t sA,
              const T* B, int64_t ldb, int64_t sB, double beta, T* C, int64_t ldc, int64_t sC, int S, int tri,
              cudaStream_t st) {
    if (m == 0 || n == 0 || S == 0) return MXF_OK;
    if constexpr (sizeof(T) == 4) {
        constexpr int BM = 128, BN = 128, BK = 16;
        dim3 grid(cdiv(n, BN), cdiv(m, BM), S);
        gemm_kernel<T, BM, BN, BK><<<grid, 256, 0, st>>>(transA, transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB,
                                                        (T)beta, C, ldc, sC, tri);
    } else {
        constexpr int BM = 64, BN = 64, BK = 16;
        dim3 grid(cdiv(n, BN), cdiv(m, BM), S);
        gemm_kernel<T, BM, BN, BK><<<grid, 256, 0, st>>>(transA, transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB,
                                                        (T)beta, C, ldc, sC, tri);
    }
    return after_launch();
}

// ------------------------------------------------------------------------------------------------------------
// Skinny products (n <= 8 columns): the matrix-vector shaped pieces of the SVGP bound -- A^T (L^-1 mu), A Y, Phi mt
// (svgp_regression.py:82,89) -- are bandwidth-bound passes over A, not tile GEMMs.
//   non-transposed A (m x k): one warp per output row, lanes stride over k, warp-shuffle reduction;
//   transposed A (stored k x m): 32 consecutive output rows per CTA (coalesced 128-byte reads of A), the 8 warps
//   split k, partial sums combined through shared memory (deterministic).
// ------------------------------------------------------------------------------------------------------------
template <typename T, int NCOL>
__global__ void __launch_bounds__(256)
gemm_skinny_n_kernel(int transB, int m, int n, int k, T alpha, const T* __restrict__ A, int64_t lda, int64_t sA,
                     const T* __restrict__ B, int64_t ldb, int64_t sB, T beta, T* __restrict__ C, int64_t ldc,
                     int64_t sC) {
    const int s = blockIdx.y;
    A += (int64_t)s * sA; B += (int64_t)s * sB; C += (int64_t)s * sC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= m) return;
    T acc[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) acc[j] = T(0);
    const T* arow = A + (int64_t)i * lda;
    for (int kk = lane; kk < k; kk += 32) {
        const T a = arow[kk];
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (j < n) acc[j] = fma(a, transB ? B[(int64_t)j * ldb + kk] : B[(int64_t)kk * ldb + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NCOL; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (j < n) {
                T* dst = C + (int64_t)i * ldc + j;
                T v = alpha * acc[j];
                if (beta != T(0)) v = fma(beta, *dst, v);
                *dst = v;
            }
    }
}

// Long-k variant (k >= 4096, e.g. b += L^-1 K(Z, X_c) Y_c over a streamed block of 28k rows): a CTA owns 4 output rows and
// all 256 threads stride over k with 16-byte loads, 4 rows x 2 unrolled steps = 8 independent loads in flight per thread,
// so the pass over A runs at memory speed instead of on one warp's load latency chain.  Requires lda % 4 == 0 and A
// 16-byte aligned (fp32); B is read through the read-only path (it is re-read by every CTA and stays in L1/L2).
template <int NCOL>
__global__ void __launch_bounds__(256)
gemm_skinny_longk_kernel(int transB, int m, int n, int k, float alpha, const float* __restrict__ A, int64_t lda, int64_t sA,
                         const float* __restrict__ B, int64_t ldb, int64_t sB, float beta, float* __restrict__ C,
                         int64_t ldc, int64_t sC) {
    constexpr int RPC = 4;
    __shared__ float part[8][RPC][NCOL];
    const int s = blockIdx.y;
    A += (int64_t)s * sA; B += (int64_t)s * sB; C += (int64_t)s * sC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * RPC;
    float acc[RPC][NCOL];
#pragma unroll
    for (int r = 0; r < RPC; ++r)
#pragma unroll
        for (int j = 0; j < NCOL; ++j) acc[r][j] = 0.f;
    const float* arow[RPC];
#pragma unroll
    for (int r = 0; r < RPC; ++r) arow[r] = A + (int64_t)min(i0 + r, m - 1) * lda;
    const int k4 = k & ~3;
#pragma unroll 2
    for (int kk = threadIdx.x * 4; kk < k4; kk += 1024) {
        float4 a[RPC];
#pragma unroll
        for (int r = 0; r < RPC; ++r) a[r] = __ldcs(reinterpret_cast<const float4*>(arow[r] + kk));
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            if (j < n) {
                float b0, b1, b2, b3;
                if (transB) {
                    const float* bp = B + (int64_t)j * ldb + kk;
                    b0 = __ldg(bp); b1 = __ldg(bp + 1); b2 = __ldg(bp + 2); b3 = __ldg(bp + 3);
                } else {
                    const float* bp = B + (int64_t)kk * ldb + j;
                    b0 = __ldg(bp); b1 = __ldg(bp + ldb); b2 = __ldg(bp + 2 * ldb); b3 = __ldg(bp + 3 * ldb);
                }
#pragma unroll
                for (int r = 0; r < RPC; ++r)
                    acc[r][j] = fmaf(a[r].x, b0, fmaf(a[r].y, b1, fmaf(a[r].z, b2, fmaf(a[r].w, b3, acc[r][j]))));
            }
        }
    }
    for (int kk = k4 + threadIdx.x; kk < k; kk += 256) {           // tail (k % 4 elements)
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (j < n) {
                const float b = transB ? B[(int64_t)j * ldb + kk] : B[(int64_t)kk * ldb + j];
#pragma unroll
                for (int r = 0; r < RPC; ++r) acc[r][j] = fmaf(arow[r][kk], b, acc[r][j]);
            }
    }
#pragma unroll
    for (int r = 0; r < RPC; ++r)
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const float v = warp_sum(acc[r][j]);
            if (lane == 0) part[warp][r][j] = v;
        }
    __syncthreads();
    if (threadIdx.x < RPC * NCOL) {
        const int r = threadIdx.x / NCOL, j = threadIdx.x - r * NCOL;
        if (i0 + r < m && j < n) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += part[w][r][j];
            float* dst = C + (int64_t)(i0 + r) * ldc + j;
            float v = alpha * t;
            if (beta != 0.f) v = fmaf(beta, *dst, v);
            *dst = v;
        }
    }
}

template <typename T, int NCOL>
__global__ void __launch_bounds__(256)
gemm_skinny_t_kernel(int transB, int m, int n, int k, T alpha, const T* __restrict__ A, int64_t lda, int64_t sA,
                     const T* __restrict__ B, int64_t ldb, int64_t sB, T beta, T* __restrict__ C, int64_t ldc,
                     int64_t sC) {
    __shared__ T part[8][NCOL][33];
    const int s = blockIdx.y;
    A += (int64_t)s * sA; B += (int64_t)s * sB; C += (int64_t)s * sC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 32 + lane;
    T acc[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) acc[j] = T(0);
    if (i < m) {
        for (int kk = warp; kk < k; kk += 8) {
            const T a = A[(int64_t)kk * lda + i];
#pragma unroll
            for (int j = 0; j < NCOL; ++j)
                if (j < n) acc[j] = fma(a, transB ? B[(int64_t)j * ldb + kk] : B[(int64_t)kk * ldb + j], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NCOL; ++j) part[warp][j][lane] = acc[j];
    __syncthreads();
    if (warp == 0 && i < m) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (j < n) {
                T t = T(0);
#pragma unroll
                for (int w = 0; w < 8; ++w) t += part[w][j][lane];
                T* dst = C + (int64_t)i * ldc + j;
                T v = alpha * t;
                if (beta != T(0)) v = fma(beta, *dst, v);
                *dst = v;
            }
    }
}

// k <= 8 (outer products such as  Kuf_bar += (g s beta w) Y^T ): one coalesced pass over C.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_rank_k_kernel(int transA, int transB, int m, int n, int k, T alpha, const T* __restrict__ A, int64_t lda, int64_t sA,
                   const T* __restrict__ B, int64_t ldb, int64_t sB, T beta, T* __restrict__ C, int64_t ldc, int64_t sC) {
    const int s = blockIdx.z;
    A += (int64_t)s * sA; B += (int64_t)s * sB; C += (int64_t)s * sC;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    T b[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) b[kk] = (kk < k) ? (transB ? B[(int64_t)j * ldb + kk] : B[(int64_t)kk * ldb + j]) : T(0);
    for (int i = blockIdx.y; i < m; i += gridDim.y) {
        T acc = T(0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            if (kk < k) acc = fma(transA ? A[(int64_t)kk * lda + i] : A[(int64_t)i * lda + kk], b[kk], acc);
        T* dst = C + (int64_t)i * ldc + j;
        T v = alpha * acc;
        if (beta != T(0)) v = fma(beta, *dst, v);
        *dst = v;
    }
}

template <typename T>
static int gemm_skinny(int transA, int transB, int m, int n, int k, double alpha, const T* A, int64_t lda, int64_t sA,
                       const T* B, int64_t ldb, int64_t sB, double beta, T* C, int64_t ldc, int64_t sC, int S,
                       cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        if (!transA && k >= 4096 && (lda & 3) == 0 && (sA & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
            dim3 grid(cdiv(m, 4), S);
            if (n <= 1) gemm_skinny_longk_kernel<1><<<grid, 256, 0, st>>>(transB, m, n, k, (float)alpha, A, lda, sA, B, ldb, sB, (float)beta, C, ldc, sC);
            else if (n <= 4) gemm_skinny_longk_kernel<4><<<grid, 256, 0, st>>>(transB, m, n, k, (float)alpha, A, lda, sA, B, ldb, sB, (float)beta, C, ldc, sC);
            else gemm_skinny_longk_kernel<8><<<grid, 256, 0, st>>>(transB, m, n, k, (float)alpha, A, lda, sA, B, ldb, sB, (float)beta, C, ldc, sC);
            return after_launch();
        }
    }
    if (!transA) {
        dim3 grid(cdiv(m, 8), S);
        if (n <= 1) gemm_skinny_n_kernel<T, 1><<<grid, 256, 0, st>>>(transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
        else if (n <= 4) gemm_skinny_n_kernel<T, 4><<<grid, 256, 0, st>>>(transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
        else gemm_skinny_n_kernel<T, 8><<<grid, 256, 0, st>>>(transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
    } else {
        dim3 grid(cdiv(m, 32), S);
        if (n <= 1) gemm_skinny_t_kernel<T, 1><<<grid, 256, 0, st>>>(transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
        else if (n <= 4) gemm_skinny_t_kernel<T, 4><<<grid, 256, 0, st>>>(transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
        else gemm_skinny_t_kernel<T, 8><<<grid, 256, 0, st>>>(transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
    }
    return after_launch();
}

int gemm_tc_f32(int transA, int transB, int m, int n, int k, double alpha, const float* A, int64_t lda, int64_t sA,
                const float* B, int64_t ldb, int64_t sB, double beta, float* C, int64_t ldc, int64_t sC, int S, int tri,
                int wide, cudaStream_t st);

// Tensor-core path (tcgen05, 3xTF32) for FP32 problems it supports, FMA-pipe kernel otherwise (FP64, transposed A,
// unaligned leading dimensions, tiny problems).
template <typename T>
int gemm_any(int transA, int transB, int m, int n, int k, double alpha, const T* A, int64_t lda, int64_t sA,
             const T* B, int64_t ldb, int64_t sB, double beta, T* C, int64_t ldc, int64_t sC, int S, int tri,
             cudaStream_t st, int wide) {
    if (m == 0 || n == 0 || S == 0) return MXF_OK;
    if (k <= 8 && !(tri & 1) && (int64_t)m * n >= 4096) {
        dim3 grid(cdiv(n, 256), std::min(m, 2048), S);
        gemm_rank_k_kernel<T><<<grid, 256, 0, st>>>(transA, transB, m, n, k, (T)alpha, A, lda, sA, B, ldb, sB, (T)beta, C, ldc, sC);
        return after_launch();
    }
    if (n <= 8 && k >= 32 && !(tri & 1) && (const void*)C != (const void*)A && (const void*)C != (const void*)B)
        return gemm_skinny<T>(transA, transB, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, S, st);
    if constexpr (sizeof(T) == 4) {
        if ((int64_t)m * n >= 64 * 64 && k >= 16) {
            int rc = gemm_tc_f32(transA, transB, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, S, tri, wide,
                                 st);
            if (rc != MXF_ENOTIMPL) return rc;
        }
    }
    return gemm_simt<T>(transA, transB, m, n, k, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, S, tri, st);
}

template int gemm_any<float>(int, int, int, int, int, double, const float*, int64_t, int64_t, const float*, int64_t,
                             int64_t, double, float*, int64_t, int64_t, int, int, cudaStream_t, int);
template int gemm_any<double>(int, int, int, int, int, double, const double*, int64_t, int64_t, const double*, int64_t,
                              int64_t, double, double*, int64_t, int64_t, int, int, cudaStream_t, int);

template int gemm_simt<float>(int, int, int, int, int, double, const float*, int64_t, int64_t, const float*,
                              int64_t, int64_t, double, float*, int64_t, int64_t, int, int, cudaStream_t);
template int gemm_simt<double>(int, int, int, int, int, double, const double*, int64_t, int64_t, const double*,
                               int64_t, int64_t, double, double*, int64_t, int64_t, int, int, cudaStream_t);

}  // namespace mxf

using namespace mxf;

extern "C" int mxf_gemm(int dtype, int transA, int transB, int m, int n, int k, double alpha, const void* A,
                        int64_t lda, int64_t sA, const void* B, int64_t ldb, int64_t sB, double beta, void* C,
                        int64_t ldc, int64_t sC, int S, int tri, void* stream) {
    if (m < 0 || n < 0 || k < 0 || S < 0 || !C) return MXF_EINVAL;
    if (k > 0 && (!A || !B)) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return gemm_any<T>(transA, transB, m, n, k, alpha, (const T*)A, lda, sA,
                                                  (const T*)B, ldb, sB, beta, (T*)C, ldc, sC, S, tri,
                                                  (cudaStream_t)stream, 0));
}
