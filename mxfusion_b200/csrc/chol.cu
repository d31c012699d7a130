// Blocked Cholesky factorisation and triangular solves (replaces mx.nd.linalg.potrf / trsm:
// modules/gp_modules/svgp_regression.py:83-87,92; gp_regression.py:61,66).
//
// potrf: right-looking, block size NB = 64.  Per panel ONE kernel factors the NB x NB diagonal block
// in shared memory (every CTA of the panel does it redundantly -- it is ~90 kFLOP -- so no second
// launch or grid-wide barrier is needed) and then solves its 128-row slice of the panel, one thread
// per row with the row held in registers (X L11^T = A21).  The trailing update A22 -= L21 L21^T is a
// lower-tiles-only GEMM (gemm.cu / gemm_tc.cu).  Row-major lower factor, strict upper triangle
// zeroed (MXNet convention).  A non-positive pivot is reported through info[s] (1-based).
//
// trsm (left, lower, no-transpose / transpose): per block row one kernel solves the NB x NB diagonal
// system for 128 right-hand-side columns per CTA, one thread per column (coalesced along the
// row-major B), followed by a GEMM update of the remaining block rows.
#include "common.cuh"

namespace mxf {

template <typename T>
int gemm_simt(int transA, int transB, int m, int n, int k, double alpha, const T* A, int64_t lda, int64_t sA,
              const T* B, int64_t ldb, int64_t sB, double beta, T* C, int64_t ldc, int64_t sC, int S, int tri,
              cudaStream_t st);

constexpr int CH_NB = 64;
constexpr int CH_THREADS = 128;

// Factor the (padded) NB x NB block held in Ds (row stride NB+1) in place; dg receives the pivots.
// Returns through *bad the 1-based index of the first non-positive pivot (0 if none) -- same in all threads.
template <typename T, int NB>
__device__ __forceinline__ void factor_block_smem(T* Ds, T* dg, int* bad_s) {
    constexpr int LD = NB + 1;
    const int tid = threadIdx.x;
    if (tid == 0) *bad_s = 0;
    for (int j = 0; j < NB; ++j) {
        __syncthreads();
        const T d = Ds[j * LD + j];
        const T piv = Num<T>::sqrt_(d);
        const T inv = T(1) / piv;
        if (tid == 0) {
            dg[j] = piv;
            if (!(d > T(0)) && *bad_s == 0) *bad_s = j + 1;
        }
        for (int i = j + 1 + tid; i < NB; i += CH_THREADS) Ds[i * LD + j] *= inv;
        __syncthreads();
        const int rem = NB - j - 1;
        for (int e = tid; e < rem * rem; e += CH_THREADS) {
            const int i = j + 1 + e / rem, t = j + 1 + e % rem;
            if (t <= i) Ds[i * LD + t] = fma(-Ds[i * LD + j], Ds[t * LD + j], Ds[i * LD + t]);
        }
    }
    __syncthreads();
}

template <typename T, int NB>
__global__ void __launch_bounds__(CH_THREADS)
potrf_panel_kernel(T* __restrict__ A, int64_t lda, int64_t sA, int n, int k0, int* __restrict__ info) {
    constexpr int LD = NB + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Ds = reinterpret_cast<T*>(smem_raw);        // [NB][LD]
    T* dg = Ds + NB * LD;                          // [NB]
    T* Rs = dg + NB;                               // [CH_THREADS][LD]
    __shared__ int bad_s;

    const int s = blockIdx.y;
    T* As = A + (int64_t)s * sA;
    const int nbk = min(NB, n - k0);
    const int tid = threadIdx.x;

    for (int e = tid; e < NB * NB; e += CH_THREADS) {
        const int r = e / NB, c = e % NB;
        T v = (r == c) ? T(1) : T(0);
        if (r < nbk && c < nbk) v = As[(int64_t)(k0 + r) * lda + k0 + c];
        Ds[r * LD + c] = v;
    }
    factor_block_smem<T, NB>(Ds, dg, &bad_s);

    if (blockIdx.x == 0) {
        if (tid == 0 && bad_s != 0 && info) atomicCAS(&info[s], 0, k0 + bad_s);
        // write the factor, zero everything to the right of the diagonal in these rows
        const int width = n - k0;
        for (int e = tid; e < nbk * width; e += CH_THREADS) {
            const int r = e / width, c = e % width;
            T v = T(0);
            if (c < r) v = Ds[r * LD + c];
            else if (c == r) v = dg[r];
            As[(int64_t)(k0 + r) * lda + k0 + c] = v;
        }
    }

    const int rows_below = n - k0 - nbk;
    if (rows_below <= 0) return;          // uniform across the CTA
    const int rbase = k0 + nbk + blockIdx.x * CH_THREADS;
    const int rcnt = min(CH_THREADS, n - rbase);
    if (rcnt <= 0) return;
    for (int e = tid; e < rcnt * NB; e += CH_THREADS) {
        const int r = e / NB, c = e % NB;
        Rs[r * LD + c] = As[(int64_t)(rbase + r) * lda + k0 + c];
    }
    __syncthreads();
    if (tid < rcnt) {
        T x[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) x[c] = Rs[tid * LD + c];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            x[c] = x[c] / dg[c];
#pragma unroll
            for (int t = c + 1; t < NB; ++t) x[t] = fma(-x[c], Ds[t * LD + c], x[t]);
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) Rs[tid * LD + c] = x[c];
    }
    __syncthreads();
    for (int e = tid; e < rcnt * NB; e += CH_THREADS) {
        const int r = e / NB, c = e % NB;
        As[(int64_t)(rbase + r) * lda + k0 + c] = Rs[r * LD + c];
    }
}

template <typename T>
__global__ void zero_info_kernel(int* info, int S) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) info[i] = 0;
}

template <typename T>
static int potrf_impl(T* A, int64_t lda, int64_t sA, int S, int n, int* info, cudaStream_t st) {
    constexpr int NB = CH_NB;
    constexpr int LD = NB + 1;
    const size_t smem = sizeof(T) * ((size_t)NB * LD + NB + (size_t)CH_THREADS * LD);
    auto k = potrf_panel_kernel<T, NB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int launches = 0;
    if (info) { zero_info_kernel<T><<<cdiv(S, 128), 128, 0, st>>>(info, S); ++launches; }
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nbk = std::min(NB, n - k0);
        const int below = n - k0 - nbk;
        dim3 grid(std::max(1, cdiv(below, CH_THREADS)), S);
        k<<<grid, CH_THREADS, smem, st>>>(A, lda, sA, n, k0, info);
        ++launches;
        if (below > 0) {
            T* L21 = A + (int64_t)(k0 + nbk) * lda + k0;
            T* A22 = A + (int64_t)(k0 + nbk) * lda + (k0 + nbk);
            int rc = gemm_simt<T>(0, 1, below, below, nbk, -1.0, L21, lda, sA, L21, lda, sA, 1.0, A22, lda, sA, S, 1, st);
            if (rc != MXF_OK) return rc;
        }
    }
    return after_launch(launches);
}

// ------------------------------------------------------------------------------------------
template <typename T, int NB, bool TRANS>
__global__ void __launch_bounds__(CH_THREADS)
trsm_diag_kernel(const T* __restrict__ A, int64_t lda, int64_t sA, T* __restrict__ B, int64_t ldb,
                 int64_t sB, int n, int nrhs, int k0) {
    constexpr int LD = NB + 1;
    __shared__ T Ds[NB * LD];
    const int s = blockIdx.y;
    const T* As = A + (int64_t)s * sA;
    T* Bs = B + (int64_t)s * sB;
    const int nbk = min(NB, n - k0);
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += CH_THREADS) {
        const int r = e / NB, c = e % NB;
        T v = (r == c) ? T(1) : T(0);
        if (r < nbk && c < nbk && c <= r) v = As[(int64_t)(k0 + r) * lda + k0 + c];
        Ds[r * LD + c] = v;
    }
    __syncthreads();
    const int col = blockIdx.x * CH_THREADS + tid;
    if (col >= nrhs) return;
    T x[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) x[r] = r < nbk ? Bs[(int64_t)(k0 + r) * ldb + col] : T(0);
    if (!TRANS) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            x[c] = x[c] / Ds[c * LD + c];
#pragma unroll
            for (int r = c + 1; r < NB; ++r) x[r] = fma(-Ds[r * LD + c], x[c], x[r]);
        }
    } else {
#pragma unroll
        for (int c = NB - 1; c >= 0; --c) {
            x[c] = x[c] / Ds[c * LD + c];
#pragma unroll
            for (int r = 0; r < c; ++r) x[r] = fma(-Ds[c * LD + r], x[c], x[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r)
        if (r < nbk) Bs[(int64_t)(k0 + r) * ldb + col] = x[r];
}

template <typename T>
__global__ void scale_rows_kernel(T* B, int64_t ldb, int64_t sB, int rows, int cols, T alpha) {
    const int s = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < cols && r < rows) B[(int64_t)s * sB + (int64_t)r * ldb + c] *= alpha;
}

template <typename T>
static int trsm_impl(int transpose, int n, int nrhs, double alpha, const T* A, int64_t lda, int64_t sA, T* B,
                     int64_t ldb, int64_t sB, int S, cudaStream_t st) {
    constexpr int NB = CH_NB;
    if (n == 0 || nrhs == 0 || S == 0) return MXF_OK;
    int launches = 0;
    if (alpha != 1.0) {
        dim3 g(cdiv(nrhs, 256), n, S);
        if (n > 65535) return MXF_ENOTIMPL;
        scale_rows_kernel<T><<<g, 256, 0, st>>>(B, ldb, sB, n, nrhs, (T)alpha);
        ++launches;
    }
    dim3 grid(cdiv(nrhs, CH_THREADS), S);
    if (!transpose) {
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int nbk = std::min(NB, n - k0);
            trsm_diag_kernel<T, NB, false><<<grid, CH_THREADS, 0, st>>>(A, lda, sA, B, ldb, sB, n, nrhs, k0);
            ++launches;
            const int below = n - k0 - nbk;
            if (below > 0) {
                int rc = gemm_simt<T>(0, 0, below, nrhs, nbk, -1.0, A + (int64_t)(k0 + nbk) * lda + k0, lda, sA,
                                      B + (int64_t)k0 * ldb, ldb, sB, 1.0, B + (int64_t)(k0 + nbk) * ldb, ldb, sB, S,
                                      0, st);
                if (rc != MXF_OK) return rc;
            }
        }
    } else {
        const int last = ((n - 1) / NB) * NB;
        for (int k0 = last; k0 >= 0; k0 -= NB) {
            const int nbk = std::min(NB, n - k0);
            trsm_diag_kernel<T, NB, true><<<grid, CH_THREADS, 0, st>>>(A, lda, sA, B, ldb, sB, n, nrhs, k0);
            ++launches;
            if (k0 > 0) {
                // B[0:k0,:] -= L[k0:k0+nbk, 0:k0]^T  B[k0:k0+nbk,:]
                int rc = gemm_simt<T>(1, 0, k0, nrhs, nbk, -1.0, A + (int64_t)k0 * lda, lda, sA,
                                      B + (int64_t)k0 * ldb, ldb, sB, 1.0, B, ldb, sB, S, 0, st);
                if (rc != MXF_OK) return rc;
            }
        }
    }
    return after_launch(launches);
}

}  // namespace mxf

using namespace mxf;

extern "C" int mxf_potrf(int dtype, void* A, int64_t lda, int64_t sA, int S, int n, int* info, void* stream) {
    if (!A || n < 0 || S < 0 || lda < n) return MXF_EINVAL;
    if (n == 0 || S == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return potrf_impl<T>((T*)A, lda, sA, S, n, info, (cudaStream_t)stream));
}

extern "C" int mxf_trsm(int dtype, int transpose, int n, int nrhs, double alpha, const void* A, int64_t lda,
                        int64_t sA, void* B, int64_t ldb, int64_t sB, int S, void* stream) {
    if (!A || !B || n < 0 || nrhs < 0 || S < 0 || lda < n || ldb < nrhs) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return trsm_impl<T>(transpose, n, nrhs, alpha, (const T*)A, lda, sA, (T*)B, ldb, sB,
                                                  S, (cudaStream_t)stream));
}
