// Cholesky factorisation and triangular solves organised around tensor-core GEMMs.
//
// potrf (linalg.potrf, svgp_regression.py:83-84, gp_regression.py:61): blocked right-looking, block NB (128 for
// f32, 64 for f64).  Per block column:
//   1. one CTA factors the NB x NB diagonal block in shared memory (32-wide sub-panels: a warp-register Cholesky
//      of the 32 x 32 pivot block, one thread per row for the sub-panel solve, register-tiled rank-32 update) and
//      also forms  W = L11^-1  (needed right away and kept for every later solve with this factor);
//   2. the panel  L21 = A21 W^T  is a GEMM (no per-row substitution);
//   3. the trailing update  A22 -= L21 L21^T  is a lower-tiles-only GEMM.
// 2 and 3 run on the tcgen05 kernel of gemm_tc.cu for f32.
//
// The by-product of potrf is the "pack" of the factor:  [ Dinv | DinvT | LT ]  = inverses of the NB x NB diagonal
// blocks, their transposes, and L^T.  With it linalg.trsm (svgp_regression.py:85-87,92; gp_regression.py:66) is a
// chain of GEMMs too:  X_k = Dinv_k B_k ;  B_below -= L_below,k X_k   (and the mirrored chain with DinvT / LT for the
// transposed solve).  Inverting only the diagonal blocks (not L) keeps the solve backward-stable in the blocks'
// condition numbers, the standard GPU trsm formulation.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "chol_dag.cuh"

extern "C" int mxf_transpose(int dtype, const void* A, int64_t lda, int64_t sA, void* out, int64_t ldo, int64_t sO,
                             int S, int m, int n, void* stream);

namespace mxf {

template <typename T>
int gemm_any(int transA, int transB, int m, int n, int k, double alpha, const T* A, int64_t lda, int64_t sA,
             const T* B, int64_t ldb, int64_t sB, double beta, T* C, int64_t ldc, int64_t sC, int S, int tri,
             cudaStream_t st, int wide);

template <typename T> struct TriBlock { static constexpr int NB = 128; };
template <> struct TriBlock<double> { static constexpr int NB = 64; };

// Largest inverse block built hierarchically from the NB blocks (power-of-two multiple of NB dividing n).
inline int tri_inv_max() {
    static int v = [] {
        const char* e = getenv("MXF_TRI_INV");
        int x = e ? atoi(e) : 1024;
        return x < 64 ? 64 : x;
    }();
    return v;
}

template <typename T>
struct PackLayout {
    int n, nblk, ldt, top, nlvl;
    int64_t dinv, dinvT, lt, lvl[8], topT, scratch, total;       // element offsets within one sample's pack
    explicit PackLayout(int n_) : n(n_) {
        constexpr int NB = TriBlock<T>::NB;
        nblk = (n + NB - 1) / NB;
        ldt = (n + 3) & ~3;
        dinv = 0;
        dinvT = (int64_t)nblk * NB * NB;
        lt = 2 * dinvT;
        int64_t off = (lt + (int64_t)n * ldt + 3) & ~(int64_t)3;
        top = NB;
        nlvl = 1;
        lvl[0] = dinv;
        topT = dinvT;
        scratch = 0;
        if (n % NB == 0) {
            while (top * 2 <= tri_inv_max() && n % (top * 2) == 0 && nlvl < 8) {
                top *= 2;
                lvl[nlvl++] = off;
                off += (int64_t)n * top;
            }
            if (top > NB) {
                topT = off;
                off += (int64_t)n * top;
                scratch = off;
                off += (int64_t)n * top / 2;
            }
        }
        total = (off + 3) & ~(int64_t)3;
    }
    int nq() const { return (n + top - 1) / top; }
    int64_t wtop() const { return lvl[nlvl - 1]; }      // top-level inverse blocks (top x top each, row stride top)
    int64_t wtopT() const { return topT; }
};

// f32: the factor's explicit inverse comes out of the single-launch tile-dataflow kernel (chol_dag.cu) for any
// n <= 1024, so the "top level" is ONE block W = L^-1 with row stride top = n rounded up to 4; larger factors are cut
// into 1024-blocks (the last one may be ragged).  No intermediate levels, no scratch; DG_SYNC_INTS ints of flags at the end.
template <>
struct PackLayout<float> {
    int n, nblk, ldt, top;
    int64_t dinv, dinvT, lt, wtop_, wtopT_, sync, total;
    explicit PackLayout(int n_) : n(n_) {
        constexpr int NB = 128;
        nblk = (n + NB - 1) / NB;
        ldt = (n + 3) & ~3;
        top = n <= DG_MAXN ? std::max(4, ldt) : DG_MAXN;
        dinv = 0;
        dinvT = (int64_t)nblk * NB * NB;
        lt = 2 * dinvT;
        wtop_ = (lt + (int64_t)n * ldt + 3) & ~(int64_t)3;
        wtopT_ = wtop_ + (int64_t)nq() * top * top;
        sync = wtopT_ + (int64_t)nq() * top * top;
        total = (sync + DG_SYNC_INTS + 3) & ~(int64_t)3;
    }
    int nq() const { return (n + top - 1) / top; }
    int64_t wtop() const { return wtop_; }
    int64_t wtopT() const { return wtopT_; }
};

constexpr int PD_THREADS = 512;
__device__ long long* g_prof = nullptr;          // debug: per-phase clock64 stamps of potrf_diag_kernel (thread 0)
#define PD_STAMP(i) do { if (g_prof && threadIdx.x == 0 && blockIdx.x == 0) g_prof[i] = clock64(); } while (0)

// 32 x 32 Cholesky in the registers of one warp: lane i holds row i (r[0..31]); on exit r[c] (c <= i) is L[i][c].
// Column steps are instantiated by template recursion so that every r[] index is a compile-time constant (a runtime
// index would push the array to local memory and serialise the whole factorisation on L1 latency).
template <typename T> __device__ __forceinline__ T inv_sqrt_(T d);
template <> __device__ __forceinline__ float inv_sqrt_<float>(float d) {
    float y = rsqrtf(d);
    return y * fmaf(-0.5f * d * y, y, 1.5f);          // one Newton step: full fp32 accuracy
}
template <> __device__ __forceinline__ double inv_sqrt_<double>(double d) { return 1.0 / sqrt(d); }

template <typename T, int C>
struct CholStep {
    static __device__ __forceinline__ void run(T (&r)[32], int lane, int& bad) {
        const T d = __shfl_sync(0xffffffffu, r[C], C);
        if (!(d > T(0)) && bad == 0) bad = C + 1;
        const T inv = inv_sqrt_<T>(d);
        // branch-free: lanes above the diagonal carry don't-care values that are never read back
        r[C] = ((lane == C) ? d : r[C]) * inv;
#pragma unroll
        for (int t = C + 1; t < 32; ++t) {
            const T ltc = __shfl_sync(0xffffffffu, r[C], t);
            r[t] = fma(-r[C], ltc, r[t]);
        }
        CholStep<T, C + 1>::run(r, lane, bad);
    }
};
template <typename T>
struct CholStep<T, 32> {
    static __device__ __forceinline__ void run(T (&)[32], int, int&) {}
};

// Returns the 1-based index of the first non-positive pivot (0 if none), identical in all lanes.
template <typename T>
__device__ __forceinline__ int warp_chol32(T (&r)[32], int lane) {
    int bad = 0;
    CholStep<T, 0>::run(r, lane, bad);
    return bad;
}

// Column `lane` of the inverse of the 32 x 32 lower-triangular block at Dblk (row stride LD): forward substitution with
// x[] in registers (template recursion again for static indexing).
template <typename T, int I, int LD>
struct InvStep {
    static __device__ __forceinline__ void run(T (&x)[32], const T* Dblk, const T* idg, int lane) {
        T acc = (I == lane) ? T(1) : T(0);
#pragma unroll
        for (int l = 0; l < I; ++l) acc = fma(-Dblk[I * LD + l], x[l], acc);
        x[I] = acc * idg[I];
        InvStep<T, I + 1, LD>::run(x, Dblk, idg, lane);
    }
};
template <typename T, int LD>
struct InvStep<T, 32, LD> {
    static __device__ __forceinline__ void run(T (&)[32], const T*, const T*, int) {}
};

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
};
template <> struct Vec4<double> {
    static __device__ __forceinline__ void ld(const double* p, double (&v)[4]) {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};

// Off-diagonal blocks of W = L^-1 for an NB x NB lower factor held in D, diagonal 32 x 32 blocks of W already in
// place:  block row bi:  Tmp = L[bi, 0:bi] W[0:bi, 0:bi]  then  W[bi, 0:bi] = -W[bi, bi] Tmp.
// Mapping: a warp owns 4 output rows, its lanes 32 consecutive output columns (B operand conflict-free, A operand a
// broadcast).  256 threads.
template <typename T, int NB>
__device__ __forceinline__ void inverse_offdiag(const T* D, T* W, T* Tmp) {
    constexpr int LD = NB + 1;
    constexpr int NS = NB / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    for (int bi = 1; bi < NS; ++bi) {
        // work items: (row group of 4 rows) x (column group of 32 columns), spread over the warps
        for (int item = warp; item < 8 * bi; item += nwarps) {
            const int i0 = 4 * (item & 7), cg = item >> 3;
            T acc[4] = {T(0), T(0), T(0), T(0)};
            const int col = 32 * cg + lane;
#pragma unroll 8
            for (int l = 32 * cg; l < 32 * bi; ++l) {
                const T b = W[l * LD + col];
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a] = fma(D[(32 * bi + i0 + a) * LD + l], b, acc[a]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) Tmp[(i0 + a) * LD + col] = acc[a];
        }
        __syncthreads();
        for (int item = warp; item < 8 * bi; item += nwarps) {
            const int i0 = 4 * (item & 7), cg = item >> 3;
            T acc[4] = {T(0), T(0), T(0), T(0)};
            const int col = 32 * cg + lane;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const T b = Tmp[l * LD + col];
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a] = fma(W[(32 * bi + i0 + a) * LD + 32 * bi + l], b, acc[a]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) W[(32 * bi + i0 + a) * LD + col] = -acc[a];
        }
        __syncthreads();
    }
}

// Factor the NB x NB diagonal block at (k0, k0) of every sample; write L11 back (zeros to its right within the
// block rows are written by the caller's convention below), W = L11^-1 to dinv[kb], W^T to dinvT[kb].
template <typename T, int NB>
__global__ void __launch_bounds__(PD_THREADS)
potrf_diag_kernel(T* __restrict__ A, int64_t lda, int64_t sA, int n, int k0, int* __restrict__ info,
                  T* __restrict__ pack, int64_t pack_stride, int64_t off_dinv, int64_t off_dinvT) {
    constexpr int LD = NB + 1;
    constexpr int NS = NB / 32;                 // 32-wide sub-panels
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* D = reinterpret_cast<T*>(smem_raw);      // [NB][LD]  block being factored (lower part meaningful)
    T* W = D + NB * LD;                         // [NB][LD]  inverse
    T* idg = W + NB * LD;                       // [NB]      1 / L[i][i]
    T* Pt = idg + NB + 32 * LD + 3;             // [32][NB]  current sub-panel, transposed (k-major for the update)
    Pt = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(Pt) + 31) & ~(uintptr_t)31);
    __shared__ int bad_s;

    pdl_launch_dependents();
    const int s = blockIdx.x;
    T* As = A + (int64_t)s * sA;
    const int nbk = min(NB, n - k0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) bad_s = 0;
    pdl_wait();

    // 16-byte accesses when the block is aligned (lda % 4 == 0, 16-byte aligned base): the block is 64 KB moved by ONE
    // CTA, so the number of load / store round trips -- not bandwidth -- sets the time of this phase
    const bool vec4 = (sizeof(T) == 4) && ((lda & 3) == 0) && ((sA & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (vec4) {
#pragma unroll 8
        for (int e4 = tid; e4 < NB * NB / 4; e4 += PD_THREADS) {
            const int r = e4 / (NB / 4), c = (e4 - r * (NB / 4)) * 4;
            T v[4] = {T(0), T(0), T(0), T(0)};
            if (r < nbk && c <= r) {                       // chunks entirely above the diagonal are never read
                const float4 q = *reinterpret_cast<const float4*>(&As[(int64_t)(k0 + r) * lda + k0 + c]);
                v[0] = (T)q.x; v[1] = (T)q.y; v[2] = (T)q.z; v[3] = (T)q.w;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int cc = c + u;
                T x = (r == cc) ? T(1) : T(0);
                if (r < nbk && cc < nbk && cc <= r) x = v[u];
                D[r * LD + cc] = x;
                W[r * LD + cc] = T(0);
            }
        }
    } else {
#pragma unroll 8
        for (int e = tid; e < NB * NB; e += PD_THREADS) {
            const int r = e / NB, c = e - r * NB;
            T v = (r == c) ? T(1) : T(0);
            if (r < nbk && c < nbk && c <= r) v = As[(int64_t)(k0 + r) * lda + k0 + c];
            D[r * LD + c] = v;
            W[r * LD + c] = T(0);
        }
    }
    __syncthreads();
    PD_STAMP(0);

    for (int j = 0; j < NS; ++j) {
        const int p0 = 32 * j;
        // (a) pivot block on warp 0
        if (warp == 0) {
            T r[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = D[(p0 + lane) * LD + p0 + c];
            const int bad = warp_chol32<T>(r, lane);
#pragma unroll
            for (int c = 0; c < 32; ++c) D[(p0 + lane) * LD + p0 + c] = (c <= lane) ? r[c] : T(0);
            T dself = T(1);
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (c == lane) dself = r[c];       // L[lane][lane]; static indexing keeps r[] in registers
            idg[p0 + lane] = T(1) / dself;
            if (lane == 0 && bad != 0 && bad_s == 0) bad_s = p0 + bad;
        }
        __syncthreads();
        PD_STAMP(1 + 3 * j);
        const int rem = NB - p0 - 32;
        if (rem > 0) {
            // (b) sub-panel: rows below solve  x L_jj^T = a, one thread per row
            if (tid < rem) {
                const int row = p0 + 32 + tid;
                T x[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) x[c] = D[row * LD + p0 + c];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    x[c] *= idg[p0 + c];
#pragma unroll
                    for (int t = c + 1; t < 32; ++t) x[t] = fma(-x[c], D[(p0 + t) * LD + p0 + c], x[t]);
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    D[row * LD + p0 + c] = x[c];
                    Pt[c * NB + row] = x[c];
                }
            }
            __syncthreads();
            PD_STAMP(2 + 3 * j);
            // (c) rank-32 update of the trailing lower triangle, 4 x 4 register tiles
            const int q0 = p0 + 32;
            const int nt = rem / 4;                            // tiles per side; only tiles with tj <= ti
            for (int tile = tid; tile < nt * (nt + 1) / 2; tile += PD_THREADS) {
                int ti = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
                while (ti * (ti + 1) / 2 > tile) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
                const int tj = tile - ti * (ti + 1) / 2;
                const int i0 = q0 + 4 * ti, j0 = q0 + 4 * tj;
                T acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
#pragma unroll 8
                for (int c = 0; c < 32; ++c) {
                    T ai[4], bj[4];
                    Vec4<T>::ld(Pt + c * NB + i0, ai);
                    Vec4<T>::ld(Pt + c * NB + j0, bj);
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = fma(ai[a], bj[b], acc[a][b]);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (j0 + b <= i0 + a) D[(i0 + a) * LD + j0 + b] -= acc[a][b];
            }
            __syncthreads();
            PD_STAMP(3 + 3 * j);
        }
    }
    PD_STAMP(13);

    // ---- W = L^-1 -------------------------------------------------------------------------------------------
    // diagonal 32 x 32 blocks: warp w inverts block w; lane = column of the inverse, forward substitution
    if (warp < NS) {
        const int p0 = 32 * warp;
        T x[32];
        InvStep<T, 0, LD>::run(x, D + p0 * LD + p0, idg + p0, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) W[(p0 + i) * LD + p0 + lane] = x[i];
    }
    __syncthreads();
    PD_STAMP(14);
    inverse_offdiag<T, NB>(D, W, idg + NB);     // scratch: [32][LD]
    PD_STAMP(15);
    // ---- write back ---------------------------------------------------------------------------------------------
    if (tid == 0 && bad_s != 0 && info) atomicCAS(&info[s], 0, k0 + bad_s);
    // factor block (the strict upper triangle of A is zeroed once, after the last block: MXNet potrf convention)
    T* dv = pack + (int64_t)s * pack_stride + off_dinv + (int64_t)(k0 / NB) * NB * NB;
    T* dvT = pack + (int64_t)s * pack_stride + off_dinvT + (int64_t)(k0 / NB) * NB * NB;
#pragma unroll 8
    for (int e = tid; e < NB * NB; e += PD_THREADS) {
        const int r = e / NB, c = e - r * NB;
        if (r < nbk && c < nbk) As[(int64_t)(k0 + r) * lda + k0 + c] = (c <= r) ? D[r * LD + c] : T(0);
    }
    // (16-byte stores were measured and do not help: this phase sits at one SM's ~32 B/clk write path to L2)
#pragma unroll 8
    for (int e = tid; e < NB * NB; e += PD_THREADS) {
        const int r = e / NB, c = e - r * NB;
        dv[e] = W[r * LD + c];
        dvT[e] = W[c * LD + r];
    }
    PD_STAMP(16);
}

template <typename T>
static size_t diag_smem() {
    constexpr int NB = TriBlock<T>::NB;
    return sizeof(T) * ((size_t)2 * NB * (NB + 1) + NB + (size_t)32 * (NB + 1) + 8 + (size_t)32 * NB) + 64;
}

template <typename T>
__global__ void zero_upper_kernel(T* __restrict__ A, int64_t lda, int64_t sA, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.z, r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && c > r) A[(int64_t)s * sA + (int64_t)r * lda + c] = T(0);
}

template <typename T>
__global__ void zero_i32_kernel(int* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// Inverse of the diagonal blocks of an EXISTING lower factor (no factorisation): same kernel structure, reading L.
template <typename T, int NB>
__global__ void __launch_bounds__(PD_THREADS)
tri_diag_inv_kernel(const T* __restrict__ L, int64_t lda, int64_t sA, int n, T* __restrict__ pack, int64_t pack_stride,
                    int64_t off_dinv, int64_t off_dinvT) {
    constexpr int LD = NB + 1;
    constexpr int NS = NB / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* D = reinterpret_cast<T*>(smem_raw);
    T* W = D + NB * LD;
    T* idg = W + NB * LD;
    const int s = blockIdx.y, kb = blockIdx.x, k0 = kb * NB;
    const T* Ls = L + (int64_t)s * sA;
    const int nbk = min(NB, n - k0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < NB * NB; e += PD_THREADS) {
        const int r = e / NB, c = e - r * NB;
        T v = (r == c) ? T(1) : T(0);
        if (r < nbk && c < nbk && c <= r) v = Ls[(int64_t)(k0 + r) * lda + k0 + c];
        D[r * LD + c] = v;
        W[r * LD + c] = T(0);
    }
    __syncthreads();
    for (int i = tid; i < NB; i += PD_THREADS) idg[i] = T(1) / D[i * LD + i];
    __syncthreads();
    if (warp < NS) {
        const int p0 = 32 * warp;
        T x[32];
        InvStep<T, 0, LD>::run(x, D + p0 * LD + p0, idg + p0, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) W[(p0 + i) * LD + p0 + lane] = x[i];
    }
    __syncthreads();
    inverse_offdiag<T, NB>(D, W, idg + NB);
    T* dv = pack + (int64_t)s * pack_stride + off_dinv + (int64_t)kb * NB * NB;
    T* dvT = pack + (int64_t)s * pack_stride + off_dinvT + (int64_t)kb * NB * NB;
#pragma unroll 8
    for (int e = tid; e < NB * NB; e += PD_THREADS) {
        const int r = e / NB, c = e - r * NB;
        dv[e] = W[r * LD + c];
        dvT[e] = W[c * LD + r];
    }
}

template <typename T>
static int dtype_of() { return sizeof(T) == 4 ? MXF_F32 : MXF_F64; }

template <typename T>
static int build_inverse_levels(const T* L, int64_t lda, int64_t sA, int S, int n, T* pack, const PackLayout<T>& pl,
                                cudaStream_t st, int row0 = 0, int nrows = -1);

// f32: every (<= 1024)^2 diagonal block is factored AND inverted by one launch of the tile-dataflow kernel; for n > 1024
// the panel  L21 = A21 W^T  and the trailing update  A22 -= L21 L21^T  are K = 1024 tensor-core GEMMs.
static int potrf_packed_f32(float* A, int64_t lda, int64_t sA, int S, int n, int* info, float* pack, cudaStream_t st) {
    const PackLayout<float> pl(n);
    if (info) {
        cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int) * (size_t)S, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (n <= DG_MAXN)
        return dag_launch(1, A, lda, sA, n, pack, pl.total, pl.wtop(), pl.wtopT(), pl.lt, pl.dinv, pl.dinvT, pl.sync, pl.top,
                          pl.ldt, info, 0, S, st);
    const int OB = pl.top;
    for (int o0 = 0; o0 < n; o0 += OB) {
        const int nb = std::min(OB, n - o0);
        const int64_t qoff = (int64_t)(o0 / OB) * OB * OB, doff = (int64_t)(o0 / 128) * 128 * 128;
        int rc = dag_launch(1, A + (int64_t)o0 * lda + o0, lda, sA, nb, pack, pl.total, pl.wtop() + qoff, pl.wtopT() + qoff, -1,
                            pl.dinv + doff, pl.dinvT + doff, pl.sync, OB, 0, info, o0, S, st);
        if (rc != MXF_OK) return rc;
        const int below = n - o0 - nb;
        if (below > 0) {
            float* A21 = A + (int64_t)(o0 + nb) * lda + o0;
            float* A22 = A + (int64_t)(o0 + nb) * lda + (o0 + nb);
            const float* Wo = pack + pl.wtop() + qoff;
            float* scr = pack + pl.lt;                      // L^T is written last: its space is free until then
            rc = gemm_any<float>(0, 1, below, OB, OB, 1.0, A21, lda, sA, Wo, OB, pl.total, 0.0, scr, OB, pl.total, S, 0, st, 0);
            if (rc != MXF_OK) return rc;
            rc = gemm_any<float>(0, 1, below, below, OB, -1.0, scr, OB, pl.total, scr, OB, pl.total, 1.0, A22, lda, sA, S, 1, st, 0);
            if (rc != MXF_OK) return rc;
            for (int s = 0; s < S; ++s)
                cudaMemcpy2DAsync(A21 + (int64_t)s * sA, (size_t)lda * sizeof(float), scr + (int64_t)s * pl.total,
                                  (size_t)OB * sizeof(float), (size_t)OB * sizeof(float), below, cudaMemcpyDeviceToDevice, st);
        }
    }
    if (n > 65535) return MXF_ENOTIMPL;
    dim3 g(cdiv(n, 256), n, S);
    launch_pdl(zero_upper_kernel<float>, g, dim3(256), (size_t)0, st, A, lda, sA, n);
    int rc = mxf_transpose(MXF_F32, A, lda, sA, pack + pl.lt, pl.ldt, pl.total, S, n, n, st);
    if (rc != MXF_OK) return rc;
    return after_launch(1);
}

static int tri_pack_f32(const float* L, int64_t lda, int64_t sA, int S, int n, float* pack, cudaStream_t st) {
    const PackLayout<float> pl(n);
    const int OB = pl.top;
    for (int o0 = 0; o0 < n; o0 += OB) {
        const int nb = std::min(OB, n - o0);
        const int64_t qoff = (int64_t)(o0 / OB) * OB * OB, doff = (int64_t)(o0 / 128) * 128 * 128;
        int rc = dag_launch(0, const_cast<float*>(L) + (int64_t)o0 * lda + o0, lda, sA, nb, pack, pl.total, pl.wtop() + qoff,
                            pl.wtopT() + qoff, -1, pl.dinv + doff, pl.dinvT + doff, pl.sync, OB, 0, nullptr, 0, S, st);
        if (rc != MXF_OK) return rc;
    }
    return mxf_transpose(MXF_F32, L, lda, sA, pack + pl.lt, pl.ldt, pl.total, S, n, n, st);
}

template <typename T>
static int potrf_packed_impl(T* A, int64_t lda, int64_t sA, int S, int n, int* info, T* pack, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        return potrf_packed_f32(A, lda, sA, S, n, info, pack, st);
    } else {
    constexpr int NB = TriBlock<T>::NB;
    const PackLayout<T> pl(n);
    const size_t smem = diag_smem<T>();
    auto k = potrf_diag_kernel<T, NB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int launches = 0;
    if (info) { zero_i32_kernel<T><<<cdiv(S, 128), 128, 0, st>>>(info, S); ++launches; }
    // Two-level blocking for large matrices: factor an OB x OB diagonal block with the NB-step algorithm, invert it
    // hierarchically (128 -> 256 -> 512), then the panel and the trailing update are GEMMs with K = OB (deep K loops keep
    // the tensor pipe busy; with K = NB the tile prologue / epilogue dominates).
    const int OB = (n > 1024 && pl.top >= 512 && n % pl.top == 0) ? pl.top : n;
    for (int o0 = 0; o0 < n; o0 += OB) {
        const int oend = std::min(n, o0 + OB);
        for (int k0 = o0; k0 < oend; k0 += NB) {
            const int nbk = std::min(NB, n - k0);
            cudaError_t le = launch_pdl(k, dim3(S), dim3(PD_THREADS), smem, st, A, lda, sA, n, k0, info, pack, pl.total, pl.dinv,
                                        pl.dinvT);
            if (le != cudaSuccess) return (int)le;
            ++launches;
            const int below = oend - k0 - nbk;
            if (below > 0) {
                T* A21 = A + (int64_t)(k0 + nbk) * lda + k0;
                T* A22 = A + (int64_t)(k0 + nbk) * lda + (k0 + nbk);
                const T* Wk = pack + pl.dinv + (int64_t)(k0 / NB) * NB * NB;
                // L21 = A21 W^T (in place; one CTA owns full rows)
                int rc = gemm_any<T>(0, 1, below, NB, NB, 1.0, A21, lda, sA, Wk, NB, pl.total, 0.0, A21, lda, sA, S, 0, st, 1);
                if (rc != MXF_OK) return rc;
                rc = gemm_any<T>(0, 1, below, below, NB, -1.0, A21, lda, sA, A21, lda, sA, 1.0, A22, lda, sA, S, 1, st, 0);
                if (rc != MXF_OK) return rc;
            }
        }
        if (OB < n) {
            int rc = build_inverse_levels<T>(A, lda, sA, S, n, pack, pl, st, o0, OB);
            if (rc != MXF_OK) return rc;
            const int below = n - oend;
            if (below > 0) {
                T* A21 = A + (int64_t)oend * lda + o0;
                T* A22 = A + (int64_t)oend * lda + oend;
                const T* Wo = pack + pl.lvl[pl.nlvl - 1] + (int64_t)(o0 / OB) * OB * OB;
                T* scr = pack + pl.lt;                      // L^T is written last: its space is free until then
                // L21 = A21 Wo^T, out of place (several column tiles share the rows of A21)
                rc = gemm_any<T>(0, 1, below, OB, OB, 1.0, A21, lda, sA, Wo, OB, pl.total, 0.0, scr, OB, pl.total, S, 0, st, 0);
                if (rc != MXF_OK) return rc;
                rc = gemm_any<T>(0, 1, below, below, OB, -1.0, scr, OB, pl.total, scr, OB, pl.total, 1.0, A22, lda, sA, S, 1, st, 0);
                if (rc != MXF_OK) return rc;
                for (int s = 0; s < S; ++s)
                    cudaMemcpy2DAsync(A21 + (int64_t)s * sA, (size_t)lda * sizeof(T), scr + (int64_t)s * pl.total,
                                      (size_t)OB * sizeof(T), (size_t)OB * sizeof(T), below, cudaMemcpyDeviceToDevice, st);
            }
        }
    }
    if (n > 1) {
        if (n > 65535) return MXF_ENOTIMPL;
        dim3 g(cdiv(n, 256), n, S);
        launch_pdl(zero_upper_kernel<T>, g, dim3(256), (size_t)0, st, A, lda, sA, n);
        ++launches;
    }
    int rc = mxf_transpose(dtype_of<T>(), A, lda, sA, pack + pl.lt, pl.ldt, pl.total, S, n, n, st);
    if (rc != MXF_OK) return rc;
    if (OB == n) {
        rc = build_inverse_levels<T>(A, lda, sA, S, n, pack, pl, st);
        if (rc != MXF_OK) return rc;
    }
    return after_launch(launches);
    }
}

template <typename T>
static int tri_pack_impl(const T* L, int64_t lda, int64_t sA, int S, int n, T* pack, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        return tri_pack_f32(L, lda, sA, S, n, pack, st);
    } else {
    constexpr int NB = TriBlock<T>::NB;
    const PackLayout<T> pl(n);
    const size_t smem = diag_smem<T>();
    auto k = tri_diag_inv_kernel<T, NB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(pl.nblk, S);
    k<<<grid, PD_THREADS, smem, st>>>(L, lda, sA, n, pack, pl.total, pl.dinv, pl.dinvT);
    int rc = mxf_transpose(dtype_of<T>(), L, lda, sA, pack + pl.lt, pl.ldt, pl.total, S, n, n, st);
    if (rc != MXF_OK) return rc;
    rc = build_inverse_levels<T>(L, lda, sA, S, n, pack, pl, st);
    if (rc != MXF_OK) return rc;
    return after_launch(1);
    }
}

// ---- hierarchical inverse blocks -------------------------------------------------------------------------------
// [[Wa, 0], [-Wc Lca Wa, Wc]] from the inverses Wa, Wc of two consecutive b x b diagonal blocks: copies + zero fill
// (the product block is written by the GEMMs).
template <typename T>
__global__ void __launch_bounds__(256)
inv_level_assemble_kernel(const T* __restrict__ src, T* __restrict__ dst, int b) {
    pdl_launch_dependents();
    pdl_wait();
    const int q = blockIdx.y;
    const T* wa = src + (int64_t)(2 * q) * b * b;
    const T* wc = wa + (int64_t)b * b;
    T* d = dst + (int64_t)q * 4 * b * b;
    const int64_t tot = (int64_t)b * b;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / b), c = (int)(e - (int64_t)r * b);
        d[(int64_t)r * 2 * b + c] = wa[e];
        d[(int64_t)r * 2 * b + b + c] = T(0);
        d[(int64_t)(b + r) * 2 * b + b + c] = wc[e];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
transpose_blocks_kernel(const T* __restrict__ src, T* __restrict__ dst, int b) {
    __shared__ T tile[32][33];
    pdl_launch_dependents();
    pdl_wait();
    const T* sp = src + (int64_t)blockIdx.z * b * b;
    T* dp = dst + (int64_t)blockIdx.z * b * b;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) tile[r][threadIdx.x] = sp[(int64_t)(by + r) * b + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) dp[(int64_t)(bx + r) * b + by + threadIdx.x] = tile[threadIdx.x][r];
}

// Builds the levels for the diagonal blocks inside rows/cols [row0, row0 + nrows) (default: the whole matrix); row0 and
// nrows are multiples of pl.top.
template <typename T>
static int build_inverse_levels(const T* L, int64_t lda, int64_t sA, int S, int n, T* pack, const PackLayout<T>& pl,
                                cudaStream_t st, int row0, int nrows) {
    if constexpr (sizeof(T) == 4) {
        return MXF_OK;
    } else {
    constexpr int NB = TriBlock<T>::NB;
    if (pl.top == NB) return MXF_OK;
    if (nrows < 0) nrows = n;
    int launches = 0;
    for (int s = 0; s < S; ++s) {
        const T* Ls = L + (int64_t)s * sA + (int64_t)row0 * lda + row0;
        T* pk = pack + (int64_t)s * pl.total;
        int b = NB;
        for (int j = 0; j + 1 < pl.nlvl; ++j, b *= 2) {
            const T* src = pk + pl.lvl[j] + (int64_t)(row0 / b) * b * b;
            T* dst = pk + pl.lvl[j + 1] + (int64_t)(row0 / (2 * b)) * 4 * b * b;
            T* t1 = pk + pl.scratch + (int64_t)(row0 / (2 * b)) * b * b;
            const int np = nrows / (2 * b);
            // T1_q = L[(2q+1)b.., 2qb..] * W_{2q}
            int rc = gemm_any<T>(0, 0, b, b, b, 1.0, Ls + (int64_t)b * lda, lda, (int64_t)2 * b * lda + 2 * b, src, b,
                                 (int64_t)2 * b * b, 0.0, t1, b, (int64_t)b * b, np, 0, st, 0);
            if (rc != MXF_OK) return rc;
            // dst_q[b.., 0..b) = -W_{2q+1} * T1_q      (W lower triangular)
            rc = gemm_any<T>(0, 0, b, b, b, -1.0, src + (int64_t)b * b, b, (int64_t)2 * b * b, t1, b, (int64_t)b * b, 0.0,
                             dst + (int64_t)b * 2 * b, 2 * b, (int64_t)4 * b * b, np, 2, st, 0);
            if (rc != MXF_OK) return rc;
            dim3 g(std::min(64, cdiv((int64_t)b * b, 256)), np);
            launch_pdl(inv_level_assemble_kernel<T>, g, dim3(256), (size_t)0, st, src, dst, b);
            ++launches;
        }
        dim3 gt(pl.top / 32, pl.top / 32, nrows / pl.top);
        const int64_t toff = (int64_t)(row0 / pl.top) * pl.top * pl.top;
        launch_pdl(transpose_blocks_kernel<T>, gt, dim3(32, 8), (size_t)0, st, (const T*)(pk + pl.lvl[pl.nlvl - 1] + toff),
                   pk + pl.topT + toff, pl.top);
        ++launches;
    }
    return after_launch(launches);
    }
}

// Out-of-place solve with the top-level inverse blocks: X = op(L)^-1 B in (n / top) block steps, each one or two large
// GEMMs.  B is used as scratch (its later block rows receive the updates); requires pl.top > NB.
template <typename T>
static int trsm_packed_oop_impl(int transpose, int n, int nrhs, const T* L, int64_t lda, int64_t sA, const T* pack,
                                int64_t sP, T* B, int64_t ldb, int64_t sB, T* X, int64_t ldx, int64_t sX, int S,
                                cudaStream_t st) {
    const PackLayout<T> pl(n);
    const int top = pl.top;
    if (!transpose) {
        for (int k0 = 0; k0 < n; k0 += top) {
            const int kb = std::min(top, n - k0);
            const T* Wk = pack + pl.wtop() + (int64_t)(k0 / top) * top * top;
            int rc = gemm_any<T>(0, 0, kb, nrhs, kb, 1.0, Wk, top, sP, B + (int64_t)k0 * ldb, ldb, sB, 0.0,
                                 X + (int64_t)k0 * ldx, ldx, sX, S, 2, st, 0);
            if (rc != MXF_OK) return rc;
            const int below = n - k0 - kb;
            if (below > 0) {
                rc = gemm_any<T>(0, 0, below, nrhs, kb, -1.0, L + (int64_t)(k0 + kb) * lda + k0, lda, sA,
                                 X + (int64_t)k0 * ldx, ldx, sX, 1.0, B + (int64_t)(k0 + kb) * ldb, ldb, sB, S, 0, st, 0);
                if (rc != MXF_OK) return rc;
            }
        }
    } else {
        const T* LT = pack + pl.lt;
        for (int k0 = ((n - 1) / top) * top; k0 >= 0; k0 -= top) {
            const int kb = std::min(top, n - k0);
            const T* Wk = pack + pl.wtopT() + (int64_t)(k0 / top) * top * top;
            int rc = gemm_any<T>(0, 0, kb, nrhs, kb, 1.0, Wk, top, sP, B + (int64_t)k0 * ldb, ldb, sB, 0.0,
                                 X + (int64_t)k0 * ldx, ldx, sX, S, 4, st, 0);
            if (rc != MXF_OK) return rc;
            if (k0 > 0) {
                rc = gemm_any<T>(0, 0, k0, nrhs, kb, -1.0, LT + k0, pl.ldt, sP, X + (int64_t)k0 * ldx, ldx, sX, 1.0, B, ldb,
                                 sB, S, 0, st, 0);
                if (rc != MXF_OK) return rc;
            }
        }
    }
    return MXF_OK;
}

template <typename T>
__global__ void scale_mat_kernel(T* B, int64_t ldb, int64_t sB, int rows, int cols, T alpha) {
    const int s = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        if (c < cols) B[(int64_t)s * sB + (int64_t)r * ldb + c] *= alpha;
}

// Few right-hand sides (nrhs <= 8): the whole solve in ONE CTA per sample -- x_k = Dinv_k b_k by one warp per row,
// then b_rest -= L_rest,k x_k (or LT for the transposed solve) with one warp per row and coalesced reads of the factor.
constexpr int TV_THREADS = 512;
template <typename T, int NB>
__global__ void __launch_bounds__(TV_THREADS)
trsv_packed_kernel(int transpose, int n, int nrhs, T alpha, const T* __restrict__ L, int64_t lda, int64_t sA,
                   const T* __restrict__ pack, int64_t sP, int64_t off_dinv, int64_t off_dinvT, int64_t off_lt, int ldt,
                   T* __restrict__ B, int64_t ldb, int64_t sB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x = reinterpret_cast<T*>(smem_raw);             // [n][8]  running right-hand side / solution
    T* xk = x + (size_t)n * 8;                          // [NB][8] solution of the current block
    const int s = blockIdx.x;
    const T* Ls = L + (int64_t)s * sA;
    const T* pk = pack + (int64_t)s * sP;
    T* Bs = B + (int64_t)s * sB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = TV_THREADS / 32;
    for (int e = tid; e < n * 8; e += TV_THREADS) {
        const int r = e >> 3, j = e & 7;
        x[e] = (j < nrhs) ? alpha * Bs[(int64_t)r * ldb + j] : T(0);
    }
    __syncthreads();
    const int nblk = (n + NB - 1) / NB;
    for (int step = 0; step < nblk; ++step) {
        const int kb = transpose ? (nblk - 1 - step) : step;
        const int k0 = kb * NB, nbk = min(NB, n - k0);
        const T* Dk = pk + (transpose ? off_dinvT : off_dinv) + (int64_t)kb * NB * NB;
        // x_k = Dk[0:nbk, 0:nbk] * x[k0:k0+nbk]
        for (int r = warp; r < nbk; r += NW) {
            T acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = T(0);
            for (int c = lane; c < nbk; c += 32) {
                const T d = Dk[r * NB + c];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fma(d, x[(k0 + c) * 8 + j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
            if (lane < 8) {
                T v = acc[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) v = (lane == j) ? acc[j] : v;
                xk[r * 8 + lane] = v;
            }
        }
        __syncthreads();
        for (int e = tid; e < nbk * 8; e += TV_THREADS) x[k0 * 8 + e] = xk[e];
        // rest -= F[rest, k0:k0+nbk] * x_k,  F = L (rows below) or LT (rows above)
        const int r_lo = transpose ? 0 : k0 + nbk, r_hi = transpose ? k0 : n;
        for (int r = r_lo + warp; r < r_hi; r += NW) {
            const T* frow = transpose ? (pk + off_lt + (int64_t)r * ldt + k0) : (Ls + (int64_t)r * lda + k0);
            T acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = T(0);
            for (int c = lane; c < nbk; c += 32) {
                const T f = frow[c];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fma(f, xk[c * 8 + j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
            if (lane < 8) {
                T v = acc[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) v = (lane == j) ? acc[j] : v;
                x[r * 8 + lane] -= v;
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < n * 8; e += TV_THREADS) {
        const int r = e >> 3, j = e & 7;
        if (j < nrhs) Bs[(int64_t)r * ldb + j] = x[e];
    }
}

template <typename T>
static int trsm_packed_impl(int transpose, int n, int nrhs, double alpha, const T* L, int64_t lda, int64_t sA,
                            const T* pack, int64_t sP, T* B, int64_t ldb, int64_t sB, int S, cudaStream_t st) {
    constexpr int NB = TriBlock<T>::NB;
    if (n == 0 || nrhs == 0 || S == 0) return MXF_OK;
    const PackLayout<T> pl(n);
    if (nrhs <= 8) {
        const size_t smem = sizeof(T) * ((size_t)n * 8 + (size_t)NB * 8);
        if (smem <= 200 * 1024) {
            auto kern = trsv_packed_kernel<T, NB>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<S, TV_THREADS, smem, st>>>(transpose, n, nrhs, (T)alpha, L, lda, sA, pack, sP, pl.dinv, pl.dinvT, pl.lt,
                                              pl.ldt, B, ldb, sB);
            return after_launch();
        }
    }
    if (alpha != 1.0) {
        dim3 g(cdiv(nrhs, 256), std::min(n, 1024), S);
        scale_mat_kernel<T><<<g, 256, 0, st>>>(B, ldb, sB, n, nrhs, (T)alpha);
        after_launch();
    }
    if (!transpose) {
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int nbk = std::min(NB, n - k0);
            const T* Dk = pack + pl.dinv + (int64_t)(k0 / NB) * NB * NB;
            T* Bk = B + (int64_t)k0 * ldb;
            int rc = gemm_any<T>(0, 0, nbk, nrhs, nbk, 1.0, Dk, NB, sP, Bk, ldb, sB, 0.0, Bk, ldb, sB, S, 0, st, 0);
            if (rc != MXF_OK) return rc;
            const int below = n - k0 - nbk;
            if (below > 0) {
                rc = gemm_any<T>(0, 0, below, nrhs, nbk, -1.0, L + (int64_t)(k0 + nbk) * lda + k0, lda, sA, Bk, ldb, sB, 1.0,
                                 B + (int64_t)(k0 + nbk) * ldb, ldb, sB, S, 0, st, 0);
                if (rc != MXF_OK) return rc;
            }
        }
    } else {
        const int last = ((n - 1) / NB) * NB;
        const T* LT = pack + pl.lt;
        for (int k0 = last; k0 >= 0; k0 -= NB) {
            const int nbk = std::min(NB, n - k0);
            const T* Dk = pack + pl.dinvT + (int64_t)(k0 / NB) * NB * NB;
            T* Bk = B + (int64_t)k0 * ldb;
            int rc = gemm_any<T>(0, 0, nbk, nrhs, nbk, 1.0, Dk, NB, sP, Bk, ldb, sB, 0.0, Bk, ldb, sB, S, 0, st, 0);
            if (rc != MXF_OK) return rc;
            if (k0 > 0) {
                // B[0:k0] -= (L[k0:k0+nbk, 0:k0])^T X_k = LT[0:k0, k0:k0+nbk] X_k
                rc = gemm_any<T>(0, 0, k0, nrhs, nbk, -1.0, LT + k0, pl.ldt, sP, Bk, ldb, sB, 1.0, B, ldb, sB, S, 0, st, 0);
                if (rc != MXF_OK) return rc;
            }
        }
    }
    return MXF_OK;
}

}  // namespace mxf

using namespace mxf;

extern "C" int mxf_debug_set_prof(void* dev_ptr) {
    long long* p = (long long*)dev_ptr;
    return (int)cudaMemcpyToSymbol(g_prof, &p, sizeof(p));
}

extern "C" int mxf_debug_set_dag_prof(void* dev_ptr) { return dag_set_prof((long long*)dev_ptr); }

extern "C" int mxf_potrf_dag_ctas(int ctas) {
    const int old = mxf::g_dag_ctas.load();
    if (ctas > 0) mxf::g_dag_ctas.store(ctas);
    return old;
}

extern "C" int mxf_tri_block(int dtype) { return dtype == MXF_F64 ? TriBlock<double>::NB : TriBlock<float>::NB; }

extern "C" size_t mxf_tri_pack_elems(int dtype, int n) {
    if (n <= 0) return 0;
    return dtype == MXF_F64 ? (size_t)PackLayout<double>(n).total : (size_t)PackLayout<float>(n).total;
}

extern "C" int mxf_tri_pack(int dtype, const void* L, int64_t lda, int64_t sA, int S, int n, void* pack, void* stream) {
    if (!L || !pack || n < 0 || S < 0 || lda < n) return MXF_EINVAL;
    if (n == 0 || S == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return tri_pack_impl<T>((const T*)L, lda, sA, S, n, (T*)pack, (cudaStream_t)stream));
}

extern "C" int mxf_potrf_packed(int dtype, void* A, int64_t lda, int64_t sA, int S, int n, int* info, void* pack,
                                void* stream) {
    if (!A || !pack || n < 0 || S < 0 || lda < n) return MXF_EINVAL;
    if (n == 0 || S == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return potrf_packed_impl<T>((T*)A, lda, sA, S, n, info, (T*)pack, (cudaStream_t)stream));
}

extern "C" int mxf_trsm_packed(int dtype, int transpose, int n, int nrhs, double alpha, const void* L, int64_t lda,
                               int64_t sA, const void* pack, int64_t sP, void* B, int64_t ldb, int64_t sB, int S,
                               void* stream) {
    if (!L || !pack || !B || n < 0 || nrhs < 0 || S < 0 || lda < n || ldb < nrhs) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return trsm_packed_impl<T>(transpose, n, nrhs, alpha, (const T*)L, lda, sA, (const T*)pack,
                                                        sP, (T*)B, ldb, sB, S, (cudaStream_t)stream));
}

extern "C" int mxf_tri_pack_layout(int dtype, int n, int64_t* out) {
    if (!out || n <= 0) return MXF_EINVAL;
    if (dtype == MXF_F64) {
        const PackLayout<double> pl(n);
        out[0] = pl.dinv; out[1] = pl.dinvT; out[2] = pl.lt; out[3] = pl.wtop(); out[4] = pl.wtopT(); out[5] = pl.top;
        out[6] = pl.ldt; out[7] = pl.nq();
    } else if (dtype == MXF_F32) {
        const PackLayout<float> pl(n);
        out[0] = pl.dinv; out[1] = pl.dinvT; out[2] = pl.lt; out[3] = pl.wtop(); out[4] = pl.wtopT(); out[5] = pl.top;
        out[6] = pl.ldt; out[7] = pl.nq();
    } else {
        return MXF_EDTYPE;
    }
    return MXF_OK;
}

extern "C" int mxf_tri_top_block(int dtype, int n) {
    if (n <= 0) return 0;
    return dtype == MXF_F64 ? PackLayout<double>(n).top : PackLayout<float>(n).top;
}

extern "C" int mxf_trsm_packed_oop(int dtype, int transpose, int n, int nrhs, const void* L, int64_t lda, int64_t sA,
                                   const void* pack, int64_t sP, void* B, int64_t ldb, int64_t sB, void* X, int64_t ldx,
                                   int64_t sX, int S, void* stream) {
    if (!L || !pack || !B || !X || n <= 0 || nrhs <= 0 || S <= 0 || lda < n || ldb < nrhs || ldx < nrhs) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    if (mxf_tri_top_block(dtype, n) <= mxf_tri_block(dtype)) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return trsm_packed_oop_impl<T>(transpose, n, nrhs, (const T*)L, lda, sA, (const T*)pack, sP,
                                                            (T*)B, ldb, sB, (T*)X, ldx, sX, S, (cudaStream_t)stream));
}
