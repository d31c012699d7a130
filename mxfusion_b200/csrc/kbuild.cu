// Covariance-matrix build for stationary kernels (RBF, Matern 1/2, 3/2, 5/2), forward and adjoint.
//
// Replaces, in one pass over the (S,N,N2) output, the reference's chain of MXNet operators
//   X/l -> gemm2|syrk * -2 -> +|a|^2 -> +|b|^2 -> [clip, sqrt] -> exp -> * variance
// (mxfusion/components/distributions/gp/kernels/stationary.py:90-107, rbf.py:71-72,
//  matern.py:84-88,116-120,148-151).  The same expanded form r2 = |a|^2+|b|^2-2a.b is used so the
// numerics follow the reference (r2 may be slightly negative; Matern clips it at 1e-14 for the
// square root only, and Matern-5/2 keeps the UNCLIPPED r2 in its 5/3 r2 term, matern.py:84-87).
//
// Forward kernel: HBM-write-bound.  Each thread owns 4 consecutive output columns (one 16-byte
// store per row for f32), keeps the scaled b-vectors of those columns in registers and walks the
// rows of the CTA tile, whose scaled a-vectors (pre-multiplied by -2) sit in shared memory and are
// read as warp-uniform broadcasts.  A warp therefore writes 512 contiguous bytes per row.
#include <cstdlib>
#include "common.cuh"

namespace mxf {

// 1 / sqrt(x) for x >= 1e-14 (never denormal here): one MUFU.RSQ, without rsqrtf's denormal fix-up instructions
template <typename T> __device__ __forceinline__ T rsqrt_pos(T x) { return Num<T>::rsqrt_(x); }
template <> __device__ __forceinline__ float rsqrt_pos<float>(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T, int KIND>
__device__ __forceinline__ T kern_value(T r2, T var) {
    if (KIND == MXF_KERN_RBF) {
        return var * Num<T>::exp_(T(-0.5) * r2);
    } else {
        T r2c = r2 > T(1e-14) ? r2 : T(1e-14);
        T R = r2c * rsqrt_pos<T>(r2c);
        if (KIND == MXF_KERN_MATERN52) {
            const T s5 = T(2.23606797749978969641);
            T poly = fma(s5, R, fma(T(5.0 / 3.0), r2, T(1)));
            return var * poly * Num<T>::exp_(-s5 * R);
        } else if (KIND == MXF_KERN_MATERN32) {
            const T s3 = T(1.73205080756887729353);
            return var * fma(s3, R, T(1)) * Num<T>::exp_(-s3 * R);
        } else {
            return var * Num<T>::exp_(-R);
        }
    }
}

// dK/d(r2) following what autograd does on the reference's expression graph:
// the clip has gradient 1 inside [1e-14, inf) and 0 below; Matern-5/2's 5/3*r2 term is
// differentiated directly.  Also returns K/var through *k_over_var.
template <typename T, int KIND>
__device__ __forceinline__ T kern_dr2(T r2, T var, T* k_over_var) {
    if (KIND == MXF_KERN_RBF) {
        T e = Num<T>::exp_(T(-0.5) * r2);
        *k_over_var = e;
        return T(-0.5) * var * e;
    } else {
        const bool inside = r2 >= T(1e-14);
        T r2c = inside ? r2 : T(1e-14);
        T R = r2c * Num<T>::rsqrt_(r2c);
        T dRdr2 = inside ? T(0.5) / R : T(0);
        if (KIND == MXF_KERN_MATERN52) {
            const T s5 = T(2.23606797749978969641);
            T e = Num<T>::exp_(-s5 * R);
            T poly = fma(s5, R, fma(T(5.0 / 3.0), r2, T(1)));
            *k_over_var = poly * e;
            // d/dR [(1+s5 R + 5/3 r2) e^{-s5 R}] with r2 held fixed, plus d/dr2 of the 5/3 r2 term
            T dR = (s5 - s5 * poly) * e;
            return var * (dR * dRdr2 + T(5.0 / 3.0) * e);
        } else if (KIND == MXF_KERN_MATERN32) {
            const T s3 = T(1.73205080756887729353);
            T e = Num<T>::exp_(-s3 * R);
            T poly = fma(s3, R, T(1));
            *k_over_var = poly * e;
            T dR = (s3 - s3 * poly) * e;
            return var * dR * dRdr2;
        } else {
            T e = Num<T>::exp_(-R);
            *k_over_var = e;
            return -var * e * dRdr2;
        }
    }
}

// Matern value with the variance folded into the exponent (fp32 streaming kernel: var * exp(-a R) = 2^(log2 var - a log2e R),
// one FFMA + one MUFU.EX2 instead of two multiplies + FMUL + MUFU); l2v = log2(var) is computed once per thread.
template <int KIND>
__device__ __forceinline__ float matern_value_l2(float r2, float l2v) {
    const float r2c = r2 > 1e-14f ? r2 : 1e-14f;
    float R, y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(R) : "f"(r2c));       // one MUFU.SQRT (r2c >= 1e-14: never denormal)
    if (KIND == MXF_KERN_MATERN52) {
        const float poly = fmaf(2.2360679774997896f, R, fmaf(5.0f / 3.0f, r2, 1.0f));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(-2.2360679774997896f * 1.4426950408889634f, R, l2v)));
        return poly * y;
    } else if (KIND == MXF_KERN_MATERN32) {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(-1.7320508075688772f * 1.4426950408889634f, R, l2v)));
        return fmaf(1.7320508075688772f, R, 1.0f) * y;
    } else {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(-1.4426950408889634f, R, l2v)));
        return y;
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
constexpr int KB_THREADS = 256;   // 8 warps: 2 column groups x 4 row groups
constexpr int KB_TN = 256;        // columns per CTA tile (2 warps x 32 lanes x 4)

template <typename T, int KIND, int DC, int RM, bool SYM>
__global__ void __launch_bounds__(KB_THREADS)
kbuild_fwd_kernel(const T* __restrict__ X, const T* __restrict__ X2, const T* __restrict__ ls,
                  int ls_len, const T* __restrict__ var, const T* __restrict__ diag_add,
                  T diag_const, T* __restrict__ out, int64_t ldo, int N, int N2, int D, int Dpad,
                  int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut,
                  int vec_ok) {
    constexpr int TM = 4 * RM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* a_s = reinterpret_cast<T*>(smem_raw);      // [TM][Dpad]  (-2 * x / l)
    T* na_s = a_s + TM * Dpad;                    // [TM]        |x/l|^2
    T* il_s = na_s + TM;                          // [Dpad]      1 / l  (0 in the padding)

    const int s = blockIdx.z;
    const T* Xs = X + (int64_t)s * sX;
    const T* X2s = X2 + (int64_t)s * sX2;
    const T* lss = ls + (int64_t)s * sLs;
    const T v = var[(int64_t)s * sVar];
    T* outs = out + (int64_t)s * sOut;
    const int i0 = blockIdx.y * TM;
    const int jt = blockIdx.x * KB_TN;

    for (int d = threadIdx.x; d < Dpad; d += KB_THREADS)
        il_s[d] = d < D ? T(1) / lss[ls_len == 1 ? 0 : d] : T(0);
    __syncthreads();
    for (int e = threadIdx.x; e < TM * Dpad; e += KB_THREADS) {
        int r = e / Dpad, d = e - r * Dpad;
        int i = i0 + r;
        T x = (i < N && d < D) ? Xs[(int64_t)i * D + d] : T(0);
        a_s[e] = T(-2) * x * il_s[d];
    }
    __syncthreads();
    if (threadIdx.x < TM) {
        T acc = 0;
        for (int d = 0; d < Dpad; ++d) { T a = a_s[threadIdx.x * Dpad + d]; acc = fma(a, a, acc); }
        na_s[threadIdx.x] = T(0.25) * acc;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = jt + ((warp & 1) * 32 + lane) * 4;
    const int r0 = (warp >> 1) * RM;
    if (j0 >= N2) return;

    T acc[RM][4];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = T(0);
    T nb[4] = {T(0), T(0), T(0), T(0)};

    for (int d0 = 0; d0 < Dpad; d0 += DC) {
        T b[4][DC];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + c;
#pragma unroll
            for (int dd = 0; dd < DC; ++dd) {
                const int d = d0 + dd;
                T x = (j < N2 && d < D) ? X2s[(int64_t)j * D + d] : T(0);
                x *= il_s[d];
                b[c][dd] = x;
                nb[c] = fma(x, x, nb[c]);
            }
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            T a[DC];
#pragma unroll
            for (int dd = 0; dd < DC; ++dd) a[dd] = a_s[(r0 + r) * Dpad + d0 + dd];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int dd = 0; dd < DC; ++dd) acc[r][c] = fma(a[dd], b[c][dd], acc[r][c]);
        }
    }

    T dadd = T(0);
    if (SYM) dadd = diag_const + (diag_add ? diag_add[(int64_t)s * sDiag] : T(0));
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        const int i = i0 + r0 + r;
        if (i >= N) break;
        const T na = na_s[r0 + r];
        T o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            T r2 = acc[r][c] + na + nb[c];
            if (SYM && i == j0 + c) r2 = T(0);       // exact on the diagonal of K(X,X) (see the streaming kernel)
            o[c] = kern_value<T, KIND>(r2, v);
            if (SYM && i == j0 + c) o[c] += dadd;
        }
        T* dst = outs + (int64_t)i * ldo + j0;
        if (vec_ok && j0 + 3 < N2) {
            if (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(dst) = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
            } else {
                *reinterpret_cast<double2*>(dst) = make_double2((double)o[0], (double)o[1]);
                *reinterpret_cast<double2*>(dst + 2) = make_double2((double)o[2], (double)o[3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (j0 + c < N2) dst[c] = o[c];
        }
    }
}

template <typename T, int KIND, int DC, int RM>
static int launch_fwd(const T* X, const T* X2, const T* ls, int ls_len, const T* var,
                      const T* diag_add, double diag_const, T* out, int64_t ldo, int S, int N,
                      int N2, int D, int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar,
                      int64_t sDiag, int64_t sOut, cudaStream_t st) {
    constexpr int TM = 4 * RM;
    const int Dpad = (D + DC - 1) / DC * DC;
    const size_t smem = sizeof(T) * ((size_t)TM * Dpad + TM + Dpad);
    if (smem > 200 * 1024) return MXF_ENOTIMPL;
    const bool sym = (X2 == nullptr);
    const T* X2e = sym ? X : X2;
    const int64_t sX2e = sym ? sX : sX2;
    dim3 grid(cdiv(N2, KB_TN), cdiv(N, TM), S);
    const int vec_ok = (ldo % 4 == 0) && (sOut % 4 == 0) && (((uintptr_t)out) % 16 == 0);
    if (sym) {
        auto k = kbuild_fwd_kernel<T, KIND, DC, RM, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, KB_THREADS, smem, st>>>(X, X2e, ls, ls_len, var, diag_add, (T)diag_const, out, ldo, N, N2, D,
                                          Dpad, sX, sX2e, sLs, sVar, sDiag, sOut, vec_ok);
    } else {
        auto k = kbuild_fwd_kernel<T, KIND, DC, RM, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, KB_THREADS, smem, st>>>(X, X2e, ls, ls_len, var, diag_add, (T)diag_const, out, ldo, N, N2, D,
                                          Dpad, sX, sX2e, sLs, sVar, sDiag, sOut, vec_ok);
    }
    return after_launch();
}

// ------------------------------------------------------------------------------------------
// forward, streaming variant (D <= 16): the kernel behind the K(X,Z) HBM figure.
//
// A CTA (8 warps arranged WR x WC) owns WC*128 output columns and `chunk_rows` output rows.  Each warp keeps the
// scaled vectors of its 128 columns (4 per lane) in registers for the whole CTA lifetime and walks the rows RM at a
// time; the scaled row vectors of the chunk sit in shared memory (staged once, read as warp-uniform 16-byte
// broadcasts).  Every store is a 16-byte streaming store, a warp writes 512 contiguous bytes per row.
// RBF folds all constants into the operands:  K = 2^(log2 var - c r2),  c = log2(e)/2, so an element costs
// 1 FADD + D FFMA + 1 MUFU.EX2.
// ------------------------------------------------------------------------------------------
template <int DP> struct KsRM { static constexpr int value = DP <= 8 ? 8 : 4; };

// One MUFU.EX2 (max relative error 2^-22.5, results below 2^-126 flushed to zero) instead of exp2f's four-instruction
// sequence with denormal scaling: the streaming kernel is issue-bound (ncu: 69 % issue-active at 19 instructions per
// element), and a covariance below 1e-38 is zero for every consumer on this path.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T> __device__ __forceinline__ void store4_stream(T* dst, const T (&o)[4]);
template <> __device__ __forceinline__ void store4_stream<float>(float* dst, const float (&o)[4]) {
    __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
}
template <> __device__ __forceinline__ void store4_stream<double>(double* dst, const double (&o)[4]) {
    __stcs(reinterpret_cast<double2*>(dst), make_double2(o[0], o[1]));
    __stcs(reinterpret_cast<double2*>(dst + 2), make_double2(o[2], o[3]));
}

template <typename T, int KIND, int DP, int WC, bool SYM>
__global__ void __launch_bounds__(256, 2)
kbuild_fwd_stream_kernel(const T* __restrict__ X, const T* __restrict__ X2, const T* __restrict__ ls, int ls_len,
                         const T* __restrict__ var, const T* __restrict__ diag_add, T diag_const,
                         T* __restrict__ out, int64_t ldo, int N, int N2, int D, int chunk_rows, int64_t sX,
                         int64_t sX2, int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut, int vec_ok) {
    constexpr int WR = 8 / WC;
    constexpr int KS_RM = KsRM<DP>::value;
    constexpr bool RBF_FOLD = (KIND == MXF_KERN_RBF);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* a_s = reinterpret_cast<T*>(smem_raw);          // [chunk_rows][DP]  -2 * scaled row vectors
    T* na_s = a_s + (size_t)chunk_rows * DP;          // [chunk_rows]      |scaled row|^2
    T* sc_s = na_s + chunk_rows;                      // [DP]              per-dimension scale (0 in the padding)

    const int s = blockIdx.z;
    const T* Xs = X + (int64_t)s * sX;
    const T* X2s = X2 + (int64_t)s * sX2;
    const T* lss = ls + (int64_t)s * sLs;
    const T v = var[(int64_t)s * sVar];
    T* outs = out + (int64_t)s * sOut;
    const int i0 = blockIdx.y * chunk_rows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wc = warp % WC, wr = warp / WC;
    const int j0 = (blockIdx.x * WC + wc) * 128 + lane * 4;
    const T csq = RBF_FOLD ? T(0.84932180028801904272) : T(1);      // sqrt(log2(e)/2)

    for (int d = threadIdx.x; d < DP; d += 256) sc_s[d] = d < D ? csq / lss[ls_len == 1 ? 0 : d] : T(0);
    __syncthreads();
    const int nrows = min(chunk_rows, N - i0);
    for (int e = threadIdx.x; e < chunk_rows * DP; e += 256) {
        const int r = e / DP, d = e - r * DP;
        const T x = (r < nrows && d < D) ? Xs[(int64_t)(i0 + r) * D + d] : T(0);
        a_s[e] = T(-2) * x * sc_s[d];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < chunk_rows; r += 256) {
        T acc = 0;
#pragma unroll
        for (int d = 0; d < DP; ++d) { const T a = a_s[r * DP + d]; acc = fma(a, a, acc); }
        na_s[r] = T(0.25) * acc;
    }
    // scaled column vectors of this lane's 4 columns
    T b[4][DP];
    T nb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int j = j0 + c;
        T n2 = 0;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            T x = (j < N2 && d < D) ? X2s[(int64_t)j * D + d] : T(0);
            x *= sc_s[d];
            b[c][d] = x;
            n2 = fma(x, x, n2);
        }
        nb[c] = RBF_FOLD ? n2 - log2(v) : n2;
    }
    __syncthreads();
    if (j0 >= N2) return;
    T dadd = T(0);
    if (SYM) dadd = diag_const + (diag_add ? diag_add[(int64_t)s * sDiag] : T(0));
    const bool full4 = vec_ok && (j0 + 3 < N2);
    const float l2v = log2f((float)v);

    for (int rt = wr * KS_RM; rt < nrows; rt += WR * KS_RM) {
        T acc[KS_RM][4];
#pragma unroll
        for (int r = 0; r < KS_RM; ++r) {
            const T na = na_s[rt + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = na + nb[c];
        }
#pragma unroll
        for (int r = 0; r < KS_RM; ++r) {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                const T a = a_s[(rt + r) * DP + d];
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a, b[c][d], acc[r][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < KS_RM; ++r) {
            const int i = i0 + rt + r;
            if (rt + r >= nrows) break;
            T o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                T e = acc[r][c];
                // K(X,X): on the diagonal r2 is exactly 0; the expanded form only leaves cancellation noise there, which
                // the Matern square root would amplify (sqrt(1e-7) in f32)
                if (SYM && i == j0 + c) e = RBF_FOLD ? -log2(v) : T(0);
                if (RBF_FOLD) o[c] = (sizeof(T) == 4) ? (T)ex2_approx(-(float)e) : (T)exp2(-(double)e);
                else if (sizeof(T) == 4) o[c] = (T)matern_value_l2<KIND>((float)e, l2v);
                else o[c] = kern_value<T, KIND>(e, v);
                if (SYM && i == j0 + c) o[c] += dadd;
            }
            T* dst = outs + (int64_t)i * ldo + j0;
            if (full4) {
                store4_stream<T>(dst, o);
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (j0 + c < N2) dst[c] = o[c];
            }
        }
    }
}

template <typename T, int KIND, int DP, int WC>
static int launch_fwd_stream(const T* X, const T* X2, const T* ls, int ls_len, const T* var, const T* diag_add,
                             double diag_const, T* out, int64_t ldo, int S, int N, int N2, int D, int64_t sX,
                             int64_t sX2, int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut, cudaStream_t st) {
    constexpr int WR = 8 / WC;
    constexpr int KS_RM = KsRM<DP>::value;
    const bool sym = (X2 == nullptr);
    const T* X2e = sym ? X : X2;
    const int64_t sX2e = sym ? sX : sX2;
    const int colblocks = cdiv(N2, WC * 128);
    // rows per CTA: a multiple of WR*RM, at most 256, small enough to give every SM several CTAs
    const int unit = WR * KS_RM;
    int chunk = 256;
    while (chunk > unit && (int64_t)colblocks * cdiv(N, chunk) * S < 4 * kNumSMs) chunk >>= 1;
    chunk = std::max(unit, (chunk / unit) * unit);
    const size_t smem = sizeof(T) * ((size_t)chunk * DP + chunk + DP);
    dim3 grid(colblocks, cdiv(N, chunk), S);
    if (grid.y > 65535) return MXF_ENOTIMPL;
    const int vec_ok = (ldo % 4 == 0) && (sOut % 4 == 0) && (((uintptr_t)out) % 16 == 0);
    if (sym)
        kbuild_fwd_stream_kernel<T, KIND, DP, WC, true><<<grid, 256, smem, st>>>(
            X, X2e, ls, ls_len, var, diag_add, (T)diag_const, out, ldo, N, N2, D, chunk, sX, sX2e, sLs, sVar, sDiag, sOut, vec_ok);
    else
        kbuild_fwd_stream_kernel<T, KIND, DP, WC, false><<<grid, 256, smem, st>>>(
            X, X2e, ls, ls_len, var, diag_add, (T)diag_const, out, ldo, N, N2, D, chunk, sX, sX2e, sLs, sVar, sDiag, sOut, vec_ok);
    return after_launch();
}

// ------------------------------------------------------------------------------------------
// forward, streaming variant for 8 < D <= 16 (f32): the dot products on the tensor pipe.
//
// At D = 16 the FMA formulation needs 16 FFMA + ~8 other instructions per output element and is issue-bound at
// 0.40-0.48 of the HBM peak (25 % occupancy).  The reference itself forms -2 a.b with a GEMM (stationary.py:102), and
// that is what this kernel does: every warp owns 64 output columns whose scaled vectors sit in registers as TF32
// B-fragments (hi and lo halves), walks the rows 16 at a time with the A-fragments read pre-split from shared memory, and
// issues mma.sync.m16n8k8 TF32 x 3 (a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulate: the 3xTF32 scheme of
// gemm_tc.cu).  An element then costs 48/1024 MMAs + 2 FADD + 1 MUFU (+ the Matern polynomial) + half a 64-bit
// streaming store: the kernel is back on the HBM roofline.  (K = 16 is one or two MMA steps: far below the 128 x N x 8
// tiles tcgen05 wants, and the accumulators are consumed immediately by the exponential -- the warp-level MMA is the
// right tool for this contraction; the large GEMMs of the path are tcgen05, gemm_tc.cu.)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int KM_COLS = 64;        // output columns per warp (8 n-tiles)

template <int KIND, bool SYM>
__global__ void __launch_bounds__(256, 1)
kbuild_fwd_mma_kernel(const float* __restrict__ X, const float* __restrict__ X2, const float* __restrict__ ls, int ls_len,
                      const float* __restrict__ var, const float* __restrict__ diag_add, float diag_const,
                      float* __restrict__ out, int64_t ldo, int N, int N2, int D, int chunk_rows, int64_t sX, int64_t sX2,
                      int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut, int vec2_ok) {
    constexpr int DP = 16;
    constexpr bool RBF_FOLD = (KIND == MXF_KERN_RBF);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* ah_s = reinterpret_cast<unsigned*>(smem_raw);           // [chunk_rows][DP]  tf32 hi of -2 * scaled row
    unsigned* al_s = ah_s + (size_t)chunk_rows * DP;                   // [chunk_rows][DP]  tf32 lo
    float* na_s = reinterpret_cast<float*>(al_s + (size_t)chunk_rows * DP);   // [chunk_rows]  |scaled row|^2
    float* sc_s = na_s + chunk_rows;                                   // [DP]
    float* nb_s = sc_s + DP;                                           // [8 * KM_COLS]  column norms (+ folded constants)

    const int s = blockIdx.z;
    const float* Xs = X + (int64_t)s * sX;
    const float* X2s = X2 + (int64_t)s * sX2;
    const float* lss = ls + (int64_t)s * sLs;
    const float v = var[(int64_t)s * sVar];
    float* outs = out + (int64_t)s * sOut;
    const int i0 = blockIdx.y * chunk_rows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int jw = blockIdx.x * (8 * KM_COLS) + warp * KM_COLS;         // first column of this warp
    const float csq = RBF_FOLD ? 0.84932180028801904272f : 1.0f;       // sqrt(log2(e)/2)
    const float l2v = log2f(v);

    for (int d = threadIdx.x; d < DP; d += 256) sc_s[d] = d < D ? csq / lss[ls_len == 1 ? 0 : d] : 0.f;
    __syncthreads();
    const int nrows = min(chunk_rows, N - i0);
    for (int e = threadIdx.x; e < chunk_rows * DP; e += 256) {
        const int r = e / DP, d = e - r * DP;
        const float x = (r < nrows && d < D) ? Xs[(int64_t)(i0 + r) * D + d] : 0.f;
        const float a = -2.f * x * sc_s[d];
        const unsigned hi = f2tf32(a);
        ah_s[e] = hi;
        al_s[e] = f2tf32(a - __uint_as_float(hi));
    }
    for (int r = threadIdx.x; r < chunk_rows; r += 256) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const float x = (r < nrows && d < D) ? Xs[(int64_t)(i0 + r) * D + d] * sc_s[d] : 0.f;
            acc = fmaf(x, x, acc);
        }
        na_s[r] = acc;
    }
    // column norms of the CTA's 512 columns
    for (int cidx = threadIdx.x; cidx < 8 * KM_COLS; cidx += 256) {
        const int j = blockIdx.x * (8 * KM_COLS) + cidx;
        float n2 = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const float x = (j < N2 && d < D) ? X2s[(int64_t)j * D + d] * sc_s[d] : 0.f;
            n2 = fmaf(x, x, n2);
        }
        nb_s[cidx] = RBF_FOLD ? n2 - l2v : n2;
    }
    // B fragments of this warp's 64 columns: n-tile nt, k-step ks: b0 = (k = 8 ks + t, n = g), b1 = (k = 8 ks + t + 4, n = g)
    unsigned bh[8][2][2], bl[8][2][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int j = jw + 8 * nt + g;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int d = 8 * ks + t + 4 * h;
                const float x = (j < N2 && d < D) ? X2s[(int64_t)j * D + d] * sc_s[d] : 0.f;
                const unsigned hi = f2tf32(x);
                bh[nt][ks][h] = hi;
                bl[nt][ks][h] = f2tf32(x - __uint_as_float(hi));
            }
    }
    __syncthreads();
    if (jw >= N2) return;
    float nbv[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        nbv[nt][0] = nb_s[warp * KM_COLS + 8 * nt + 2 * t];
        nbv[nt][1] = nb_s[warp * KM_COLS + 8 * nt + 2 * t + 1];
    }
    float dadd = 0.f;
    if (SYM) dadd = diag_const + (diag_add ? diag_add[(int64_t)s * sDiag] : 0.f);

    for (int rt = 0; rt < nrows; rt += 16) {
        // A fragments: a0 = (row g, k = t), a1 = (row g + 8, k = t), a2 = (row g, k = t + 4), a3 = (row g + 8, k = t + 4)
        unsigned ah[2][4], al[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int o0 = (rt + g) * DP + 8 * ks + t, o1 = (rt + g + 8) * DP + 8 * ks + t;
            ah[ks][0] = ah_s[o0]; ah[ks][1] = ah_s[o1]; ah[ks][2] = ah_s[o0 + 4]; ah[ks][3] = ah_s[o1 + 4];
            al[ks][0] = al_s[o0]; al[ks][1] = al_s[o1]; al[ks][2] = al_s[o0 + 4]; al[ks][3] = al_s[o1 + 4];
        }
        const float na0 = na_s[rt + g], na1 = na_s[rt + g + 8];
        const int ia = i0 + rt + g, ib = ia + 8;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                mma_tf32_16x8x8(c, al[ks], bh[nt][ks][0], bh[nt][ks][1]);
                mma_tf32_16x8x8(c, ah[ks], bl[nt][ks][0], bl[nt][ks][1]);
                mma_tf32_16x8x8(c, ah[ks], bh[nt][ks][0], bh[nt][ks][1]);
            }
            const int j = jw + 8 * nt + 2 * t;
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = (q < 2) ? ia : ib, jj = j + (q & 1);
                float e = c[q] + ((q < 2) ? na0 : na1) + nbv[nt][q & 1];
                if (SYM && i == jj) e = RBF_FOLD ? -l2v : 0.f;          // exact on the diagonal of K(X, X)
                if (RBF_FOLD) o[q] = ex2_approx(-e);
                else o[q] = matern_value_l2<KIND>(e, l2v);
                if (SYM && i == jj) o[q] += dadd;
            }
            if (ia < N) {
                float* dst = outs + (int64_t)ia * ldo + j;
                if (vec2_ok && j + 1 < N2) __stcs(reinterpret_cast<float2*>(dst), make_float2(o[0], o[1]));
                else { if (j < N2) dst[0] = o[0]; if (j + 1 < N2) dst[1] = o[1]; }
            }
            if (ib < N) {
                float* dst = outs + (int64_t)ib * ldo + j;
                if (vec2_ok && j + 1 < N2) __stcs(reinterpret_cast<float2*>(dst), make_float2(o[2], o[3]));
                else { if (j < N2) dst[0] = o[2]; if (j + 1 < N2) dst[1] = o[3]; }
            }
        }
    }
}

template <int KIND>
static int launch_fwd_mma(const float* X, const float* X2, const float* ls, int ls_len, const float* var, const float* diag_add,
                          double diag_const, float* out, int64_t ldo, int S, int N, int N2, int D, int64_t sX, int64_t sX2,
                          int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut, cudaStream_t st) {
    const bool sym = (X2 == nullptr);
    const float* X2e = sym ? X : X2;
    const int64_t sX2e = sym ? sX : sX2;
    const int colblocks = cdiv(N2, 8 * KM_COLS);
    int chunk = 256;
    while (chunk > 16 && (int64_t)colblocks * cdiv(N, chunk) * S < 2 * kNumSMs) chunk >>= 1;
    const size_t smem = (size_t)chunk * 16 * 8 + sizeof(float) * ((size_t)chunk + 16 + 8 * KM_COLS);
    dim3 grid(colblocks, cdiv(N, chunk), S);
    if (grid.y > 65535) return MXF_ENOTIMPL;
    const int vec2_ok = (ldo % 2 == 0) && (sOut % 2 == 0) && (((uintptr_t)out) % 8 == 0);
    if (sym)
        kbuild_fwd_mma_kernel<KIND, true><<<grid, 256, smem, st>>>(X, X2e, ls, ls_len, var, diag_add, (float)diag_const, out, ldo, N,
                                                                    N2, D, chunk, sX, sX2e, sLs, sVar, sDiag, sOut, vec2_ok);
    else
        kbuild_fwd_mma_kernel<KIND, false><<<grid, 256, smem, st>>>(X, X2e, ls, ls_len, var, diag_add, (float)diag_const, out, ldo, N,
                                                                     N2, D, chunk, sX, sX2e, sLs, sVar, sDiag, sOut, vec2_ok);
    return after_launch();
}

}  // namespace mxf
#include "kbuild_tc.cuh"
namespace mxf {

template <typename T, int KIND, int DP>
static int dispatch_fwd_stream_wc(const T* X, const T* X2, const T* ls, int ls_len, const T* var, const T* diag_add,
                                  double diag_const, T* out, int64_t ldo, int S, int N, int N2, int D, int64_t sX,
                                  int64_t sX2, int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut, cudaStream_t st) {
#define MXF_KS_ARGS X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D, sX, sX2, sLs, sVar, sDiag, sOut, st
    if (N2 > 512) return launch_fwd_stream<T, KIND, DP, 8>(MXF_KS_ARGS);
    if (N2 > 256) return launch_fwd_stream<T, KIND, DP, 4>(MXF_KS_ARGS);
    if (N2 > 128) return launch_fwd_stream<T, KIND, DP, 2>(MXF_KS_ARGS);
    return launch_fwd_stream<T, KIND, DP, 1>(MXF_KS_ARGS);
#undef MXF_KS_ARGS
}

// Measured on B200 (N=1e6, M=1024, D=16): 1.46 ms RBF / 1.74 ms Matern-5/2 = 0.43 / 0.36 of the HBM peak, against 0.48 /
// 0.40 for the FMA kernel: the legacy warp-level TF32 MMA runs at about the FP32 FMA rate on sm_100 (48 HMMA.1688 per
// 16 x 64 tile ~ 1.5k cycles per warp), so three of them per multiply-add lose to one FFMA.  Kept as an opt-in
// (MXF_KBUILD_MMA=1) and as the parity-tested starting point of a tcgen05 version; off by default.
static bool kbuild_mma_disabled() {
    static int on = [] { const char* e = getenv("MXF_KBUILD_MMA"); return (e && e[0] == '1') ? 1 : 0; }();
    return on == 0;
}

template <typename T, int KIND>
static int dispatch_fwd_dc(const T* X, const T* X2, const T* ls, int ls_len, const T* var,
                           const T* diag_add, double diag_const, T* out, int64_t ldo, int S, int N,
                           int N2, int D, int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar,
                           int64_t sDiag, int64_t sOut, cudaStream_t st) {
    constexpr int RM = sizeof(T) == 4 ? 16 : 8;
    if (N2 >= 32) {
#define MXF_KS_ARGS X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D, sX, sX2, sLs, sVar, sDiag, sOut, st
        if constexpr (sizeof(T) == 4) {
            // large cross-covariances: the dot products on tcgen05, the output through TMA stores (kbuild_tc.cuh)
            if (D <= 16 && N2 >= 128 && (int64_t)S * N * N2 >= g_kbuild_tc_min_elems.load(std::memory_order_relaxed)) {
                const int rc = launch_fwd_tc<KIND>(X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D, sX, sX2,
                                                   sLs, sVar, sDiag, sOut, st);
                if (rc != MXF_ENOTIMPL) return rc;
            }
        }
        if (D <= 4) return dispatch_fwd_stream_wc<T, KIND, 4>(MXF_KS_ARGS);
        if (D <= 8) return dispatch_fwd_stream_wc<T, KIND, 8>(MXF_KS_ARGS);
        if constexpr (sizeof(T) == 4) {
            if (D > 8 && D <= 16 && N2 >= 64 && !kbuild_mma_disabled())
                return launch_fwd_mma<KIND>(X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D, sX, sX2, sLs,
                                            sVar, sDiag, sOut, st);
        }
        if (D <= 16 && sizeof(T) == 4) return dispatch_fwd_stream_wc<T, KIND, 16>(MXF_KS_ARGS);
#undef MXF_KS_ARGS
    }
    if (D <= 4)
        return launch_fwd<T, KIND, 4, RM>(X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D,
                                          sX, sX2, sLs, sVar, sDiag, sOut, st);
    return launch_fwd<T, KIND, 8, RM>(X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D, sX,
                                      sX2, sLs, sVar, sDiag, sOut, st);
}

template <typename T>
static int dispatch_fwd(int kind, const T* X, const T* X2, const T* ls, int ls_len, const T* var,
                        const T* diag_add, double diag_const, T* out, int64_t ldo, int S, int N, int N2,
                        int D, int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar, int64_t sDiag,
                        int64_t sOut, cudaStream_t st) {
#define MXF_KB_CASE(K)                                                                                   \
    case K:                                                                                              \
        return dispatch_fwd_dc<T, K>(X, X2, ls, ls_len, var, diag_add, diag_const, out, ldo, S, N, N2, D, sX, \
                                     sX2, sLs, sVar, sDiag, sOut, st);
    switch (kind) {
        MXF_KB_CASE(MXF_KERN_RBF)
        MXF_KB_CASE(MXF_KERN_MATERN12)
        MXF_KB_CASE(MXF_KERN_MATERN32)
        MXF_KB_CASE(MXF_KERN_MATERN52)
    }
#undef MXF_KB_CASE
    return MXF_EINVAL;
}

// ------------------------------------------------------------------------------------------
// adjoint
// ------------------------------------------------------------------------------------------
// Workspace layout (all T, zero-initialised by the first kernel):
//   RS[S][N]  RB[S][N][D]  CS[S][N2]  CB[S][N2][D]  ACC[S][ls_len + 1]
// with H_ij = G_ij * dK/dr2:  RS_i = sum_j H_ij, RB_id = sum_j H_ij b'_jd, CS_j = sum_i H_ij,
// CB_jd = sum_i H_ij a'_id (a' = x/l, b' = x2/l).  Then
//   dX_id  = (2/l_d) (a'_id RS_i - RB_id),   dX2_jd = (2/l_d) (b'_jd CS_j - CB_jd),
//   dl_d   = -(2/l_d) (sum_i a'_id^2 RS_i + sum_j b'_jd^2 CS_j - 2 sum_i a'_id RB_id),
//   dvar   = sum_ij G_ij K_ij / var.
constexpr int KBB_THREADS = 128;
constexpr int KBB_TN = 128;

// COLS = false: the column operand X2 needs no gradient.  The column sums (and their atomics) are skipped and the
// lengthscale gradient is accumulated directly from  d r2_ij / d l_d = -(2 / l_d) (a'_id - b'_jd)^2.
template <typename T, int KIND, int TR, bool COLS>
__global__ void __launch_bounds__(KBB_THREADS)
kbuild_bwd_tile_kernel(const T* __restrict__ X, const T* __restrict__ X2, const T* __restrict__ ls,
                       int ls_len, const T* __restrict__ var, const T* __restrict__ G, int64_t ldg,
                       T* __restrict__ RS, T* __restrict__ RB, T* __restrict__ CS, T* __restrict__ CB,
                       T* __restrict__ ACC, int N, int N2, int D, int64_t sX, int64_t sX2,
                       int64_t sLs, int64_t sVar, int64_t sG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* a_s = reinterpret_cast<T*>(smem_raw);        // [TR][D]   a' rows
    T* bT_s = a_s + TR * D;                         // [D][KBB_TN] b' columns (transposed)
    T* h_s = bT_s + D * KBB_TN;                     // [TR][KBB_TN+1]
    T* red = h_s + TR * (KBB_TN + 1);               // [32]
    constexpr int HS = KBB_TN + 1;

    const int s = blockIdx.z;
    const T* Xs = X + (int64_t)s * sX;
    const T* X2s = X2 + (int64_t)s * sX2;
    const T* lss = ls + (int64_t)s * sLs;
    const T v = var[(int64_t)s * sVar];
    const T* Gs = G + (int64_t)s * sG;
    const int i0 = blockIdx.y * TR, j0 = blockIdx.x * KBB_TN;
    const int tid = threadIdx.x;

    for (int e = tid; e < TR * D; e += KBB_THREADS) {
        int r = e / D, d = e - r * D;
        int i = i0 + r;
        a_s[e] = i < N ? Xs[(int64_t)i * D + d] / lss[ls_len == 1 ? 0 : d] : T(0);
    }
    for (int e = tid; e < KBB_TN * D; e += KBB_THREADS) {
        int c = e / D, d = e - c * D;
        int j = j0 + c;
        bT_s[d * KBB_TN + c] = j < N2 ? X2s[(int64_t)j * D + d] / lss[ls_len == 1 ? 0 : d] : T(0);
    }
    __syncthreads();

    // phase H: thread per column
    const int j = j0 + tid;
    T nb = 0;
    for (int d = 0; d < D; ++d) { T b = bT_s[d * KBB_TN + tid]; nb = fma(b, b, nb); }
    T gk = 0;
    T dl_acc[COLS ? 1 : 16];
    if (!COLS) {
#pragma unroll
        for (int d = 0; d < 16; ++d) dl_acc[d] = T(0);
    }
    // rows in groups of 8: the 8 (coalesced) loads of G are issued before any arithmetic, so their latency overlaps
    for (int rb = 0; rb < TR; rb += 8) {
        T g8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + rb + u;
            g8[u] = (i < N && j < N2) ? Gs[(int64_t)i * ldg + j] : T(0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = rb + u, i = i0 + r;
            T h = 0;
            if (i < N && j < N2) {
                T dot = 0, na = 0;
                for (int d = 0; d < D; ++d) {
                    T a = a_s[r * D + d];
                    na = fma(a, a, na);
                    dot = fma(a, bT_s[d * KBB_TN + tid], dot);
                }
                T r2 = na + nb - T(2) * dot;
                T kv;
                T dk = kern_dr2<T, KIND>(r2, v, &kv);
                const T g = g8[u];
                h = g * dk;
                gk = fma(g, kv, gk);
                if (!COLS) {
#pragma unroll
                    for (int d = 0; d < 16; ++d)
                        if (d < D) { const T df = a_s[r * D + d] - bT_s[d * KBB_TN + tid]; dl_acc[d] = fma(h * df, df, dl_acc[d]); }
                }
            }
            h_s[r * HS + tid] = h;
        }
    }
    __syncthreads();

    if (COLS) {
        // phase A: column sums (thread per column)
        if (j < N2) {
            T cs = 0;
            for (int r = 0; r < TR; ++r) cs += h_s[r * HS + tid];
            atomicAdd(&CS[(int64_t)s * N2 + j], cs);
            for (int d = 0; d < D; ++d) {
                T cb = 0;
                for (int r = 0; r < TR; ++r) cb = fma(h_s[r * HS + tid], a_s[r * D + d], cb);
                atomicAdd(&CB[((int64_t)s * N2 + j) * D + d], cb);
            }
        }
    } else {
        // direct lengthscale gradient: one atomic per dimension per CTA
        for (int d = 0; d < D; ++d) {
            const T l = lss[ls_len == 1 ? 0 : d];
            T tot = block_sum(dl_acc[d] * (T(-2) / l), red);
            if (tid == 0) atomicAdd(&ACC[(int64_t)s * (ls_len + 1) + (ls_len == 1 ? 0 : d)], tot);
        }
    }
    // phase B: row sums.  KBB_THREADS / TR threads cooperate on one row.
    {
        constexpr int TPR = KBB_THREADS / TR;       // threads per row (power of two <= 32)
        const int r = tid / TPR, q = tid % TPR;
        const int i = i0 + r;
        T rs = 0;
        for (int c = q; c < KBB_TN; c += TPR) rs += h_s[r * HS + c];
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        if (q == 0 && i < N) atomicAdd(&RS[(int64_t)s * N + i], rs);
        for (int d = 0; d < D; ++d) {
            T rb = 0;
            for (int c = q; c < KBB_TN; c += TPR) rb = fma(h_s[r * HS + c], bT_s[d * KBB_TN + c], rb);
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) rb += __shfl_xor_sync(0xffffffffu, rb, o);
            if (q == 0 && i < N) atomicAdd(&RB[((int64_t)s * N + i) * D + d], rb);
        }
    }
    // dvar partial
    T tot = block_sum(gk, red);
    if (tid == 0) atomicAdd(&ACC[(int64_t)s * (ls_len + 1) + ls_len], tot);
}

// Register-blocked variant for D <= 16 (DP = D rounded up to 4 / 8 / 16): the thread's scaled column vector lives in
// registers, the scaled row vectors (zero-padded to DP) are read as 16-byte warp-uniform broadcasts, |a'|^2 is staged
// once per row, and -- when the column operand needs no gradient (COLS = false: the K(Z, X) adjoint of the SVGP / sparse-GP
// step, X being data) -- the column sums never leave registers: the lengthscale gradient's column part
// sum_j b'_jd^2 CS_j is formed directly and its row part comes from RS / RB in the finalise kernel.  ncu on the first
// kernel (1024 x 4096 x 8): 22 M warp instructions, issue-bound at 20 % occupancy; this one executes about a third.
template <typename T, int KIND, int TR, bool COLS, int DP>
__global__ void __launch_bounds__(KBB_THREADS)
kbuild_bwd_tile2_kernel(const T* __restrict__ X, const T* __restrict__ X2, const T* __restrict__ ls,
                        int ls_len, const T* __restrict__ var, const T* __restrict__ G, int64_t ldg,
                        T* __restrict__ RS, T* __restrict__ RB, T* __restrict__ CS, T* __restrict__ CB,
                        T* __restrict__ ACC, int N, int N2, int D, int64_t sX, int64_t sX2,
                        int64_t sLs, int64_t sVar, int64_t sG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* a_s = reinterpret_cast<T*>(smem_raw);        // [TR][DP]  a' rows, zero-padded
    T* bT_s = a_s + TR * DP;                        // [DP][KBB_TN] b' columns (transposed)
    T* h_s = bT_s + DP * KBB_TN;                    // [TR][KBB_TN+1]
    T* na_s = h_s + TR * (KBB_TN + 1);              // [TR]  |a'|^2
    T* red = na_s + TR;                             // [32 * (DP + 1)]
    constexpr int HS = KBB_TN + 1;

    const int s = blockIdx.z;
    const T* Xs = X + (int64_t)s * sX;
    const T* X2s = X2 + (int64_t)s * sX2;
    const T* lss = ls + (int64_t)s * sLs;
    const T v = var[(int64_t)s * sVar];
    const T* Gs = G + (int64_t)s * sG;
    const int i0 = blockIdx.y * TR, j0 = blockIdx.x * KBB_TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < TR * DP; e += KBB_THREADS) {
        const int r = e / DP, d = e - r * DP;
        const int i = i0 + r;
        a_s[e] = (i < N && d < D) ? Xs[(int64_t)i * D + d] / lss[ls_len == 1 ? 0 : d] : T(0);
    }
    for (int e = tid; e < KBB_TN * DP; e += KBB_THREADS) {
        const int c = e / DP, d = e - c * DP;
        const int j = j0 + c;
        bT_s[d * KBB_TN + c] = (j < N2 && d < D) ? X2s[(int64_t)j * D + d] / lss[ls_len == 1 ? 0 : d] : T(0);
    }
    __syncthreads();
    for (int r = tid; r < TR; r += KBB_THREADS) {
        T na = 0;
#pragma unroll
        for (int d = 0; d < DP; ++d) na = fma(a_s[r * DP + d], a_s[r * DP + d], na);
        na_s[r] = na;
    }
    // phase H: thread per column, its b' vector in registers
    const int j = j0 + tid;
    T b[DP];
    T nb = 0;
#pragma unroll
    for (int d = 0; d < DP; ++d) { b[d] = bT_s[d * KBB_TN + tid]; nb = fma(b[d], b[d], nb); }
    __syncthreads();
    T gk = 0, cs = 0;
    for (int rb = 0; rb < TR; rb += 8) {
        T g8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + rb + u;
            g8[u] = (i < N && j < N2) ? Gs[(int64_t)i * ldg + j] : T(0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = rb + u;
            T dot = 0;
#pragma unroll
            for (int d = 0; d < DP; ++d) dot = fma(a_s[r * DP + d], b[d], dot);
            const T r2 = na_s[r] + nb - T(2) * dot;
            T kv;
            const T dk = kern_dr2<T, KIND>(r2, v, &kv);
            const T h = g8[u] * dk;             // g8 is 0 outside the matrix, so h and the sums below are too
            gk = fma(g8[u], kv, gk);
            cs += h;
            h_s[r * HS + tid] = h;
        }
    }
    __syncthreads();

    if (COLS) {
        // phase A: column sums (thread per column)
        if (j < N2) {
            atomicAdd(&CS[(int64_t)s * N2 + j], cs);
            T cb[DP];
#pragma unroll
            for (int d = 0; d < DP; ++d) cb[d] = T(0);
            for (int r = 0; r < TR; ++r) {
                const T h = h_s[r * HS + tid];
#pragma unroll
                for (int d = 0; d < DP; ++d) cb[d] = fma(h, a_s[r * DP + d], cb[d]);
            }
#pragma unroll
            for (int d = 0; d < DP; ++d)
                if (d < D) atomicAdd(&CB[((int64_t)s * N2 + j) * D + d], cb[d]);
        }
    } else {
        // column part of the lengthscale gradient, -(2 / l_d) sum_j b'_jd^2 CS_j: warp sums, then one atomic per d per CTA
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const T t = warp_sum(cs * b[d] * b[d]);
            if (lane == 0) red[warp * DP + d] = t;
        }
        __syncthreads();
        if (tid < D) {
            T t = 0;
            for (int w = 0; w < KBB_THREADS / 32; ++w) t += red[w * DP + tid];
            const T l = lss[ls_len == 1 ? 0 : tid];
            atomicAdd(&ACC[(int64_t)s * (ls_len + 1) + (ls_len == 1 ? 0 : tid)], t * (T(-2) / l));
        }
        __syncthreads();
    }
    // phase B: row sums.  A warp owns TR / 4 rows; its lanes stride over the columns, then a warp reduction per quantity.
    {
        constexpr int RPW = TR / (KBB_THREADS / 32);
        T bb[KBB_TN / 32][DP];
#pragma unroll
        for (int q = 0; q < KBB_TN / 32; ++q)
#pragma unroll
            for (int d = 0; d < DP; ++d) bb[q][d] = bT_s[d * KBB_TN + lane + 32 * q];
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr, i = i0 + r;
            T rs = 0, rbv[DP];
#pragma unroll
            for (int d = 0; d < DP; ++d) rbv[d] = T(0);
#pragma unroll
            for (int q = 0; q < KBB_TN / 32; ++q) {
                const T h = h_s[r * HS + lane + 32 * q];
                rs += h;
#pragma unroll
                for (int d = 0; d < DP; ++d) rbv[d] = fma(h, bb[q][d], rbv[d]);
            }
            rs = warp_sum(rs);
#pragma unroll
            for (int d = 0; d < DP; ++d) rbv[d] = warp_sum(rbv[d]);
            if (i < N) {
                if (lane == 0) atomicAdd(&RS[(int64_t)s * N + i], rs);
#pragma unroll
                for (int d = 0; d < DP; ++d)
                    if (lane == d && d < D) atomicAdd(&RB[((int64_t)s * N + i) * D + d], rbv[d]);
            }
        }
    }
    // dvar partial
    const T tot = block_sum(gk, red);
    if (tid == 0) atomicAdd(&ACC[(int64_t)s * (ls_len + 1) + ls_len], tot);
}

// Finalise: rows (which = 0) or columns (which = 1).  One thread per (row, d).
template <typename T>
__global__ void kbuild_bwd_final_kernel(const T* __restrict__ Xp, const T* __restrict__ ls, int ls_len,
                                        const T* __restrict__ SUMS, const T* __restrict__ SUMB,
                                        T* __restrict__ dX, int accumulate, T* __restrict__ ACC,
                                        int n, int D, int64_t sX, int64_t sLs, int is_rows, int skip_dl) {
    __shared__ T red[32];
    const int s = blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tot = (int64_t)n * D;
    T contrib = 0;
    int d = 0;
    if (e < tot) {
        const int64_t i = e / D;
        d = (int)(e - i * D);
        const T l = ls[(int64_t)s * sLs + (ls_len == 1 ? 0 : d)];
        const T a = Xp[(int64_t)s * sX + e] / l;
        const T rs = SUMS[(int64_t)s * n + i];
        const T rb = SUMB[((int64_t)s * n) * D + e];
        if (dX) {
            T g = (T(2) / l) * (a * rs - rb);
            T* dst = dX + ((int64_t)s * n) * D + e;
            *dst = accumulate ? *dst + g : g;
        }
        // dl contribution: rows carry a'^2 RS - 2 a' RB ; columns carry b'^2 CS
        contrib = is_rows ? (a * a * rs - T(2) * a * rb) : (a * a * rs);
        contrib *= -(T(2) / l);
    }
    if (skip_dl) return;                 // uniform: the tile kernel already accumulated d/dl directly
    if (ls_len == 1) {
        T t = block_sum(contrib, red);
        if (threadIdx.x == 0) atomicAdd(&ACC[(int64_t)s * (ls_len + 1)], t);
    } else {
        // ARD: lanes hold different d; blockDim.x is a multiple of D only by luck, so use atomics
        // after a cheap same-d pre-reduction in shared memory is not worth it: D*small.
        if (e < tot) atomicAdd(&ACC[(int64_t)s * (ls_len + 1) + d], contrib);
    }
}

template <typename T>
__global__ void kbuild_bwd_emit_kernel(const T* __restrict__ ACC, const T* __restrict__ var, int64_t sVar,
                                       int ls_len, T* __restrict__ dls, T* __restrict__ dvar, int S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * (ls_len + 1)) return;
    const int s = e / (ls_len + 1), q = e - s * (ls_len + 1);
    if (q < ls_len) dls[s * ls_len + q] = ACC[e];
    else dvar[s] = ACC[e];
}

template <typename T>
__global__ void zero_kernel(T* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = T(0);
}

template <typename T>
static size_t bwd_ws_elems(int S, int N, int N2, int D, int ls_len_max) {
    return (size_t)S * ((size_t)N * (D + 1) + (size_t)N2 * (D + 1) + ls_len_max + 1);
}

template <typename T, int KIND>
static int launch_bwd(const T* X, const T* X2, const T* ls, int ls_len, const T* var, const T* G,
                      int64_t ldg, T* dX, T* dX2, T* dls, T* dvar, int S, int N, int N2, int D,
                      int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar, int64_t sG, T* ws,
                      size_t ws_bytes, cudaStream_t st) {
    constexpr int TR = sizeof(T) == 4 ? 64 : 32;
    const bool sym = (X2 == nullptr);
    const T* X2e = sym ? X : X2;
    const int64_t sX2e = sym ? sX : sX2;
    const size_t need = bwd_ws_elems<T>(S, N, N2, D, D) * sizeof(T);
    if (ws_bytes < need) return MXF_EWORKSPACE;
    T* RS = ws;
    T* RB = RS + (size_t)S * N;
    T* CS = RB + (size_t)S * N * D;
    T* CB = CS + (size_t)S * N2;
    T* ACC = CB + (size_t)S * N2 * D;
    const int64_t nz = (int64_t)(ACC - ws) + (int64_t)S * (ls_len + 1);
    zero_kernel<T><<<std::min<int64_t>(cdiv(nz, 256), 4 * kNumSMs), 256, 0, st>>>(ws, nz);

    // the column side needs its sums in memory only if X2 (or, for K(X,X), the same X through its second role) gets a
    // gradient
    const bool cols = sym || dX2 != nullptr;
    dim3 grid(cdiv(N2, KBB_TN), cdiv(N, TR), S);
    if (D <= 16) {
#define MXF_KBB2_LAUNCH(COLSV, DPV)                                                                                   \
    do {                                                                                                              \
        auto k = kbuild_bwd_tile2_kernel<T, KIND, TR, COLSV, DPV>;                                                    \
        const size_t smem2 = sizeof(T) * ((size_t)TR * DPV + (size_t)DPV * KBB_TN + (size_t)TR * (KBB_TN + 1) + TR +   \
                                          32 * (DPV + 1));                                                            \
        if (smem2 > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);      \
        k<<<grid, KBB_THREADS, smem2, st>>>(X, X2e, ls, ls_len, var, G, ldg, RS, RB, CS, CB, ACC, N, N2, D, sX, sX2e,  \
                                            sLs, sVar, sG);                                                           \
    } while (0)
        if (D <= 4) { if (cols) MXF_KBB2_LAUNCH(true, 4); else MXF_KBB2_LAUNCH(false, 4); }
        else if (D <= 8) { if (cols) MXF_KBB2_LAUNCH(true, 8); else MXF_KBB2_LAUNCH(false, 8); }
        else { if (cols) MXF_KBB2_LAUNCH(true, 16); else MXF_KBB2_LAUNCH(false, 16); }
#undef MXF_KBB2_LAUNCH
    } else {
        const size_t smem = sizeof(T) * ((size_t)TR * D + (size_t)D * KBB_TN + (size_t)TR * (KBB_TN + 1) + 32);
        if (smem > 200 * 1024) return MXF_ENOTIMPL;
        if (cols) {
            auto k = kbuild_bwd_tile_kernel<T, KIND, TR, true>;
            if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<grid, KBB_THREADS, smem, st>>>(X, X2e, ls, ls_len, var, G, ldg, RS, RB, CS, CB, ACC, N, N2, D, sX, sX2e,
                                               sLs, sVar, sG);
        } else {
            auto k = kbuild_bwd_tile_kernel<T, KIND, TR, false>;
            if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<grid, KBB_THREADS, smem, st>>>(X, X2e, ls, ls_len, var, G, ldg, RS, RB, CS, CB, ACC, N, N2, D, sX, sX2e,
                                               sLs, sVar, sG);
        }
    }
    // lengthscale gradient: the old COLS = false kernel (D > 16) accumulates all of it directly; the register-blocked one
    // only its column part, the row part comes from RS / RB below
    const bool direct_dl = !cols && D > 16;
    // rows -> dX (and dl row terms); columns -> dX2 (or accumulated into dX when symmetric)
    if (dX || !direct_dl) {
        dim3 g(cdiv((int64_t)N * D, 256), S);
        kbuild_bwd_final_kernel<T><<<g, 256, 0, st>>>(X, ls, ls_len, RS, RB, dX, 0, ACC, N, D, sX, sLs, 1, direct_dl ? 1 : 0);
    }
    if (cols) {
        dim3 g(cdiv((int64_t)N2 * D, 256), S);
        T* dst = sym ? dX : dX2;
        kbuild_bwd_final_kernel<T><<<g, 256, 0, st>>>(X2e, ls, ls_len, CS, CB, dst, sym ? 1 : 0, ACC, N2, D, sX2e,
                                                      sLs, 0, 0);
    }
    kbuild_bwd_emit_kernel<T><<<cdiv(S * (ls_len + 1), 128), 128, 0, st>>>(ACC, var, sVar, ls_len, dls, dvar, S);
    return after_launch(5);
}

}  // namespace mxf

using namespace mxf;

extern "C" int mxf_kbuild_fwd(int kind, int dtype, const void* X, const void* X2, const void* ls, int ls_len,
                              const void* var, const void* diag_add, double diag_add_const, void* out,
                              int64_t ldo, int S, int N, int N2, int D, int64_t sX, int64_t sX2, int64_t sLs,
                              int64_t sVar, int64_t sDiag, int64_t sOut, void* stream) {
    if (!X || !ls || !var || !out || S < 0 || N < 0 || N2 < 0 || D <= 0 || (ls_len != 1 && ls_len != D))
        return MXF_EINVAL;
    if (X2 == nullptr && N2 != N) return MXF_EINVAL;
    if (S == 0 || N == 0 || N2 == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return dispatch_fwd<T>(kind, (const T*)X, (const T*)X2, (const T*)ls, ls_len,
                                                     (const T*)var, (const T*)diag_add, diag_add_const, (T*)out,
                                                     ldo, S, N, N2, D, sX, sX2, sLs, sVar, sDiag, sOut,
                                                     (cudaStream_t)stream));
}

extern "C" long long mxf_kbuild_tc_threshold(long long min_elems) {
    const long long old = mxf::g_kbuild_tc_min_elems.load();
    if (min_elems >= 0) mxf::g_kbuild_tc_min_elems.store(min_elems);
    return old;
}

extern "C" size_t mxf_kbuild_bwd_workspace_bytes(int dtype, int S, int N, int N2, int D) {
    const size_t el = dtype == MXF_F64 ? 8 : 4;
    return el * (size_t)S * ((size_t)N * (D + 1) + (size_t)N2 * (D + 1) + D + 1);
}

extern "C" int mxf_kbuild_bwd(int kind, int dtype, const void* X, const void* X2, const void* ls, int ls_len,
                              const void* var, const void* G, int64_t ldg, void* dX, void* dX2, void* dls,
                              void* dvar, int S, int N, int N2, int D, int64_t sX, int64_t sX2, int64_t sLs,
                              int64_t sVar, int64_t sG, void* ws, size_t ws_bytes, void* stream) {
    if (!X || !ls || !var || !G || !dls || !dvar || !ws || D <= 0 || (ls_len != 1 && ls_len != D))
        return MXF_EINVAL;
    if (X2 == nullptr && N2 != N) return MXF_EINVAL;
    if (S == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
#define MXF_KBB_CASE(K)                                                                                        \
    case K:                                                                                                    \
        MXF_DISPATCH_DTYPE(dtype, return launch_bwd<T, K>((const T*)X, (const T*)X2, (const T*)ls, ls_len,      \
                                                          (const T*)var, (const T*)G, ldg, (T*)dX, (T*)dX2,    \
                                                          (T*)dls, (T*)dvar, S, N, N2, D, sX, sX2, sLs, sVar, \
                                                          sG, (T*)ws, ws_bytes, (cudaStream_t)stream));        \
        break;
    switch (kind) {
        MXF_KBB_CASE(MXF_KERN_RBF)
        MXF_KBB_CASE(MXF_KERN_MATERN12)
        MXF_KBB_CASE(MXF_KERN_MATERN32)
        MXF_KBB_CASE(MXF_KERN_MATERN52)
    }
#undef MXF_KBB_CASE
    return MXF_EINVAL;
}
