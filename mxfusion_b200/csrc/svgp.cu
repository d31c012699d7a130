// Pieces of the analytic SVGP-bound gradient (modules/gp_modules/svgp_regression.py:43-109 differentiated by
// hand; the reference leaves this to MXNet autograd), the device-scalar axpby used to combine them, and the
// softplus parameter transform (components/variables/var_trans.py:63-91).
#include <algorithm>
#include "common.cuh"

namespace mxf {

template <typename T>
__global__ void __launch_bounds__(256)
axpby_dev_kernel(const T* __restrict__ a, const T* __restrict__ X, int64_t sX, const T* __restrict__ b,
                 const T* __restrict__ Y, int64_t sY, T* __restrict__ out, int64_t sO, int64_t n) {
    const int s = blockIdx.y;
    const T av = a ? a[s] : T(1);
    const T bv = b ? b[s] : T(0);
    const T* xs = X + (int64_t)s * sX;
    const T* ys = Y ? Y + (int64_t)s * sY : nullptr;
    T* os = out + (int64_t)s * sO;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T v = av * xs[i];
        if (ys) v = fma(bv, ys[i], v);
        os[i] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) softplus_fwd_kernel(const T* __restrict__ x, T offset, T* __restrict__ y, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T v = x[i];
        // log(1+e^v) = max(v,0) + log1p(e^{-|v|})
        const T av = v < T(0) ? -v : v;
        y[i] = (v > T(0) ? v : T(0)) + log1p(exp(-av)) + offset;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) softplus_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, T* __restrict__ gx, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T v = x[i];
        const T e = exp(v < T(0) ? v : -v);                 // e^{-|v|}
        const T sig = v >= T(0) ? T(1) / (T(1) + e) : e / (T(1) + e);
        gx[i] = gy[i] * sig;
    }
}

// 32x32 tiles, 32x8 threads; U^T is staged through shared memory so every global access is coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
svgp_bwd_assemble_kernel(const T* __restrict__ Phi, const T* __restrict__ Tm, const T* __restrict__ U,
                         const T* __restrict__ mt, const T* __restrict__ v, const T* __restrict__ coef,
                         T* __restrict__ out, int64_t ldo, int64_t sO, int M, int P) {
    __shared__ T ut[32][33];
    const int s = blockIdx.z;
    const int64_t mo = (int64_t)s * M * M;
    const T* Phis = Phi + mo;
    const T* Ts = Tm + mo;
    const T* Us = U + mo;
    const T* mts = mt + (int64_t)s * M * P;
    const T* vs = v + (int64_t)s * M * P;
    const T* cf = coef + (int64_t)s * 6;
    T* os = out + (int64_t)s * sO;
    const T c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3], c4 = cf[4], c5 = cf[5];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ui = bx + r, uj = by + threadIdx.x;       // U[ui][uj], coalesced along uj
        ut[r][threadIdx.x] = (ui < M && uj < M) ? Us[(int64_t)ui * M + uj] : T(0);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i >= M || j >= M) continue;
        const int64_t e = (int64_t)i * M + j;
        const T ph = Phis[e], t = Ts[e], u = Us[e], uT = ut[threadIdx.x][r];
        T mm = 0, vm = 0;
        for (int p = 0; p < P; ++p) {
            const T mi = mts[i * P + p], mj = mts[j * P + p];
            mm = fma(mi, mj, mm);
            vm = fma(vs[i * P + p], mj, vm);
            vm = fma(mi, vs[j * P + p], vm);
        }
        const T dl = (i == j) ? T(1) : T(0);
        T* orow = os + (int64_t)i * ldo;
        orow[j] = c2 * mm + c0 * (t - dl) - c1 * ph + c1 * (u + uT) - c3 * vm;
        orow[M + j] = c0 * dl + c1 * ph;
        orow[2 * M + j] = -c4 * (t - dl) - c5 * mm;
    }
}


// ------------------------------------------------------------------------------------------------------------
// Scalar head / tail of the SVGP bound (svgp_regression.py:94-108): everything that happens on the ~10 reduced
// scalars of a sample.  The reference issues ~20 scalar NDArray operators here (and autograd twice as many on the way
// back); as separate launches they are a 100 us chain of 2 us kernels in a 1.4 ms step.  One thread per sample.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void svgp_bound_fwd_kernel(int S, T P, T B, T M, T scale, const T* __restrict__ sumr2,
                                      const T* __restrict__ trPhi, const T* __restrict__ trT,
                                      const T* __restrict__ trPhiT, const T* __restrict__ mm,
                                      const T* __restrict__ sldL, const T* __restrict__ sldLs,
                                      const T* __restrict__ noise, int64_t sNoise, const T* __restrict__ kvar,
                                      int64_t sKvar, T* __restrict__ logL, T* __restrict__ beta_out, T* __restrict__ Q_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const T nv = noise[s * sNoise], kv = kvar[s * sKvar];
    const T beta = T(1) / nv;
    const T Q = T(-0.5) * sumr2[s] - (T(0.5) * P * B) * kv - (T(0.5) * P) * (trPhiT[s] - trPhi[s]);
    const T data = beta * Q - (T(0.5) * B * P) * (T(1.8378770664093453) + Num<T>::log_(nv));       // :98-107
    const T neg_kl = P * (T(0.5) * M + sldLs[s] - sldL[s]) - (T(0.5) * P) * trT[s] - T(0.5) * mm[s];   // :94-96
    logL[s] = scale * data + neg_kl;                                                                  // :108
    beta_out[s] = beta;
    Q_out[s] = Q;
}

// coef[s][0..5] = {gP/2, g s P beta/2, g/2, g s beta/2, g s P beta, g s beta} (see mxf_svgp_bwd_assemble), plus
// gsb = g s beta, its negative, the noise-variance gradient and the Kff_diag part of the kernel-variance gradient.
template <typename T>
__global__ void svgp_coef_bwd_kernel(int S, T P, T B, T scale, const T* __restrict__ g, const T* __restrict__ beta,
                                     const T* __restrict__ Q, T* __restrict__ coef, T* __restrict__ gsb_out,
                                     T* __restrict__ neg_gsb_out, T* __restrict__ dnoise, T* __restrict__ dkvar_diag,
                                     T* __restrict__ neg_g_out, T* __restrict__ minus_one_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const T gg = g[s], b = beta[s];
    const T gsb = gg * (scale * b);
    coef[6 * s + 0] = gg * (T(0.5) * P);
    coef[6 * s + 1] = gsb * (T(0.5) * P);
    coef[6 * s + 2] = T(0.5) * gg;
    coef[6 * s + 3] = T(0.5) * gsb;
    coef[6 * s + 4] = gsb * P;
    coef[6 * s + 5] = gsb;
    gsb_out[s] = gsb;
    neg_gsb_out[s] = -gsb;
    dnoise[s] = gg * scale * (-b * b * Q[s] - (T(0.5) * B * P) * b);
    dkvar_diag[s] = -gsb * (T(0.5) * P * B);                          // Kff_diag term (:100)
    neg_g_out[s] = -gg;
    minus_one_out[s] = T(-1);
}


// ------------------------------------------------------------------------------------------------------------
// Multi-tensor parameter plumbing over the flat parameter bucket (inference_parameters.py / inference_alg.py:79-80 /
// gluon Trainer): the reference transforms every constrained parameter with its own softrelu launch on the way in and
// autograd accumulates one gradient per parameter on the way out (plus the softplus adjoints).  Here one launch
// transforms all constrained segments of the flat buffer, and one launch gathers every parameter's gradient into the
// flat gradient bucket, applying the softplus chain rule where the segment is constrained.
// ------------------------------------------------------------------------------------------------------------
constexpr int MT_MAX = 24;
template <typename T>
struct MtTable {
    const T* src[MT_MAX];       // gradient w.r.t. the (transformed) value, or nullptr (no gradient: zeros)
    int64_t off[MT_MAX];        // segment offset in the flat buffers
    int64_t n[MT_MAX];          // segment length
    T offset[MT_MAX];           // softplus offset
    int kind[MT_MAX];           // 0: identity, 1: softplus
    int count;
};

template <typename T>
__global__ void __launch_bounds__(256) mt_transform_kernel(MtTable<T> tb, const T* __restrict__ flat, T* __restrict__ tflat) {
    const int t = blockIdx.y;
    if (t >= tb.count) return;
    const T* x = flat + tb.off[t];
    T* y = tflat + tb.off[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tb.n[t]; i += stride) {
        const T v = x[i];
        if (tb.kind[t] == 1) {
            const T av = v < T(0) ? -v : v;
            y[i] = (v > T(0) ? v : T(0)) + log1p(exp(-av)) + tb.offset[t];
        } else {
            y[i] = v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) mt_pack_grads_kernel(MtTable<T> tb, const T* __restrict__ flat, T* __restrict__ gflat) {
    const int t = blockIdx.y;
    if (t >= tb.count) return;
    const T* x = flat + tb.off[t];
    const T* g = tb.src[t];
    T* out = gflat + tb.off[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tb.n[t]; i += stride) {
        T gv = g ? g[i] : T(0);
        if (g && tb.kind[t] == 1) {
            const T v = x[i];
            const T e = exp(v < T(0) ? v : -v);
            gv *= v >= T(0) ? T(1) / (T(1) + e) : e / (T(1) + e);
        }
        out[i] = gv;
    }
}

static inline int grid1d(int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)8 * kNumSMs));
}

}  // namespace mxf

using namespace mxf;

extern "C" int mxf_axpby_dev(int dtype, const void* a, const void* X, int64_t sX, const void* b, const void* Y,
                             int64_t sY, void* out, int64_t sO, int S, int64_t n, void* stream) {
    if (!X || !out || S < 0 || n < 0 || (b && !Y)) return MXF_EINVAL;
    if (S == 0 || n == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    dim3 grid(grid1d(n), S);
    MXF_DISPATCH_DTYPE(dtype, axpby_dev_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)a, (const T*)X, sX, (const T*)b, (const T*)Y, sY, (T*)out, sO, n));
    return after_launch();
}

// out[s][r][c] = a[s] X[s][r][c] + b[s] Y[s][r][c] on strided (rows x cols) views: the blocks of the solve buffers are
// combined where they lie (no contiguous copies).  a == NULL: 1, b == NULL: 1 (Y == NULL: no second term).
template <typename T>
__global__ void __launch_bounds__(256)
axpby2d_kernel(const T* __restrict__ a, const T* __restrict__ X, int64_t ldx, int64_t sX, const T* __restrict__ b,
               const T* __restrict__ Y, int64_t ldy, int64_t sY, T* __restrict__ out, int64_t ldo, int64_t sO, int rows,
               int cols) {
    const int s = blockIdx.z;
    const T av = a ? a[s] : T(1);
    const T bv = b ? b[s] : T(1);
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
            T v = av * X[(int64_t)s * sX + (int64_t)r * ldx + c];
            if (Y) v = fma(bv, Y[(int64_t)s * sY + (int64_t)r * ldy + c], v);
            out[(int64_t)s * sO + (int64_t)r * ldo + c] = v;
        }
}

extern "C" int mxf_axpby2d(int dtype, const void* a, const void* X, int64_t ldx, int64_t sX, const void* b, const void* Y,
                           int64_t ldy, int64_t sY, void* out, int64_t ldo, int64_t sO, int S, int rows, int cols,
                           void* stream) {
    if (!X || !out || S < 0 || rows < 0 || cols < 0) return MXF_EINVAL;
    if (S == 0 || rows == 0 || cols == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    dim3 grid((unsigned)std::min<int64_t>((cols + 255) / 256, 64), (unsigned)std::min(rows, 2048), (unsigned)S);
    MXF_DISPATCH_DTYPE(dtype, axpby2d_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)a, (const T*)X, ldx, sX, (const T*)b, (const T*)Y, ldy, sY, (T*)out, ldo, sO, rows,
                                  cols));
    return after_launch();
}

extern "C" int mxf_softplus_fwd(int dtype, const void* x, double offset, void* y, int64_t n, void* stream) {
    if (!x || !y || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, softplus_fwd_kernel<T><<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T)offset,
                                                                                                 (T*)y, n));
    return after_launch();
}

extern "C" int mxf_softplus_bwd(int dtype, const void* x, const void* gy, void* gx, int64_t n, void* stream) {
    if (!x || !gy || !gx || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, softplus_bwd_kernel<T><<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)x, (const T*)gy, (T*)gx, n));
    return after_launch();
}

extern "C" int mxf_svgp_bwd_assemble(int dtype, const void* Phi, const void* T_, const void* U, const void* mt,
                                     const void* v, const void* coef, void* out, int64_t ldo, int64_t sO, int S,
                                     int M, int P, void* stream) {
    if (!Phi || !T_ || !U || !mt || !v || !coef || !out || S < 0 || M < 0 || P < 0 || ldo < 3 * (int64_t)M) return MXF_EINVAL;
    if (S == 0 || M == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    dim3 grid(cdiv(M, 32), cdiv(M, 32), S);
    MXF_DISPATCH_DTYPE(dtype, svgp_bwd_assemble_kernel<T><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
                                  (const T*)Phi, (const T*)T_, (const T*)U, (const T*)mt, (const T*)v,
                                  (const T*)coef, (T*)out, ldo, sO, M, P));
    return after_launch();
}

extern "C" int mxf_svgp_bound_fwd(int dtype, int S, int P, int B, int M, double scale, const void* sumr2,
                                  const void* trPhi, const void* trT, const void* trPhiT, const void* mm,
                                  const void* sldL, const void* sldLs, const void* noise, int64_t sNoise,
                                  const void* kvar, int64_t sKvar, void* logL, void* beta, void* Q, void* stream) {
    if (!sumr2 || !trPhi || !trT || !trPhiT || !mm || !sldL || !sldLs || !noise || !kvar || !logL || !beta || !Q || S < 0)
        return MXF_EINVAL;
    if (S == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, svgp_bound_fwd_kernel<T><<<cdiv(S, 128), 128, 0, (cudaStream_t)stream>>>(
                                  S, (T)P, (T)B, (T)M, (T)scale, (const T*)sumr2, (const T*)trPhi, (const T*)trT,
                                  (const T*)trPhiT, (const T*)mm, (const T*)sldL, (const T*)sldLs, (const T*)noise,
                                  sNoise, (const T*)kvar, sKvar, (T*)logL, (T*)beta, (T*)Q));
    return after_launch();
}

extern "C" int mxf_svgp_coef_bwd(int dtype, int S, int P, int B, double scale, const void* g, const void* beta,
                                 const void* Q, void* coef, void* gsb, void* neg_gsb, void* dnoise, void* dkvar_diag,
                                 void* neg_g, void* minus_one, void* stream) {
    if (!g || !beta || !Q || !coef || !gsb || !neg_gsb || !dnoise || !dkvar_diag || !neg_g || !minus_one || S < 0)
        return MXF_EINVAL;
    if (S == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, svgp_coef_bwd_kernel<T><<<cdiv(S, 128), 128, 0, (cudaStream_t)stream>>>(
                                  S, (T)P, (T)B, (T)scale, (const T*)g, (const T*)beta, (const T*)Q, (T*)coef, (T*)gsb,
                                  (T*)neg_gsb, (T*)dnoise, (T*)dkvar_diag, (T*)neg_g, (T*)minus_one));
    return after_launch();
}

template <typename T>
static int mt_launch(int which, int count, const void* const* src, const int64_t* off, const int64_t* n, const int* kind,
                     const double* offset, const void* flat, void* out, cudaStream_t st) {
    for (int base = 0; base < count; base += MT_MAX) {
        MtTable<T> tb;
        tb.count = std::min(MT_MAX, count - base);
        int64_t nmax = 1;
        for (int i = 0; i < MT_MAX; ++i) {
            const bool on = i < tb.count;
            tb.src[i] = (on && src) ? static_cast<const T*>(src[base + i]) : nullptr;
            tb.off[i] = on ? off[base + i] : 0;
            tb.n[i] = on ? n[base + i] : 0;
            tb.kind[i] = on ? kind[base + i] : 0;
            tb.offset[i] = (on && offset) ? (T)offset[base + i] : T(0);
            if (on) {
                if (tb.n[i] < 0 || tb.off[i] < 0 || (tb.kind[i] != 0 && tb.kind[i] != 1)) return MXF_EINVAL;
                nmax = std::max(nmax, tb.n[i]);
            }
        }
        dim3 grid((unsigned)std::min<int64_t>(cdiv(nmax, 256), 2 * kNumSMs), tb.count);
        if (which == 0)
            mt_transform_kernel<T><<<grid, 256, 0, st>>>(tb, static_cast<const T*>(flat), static_cast<T*>(out));
        else
            mt_pack_grads_kernel<T><<<grid, 256, 0, st>>>(tb, static_cast<const T*>(flat), static_cast<T*>(out));
        int rc = after_launch();
        if (rc != MXF_OK) return rc;
    }
    return MXF_OK;
}

extern "C" int mxf_params_transform(int dtype, int count, const int64_t* off, const int64_t* n, const int* kind,
                                    const double* offset, const void* flat, void* tflat, void* stream) {
    if (count < 0 || (count > 0 && (!off || !n || !kind || !flat || !tflat))) return MXF_EINVAL;
    if (count == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, return mt_launch<T>(0, count, nullptr, off, n, kind, offset, flat, tflat,
                                                  (cudaStream_t)stream));
}

extern "C" int mxf_params_pack_grads(int dtype, int count, const void* const* grads, const int64_t* off,
                                     const int64_t* n, const int* kind, const void* flat, void* gflat, void* stream) {
    if (count < 0 || (count > 0 && (!grads || !off || !n || !kind || !flat || !gflat))) return MXF_EINVAL;
    if (count == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, return mt_launch<T>(1, count, grads, off, n, kind, nullptr, flat, gflat,
                                                  (cudaStream_t)stream));
}
