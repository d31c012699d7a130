// Fused dense-tanh network with SAMPLED weights: forward and adjoint.
//
// Replaces the reference's per-sample Python loop over a Gluon block
// (mxfusion/components/functions/function_evaluation.py:72-96 -> mxfusion_gluon_function.py:97-111, 166-194): with
// weights drawn from q(w) every Monte-Carlo sample s runs its own small MLP over the minibatch, S separate forward
// passes (and S autograd tapes) of ~60 MXNet operator launches each.  Here one launch evaluates
//     out[s, r, :] = W_L^s tanh( ... tanh(W_1^s x[s|0, r, :] + b_1^s) ... ) + b_L^s
// for all samples and rows, and one launch computes all weight / bias gradients (BASELINE config 4: 1-50-50-1,
// S = 3, B = 4096).  Widths <= 64, <= 4 dense layers; torch.nn.Linear layout W (out, in), y = W x + b.
//
// Mapping: a CTA owns ROWS rows of ONE sample and runs 4 x ROWS threads: thread (p, r) owns row r and every fourth group
// of 4 output units (p is warp-uniform).  The sample's weights sit in shared memory (read as warp-uniform broadcasts,
// 16 bytes at a time), activations in shared memory as [unit][row] with an odd row stride (conflict-free both for
// "thread = row" and for the "thread = 4 x 4 tile of dW" reduction).
// The adjoint recomputes the forward pass (cheaper than storing S x B x width activations in HBM), back-propagates per
// row, reduces  dW_l = delta_l act_{l-1}^T  over the CTA's rows in registers (4 x 4 tiles) and adds the CTA's partial sums
// to HBM with atomics (buffers zeroed by the caller; a weight shared by all samples has stride 0 and simply receives
// the contributions of every sample).
#include "common.cuh"

namespace mxf {

constexpr int MLP_MAXW = 64;
constexpr int MLP_MAXL = 4;
constexpr int MLP_SPLIT = 4;               // threads per row (each takes every fourth group of 4 output units)
template <typename T> struct MlpRows { static constexpr int value = 64; };     // rows per CTA (threads = 4 x rows; 32 measured slower: weight staging and atomics double)
template <> struct MlpRows<double> { static constexpr int value = 32; };       // f64: half, to fit shared memory
constexpr int MLP_LDW = MLP_MAXW;          // row stride of the staged weight matrices

template <typename T>
struct MlpArgs {
    const T* W[MLP_MAXL];
    const T* b[MLP_MAXL];
    T* dW[MLP_MAXL];
    T* db[MLP_MAXL];
    int64_t sW[MLP_MAXL], sb[MLP_MAXL];    // sample strides in elements (0: shared by all samples)
    int width[MLP_MAXL + 1];
    int n_layers;
};

template <typename T> __device__ __forceinline__ T tanh_(T x);
template <> __device__ __forceinline__ float tanh_<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double tanh_<double>(double x) { return tanh(x); }

// Stage layer l of sample s: Wt[k][j] = W[j][k] (k-major, for the forward product) and, if Wj != nullptr,
// Wj[j][k] = W[j][k] (for the back-propagation product); bias to bs.
template <typename T>
__device__ __forceinline__ void stage_layer(const MlpArgs<T>& a, int l, int s, T* Wt, T* Wj, T* bs) {
    const int in = a.width[l], out = a.width[l + 1];
    const T* W = a.W[l] + (int64_t)s * a.sW[l];
    for (int e = threadIdx.x; e < MLP_MAXW * MLP_LDW; e += blockDim.x) {
        const int j = e / MLP_LDW, k = e - j * MLP_LDW;
        const T w = (j < out && k < in) ? W[(int64_t)j * in + k] : T(0);
        Wt[k * MLP_LDW + j] = w;
        if (Wj) Wj[j * MLP_LDW + k] = w;
    }
    const T* b = a.b[l] ? a.b[l] + (int64_t)s * a.sb[l] : nullptr;
    for (int j = threadIdx.x; j < MLP_MAXW; j += blockDim.x) bs[j] = (b && j < out) ? b[j] : T(0);
}

// One dense layer for this thread's row: dst[j][r] = act(bias[j] + sum_k Wt[k][j] src[k][r]).
template <typename T, bool TANH, int MLP_RS>
__device__ __forceinline__ void dense_row(const T* __restrict__ Wt, const T* __restrict__ bs, const T* __restrict__ src,
                                          T* __restrict__ dst, int in, int out, int r, int p) {
    for (int j0 = 4 * p; j0 < out; j0 += 4 * MLP_SPLIT) {
        T acc[4] = {bs[j0], bs[j0 + 1], bs[j0 + 2], bs[j0 + 3]};
        for (int k = 0; k < in; ++k) {
            const T x = src[k * MLP_RS + r];
            const T* w = Wt + k * MLP_LDW + j0;
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fma(w[u], x, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j0 + u < out) dst[(j0 + u) * MLP_RS + r] = TANH ? tanh_<T>(acc[u]) : acc[u];
    }
}

template <typename T>
__global__ void __launch_bounds__(MLP_SPLIT * MlpRows<T>::value)
mlp_tanh_fwd_kernel(MlpArgs<T> a, const T* __restrict__ x, int64_t sx, T* __restrict__ out, int B) {
    constexpr int MLP_ROWS = MlpRows<T>::value, MLP_RS = MLP_ROWS + 1;   // odd row stride of the [unit][row] arrays
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Wt = reinterpret_cast<T*>(smem_raw);                 // [MAXW][LDW]
    T* bs = Wt + MLP_MAXW * MLP_LDW;                        // [MAXW]
    T* act0 = bs + MLP_MAXW;                                // [MAXW][RS]
    T* act1 = act0 + MLP_MAXW * MLP_RS;
    const int s = blockIdx.y, r = threadIdx.x % MLP_ROWS, p = threadIdx.x / MLP_ROWS, row = blockIdx.x * MLP_ROWS + r;
    const bool live = row < B;
    const int w0 = a.width[0];
    const T* xr = x + (int64_t)s * sx + (int64_t)row * w0;
    for (int k = p; k < w0; k += MLP_SPLIT) act0[k * MLP_RS + r] = live ? xr[k] : T(0);
    T* src = act0;
    T* dst = act1;
    for (int l = 0; l < a.n_layers; ++l) {
        __syncthreads();
        stage_layer<T>(a, l, s, Wt, nullptr, bs);
        __syncthreads();
        if (l + 1 < a.n_layers) dense_row<T, true, MLP_RS>(Wt, bs, src, dst, a.width[l], a.width[l + 1], r, p);
        else dense_row<T, false, MLP_RS>(Wt, bs, src, dst, a.width[l], a.width[l + 1], r, p);
        T* t = src; src = dst; dst = t;
    }
    __syncthreads();
    const int wo = a.width[a.n_layers];
    if (live) {
        T* o = out + ((int64_t)s * B + row) * wo;
        for (int j = p; j < wo; j += MLP_SPLIT) o[j] = src[j * MLP_RS + r];
    }
}

template <typename T>
__global__ void __launch_bounds__(MLP_SPLIT * MlpRows<T>::value)
mlp_tanh_bwd_kernel(MlpArgs<T> a, const T* __restrict__ x, int64_t sx, const T* __restrict__ gout, int B) {
    constexpr int MLP_ROWS = MlpRows<T>::value, MLP_RS = MLP_ROWS + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Wt = reinterpret_cast<T*>(smem_raw);                 // [MAXW][LDW]  k-major
    T* Wj = Wt + MLP_MAXW * MLP_LDW;                        // [MAXW][LDW]  j-major
    T* bs = Wj + MLP_MAXW * MLP_LDW;                        // [MAXW]
    T* act = bs + MLP_MAXW;                                 // [MAXL][MAXW][RS]: act[l] = input of layer l
    T* d0 = act + MLP_MAXL * MLP_MAXW * MLP_RS;             // [MAXW][RS]  delta ping
    T* d1 = d0 + MLP_MAXW * MLP_RS;                         // [MAXW][RS]  delta pong
    const int s = blockIdx.y, r = threadIdx.x % MLP_ROWS, p = threadIdx.x / MLP_ROWS, row = blockIdx.x * MLP_ROWS + r;
    const bool live = row < B;
    const int L = a.n_layers;
    // ---- forward recomputation, keeping every layer's input -------------------------------------------------
    const int w0 = a.width[0];
    const T* xr = x + (int64_t)s * sx + (int64_t)row * w0;
    for (int k = p; k < w0; k += MLP_SPLIT) act[k * MLP_RS + r] = live ? xr[k] : T(0);
    for (int l = 0; l + 1 < L; ++l) {
        __syncthreads();
        stage_layer<T>(a, l, s, Wt, nullptr, bs);
        __syncthreads();
        dense_row<T, true, MLP_RS>(Wt, bs, act + l * MLP_MAXW * MLP_RS, act + (l + 1) * MLP_MAXW * MLP_RS, a.width[l],
                                   a.width[l + 1], r, p);
    }
    // ---- delta of the output layer = upstream gradient -------------------------------------------------------
    const int wo = a.width[L];
    T* dcur = d0;
    T* dnext = d1;
    {
        const T* g = gout + ((int64_t)s * B + row) * wo;
        for (int j = p; j < MLP_MAXW; j += MLP_SPLIT) dcur[j * MLP_RS + r] = (live && j < wo) ? g[j] : T(0);
    }
    for (int l = L - 1; l >= 0; --l) {
        const int in = a.width[l], out = a.width[l + 1];
        const T* ain = act + l * MLP_MAXW * MLP_RS;
        __syncthreads();                                    // dcur complete (all rows), previous Wj no longer read
        if (l > 0) stage_layer<T>(a, l, s, Wt, Wj, bs);
        // dW_l[j][k] += sum_r dcur[j][r] ain[k][r]  (4 x 4 tile per thread), db_l[j] += sum_r dcur[j][r]
        {
            const int tid = threadIdx.x;
            const int k0 = 4 * (tid & 15);
            for (int j0 = 4 * (tid >> 4); j0 < out && k0 < in; j0 += MLP_SPLIT * MLP_ROWS / 4) {
                T acc[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[u][v] = T(0);
#pragma unroll 4
                for (int q = 0; q < MLP_ROWS; ++q) {
                    T dj[4], ak[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) dj[u] = dcur[(j0 + u) * MLP_RS + q];
#pragma unroll
                    for (int v = 0; v < 4; ++v) ak[v] = ain[(k0 + v) * MLP_RS + q];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) acc[u][v] = fma(dj[u], ak[v], acc[u][v]);
                }
                T* dW = a.dW[l] + (int64_t)s * a.sW[l];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        if (j0 + u < out && k0 + v < in) atomicAdd(dW + (int64_t)(j0 + u) * in + k0 + v, acc[u][v]);
            }
            for (int j = tid; a.db[l] && j < out; j += MLP_SPLIT * MLP_ROWS) {
                T sum = T(0);
                for (int q = 0; q < MLP_ROWS; ++q) sum += dcur[j * MLP_RS + q];
                atomicAdd(a.db[l] + (int64_t)s * a.sb[l] + j, sum);
            }
        }
        if (l == 0) break;
        __syncthreads();                                    // Wj staged
        // delta of the layer below: dnext[k][r] = (sum_j W[j][k] dcur[j][r]) (1 - ain[k][r]^2)
        for (int k0 = 4 * p; k0 < in; k0 += 4 * MLP_SPLIT) {
            T acc[4] = {T(0), T(0), T(0), T(0)};
            for (int j = 0; j < out; ++j) {
                const T d = dcur[j * MLP_RS + r];
                const T* w = Wj + j * MLP_LDW + k0;
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = fma(w[u], d, acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const T h = ain[(k0 + u) * MLP_RS + r];
                dnext[(k0 + u) * MLP_RS + r] = (k0 + u < in) ? acc[u] * (T(1) - h * h) : T(0);
            }
        }
        for (int k = ((in + 3) & ~3) + p; k < MLP_MAXW; k += MLP_SPLIT) dnext[k * MLP_RS + r] = T(0);
        T* t = dcur; dcur = dnext; dnext = t;
    }
}

template <typename T>
static int fill_args(MlpArgs<T>& a, int n_layers, const int* widths, const void* const* W, const int64_t* sW,
                     const void* const* b, const int64_t* sb, void* const* dW, void* const* db) {
    if (n_layers < 1 || n_layers > MLP_MAXL) return MXF_ENOTIMPL;
    for (int l = 0; l <= n_layers; ++l)
        if (widths[l] < 1 || widths[l] > MLP_MAXW) return MXF_ENOTIMPL;
    a.n_layers = n_layers;
    for (int l = 0; l <= MLP_MAXL; ++l) a.width[l] = l <= n_layers ? widths[l] : 0;
    for (int l = 0; l < MLP_MAXL; ++l) {
        const bool on = l < n_layers;
        a.W[l] = on ? static_cast<const T*>(W[l]) : nullptr;
        a.b[l] = (on && b) ? static_cast<const T*>(b[l]) : nullptr;
        a.dW[l] = (on && dW) ? static_cast<T*>(dW[l]) : nullptr;
        a.db[l] = (on && db) ? static_cast<T*>(db[l]) : nullptr;
        a.sW[l] = on ? sW[l] : 0;
        a.sb[l] = (on && sb) ? sb[l] : 0;
        if (on && !a.W[l]) return MXF_EINVAL;
    }
    return MXF_OK;
}

template <typename T>
static int mlp_fwd_impl(int n_layers, const int* widths, const void* x, int64_t sx, const void* const* W,
                        const int64_t* sW, const void* const* b, const int64_t* sb, void* out, int S, int B,
                        cudaStream_t st) {
    MlpArgs<T> a;
    int rc = fill_args<T>(a, n_layers, widths, W, sW, b, sb, nullptr, nullptr);
    if (rc != MXF_OK) return rc;
    constexpr int MLP_ROWS = MlpRows<T>::value, MLP_RS = MLP_ROWS + 1;
    const size_t smem = sizeof(T) * (MLP_MAXW * MLP_LDW + MLP_MAXW + 2 * MLP_MAXW * MLP_RS);
    auto k = mlp_tanh_fwd_kernel<T>;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    dim3 grid(cdiv(B, MLP_ROWS), S);
    k<<<grid, MLP_SPLIT * MLP_ROWS, smem, st>>>(a, static_cast<const T*>(x), sx, static_cast<T*>(out), B);
    return after_launch();
}

template <typename T>
static int mlp_bwd_impl(int n_layers, const int* widths, const void* x, int64_t sx, const void* const* W,
                        const int64_t* sW, const void* const* b, const int64_t* sb, const void* gout, void* const* dW,
                        void* const* db, int S, int B, cudaStream_t st) {
    MlpArgs<T> a;
    int rc = fill_args<T>(a, n_layers, widths, W, sW, b, sb, dW, db);
    if (rc != MXF_OK) return rc;
    for (int l = 0; l < n_layers; ++l)
        if (!a.dW[l]) return MXF_EINVAL;
    constexpr int MLP_ROWS = MlpRows<T>::value, MLP_RS = MLP_ROWS + 1;
    const size_t smem = sizeof(T) * (2 * MLP_MAXW * MLP_LDW + MLP_MAXW + (MLP_MAXL + 2) * MLP_MAXW * MLP_RS);
    auto k = mlp_tanh_bwd_kernel<T>;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    dim3 grid(cdiv(B, MLP_ROWS), S);
    k<<<grid, MLP_SPLIT * MLP_ROWS, smem, st>>>(a, static_cast<const T*>(x), sx, static_cast<const T*>(gout), B);
    return after_launch();
}

}  // namespace mxf

extern "C" int mxf_mlp_tanh_fwd(int dtype, int n_layers, const int* widths, const void* x, int64_t sx,
                                const void* const* W, const int64_t* sW, const void* const* b, const int64_t* sb,
                                void* out, int S, int B, void* stream) {
    if (!widths || !x || !W || !sW || !out || S < 1 || B < 1) return MXF_EINVAL;
    MXF_DISPATCH_DTYPE(dtype, return mxf::mlp_fwd_impl<T>(n_layers, widths, x, sx, W, sW, b, sb, out, S, B,
                                                          (cudaStream_t)stream));
}

extern "C" int mxf_mlp_tanh_bwd(int dtype, int n_layers, const int* widths, const void* x, int64_t sx,
                                const void* const* W, const int64_t* sW, const void* const* b, const int64_t* sb,
                                const void* gout, void* const* dW, void* const* db, int S, int B, void* stream) {
    if (!widths || !x || !W || !sW || !gout || !dW || S < 1 || B < 1) return MXF_EINVAL;
    MXF_DISPATCH_DTYPE(dtype, return mxf::mlp_bwd_impl<T>(n_layers, widths, x, sx, W, sW, b, sb, gout, dW, db, S, B,
                                                          (cudaStream_t)stream));
}
