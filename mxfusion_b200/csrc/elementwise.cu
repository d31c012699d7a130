// Reductions, small matrix utilities, the Normal-distribution MC-ELBO pieces, the fused Adam update and
// the minibatch row gather.  Each replaces a run of elementwise/reduce MXNet operators on the reference's
// hot path; the reference line ranges are given at each entry point in include/mxf_b200.h.
#include <algorithm>
#include "common.cuh"

namespace mxf {

// ------------------------------------------------------------------------------------------
// reductions: out[s] = scale * sum_{r,c} f(a[s][r][c], b[s][r][c])
// ------------------------------------------------------------------------------------------
template <typename T, int OP>
__device__ __forceinline__ T red_term(T a, T b) {
    if (OP == MXF_RED_SUM) return a;
    if (OP == MXF_RED_SUMSQ) return a * a;
    if (OP == MXF_RED_DOT) return a * b;
    if (OP == MXF_RED_SUMSQDIFF) return (a - b) * (a - b);
    return Num<T>::log_(a);
}

template <typename T, int OP>
__global__ void __launch_bounds__(256)
reduce_kernel(const T* __restrict__ a, int64_t lda, int64_t sA, const T* __restrict__ b, int64_t ldb,
              int64_t sB, int64_t rows, int64_t cols, int64_t rows_per_block, T scale, T* __restrict__ out,
              int use_atomic) {
    __shared__ T red[32];
    const int s = blockIdx.y;
    const T* as = a + (int64_t)s * sA;
    constexpr bool TWO = (OP == MXF_RED_DOT || OP == MXF_RED_SUMSQDIFF);
    const T* bs = TWO ? b + (int64_t)s * sB : nullptr;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(rows, r0 + rows_per_block);
    T acc = 0;
    // 16-byte loads when every row start is aligned (rows are the contiguous chunks chosen by the launcher)
    const bool vec = (sizeof(T) == 4) && (cols % 4 == 0) && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(as) & 15) == 0) &&
                     (!TWO || ((ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(bs) & 15) == 0)));
    for (int64_t r = r0; r < r1; ++r) {
        const T* ar = as + r * lda;
        const T* br = TWO ? bs + r * ldb : nullptr;
        if (vec) {
            T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            for (int64_t c = 4 * (int64_t)threadIdx.x; c < cols; c += 4 * (int64_t)blockDim.x) {
                const float4 x = *reinterpret_cast<const float4*>(ar + c);
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                if (TWO) y = *reinterpret_cast<const float4*>(br + c);
                a0 += red_term<T, OP>((T)x.x, (T)y.x); a1 += red_term<T, OP>((T)x.y, (T)y.y);
                a2 += red_term<T, OP>((T)x.z, (T)y.z); a3 += red_term<T, OP>((T)x.w, (T)y.w);
            }
            acc += (a0 + a1) + (a2 + a3);
        } else {
            for (int64_t c = threadIdx.x; c < cols; c += blockDim.x)
                acc += red_term<T, OP>(ar[c], TWO ? br[c] : T(0));
        }
    }
    T tot = block_sum(acc, red);
    if (threadIdx.x == 0) {
        if (use_atomic) atomicAdd(&out[s], scale * tot);
        else out[s] = scale * tot;
    }
}

template <typename T>
__global__ void fill_kernel(T* p, int64_t n, T v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <typename T>
static int reduce_impl(int op, const T* a, int64_t lda, int64_t sA, const T* b, int64_t ldb, int64_t sB, int S,
                       int64_t rows, int64_t cols, double scale, T* out, cudaStream_t st) {
    if (S == 0) return MXF_OK;
    // collapse contiguous matrices into long rows so that a block streams a contiguous range
    const bool two = (op == MXF_RED_DOT || op == MXF_RED_SUMSQDIFF);
    if (lda == cols && (!two || ldb == cols) && rows > 1) {
        const int64_t total = rows * cols;
        int64_t chunk = 8192;
        while (total % chunk != 0 && chunk > 1) chunk >>= 1;
        if (chunk >= 256) { rows = total / chunk; cols = chunk; lda = chunk; ldb = chunk; }
    }
    int64_t nblk = std::min<int64_t>(std::max<int64_t>(1, rows), (int64_t)8 * kNumSMs);
    int64_t rpb = (rows + nblk - 1) / nblk;
    if (rpb < 1) rpb = 1;
    nblk = std::max<int64_t>(1, (rows + rpb - 1) / rpb);
    const int use_atomic = nblk > 1;
    int launches = 1;
    if (use_atomic) { fill_kernel<T><<<cdiv(S, 128), 128, 0, st>>>(out, S, T(0)); ++launches; }
    dim3 grid((unsigned)nblk, S);
#define MXF_RED_CASE(OP)                                                                                      \
    case OP:                                                                                                  \
        reduce_kernel<T, OP><<<grid, 256, 0, st>>>(a, lda, sA, b, ldb, sB, rows, cols, rpb, (T)scale, out, use_atomic); \
        break;
    switch (op) {
        MXF_RED_CASE(MXF_RED_SUM)
        MXF_RED_CASE(MXF_RED_SUMSQ)
        MXF_RED_CASE(MXF_RED_DOT)
        MXF_RED_CASE(MXF_RED_SUMLOG)
        MXF_RED_CASE(MXF_RED_SUMSQDIFF)
        default: return MXF_EINVAL;
    }
#undef MXF_RED_CASE
    return after_launch(launches);
}

template <typename T>
__global__ void sumlogdiag_kernel(const T* __restrict__ A, int64_t lda, int64_t sA, int n, T* __restrict__ out) {
    __shared__ T red[32];
    const int s = blockIdx.x;
    T acc = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T d = A[(int64_t)s * sA + (int64_t)i * lda + i];
        acc += Num<T>::log_(d < T(0) ? -d : d);
    }
    T tot = block_sum(acc, red);
    if (threadIdx.x == 0) out[s] = tot;
}

template <typename T>
__global__ void add_diag_kernel(T* __restrict__ A, int64_t lda, int64_t sA, const T* __restrict__ d, int64_t sD,
                                T c, int n) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(int64_t)s * sA + (int64_t)i * lda + i] += (d ? d[(int64_t)s * sD + i] : T(0)) + c;
}

template <typename T>
__global__ void get_diag_kernel(const T* __restrict__ A, int64_t lda, int64_t sA, T* __restrict__ out, int64_t sO,
                                int n) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(int64_t)s * sO + i] = A[(int64_t)s * sA + (int64_t)i * lda + i];
}

// MODE 0: copy_ltu (symmetric from lower), 1: alpha*(A + A^T), 2: tril keep diag, 3: strictly lower, 4: transpose
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
square_tile_kernel(const T* __restrict__ A, int64_t lda, int64_t sA, T* __restrict__ out, int64_t ldo, int64_t sO,
                   int m, int n, T alpha) {
    // out is (n x m) for MODE 4, else (m x n) with m == n.  32x32 tiles, 32x8 threads.
    __shared__ T tile[32][33];
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.z;
    const T* As = A + (int64_t)s * sA;
    T* Os = out + (int64_t)s * sO;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // output tile origin: rows by.., cols bx..
    // stage the transposed source tile: tile[c][r] = A[bx + c][by + r]  (A^T restricted to the out tile)
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ai = bx + r, aj = by + threadIdx.x;       // read A[ai][aj] coalesced along aj
        const int rows_a = (MODE == 4) ? m : m, cols_a = (MODE == 4) ? n : n;
        tile[r][threadIdx.x] = (ai < rows_a && aj < cols_a) ? As[(int64_t)ai * lda + aj] : T(0);
    }
    __syncthreads();
    const int orow_lim = (MODE == 4) ? n : m, ocol_lim = (MODE == 4) ? m : n;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;          // output element (i, j)
        if (i >= orow_lim || j >= ocol_lim) continue;
        const T at = tile[threadIdx.x][r];                   // A[j][i]
        T v;
        if (MODE == 4) {
            v = at;
        } else {
            const T aa = As[(int64_t)i * lda + j];           // A[i][j]
            if (MODE == 0) v = (j <= i) ? aa : at;
            else if (MODE == 1) v = alpha * (aa + at);
            else if (MODE == 2) v = (j <= i) ? aa : T(0);
            else v = (j < i) ? aa : T(0);
        }
        Os[(int64_t)i * ldo + j] = v;
    }
}

template <typename T, int MODE>
static int square_tile_launch(const T* A, int64_t lda, int64_t sA, T* out, int64_t ldo, int64_t sO, int S, int m,
                              int n, double alpha, cudaStream_t st) {
    if (S == 0 || m == 0 || n == 0) return MXF_OK;
    const int orows = (MODE == 4) ? n : m, ocols = (MODE == 4) ? m : n;
    dim3 grid(cdiv(ocols, 32), cdiv(orows, 32), S);
    launch_pdl(square_tile_kernel<T, MODE>, grid, dim3(32, 8), (size_t)0, st, A, lda, sA, out, ldo, sO, m, n, (T)alpha);
    return after_launch();
}

// ------------------------------------------------------------------------------------------
// Normal distribution
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
normal_logpdf_sum_kernel(const T* __restrict__ x, int64_t sX, const T* __restrict__ m, int64_t sM,
                         const T* __restrict__ v, int64_t sV, int S, int64_t n, T scale, T* __restrict__ out) {
    __shared__ T red[32];
    const T c0 = T(-0.91893853320467274178);   // -0.5 log(2 pi)
    T acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = (sizeof(T) == 4) && (n % 4 == 0) && (sX % 4 == 0) && (sM % 4 == 0) && (sV % 4 == 0) &&
                     (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0);
    if (vec) {
        T a4[4] = {T(0), T(0), T(0), T(0)};
        for (int64_t i = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * stride) {
            float4 mv, vv;
            if (sM == 0) mv = *reinterpret_cast<const float4*>(m + i);
            if (sV == 0) vv = *reinterpret_cast<const float4*>(v + i);
            for (int s = 0; s < S; ++s) {
                const float4 xv = *reinterpret_cast<const float4*>(x + s * sX + i);
                if (sM != 0) mv = *reinterpret_cast<const float4*>(m + s * sM + i);
                if (sV != 0) vv = *reinterpret_cast<const float4*>(v + s * sV + i);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w}, vs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float d = xs[u] - ms[u];
                    a4[u] += (T)((float)c0 - 0.5f * logf(vs[u]) - d * d / (2.0f * vs[u]));
                }
            }
        }
        acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    } else
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        for (int s = 0; s < S; ++s) {
            const T xv = x[s * sX + i], mv = m[s * sM + i], vv = v[s * sV + i];
            const T d = xv - mv;
            acc += c0 - T(0.5) * Num<T>::log_(vv) - d * d / (T(2) * vv);
        }
    }
    T tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, tot * scale / T(S));
}

template <typename T>
__global__ void __launch_bounds__(256)
normal_logpdf_sum_bwd_kernel(const T* __restrict__ x, int64_t sX, const T* __restrict__ m, int64_t sM,
                             const T* __restrict__ v, int64_t sV, int S, int64_t n, T scale,
                             const T* __restrict__ gout, T* __restrict__ gx, T* __restrict__ gm,
                             T* __restrict__ gv) {
    const T g = gout[0] * scale / T(S);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T ax = 0, am = 0, av = 0;
        for (int s = 0; s < S; ++s) {
            const T xv = x[s * sX + i], mv = m[s * sM + i], vv = v[s * sV + i];
            const T d = xv - mv;
            const T dx = -d / vv * g;                                         // d/dx
            const T dv = (T(-0.5) / vv + T(0.5) * d * d / (vv * vv)) * g;     // d/dv
            if (gx) { if (sX) gx[s * n + i] = dx; else ax += dx; }
            if (gm) { if (sM) gm[s * n + i] = -dx; else am -= dx; }
            if (gv) { if (sV) gv[s * n + i] = dv; else av += dv; }
        }
        if (gx && !sX) gx[i] = ax;
        if (gm && !sM) gm[i] = am;
        if (gv && !sV) gv[i] = av;
    }
}

// Philox4x32-10 (Salmon et al. 2011), counter = (idx_lo, idx_hi, offset_lo, offset_hi), key = seed.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ void philox4x32_10(uint64_t idx, uint64_t offset, uint64_t seed, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

template <typename T>
__global__ void __launch_bounds__(256)
normal_reparam_kernel(const T* __restrict__ eps, const T* __restrict__ m, int64_t sM, const T* __restrict__ v,
                      int64_t sV, int S, int64_t n, uint64_t seed, uint64_t offset, const int* __restrict__ step_counter,
                      T* __restrict__ w, T* __restrict__ eps_out) {
    // A launch captured in a CUDA graph replays with the SAME host-side (seed, offset): the optimiser's device step
    // counter goes into the high word of the Philox counter so that every replayed step draws fresh noise.
    if (step_counter) offset += (uint64_t)(uint32_t)(*step_counter) << 32;
    const int64_t total = (int64_t)S * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (eps) {
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
            const int64_t s = e / n, i = e - s * n;
            w[e] = fma(eps[e], Num<T>::sqrt_(v[s * sV + i]), m[s * sM + i]);
        }
        return;
    }
    // four normals per Philox call: thread handles elements 4q .. 4q+3 of the flattened (S, n) array
    auto normals4 = [&](uint64_t q, float (&z)[4]) {
        uint32_t r[4];
        philox4x32_10(q, offset, seed, r);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float u1 = ((float)r[2 * h] + 1.0f) * 2.3283064365386963e-10f;       // (0,1]
            const float u2 = (float)r[2 * h + 1] * 2.3283064365386963e-10f;            // [0,1)
            const float rad = sqrtf(-2.0f * __logf(u1));
            float sn, cs;
            __sincosf(6.283185307179586f * u2, &sn, &cs);
            z[2 * h] = rad * cs;
            z[2 * h + 1] = rad * sn;
        }
    };
    const bool vec = (sizeof(T) == 4) && (n % 4 == 0) && (sM % 4 == 0) && (sV % 4 == 0) &&
                     (((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(w) |
                        reinterpret_cast<uintptr_t>(eps_out)) & 15) == 0);
    if (vec) {
        // rows are whole numbers of quads: no division; mean / variance are loaded once when shared by the samples
        const int64_t nq = n / 4;
        for (int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += stride) {
            float4 mv, sd;
            if (sM == 0) mv = *reinterpret_cast<const float4*>(m + 4 * qi);
            if (sV == 0) {
                const float4 vv = *reinterpret_cast<const float4*>(v + 4 * qi);
                sd = make_float4(sqrtf(vv.x), sqrtf(vv.y), sqrtf(vv.z), sqrtf(vv.w));
            }
            for (int s = 0; s < S; ++s) {
                float z[4];
                normals4((uint64_t)(s * nq + qi), z);
                if (sM != 0) mv = *reinterpret_cast<const float4*>(m + s * sM + 4 * qi);
                if (sV != 0) {
                    const float4 vv = *reinterpret_cast<const float4*>(v + s * sV + 4 * qi);
                    sd = make_float4(sqrtf(vv.x), sqrtf(vv.y), sqrtf(vv.z), sqrtf(vv.w));
                }
                const float4 o = make_float4(fmaf(z[0], sd.x, mv.x), fmaf(z[1], sd.y, mv.y), fmaf(z[2], sd.z, mv.z),
                                             fmaf(z[3], sd.w, mv.w));
                __stcs(reinterpret_cast<float4*>(w + s * n + 4 * qi), o);
                if (eps_out) __stcs(reinterpret_cast<float4*>(eps_out + s * n + 4 * qi), make_float4(z[0], z[1], z[2], z[3]));
            }
        }
        return;
    }
    const int64_t quads = (total + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
        float z[4];
        normals4((uint64_t)q, z);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int64_t e = 4 * q + t;
            if (e < total) {
                const int64_t s = e / n, i = e - s * n;
                const T ev = (T)z[t];
                if (eps_out) eps_out[e] = ev;
                w[e] = fma(ev, Num<T>::sqrt_(v[s * sV + i]), m[s * sM + i]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Adam, gather
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
adam_kernel(T* __restrict__ w, const T* __restrict__ g, T* __restrict__ m, T* __restrict__ v, int64_t n,
            double lr, double b1, double b2, double eps, double rescale, const int* __restrict__ step_count) {
    const int t = step_count[0] + 1;
    const double lr_t = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
    const T lrt = (T)lr_t, B1 = (T)b1, B2 = (T)b2, E = (T)eps, R = (T)rescale;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T gi = g[i] * R;
        const T mi = B1 * m[i] + (T(1) - B1) * gi;
        const T vi = B2 * v[i] + (T(1) - B2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        w[i] -= lrt * mi / (Num<T>::sqrt_(vi) + E);
    }
}

// mx.optimizer.SGD (the Trainer('sgd') of grad_based_inference.py:67): w -= lr * rescale * g, or with momentum
// mom = momentum * mom - lr * rescale * g; w += mom  (weight decay 0, as the reference never sets it)
template <typename T>
__global__ void __launch_bounds__(256)
sgd_kernel(T* __restrict__ w, const T* __restrict__ g, T* __restrict__ mom, int64_t n, double lr, double momentum, double rescale) {
    const T L = (T)(lr * rescale), MU = (T)momentum;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mom) {
            const T mi = MU * mom[i] - L * g[i];
            mom[i] = mi;
            w[i] += mi;
        } else {
            w[i] -= L * g[i];
        }
    }
}

__global__ void incr_kernel(int* c) { c[0] += 1; }

template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const T* __restrict__ src, int64_t cols, const int64_t* __restrict__ idx,
                   const int64_t* __restrict__ off, int64_t rows, T* __restrict__ out) {
    const int64_t o = off ? off[0] : 0;
    const int64_t total = rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t r = e / cols, c = e - r * cols;
        out[e] = src[idx[o + r] * cols + c];
    }
}


// Adjoint of the reparameterised draw w = eps sqrt(v) + m:  gm = gw (summed over the samples when m is shared),
// gv = gw eps / (2 sqrt(v)) (summed likewise).  One launch instead of the five elementwise / reduction operators autograd
// would issue per weight tensor (normal.py:89-92 differentiated).
template <typename T>
__global__ void __launch_bounds__(256)
normal_reparam_bwd_kernel(const T* __restrict__ gw, const T* __restrict__ eps, const T* __restrict__ v, int64_t sM,
                          int64_t sV, int S, int64_t n, T* __restrict__ gm, T* __restrict__ gv) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T am = 0, av = 0;
        for (int s = 0; s < S; ++s) {
            const T g = gw[s * n + i];
            const T dv = g * eps[s * n + i] * (T(0.5) * Num<T>::rsqrt_(v[s * sV + i]));
            if (gm) { if (sM) gm[s * n + i] = g; else am += g; }
            if (gv) { if (sV) gv[s * n + i] = dv; else av += dv; }
        }
        if (gm && !sM) gm[i] = am;
        if (gv && !sV) gv[i] = av;
    }
}


// ------------------------------------------------------------------------------------------------------------
// Multi-tensor Normal log-density: all Normal factors of a factor-graph walk (the priors and the variational factors of
// every weight tensor of a mean-field BNN, plus the likelihood; factor_graph.py:192-238 sums them one operator chain at a
// time) in ONE launch forwards and ONE launch backwards.  Entry t: x, m, v with sample strides (0 = shared by the
// samples) and a per-operand "scalar" flag (a constant prior mean / variance of shape (1,) is read in place instead of
// being broadcast into a weight-shaped copy).  out[0] += sum_t scale_t / S_t sum_{s,i} log N(x | m, v).
// ------------------------------------------------------------------------------------------------------------
constexpr int NL_MAX = 16;
template <typename T>
struct NlTable {
    const T* x[NL_MAX]; const T* m[NL_MAX]; const T* v[NL_MAX];
    T* gx[NL_MAX]; T* gm[NL_MAX]; T* gv[NL_MAX];
    int64_t sX[NL_MAX], sM[NL_MAX], sV[NL_MAX], n[NL_MAX];
    T scale[NL_MAX];
    int S[NL_MAX];
    unsigned char scalar[NL_MAX];        // bit 0: x, bit 1: m, bit 2: v is a single element per sample
    int count;
};

template <typename T>
__global__ void __launch_bounds__(256) normal_logpdf_multi_kernel(NlTable<T> tb, T* __restrict__ out) {
    __shared__ T red[32];
    const int t = blockIdx.y;
    if (t >= tb.count) return;
    const T* x = tb.x[t]; const T* m = tb.m[t]; const T* v = tb.v[t];
    const int64_t sX = tb.sX[t], sM = tb.sM[t], sV = tb.sV[t], n = tb.n[t];
    const int S = tb.S[t];
    const int fx = tb.scalar[t] & 1, fm = tb.scalar[t] & 2, fv = tb.scalar[t] & 4;
    const T c0 = T(-0.91893853320467274178);
    T acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        for (int s = 0; s < S; ++s) {
            const T xv = x[s * sX + (fx ? 0 : i)], mv = m[s * sM + (fm ? 0 : i)], vv = v[s * sV + (fv ? 0 : i)];
            const T d = xv - mv;
            acc += c0 - T(0.5) * Num<T>::log_(vv) - d * d / (T(2) * vv);
        }
    }
    const T tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, tot * tb.scale[t] / T(S));
}

// Adjoint: per-entry gx / gm / gv (nullptr = not needed).  Scalar operands never receive a gradient here (the caller
// keeps such entries out of the batch).
template <typename T>
__global__ void __launch_bounds__(256) normal_logpdf_multi_bwd_kernel(NlTable<T> tb, const T* __restrict__ gout) {
    const int t = blockIdx.y;
    if (t >= tb.count) return;
    const T* x = tb.x[t]; const T* m = tb.m[t]; const T* v = tb.v[t];
    T* gx = tb.gx[t]; T* gm = tb.gm[t]; T* gv = tb.gv[t];
    const int64_t sX = tb.sX[t], sM = tb.sM[t], sV = tb.sV[t], n = tb.n[t];
    const int S = tb.S[t];
    const int fx = tb.scalar[t] & 1, fm = tb.scalar[t] & 2, fv = tb.scalar[t] & 4;
    const T g = gout[0] * tb.scale[t] / T(S);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T ax = 0, am = 0, av = 0;
        for (int s = 0; s < S; ++s) {
            const T xv = x[s * sX + (fx ? 0 : i)], mv = m[s * sM + (fm ? 0 : i)], vv = v[s * sV + (fv ? 0 : i)];
            const T d = xv - mv;
            const T dx = -d / vv * g;
            const T dv = (T(-0.5) / vv + T(0.5) * d * d / (vv * vv)) * g;
            if (gx) { if (sX) gx[s * n + i] = dx; else ax += dx; }
            if (gm) { if (sM) gm[s * n + i] = -dx; else am -= dx; }
            if (gv) { if (sV) gv[s * n + i] = dv; else av += dv; }
        }
        if (gx && !sX) gx[i] = ax;
        if (gm && !sM) gm[i] = am;
        if (gv && !sV) gv[i] = av;
    }
}


// ------------------------------------------------------------------------------------------------------------
// Multi-tensor reparameterised draw and its adjoint: the draws of all (independent) Normal factors of a posterior graph
// walk -- one weight tensor each in a mean-field BNN -- in ONE launch each way.  Entry t has its own Philox stream
// (seed, offset_t); the device step counter is mixed in as for the single-tensor kernel.
// ------------------------------------------------------------------------------------------------------------
constexpr int RP_MAX = 16;
template <typename T>
struct RpTable {
    const T* m[RP_MAX]; const T* v[RP_MAX];
    T* w[RP_MAX]; T* eps[RP_MAX];                 // forward: outputs; adjoint: w = upstream gradient (read), eps (read)
    T* gm[RP_MAX]; T* gv[RP_MAX];
    int64_t sM[RP_MAX], sV[RP_MAX], n[RP_MAX];
    unsigned long long offset[RP_MAX];
    int S[RP_MAX];
    int count;
};

template <typename T>
__global__ void __launch_bounds__(256)
normal_reparam_multi_kernel(RpTable<T> tb, uint64_t seed, const int* __restrict__ step_counter) {
    const int t = blockIdx.y;
    if (t >= tb.count) return;
    uint64_t offset = tb.offset[t];
    if (step_counter) offset += (uint64_t)(uint32_t)(*step_counter) << 32;
    const T* m = tb.m[t]; const T* v = tb.v[t];
    T* w = tb.w[t]; T* eps_out = tb.eps[t];
    const int64_t sM = tb.sM[t], sV = tb.sV[t], n = tb.n[t];
    const int64_t total = (int64_t)tb.S[t] * n, quads = (total + 3) / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
        uint32_t r[4];
        philox4x32_10((uint64_t)q, offset, seed, r);
        float z[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float u1 = ((float)r[2 * h] + 1.0f) * 2.3283064365386963e-10f;
            const float u2 = (float)r[2 * h + 1] * 2.3283064365386963e-10f;
            const float rad = sqrtf(-2.0f * __logf(u1));
            float sn, cs;
            __sincosf(6.283185307179586f * u2, &sn, &cs);
            z[2 * h] = rad * cs;
            z[2 * h + 1] = rad * sn;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t e = 4 * q + u;
            if (e < total) {
                const int64_t sidx = e / n, i = e - sidx * n;
                const T ev = (T)z[u];
                if (eps_out) eps_out[e] = ev;
                w[e] = fma(ev, Num<T>::sqrt_(v[sidx * sV + i]), m[sidx * sM + i]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) normal_reparam_multi_bwd_kernel(RpTable<T> tb) {
    const int t = blockIdx.y;
    if (t >= tb.count) return;
    const T* gw = tb.w[t]; const T* eps = tb.eps[t]; const T* v = tb.v[t];
    T* gm = tb.gm[t]; T* gv = tb.gv[t];
    const int64_t sM = tb.sM[t], sV = tb.sV[t], n = tb.n[t];
    const int S = tb.S[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T am = 0, av = 0;
        for (int s = 0; s < S; ++s) {
            const T g = gw[s * n + i];
            const T dv = g * eps[s * n + i] * (T(0.5) * Num<T>::rsqrt_(v[s * sV + i]));
            if (gm) { if (sM) gm[s * n + i] = g; else am += g; }
            if (gv) { if (sV) gv[s * n + i] = dv; else av += dv; }
        }
        if (gm && !sM) gm[i] = am;
        if (gv && !sV) gv[i] = av;
    }
}

static inline int grid_for(int64_t n, int threads = 256) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)8 * kNumSMs));
}

}  // namespace mxf

using namespace mxf;

std::atomic<uint64_t> mxf::g_launches{0};

extern "C" int mxf_version(void) { return 100; }
extern "C" uint64_t mxf_launch_count(void) { return g_launches.load(); }

extern "C" int mxf_reduce(int op, int dtype, const void* a, int64_t lda, int64_t sA, const void* b, int64_t ldb,
                          int64_t sB, int S, int64_t rows, int64_t cols, double scale, void* out, void* stream) {
    if (!a || !out || rows < 0 || cols < 0 || S < 0 || ((op == MXF_RED_DOT || op == MXF_RED_SUMSQDIFF) && !b))
        return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return reduce_impl<T>(op, (const T*)a, lda, sA, (const T*)b, ldb, sB, S, rows, cols,
                                                    scale, (T*)out, (cudaStream_t)stream));
}

extern "C" int mxf_sumlogdiag(int dtype, const void* A, int64_t lda, int64_t sA, int S, int n, void* out,
                              void* stream) {
    if (!A || !out || n < 0 || S < 0) return MXF_EINVAL;
    if (S == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, sumlogdiag_kernel<T><<<S, 256, 0, (cudaStream_t)stream>>>((const T*)A, lda, sA, n,
                                                                                       (T*)out));
    return after_launch();
}

extern "C" int mxf_add_diag(int dtype, void* A, int64_t lda, int64_t sA, const void* d, int64_t sD, double c,
                            int S, int n, void* stream) {
    if (!A || n < 0 || S < 0) return MXF_EINVAL;
    if (S == 0 || n == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    dim3 grid(cdiv(n, 256), S);
    MXF_DISPATCH_DTYPE(dtype, add_diag_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((T*)A, lda, sA,
                                                                                        (const T*)d, sD, (T)c, n));
    return after_launch();
}

extern "C" int mxf_get_diag(int dtype, const void* A, int64_t lda, int64_t sA, void* out, int64_t sO, int S, int n,
                            void* stream) {
    if (!A || !out || n < 0 || S < 0) return MXF_EINVAL;
    if (S == 0 || n == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    dim3 grid(cdiv(n, 256), S);
    MXF_DISPATCH_DTYPE(dtype, get_diag_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)A, lda, sA,
                                                                                        (T*)out, sO, n));
    return after_launch();
}

extern "C" int mxf_copy_ltu(int dtype, const void* P, int64_t ldp, int64_t sP, void* out, int64_t ldo, int64_t sO,
                            int S, int n, void* stream) {
    if (!P || !out || P == out) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return (square_tile_launch<T, 0>((const T*)P, ldp, sP, (T*)out, ldo, sO, S, n, n, 1.0,
                                                               (cudaStream_t)stream)));
}

// out = symmetric matrix whose lower triangle (diagonal included) is the SUM of the lower triangles of the `parts`
// matrices P[0..parts) (stride sPart): the split-K partial products of Phi = A A^T (svgp_regression.py:89-90 folded,
// ops.py) are added and mirrored in one pass instead of a reduction launch + a mirror launch.
template <typename T>
__global__ void __launch_bounds__(256)
copy_ltu_sum_kernel(const T* __restrict__ P, int64_t ldp, int64_t sPart, int parts, T* __restrict__ out, int64_t ldo, int n) {
    __shared__ T tile[32][33];
    pdl_launch_dependents();
    pdl_wait();
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;       // output tile: rows by.., cols bx..
    if (bx > by + 31) {
        // strictly-upper tile: the transposed lower tile (rows bx.., cols by..)
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int ai = bx + r, aj = by + threadIdx.x;
            T v = T(0);
            if (ai < n && aj < n)
                for (int g = 0; g < parts; ++g) v += P[(int64_t)g * sPart + (int64_t)ai * ldp + aj];
            tile[r][threadIdx.x] = v;
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int i = by + r, j = bx + threadIdx.x;
            if (i < n && j < n) out[(int64_t)i * ldo + j] = tile[threadIdx.x][r];
        }
        return;
    }
    // lower or diagonal tile
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;
        T v = T(0);
        if (i < n && j < n && j <= i)
            for (int g = 0; g < parts; ++g) v += P[(int64_t)g * sPart + (int64_t)i * ldp + j];
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i >= n || j >= n) continue;
        out[(int64_t)i * ldo + j] = (j <= i) ? tile[r][threadIdx.x] : tile[threadIdx.x][r];      // diagonal tile: mirror in place
    }
}

extern "C" int mxf_copy_ltu_sum(int dtype, const void* P, int64_t ldp, int64_t sPart, int parts, void* out, int64_t ldo,
                                int n, void* stream) {
    if (!P || !out || P == out || parts <= 0 || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, {
        dim3 grid(cdiv(n, 32), cdiv(n, 32));
        launch_pdl(copy_ltu_sum_kernel<T>, grid, dim3(32, 8), (size_t)0, (cudaStream_t)stream, (const T*)P, ldp, sPart, parts,
                   (T*)out, ldo, n);
        return after_launch();
    });
}

// Strided batched 2-D copy (src == NULL: zero fill): the small right-hand-side blocks that ride along in the solve
// buffers ([Kuf | Ls | mu], ops.py) without a tensor-library launch.
template <typename T>
__global__ void __launch_bounds__(256)
copy2d_kernel(const T* __restrict__ src, int64_t lds, int64_t sS, T* __restrict__ dst, int64_t ldd, int64_t sD, int rows,
              int cols) {
    pdl_launch_dependents();
    pdl_wait();
    const int s = blockIdx.z;
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
            dst[(int64_t)s * sD + (int64_t)r * ldd + c] = src ? src[(int64_t)s * sS + (int64_t)r * lds + c] : T(0);
}

extern "C" int mxf_copy2d(int dtype, const void* src, int64_t lds, int64_t sS, void* dst, int64_t ldd, int64_t sD, int S,
                          int rows, int cols, void* stream) {
    if (!dst || S < 0 || rows < 0 || cols < 0) return MXF_EINVAL;
    if (S == 0 || rows == 0 || cols == 0) return MXF_OK;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, {
        dim3 grid(std::min(cdiv(cols, 256), 64), std::min(rows, 1024), S);
        launch_pdl(copy2d_kernel<T>, grid, dim3(256), (size_t)0, (cudaStream_t)stream, (const T*)src, lds, sS, (T*)dst, ldd, sD,
                   rows, cols);
        return after_launch();
    });
}

// acc[0] = max(acc[0], max_i info[i]): potrf's device-side failure record (first non-positive pivot per sample) folded
// into one int per device without a host synchronisation; the training loop reads it every few steps.
__global__ void info_max_kernel(int* __restrict__ acc, const int* __restrict__ info, int n) {
    int m = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = max(m, info[i] != 0 ? abs(info[i]) : 0);
    if (m != 0) atomicMax(acc, m);
}

extern "C" int mxf_info_max(int* acc, const int* info, int n, void* stream) {
    if (!acc || !info || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    info_max_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(acc, info, n);
    return after_launch();
}

extern "C" int mxf_symmetrize(int dtype, double alpha, const void* A, int64_t lda, int64_t sA, void* out,
                              int64_t ldo, int64_t sO, int S, int n, void* stream) {
    if (!A || !out || A == out) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return (square_tile_launch<T, 1>((const T*)A, lda, sA, (T*)out, ldo, sO, S, n, n,
                                                               alpha, (cudaStream_t)stream)));
}

extern "C" int mxf_tril(int dtype, int mode, const void* A, int64_t lda, int64_t sA, void* out, int64_t ldo,
                        int64_t sO, int S, int n, void* stream) {
    if (!A || !out || A == out || (mode != 0 && mode != 1)) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    if (mode == 0)
        MXF_DISPATCH_DTYPE(dtype, return (square_tile_launch<T, 2>((const T*)A, lda, sA, (T*)out, ldo, sO, S, n, n,
                                                                   1.0, (cudaStream_t)stream)));
    MXF_DISPATCH_DTYPE(dtype, return (square_tile_launch<T, 3>((const T*)A, lda, sA, (T*)out, ldo, sO, S, n, n, 1.0,
                                                               (cudaStream_t)stream)));
}

extern "C" int mxf_transpose(int dtype, const void* A, int64_t lda, int64_t sA, void* out, int64_t ldo, int64_t sO,
                             int S, int m, int n, void* stream) {
    if (!A || !out || A == out) return MXF_EINVAL;
    if (S > 65535) return MXF_ENOTIMPL;
    MXF_DISPATCH_DTYPE(dtype, return (square_tile_launch<T, 4>((const T*)A, lda, sA, (T*)out, ldo, sO, S, m, n, 1.0,
                                                               (cudaStream_t)stream)));
}

extern "C" int mxf_normal_logpdf_sum(int dtype, const void* x, int64_t sX, const void* m, int64_t sM,
                                     const void* v, int64_t sV, int S, int64_t n, double scale, void* out,
                                     void* stream) {
    if (!x || !m || !v || !out || S <= 0 || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, normal_logpdf_sum_kernel<T><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)x, sX, (const T*)m, sM, (const T*)v, sV, S, n, (T)scale, (T*)out));
    return after_launch();
}

extern "C" int mxf_normal_logpdf_sum_bwd(int dtype, const void* x, int64_t sX, const void* m, int64_t sM,
                                         const void* v, int64_t sV, int S, int64_t n, double scale,
                                         const void* gout, void* gx, void* gm, void* gv, void* stream) {
    if (!x || !m || !v || !gout || S <= 0 || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, normal_logpdf_sum_bwd_kernel<T><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)x, sX, (const T*)m, sM, (const T*)v, sV, S, n, (T)scale,
                                  (const T*)gout, (T*)gx, (T*)gm, (T*)gv));
    return after_launch();
}

extern "C" int mxf_normal_reparam(int dtype, const void* eps, const void* m, int64_t sM, const void* v, int64_t sV,
                                  int S, int64_t n, uint64_t seed, uint64_t offset, const int* step_counter, void* w,
                                  void* eps_out, void* stream) {
    if (!m || !v || !w || S <= 0 || n < 0) return MXF_EINVAL;
    if (n == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, normal_reparam_kernel<T><<<grid_for((int64_t)S * n), 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)eps, (const T*)m, sM, (const T*)v, sV, S, n, seed, offset, step_counter,
                                  (T*)w, (T*)eps_out));
    return after_launch();
}

extern "C" int mxf_adam_step(int dtype, void* w, const void* g, void* m, void* v, int64_t n, double lr,
                             double beta1, double beta2, double eps, double rescale, int* step_count,
                             void* stream) {
    if (!w || !g || !m || !v || !step_count || n < 0) return MXF_EINVAL;
    if (n > 0)
        MXF_DISPATCH_DTYPE(dtype, adam_kernel<T><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
                                      (T*)w, (const T*)g, (T*)m, (T*)v, n, lr, beta1, beta2, eps, rescale,
                                      step_count));
    incr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_count);
    return after_launch(2);
}

extern "C" int mxf_sgd_step(int dtype, void* w, const void* g, void* mom, int64_t n, double lr, double momentum,
                            double rescale, int* step_count, void* stream) {
    if (!w || !g || !step_count || n < 0 || (momentum != 0.0 && !mom)) return MXF_EINVAL;
    if (n > 0)
        MXF_DISPATCH_DTYPE(dtype, sgd_kernel<T><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
                                      (T*)w, (const T*)g, momentum != 0.0 ? (T*)mom : (T*)nullptr, n, lr, momentum, rescale));
    incr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_count);
    return after_launch(2);
}

extern "C" int mxf_gather_rows(int dtype, const void* src, int64_t cols, const int64_t* idx, const int64_t* off,
                               int64_t rows, void* out, void* stream) {
    if (!src || !idx || !out || rows < 0 || cols <= 0) return MXF_EINVAL;
    if (rows == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, gather_rows_kernel<T><<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)src, cols, idx, off, rows, (T*)out));
    return after_launch();
}

extern "C" int mxf_normal_reparam_bwd(int dtype, const void* gw, const void* eps, const void* v, int64_t sM, int64_t sV,
                                      int S, int64_t n, void* gm, void* gv, void* stream) {
    if (!gw || !eps || !v || S <= 0 || n < 0) return MXF_EINVAL;
    if (n == 0 || (!gm && !gv)) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, normal_reparam_bwd_kernel<T><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
                                  (const T*)gw, (const T*)eps, (const T*)v, sM, sV, S, n, (T*)gm, (T*)gv));
    return after_launch();
}

template <typename T>
static int nl_multi(int count, const void* const* x, const void* const* m, const void* const* v, const int64_t* sX,
                    const int64_t* sM, const int64_t* sV, const int64_t* n, const int* S, const int* scalar,
                    const double* scale, void* out, const void* gout, void* const* gx, void* const* gm, void* const* gv,
                    cudaStream_t st) {
    for (int base = 0; base < count; base += NL_MAX) {
        NlTable<T> tb;
        tb.count = std::min(NL_MAX, count - base);
        int64_t nmax = 1;
        for (int i = 0; i < NL_MAX; ++i) {
            const bool on = i < tb.count;
            const int k = base + i;
            tb.x[i] = on ? static_cast<const T*>(x[k]) : nullptr;
            tb.m[i] = on ? static_cast<const T*>(m[k]) : nullptr;
            tb.v[i] = on ? static_cast<const T*>(v[k]) : nullptr;
            tb.gx[i] = (on && gx) ? static_cast<T*>(gx[k]) : nullptr;
            tb.gm[i] = (on && gm) ? static_cast<T*>(gm[k]) : nullptr;
            tb.gv[i] = (on && gv) ? static_cast<T*>(gv[k]) : nullptr;
            tb.sX[i] = on ? sX[k] : 0; tb.sM[i] = on ? sM[k] : 0; tb.sV[i] = on ? sV[k] : 0;
            tb.n[i] = on ? n[k] : 0;
            tb.S[i] = on ? S[k] : 1;
            tb.scalar[i] = on ? (unsigned char)scalar[k] : 0;
            tb.scale[i] = on ? (T)scale[k] : T(0);
            if (on) {
                if (!tb.x[i] || !tb.m[i] || !tb.v[i] || tb.n[i] < 0 || tb.S[i] < 1) return MXF_EINVAL;
                if (gout && ((tb.gx[i] && (tb.scalar[i] & 1)) || (tb.gm[i] && (tb.scalar[i] & 2)) ||
                             (tb.gv[i] && (tb.scalar[i] & 4))))
                    return MXF_ENOTIMPL;
                nmax = std::max(nmax, tb.n[i]);
            }
        }
        dim3 grid((unsigned)std::min<int64_t>(cdiv(nmax, 256), 2 * kNumSMs), tb.count);
        if (!gout) normal_logpdf_multi_kernel<T><<<grid, 256, 0, st>>>(tb, static_cast<T*>(out));
        else normal_logpdf_multi_bwd_kernel<T><<<grid, 256, 0, st>>>(tb, static_cast<const T*>(gout));
        int rc = after_launch();
        if (rc != MXF_OK) return rc;
    }
    return MXF_OK;
}

extern "C" int mxf_normal_logpdf_multi(int dtype, int count, const void* const* x, const void* const* m,
                                       const void* const* v, const int64_t* sX, const int64_t* sM, const int64_t* sV,
                                       const int64_t* n, const int* S, const int* scalar, const double* scale, void* out,
                                       void* stream) {
    if (count < 0 || (count > 0 && (!x || !m || !v || !sX || !sM || !sV || !n || !S || !scalar || !scale || !out)))
        return MXF_EINVAL;
    if (count == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, return nl_multi<T>(count, x, m, v, sX, sM, sV, n, S, scalar, scale, out, nullptr, nullptr,
                                                 nullptr, nullptr, (cudaStream_t)stream));
}

extern "C" int mxf_normal_logpdf_multi_bwd(int dtype, int count, const void* const* x, const void* const* m,
                                           const void* const* v, const int64_t* sX, const int64_t* sM, const int64_t* sV,
                                           const int64_t* n, const int* S, const int* scalar, const double* scale,
                                           const void* gout, void* const* gx, void* const* gm, void* const* gv,
                                           void* stream) {
    if (count < 0 || (count > 0 && (!x || !m || !v || !sX || !sM || !sV || !n || !S || !scalar || !scale || !gout ||
                                    !gx || !gm || !gv)))
        return MXF_EINVAL;
    if (count == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, return nl_multi<T>(count, x, m, v, sX, sM, sV, n, S, scalar, scale, nullptr, gout, gx, gm,
                                                 gv, (cudaStream_t)stream));
}

template <typename T>
static int rp_multi(int bwd, int count, const void* const* m, const void* const* v, const int64_t* sM, const int64_t* sV,
                    const int64_t* n, const int* S, uint64_t seed, const uint64_t* offset, const int* step_counter,
                    void* const* w, void* const* eps, void* const* gm, void* const* gv, cudaStream_t st) {
    for (int base = 0; base < count; base += RP_MAX) {
        RpTable<T> tb;
        tb.count = std::min(RP_MAX, count - base);
        int64_t work = 1;
        for (int i = 0; i < RP_MAX; ++i) {
            const bool on = i < tb.count;
            const int k = base + i;
            tb.m[i] = (on && m) ? static_cast<const T*>(m[k]) : nullptr;
            tb.v[i] = on ? static_cast<const T*>(v[k]) : nullptr;
            tb.w[i] = on ? static_cast<T*>(w[k]) : nullptr;
            tb.eps[i] = (on && eps) ? static_cast<T*>(eps[k]) : nullptr;
            tb.gm[i] = (on && gm) ? static_cast<T*>(gm[k]) : nullptr;
            tb.gv[i] = (on && gv) ? static_cast<T*>(gv[k]) : nullptr;
            tb.sM[i] = on ? sM[k] : 0; tb.sV[i] = on ? sV[k] : 0; tb.n[i] = on ? n[k] : 0;
            tb.S[i] = on ? S[k] : 1;
            tb.offset[i] = (on && offset) ? offset[k] : 0;
            if (on) {
                if (!tb.v[i] || !tb.w[i] || tb.n[i] < 0 || tb.S[i] < 1 || (!bwd && !tb.m[i]) || (bwd && !tb.eps[i]))
                    return MXF_EINVAL;
                work = std::max(work, bwd ? tb.n[i] : (tb.n[i] * tb.S[i] + 3) / 4);
            }
        }
        dim3 grid((unsigned)std::min<int64_t>(cdiv(work, 256), 2 * kNumSMs), tb.count);
        if (!bwd) normal_reparam_multi_kernel<T><<<grid, 256, 0, st>>>(tb, seed, step_counter);
        else normal_reparam_multi_bwd_kernel<T><<<grid, 256, 0, st>>>(tb);
        int rc = after_launch();
        if (rc != MXF_OK) return rc;
    }
    return MXF_OK;
}

extern "C" int mxf_normal_reparam_multi(int dtype, int count, const void* const* m, const void* const* v,
                                        const int64_t* sM, const int64_t* sV, const int64_t* n, const int* S,
                                        uint64_t seed, const uint64_t* offset, const int* step_counter, void* const* w,
                                        void* const* eps_out, void* stream) {
    if (count < 0 || (count > 0 && (!m || !v || !sM || !sV || !n || !S || !offset || !w))) return MXF_EINVAL;
    if (count == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, return rp_multi<T>(0, count, m, v, sM, sV, n, S, seed, offset, step_counter, w, eps_out,
                                                 nullptr, nullptr, (cudaStream_t)stream));
}

extern "C" int mxf_normal_reparam_multi_bwd(int dtype, int count, const void* const* gw, const void* const* eps,
                                            const void* const* v, const int64_t* sM, const int64_t* sV, const int64_t* n,
                                            const int* S, void* const* gm, void* const* gv, void* stream) {
    if (count < 0 || (count > 0 && (!gw || !eps || !v || !sM || !sV || !n || !S || !gm || !gv))) return MXF_EINVAL;
    if (count == 0) return MXF_OK;
    MXF_DISPATCH_DTYPE(dtype, return rp_multi<T>(1, count, nullptr, v, sM, sV, n, S, 0, nullptr, nullptr,
                                                 const_cast<void* const*>(gw), const_cast<void* const*>(eps), gm, gv,
                                                 (cudaStream_t)stream));
}
