// Covariance-matrix build K(X, X2) for D <= 16, f32, with the -2 a.b term on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, 3xTF32 split, accumulators in TMEM) and the output tile written by TMA stores.
// Included by kbuild.cu (after the kernel-value helpers it uses).
//
// Why: the FMA formulation (kbuild_fwd_stream_kernel) costs D FFMA + ~8 other instructions per output element; at D = 16
// (and for the Matern kernels at D = 8) it is issue-bound at 0.40-0.66 of the HBM write roofline.  The reference itself forms
// the cross term with a GEMM (stationary.py:102: `F.linalg.gemm2(X, X2, transpose_b=True) * -2`); here that GEMM runs on the
// tensor pipe and an output element costs its share of a tcgen05.ld, two FADDs, the exponential (+ the Matern polynomial) and
// a quarter of a 16-byte shared store: the kernel is back on the HBM roofline.
//
// Structure (persistent, one CTA per SM, 10 warps):
//   * a CTA owns one (sample, 512-column group) -- 256-column groups when 512-wide ones would leave SMs idle -- and walks row
//     tiles of 128 rows t = blockIdx.x, + gridDim.x, ...
//   * the scaled column vectors b' = c x2 / l of its 512 columns sit in shared memory for the CTA's lifetime as TWO K-major
//     128B-swizzled operand tiles of 256 rows; a row is [hi (8 KS floats) | lo (8 KS floats)]: the tf32 head and the
//     remainder of the 3xTF32 split share one 128-byte swizzle row (KS = 1 for D <= 8, 2 for D <= 16);
//   * warp 8 (producer) stages the row tile: loads 128 rows of X, scales by -2 c / l, splits, writes the swizzled A tile
//     (double-buffered) and the row norms;
//   * warp 9, one thread: per 256-column block 3 KS tcgen05.mma (128 x 256 x 8:  a_lo b_hi + a_hi b_lo + a_hi b_hi) into
//     one of two 256-column TMEM accumulators, tcgen05.commit to the epilogue's barrier;
//   * warps 0..7 (epilogue): warp w reads lane quarter w % 4, columns (w / 4) * 128 .. + 127 of the accumulator, 32 columns
//     at a time (tcgen05.ld 32x32b.x32), forms r2 = |a|^2 + |b|^2 - 2 a.b and the kernel value exactly as the streaming
//     kernel does, writes the 32 x 32 block into a 128B-swizzled staging buffer (conflict-free 16-byte shared stores) and
//     one lane issues the TMA store (cp.async.bulk.tensor, bulk-group completion, two staging buffers per warp).
//     Ragged edges (N % 128, N2 % 32) are clipped by the TMA unit (N2 % 4 == 0 required: it clips 16-byte chunks).
//   * SYM (X2 == X): the diagonal gets the exact kernel value at r2 = 0 plus `diag_add + diag_const` (the `+ eye * noise`,
//     `+ eye * jitter` of gp_regression.py:55-60 / svgp_regression.py:70-72), exactly as the streaming kernel does.
#pragma once
#include "tc_common.cuh"

namespace mxf {

constexpr int KT_THREADS = 320;
constexpr int KT_BM = 128;                 // rows per tile
constexpr int KT_BN = 256;                 // columns per MMA / TMEM accumulator
constexpr int KT_GROUP = 512;              // columns per CTA (two accumulators)
constexpr int KT_B_BYTES = KT_GROUP * 128;             // 64 KB
constexpr int KT_A_BYTES = KT_BM * 128;                // 16 KB per buffer
constexpr int KT_STAGE_BYTES = 32 * 128;               // one 32 x 32 fp32 block
constexpr int KT_OFF_A = KT_B_BYTES;
constexpr int KT_OFF_STAGE = KT_OFF_A + 2 * KT_A_BYTES;
constexpr int KT_OFF_NB = KT_OFF_STAGE + 16 * KT_STAGE_BYTES;
constexpr int KT_OFF_NA = KT_OFF_NB + KT_GROUP * 4;
constexpr int KT_OFF_SC = KT_OFF_NA + 4 * KT_BM * 4;
constexpr int KT_OFF_BAR = KT_OFF_SC + 16 * 4;
constexpr int KT_SMEM = KT_OFF_BAR + 128 + 1024 /*align slack*/;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* tmap, uint32_t src, int c0, int c1, int c2, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// tcgen05.ld without the wait, and a wait that carries the destination registers (so that no consumer of `v` can be
// scheduled above it): lets the load of the next 32 columns fly while the current ones go through the exponential
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// one row (<= 16 scaled values, already multiplied by `mul * sc[d]`) -> [hi | lo] chunks of a 128B-swizzled operand row
template <int KS>
__device__ __forceinline__ void kt_write_row(uint8_t* tile, int row, const float (&x)[16]) {
    uint8_t* rp = tile + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 2 * KS; ++c) {
        float4 h, l;
        h.x = to_tf32(x[4 * c]); h.y = to_tf32(x[4 * c + 1]); h.z = to_tf32(x[4 * c + 2]); h.w = to_tf32(x[4 * c + 3]);
        l.x = x[4 * c] - h.x; l.y = x[4 * c + 1] - h.y; l.z = x[4 * c + 2] - h.z; l.w = x[4 * c + 3] - h.w;
        *reinterpret_cast<float4*>(rp + ((c ^ sw) << 4)) = h;
        *reinterpret_cast<float4*>(rp + (((2 * KS + c) ^ sw) << 4)) = l;
    }
}

// D values of row `r` of a row-major (rows x D) matrix, zero-padded to 16; `vec`: D == 16 and 16-byte aligned rows
__device__ __forceinline__ void kt_load_row(const float* __restrict__ base, int64_t r, int D, bool valid, bool vec, float (&x)[16]) {
#pragma unroll
    for (int d = 0; d < 16; ++d) x[d] = 0.f;
    if (!valid) return;
    const float* p = base + r * D;
    if (vec) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + c);
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < 16; ++d)
            if (d < D) x[d] = __ldg(p + d);
    }
}

template <int KIND, int KS, bool SYM>
__global__ void __launch_bounds__(KT_THREADS, 1)
kbuild_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ X, const float* __restrict__ X2,
                     const float* __restrict__ ls, int ls_len, const float* __restrict__ var, const float* __restrict__ diag_add,
                     float diag_const, int N, int N2, int D, int row_tiles, int group_cols, int64_t sX, int64_t sX2, int64_t sLs,
                     int64_t sVar, int64_t sDiag, int vecX, int vecX2, int evict_first) {
    constexpr bool RBF_FOLD = (KIND == MXF_KERN_RBF);
    extern __shared__ __align__(16) uint8_t kt_smem_raw[];
    const uint32_t base = (smem_u32(kt_smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = kt_smem_raw + (base - smem_u32(kt_smem_raw));
    float* nb_s = reinterpret_cast<float*>(bp + KT_OFF_NB);
    float* na_s = reinterpret_cast<float*>(bp + KT_OFF_NA);
    float* sc_s = reinterpret_cast<float*>(bp + KT_OFF_SC);
    const uint32_t bars = base + KT_OFF_BAR;
    auto bar_a_full = [&](int b) { return bars + 8u * b; };
    auto bar_a_empty = [&](int b) { return bars + 16u + 8u * b; };
    auto bar_t_full = [&](int b) { return bars + 32u + 8u * b; };
    auto bar_t_empty = [&](int b) { return bars + 48u + 8u * b; };
    const uint32_t tmem_slot = bars + 64u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + KT_OFF_BAR + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.z;
    const int colbase = blockIdx.y * group_cols;          // group_cols: 512, or 256 when 512-wide groups would leave SMs idle
    const int cols_here = min(group_cols, N2 - colbase);
    const int nblk = (cols_here + KT_BN - 1) / KT_BN;
    const float* Xs = X + (int64_t)s * sX;
    const float* X2s = X2 + (int64_t)s * sX2;
    const float* lss = ls + (int64_t)s * sLs;
    const float v = var[(int64_t)s * sVar];
    const float l2v = log2f(v);
    const float csq = RBF_FOLD ? 0.84932180028801904272f : 1.0f;       // sqrt(log2(e)/2)

    if (threadIdx.x == 32) prefetch_tmap(&tmOut);
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_a_full(b), 32);
            mbar_init(bar_a_empty(b), 1);
            mbar_init(bar_t_full(b), 1);
            mbar_init(bar_t_empty(b), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x < 16) sc_s[threadIdx.x] = (int)threadIdx.x < D ? csq / lss[ls_len == 1 ? 0 : threadIdx.x] : 0.f;
    __syncthreads();
    // resident B operand: the CTA's 512 scaled column vectors + their norms (with the RBF constant folded in)
    for (int c = threadIdx.x; c < nblk * KT_BN; c += KT_THREADS) {
        const int j = colbase + c;
        float x[16];
        kt_load_row(X2s, j, D, j < N2, vecX2 != 0, x);
        float n2 = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) { x[d] *= sc_s[d]; n2 = fmaf(x[d], x[d], n2); }
        nb_s[c] = RBF_FOLD ? n2 - l2v : n2;
        kt_write_row<KS>(bp + (c >> 8) * (KT_BN * 128), c & 255, x);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 8) {
        // ---------------------------------------------------------------- producer of the row tiles
        int i = 0;
        for (int t = blockIdx.x; t < row_tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            mbar_wait(bar_a_empty(buf), (uint32_t)(((i >> 1) & 1) ^ 1));
            uint8_t* tile = bp + KT_OFF_A + buf * KT_A_BYTES;
            float* na = na_s + (i & 3) * KT_BM;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = lane + 32 * rr;
                const int64_t gi = (int64_t)t * KT_BM + r;
                float x[16];
                kt_load_row(Xs, gi, D, gi < N, vecX != 0, x);
                float n2 = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) { x[d] *= -2.f * sc_s[d]; n2 = fmaf(x[d], x[d], n2); }
                na[r] = 0.25f * n2;
                kt_write_row<KS>(tile, r, x);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_a_full(buf));
        }
    } else if (warp == 9) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT_BN >> 3) << 17) | ((uint32_t)(KT_BM >> 4) << 24);
            int i = 0;
            uint32_t u = 0;
            for (int t = blockIdx.x; t < row_tiles; t += gridDim.x, ++i) {
                const int buf = i & 1;
                mbar_wait(bar_a_full(buf), (uint32_t)((i >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = base + KT_OFF_A + buf * KT_A_BYTES;
                for (int nb = 0; nb < nblk; ++nb, ++u) {
                    const uint32_t tb = u & 1u;
                    mbar_wait(bar_t_empty(tb), ((u >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b = base + nb * (KT_BN * 128);
                    const uint32_t d = tmem + tb * KT_BN;
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) {
                        const uint64_t dah = smem_desc(a + kk * 32, 16, 1024), dal = smem_desc(a + (KS + kk) * 32, 16, 1024);
                        const uint64_t dbh = smem_desc(b + kk * 32, 16, 1024), dbl = smem_desc(b + (KS + kk) * 32, 16, 1024);
                        umma_tf32(d, dal, dbh, idesc, kk != 0);         // small terms first
                        umma_tf32(d, dah, dbl, idesc, 1);
                        umma_tf32(d, dah, dbh, idesc, 1);
                    }
                    umma_commit(bar_t_full(tb));
                }
                umma_commit(bar_a_empty(buf));
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue: TMEM -> kernel value -> smem -> TMA store
        const int q = warp & 3, h = warp >> 2;
        uint64_t pol = 0;
        if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        const uint32_t stage0 = base + KT_OFF_STAGE + warp * 2 * KT_STAGE_BYTES;
        uint8_t* stage0_ptr = bp + KT_OFF_STAGE + warp * 2 * KT_STAGE_BYTES;
        const int sw = lane & 7;
        float dadd = 0.f;
        if (SYM) dadd = diag_const + (diag_add ? diag_add[(int64_t)s * sDiag] : 0.f);
        int i = 0;
        uint32_t u = 0, sb = 0;
        for (int t = blockIdx.x; t < row_tiles; t += gridDim.x, ++i) {
            const int row0 = t * KT_BM + q * 32;
            float na = 0.f;
            for (int nb = 0; nb < nblk; ++nb, ++u) {
                const uint32_t tb = u & 1u;
                mbar_wait(bar_t_full(tb), (u >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (nb == 0) na = na_s[(i & 3) * KT_BM + q * 32 + lane];
                const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + tb * KT_BN + h * 128;
                const int colw = nb * KT_BN + h * 128;                      // first column of this warp, within the group
                const bool rows_live = row0 < N;
                uint32_t acc[2][32];
                if (rows_live && colbase + colw < N2) tmem_ld32_nowait(tcol, acc[0]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col0 = colw + j * 32;
                    const bool live = rows_live && (colbase + col0 < N2);   // warp-uniform
                    if (live) tmem_ld_wait(acc[j & 1]);
                    if (j < 3) {
                        if (rows_live && colbase + col0 + 32 < N2) tmem_ld32_nowait(tcol + (j + 1) * 32, acc[(j + 1) & 1]);
                    } else {
                        // this warp's last read of the accumulator has landed: hand it back to the MMA warp
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_t_empty(tb));
                    }
                    if (!live) continue;
                    if (lane == 0) bulk_wait_read<1>();       // the store issued two blocks ago has left its staging buffer
                    __syncwarp();
                    uint8_t* st = stage0_ptr + sb * KT_STAGE_BYTES + lane * 128;
                    // all 32 column norms first, then 32 independent value chains, then the 8 stores: shared-memory loads and
                    // stores interleaved per 4 elements would serialise the chains (the compiler must assume they alias)
                    float nbq[32], o[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 nbv = *reinterpret_cast<const float4*>(nb_s + col0 + 4 * c);
                        nbq[4 * c] = nbv.x; nbq[4 * c + 1] = nbv.y; nbq[4 * c + 2] = nbv.z; nbq[4 * c + 3] = nbv.w;
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float r2 = (__uint_as_float(acc[j & 1][e]) + na) + nbq[e];
                        if (RBF_FOLD) o[e] = ex2_approx(-r2);
                        else o[e] = matern_value_l2<KIND>(r2, l2v);
                    }
                    if (SYM) {
                        // K(X, X): on the diagonal r2 is exactly 0 (the expanded form only leaves cancellation noise there,
                        // which the Matern square root would amplify), and the `+ eye * (noise + jitter)` is folded in
                        const int dcol = row0 + lane - (colbase + col0);        // this lane's diagonal column within the block
                        if (dcol >= 0 && dcol < 32) {
                            const float kd = RBF_FOLD ? ex2_approx(l2v) : matern_value_l2<KIND>(0.f, l2v);
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (e == dcol) o[e] = kd + dadd;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(st + ((c ^ sw) << 4)) = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t src = stage0 + sb * KT_STAGE_BYTES;
                        if (evict_first) tma_store_3d_hint(&tmOut, src, colbase + col0, row0, s, pol);
                        else tma_store_3d(&tmOut, src, colbase + col0, row0, s);
                        bulk_commit();
                    }
                    sb ^= 1u;
                }
            }
        }
        if (lane == 0) bulk_wait<0>();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// Output elements from which mxf_kbuild_fwd takes this path (smaller outputs: the per-CTA prologue -- 512 column vectors
// staged, TMEM allocated -- is not amortised and the streaming FMA kernel wins).  mxf_kbuild_tc_threshold() changes it.
static std::atomic<long long> g_kbuild_tc_min_elems{1ll << 20};

template <int KIND>
static int launch_fwd_tc(const float* X, const float* X2, const float* ls, int ls_len, const float* var, const float* diag_add,
                         double diag_const, float* out, int64_t ldo, int S, int N, int N2, int D, int64_t sX, int64_t sX2,
                         int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut, cudaStream_t st) {
    if (D > 16 || N2 < 32 || N < 1 || S > 65535) return MXF_ENOTIMPL;
    // TMA: 16-byte aligned base and strides; and the unit clips a box at 16-byte granularity along the row (measured: with
    // N2 % 4 != 0 the elements up to the next multiple of 4 are overwritten), so N2 must be a multiple of 4 as well
    if ((ldo & 3) || (N2 & 3) || (S > 1 && (sOut & 3)) || (reinterpret_cast<uintptr_t>(out) & 15)) return MXF_ENOTIMPL;
    CUtensorMap tm;
    if (!make_map(&tm, out, N, N2, ldo, sOut, S, 32, 32)) return MXF_ENOTIMPL;
    const bool sym = (X2 == nullptr);
    const float* X2e = sym ? X : X2;
    const int64_t sX2e = sym ? sX : sX2;
    const int row_tiles = cdiv(N, KT_BM);
    // 512-column groups amortise the row-tile staging over two accumulators; 256-column groups when that would leave SMs idle
    const int group_cols = ((int64_t)cdiv(N2, KT_GROUP) * row_tiles * S >= kNumSMs) ? KT_GROUP : KT_BN;
    const int groups = cdiv(N2, group_cols);
    if (groups > 65535) return MXF_ENOTIMPL;
    int per_group = std::max(1, kNumSMs / (groups * S));
    per_group = std::min(per_group, row_tiles);
    const int vecX = (D == 16) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((sX & 3) == 0);
    const int vecX2 = (D == 16) && ((reinterpret_cast<uintptr_t>(X2e) & 15) == 0) && ((sX2e & 3) == 0);
    const int evict_first = ((int64_t)S * N * N2 * 4 > (64ll << 20)) ? 1 : 0;     // larger than what L2 can keep for a consumer
    dim3 grid(per_group, groups, S);
    auto launch = [&](auto kern, bool& attr_set) -> int {
        if (!attr_set) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM) != cudaSuccess)
                return (int)cudaGetLastError();
            attr_set = true;
        }
        kern<<<grid, KT_THREADS, KT_SMEM, st>>>(tm, X, X2e, ls, ls_len, var, diag_add, (float)diag_const, N, N2, D, row_tiles,
                                                group_cols, sX, sX2e, sLs, sVar, sDiag, vecX, vecX2, evict_first);
        return after_launch();
    };
    static bool set[4] = {false, false, false, false};       // per KIND instantiation, one flag per kernel
    if (D <= 8) return sym ? launch(kbuild_fwd_tc_kernel<KIND, 1, true>, set[0]) : launch(kbuild_fwd_tc_kernel<KIND, 1, false>, set[1]);
    return sym ? launch(kbuild_fwd_tc_kernel<KIND, 2, true>, set[2]) : launch(kbuild_fwd_tc_kernel<KIND, 2, false>, set[3]);
}

}  // namespace mxf
