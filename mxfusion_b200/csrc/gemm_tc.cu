// FP32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), FP32-accurate through the 3xTF32
// split:   C = alpha * A * op(B) + beta * C,   A (m x k) row-major,
//          op(B) = B^T with B (n x k) row-major ("NT", both operands K-major), or
//          op(B) = B   with B (k x n) row-major ("NN", B operand MN-major).
//
// This is the update engine of the blocked potrf / trsm (trailing update A22 -= L21 L21^T is the NT
// form with A == B) and of every dense product of the SVGP bound and its gradient
// (svgp_regression.py:76,82,89,90 and their adjoints).
//
// Structure (one 128 x BN output tile per CTA, 6 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the raw FP32 A / B tiles (128B-swizzled)
//               into a STAGES-deep shared-memory ring, completion on mbarriers;
//   warps 2..5  converters: split every raw element x into hi = tf32(x) (written in place) and
//               lo = tf32(x - hi) (second buffer, same swizzled position), then, after the main loop,
//               the epilogue: tcgen05.ld of the accumulator, alpha/beta, store to global;
//   warp 1      one elected thread issues, per 8-wide K slice,  D += Ahi Bhi + Ahi Blo + Alo Bhi  with
//               tcgen05.mma.kind::tf32 (accumulator in TMEM), and frees ring slots with tcgen05.commit.
// tcgen05 has no FP32-input kind; a single TF32 pass (10-bit mantissa) would break the stated FP32
// tolerance on the ill-conditioned Kuu of a GP, hence the split (error ~ 3 * 2^-22 per product).
#include "tc_common.cuh"

namespace mxf {


constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                  // 32 floats = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;
constexpr int TC_CONV_THREADS = 128;

template <int BN>
struct TcCfg {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;               // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);     // [A hi | B hi | A lo | B lo]
    static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128) ? 3 : 4;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// B_MN = false: B (n x k) row-major, K-major operand.   B_MN = true: B (k x n) row-major, MN-major operand.
template <int BN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               float* __restrict__ C, int64_t ldc, int64_t sC, int m, int n, int k, float alpha, float beta,
               int tri, int batchA, int batchB) {
    using Cfg = TcCfg<BN>;
    pdl_launch_dependents();
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    if ((tri & 1) && n0 > m0 + TC_BM - 1) return;    // tile strictly above the diagonal: nothing to do

    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + Cfg::STAGES * Cfg::STAGE_BYTES;     // full[S], ready[S], empty[S], tmem_full, tmem slot
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_ready = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * Cfg::STAGES + s); };
    const uint32_t bar_tmem = bars + 8u * (3 * Cfg::STAGES);
    const uint32_t tmem_slot = bar_tmem + 8u;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (3 * Cfg::STAGES) + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // K range of this tile: all of K, or -- when A is known to be lower (tri & 2) / upper (tri & 4) triangular, as the
    // inverted diagonal blocks of a Cholesky factor are -- only the part where this row block of A is non-zero
    const int kb_begin = (tri & 4) ? m0 / TC_BK : 0;
    const int k_end = (tri & 2) ? min(k, m0 + TC_BM) : k;
    const int kb_end = (k_end + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 32) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_ready(s), TC_CONV_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot_ptr;
    pdl_wait();            // everything above touched no global memory; from here on the predecessor's output is read

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const int zA = batchA ? (int)blockIdx.z : 0, zB = batchB ? (int)blockIdx.z : 0;
            for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_empty(s), ph ^ 1u);
                const uint32_t st = base + s * Cfg::STAGE_BYTES;
                mbar_expect_tx(bar_full(s), Cfg::A_BYTES + Cfg::B_BYTES);
                tma_load_3d(st, &tmA, bar_full(s), kb * TC_BK, m0, zA);
                if (!B_MN) {
                    tma_load_3d(st + Cfg::A_BYTES, &tmB, bar_full(s), kb * TC_BK, n0, zB);
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c)    // one 32-column (128-byte) slab of B per box
                        tma_load_3d(st + Cfg::A_BYTES + c * (TC_BK * 128), &tmB, bar_full(s), n0 + c * 32, kb * TC_BK, zB);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
            // a_major [15]=0 (K), b_major [16], N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
                mbar_wait(bar_ready(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = base + s * Cfg::STAGE_BYTES, b_hi = a_hi + Cfg::A_BYTES;
                const uint32_t a_lo = a_hi + Cfg::A_BYTES + Cfg::B_BYTES, b_lo = a_lo + Cfg::A_BYTES;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {
                    // K-major, 128B swizzle: 8-row groups 1024 B apart; a K slice of 8 floats is 32 B along the row.
                    const uint64_t dah = smem_desc(a_hi + kk * 32, 16, 1024), dal = smem_desc(a_lo + kk * 32, 16, 1024);
                    uint64_t dbh, dbl;
                    if (!B_MN) {
                        dbh = smem_desc(b_hi + kk * 32, 16, 1024);
                        dbl = smem_desc(b_lo + kk * 32, 16, 1024);
                    } else {
                        // MN-major, 128B swizzle with 32B atoms: K rows are 128 B (32 columns of N) apart, the swizzle
                        // pattern repeats every 4 K-rows (SBO = 512 B); a K slice of 8 is 1024 B; the next 32 columns
                        // of N are one slab (LBO = TC_BK * 128 B) further on.
                        dbh = smem_desc(b_hi + kk * 1024, TC_BK * 128, 512, 1);
                        dbl = smem_desc(b_lo + kk * 1024, TC_BK * 128, 512, 1);
                    }
                    umma_tf32(tmem_d, dal, dbh, idesc, (it | kk) != 0);     // small terms first
                    umma_tf32(tmem_d, dah, dbl, idesc, 1);
                    umma_tf32(tmem_d, dah, dbh, idesc, 1);
                }
                umma_commit(bar_empty(s));       // slot reusable once these MMAs have read it
            }
            umma_commit(bar_tmem);               // accumulator complete
        }
    } else {
        // ------------------------------------------------------------------ converters, then epilogue
        const int t = threadIdx.x - 64;          // 0..127
        constexpr int VEC = (Cfg::A_BYTES + Cfg::B_BYTES) / 16;     // 16-byte vectors per stage (hi region)
        for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
            const int s = it % Cfg::STAGES;
            const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
            mbar_wait(bar_full(s), ph);
            float4* hi = reinterpret_cast<float4*>(base_ptr + s * Cfg::STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(base_ptr + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll 4
            for (int i = t; i < VEC; i += TC_CONV_THREADS) {
                const float4 x = hi[i];
                float4 h, l;
                h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
                l.x = lo_tf32(x.x - h.x); l.y = lo_tf32(x.y - h.y); l.z = lo_tf32(x.z - h.z); l.w = lo_tf32(x.w - h.w);
                hi[i] = h;
                lo[i] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
            mbar_arrive(bar_ready(s));
        }
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        float* Cb = C + (int64_t)blockIdx.z * sC;
        const bool vec_ok = ((ldc & 3) == 0) && ((sC & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
            if (row < m) {
                float* crow = Cb + (int64_t)row * ldc;
                const int c0 = n0 + j * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const int c = c0 + e;
                    if (c >= n) break;
                    float o0 = alpha * __uint_as_float(v[e]), o1 = alpha * __uint_as_float(v[e + 1]);
                    float o2 = alpha * __uint_as_float(v[e + 2]), o3 = alpha * __uint_as_float(v[e + 3]);
                    if (vec_ok && c + 3 < n) {
                        float4* dst = reinterpret_cast<float4*>(crow + c);
                        if (beta != 0.f) {
                            const float4 old = *dst;
                            o0 = fmaf(beta, old.x, o0); o1 = fmaf(beta, old.y, o1);
                            o2 = fmaf(beta, old.z, o2); o3 = fmaf(beta, old.w, o3);
                        }
                        *dst = make_float4(o0, o1, o2, o3);
                    } else {
                        const float o[4] = {o0, o1, o2, o3};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + u < n) crow[c + u] = (beta != 0.f) ? fmaf(beta, crow[c + u], o[u]) : o[u];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Variant with the A operand in TENSOR MEMORY ("TS" form of tcgen05.mma).  The plain kernel above is
// shared-memory-bandwidth bound: per K step the split writes A_hi/A_lo/B_hi/B_lo and the three MMAs read each of them
// again.  Here the converter warps split their row of A in registers and write A_hi / A_lo straight to TMEM
// (tcgen05.st); shared memory then only carries the raw landing tile of A and B_hi / B_lo, and the MMAs read only B
// from it (per K step, BN = 128: 144 KB instead of 224 KB).  Used when the K loop is long enough to pay for the
// larger prologue (k >= 256).
// ------------------------------------------------------------------------------------------------

template <int BN>
struct TaCfg {
    // Two decoupled shared-memory rings: RAW (TMA landing tiles [A raw | B raw], R deep) and CONV ([B hi | B lo], C
    // deep; the matching A hi / A lo live in tensor memory, one 64-column slot per CONV stage).  A RAW slot is handed
    // back to the TMA producer as soon as the converter warps have read it -- not when the MMAs that consume the converted
    // copy retire -- so the next loads are in flight a whole K step earlier (the in-place layout left the converters
    // waiting on TMA for a third of their samples, profiles/r1b).
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;               // raw landing tile of A (16 KB)
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int RAW_BYTES = A_BYTES + B_BYTES;             // [A raw | B raw]
    static constexpr int CONV_BYTES = 2 * B_BYTES;                  // [B hi | B lo]
    static constexpr int R = (BN == 256) ? 2 : 4;
    static constexpr int C = (BN == 256) ? 2 : (BN == 128) ? 3 : 4;
    static constexpr int CONV0 = R * RAW_BYTES;
    static constexpr int BARS0 = CONV0 + C * CONV_BYTES;
    static constexpr int TMEM_A0 = BN;                              // accumulator in columns [0, BN); A slots after it
    static constexpr int SMEM = BARS0 + 1024 + 256;
    static_assert(BN + C * 64 <= 512, "tensor memory budget");
    static_assert(SMEM <= 232448, "shared memory budget");
};

constexpr int TA_THREADS = 320;            // producer warp, MMA warp, 8 converter / epilogue warps
constexpr int TA_CONV_THREADS = 256;

template <int BN, bool B_MN>
__global__ void __launch_bounds__(TA_THREADS, 1)
gemm_tc_ta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ C, int64_t ldc, int64_t sC, int m, int n, int k, float alpha, float beta,
                  int tri, int batchA, int batchB) {
    using Cfg = TaCfg<BN>;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    if ((tri & 1) && n0 > m0 + TC_BM - 1) return;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + Cfg::BARS0;
    auto bar_full = [&](int s) { return bars + 8u * s; };                            // RAW slot landed (TMA tx)
    auto bar_rawfree = [&](int s) { return bars + 8u * (Cfg::R + s); };              // RAW slot read by all converters
    auto bar_ready = [&](int s) { return bars + 8u * (2 * Cfg::R + s); };            // CONV slot (+ TMEM A slot) written
    auto bar_empty = [&](int s) { return bars + 8u * (2 * Cfg::R + Cfg::C + s); };   // CONV slot consumed by the MMAs
    const uint32_t bar_tmem = bars + 8u * (2 * Cfg::R + 2 * Cfg::C);
    const uint32_t tmem_slot = bar_tmem + 8u;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::BARS0 + 8 * (2 * Cfg::R + 2 * Cfg::C) + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_begin = (tri & 4) ? m0 / TC_BK : 0;
    const int k_end = (tri & 2) ? min(k, m0 + TC_BM) : k;
    const int kb_end = (k_end + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 32) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::R; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_rawfree(s), TA_CONV_THREADS);
        }
        for (int s = 0; s < Cfg::C; ++s) {
            mbar_init(bar_ready(s), TA_CONV_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            const int zA = batchA ? (int)blockIdx.z : 0, zB = batchB ? (int)blockIdx.z : 0;
            for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
                const int s = it % Cfg::R;
                const uint32_t ph = (uint32_t)(it / Cfg::R) & 1u;
                mbar_wait(bar_rawfree(s), ph ^ 1u);
                const uint32_t st = base + s * Cfg::RAW_BYTES;
                mbar_expect_tx(bar_full(s), Cfg::A_BYTES + Cfg::B_BYTES);
                tma_load_3d(st, &tmA, bar_full(s), kb * TC_BK, m0, zA);
                if (!B_MN) {
                    tma_load_3d(st + Cfg::A_BYTES, &tmB, bar_full(s), kb * TC_BK, n0, zB);
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c)
                        tma_load_3d(st + Cfg::A_BYTES + c * (TC_BK * 128), &tmB, bar_full(s), n0 + c * 32, kb * TC_BK, zB);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
                const int s = it % Cfg::C;
                const uint32_t ph = (uint32_t)(it / Cfg::C) & 1u;
                mbar_wait(bar_ready(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = base + Cfg::CONV0 + s * Cfg::CONV_BYTES, b_lo = b_hi + Cfg::B_BYTES;
                const uint32_t a_hi = tmem_base + (uint32_t)(Cfg::TMEM_A0 + s * 64), a_lo = a_hi + 32u;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {
                    uint64_t dbh, dbl;
                    if (!B_MN) {
                        dbh = smem_desc(b_hi + kk * 32, 16, 1024);
                        dbl = smem_desc(b_lo + kk * 32, 16, 1024);
                    } else {
                        dbh = smem_desc(b_hi + kk * 1024, TC_BK * 128, 512, 1);
                        dbl = smem_desc(b_lo + kk * 1024, TC_BK * 128, 512, 1);
                    }
                    umma_tf32_ts(tmem_base, a_lo + kk * 8, dbh, idesc, (it | kk) != 0);
                    umma_tf32_ts(tmem_base, a_hi + kk * 8, dbl, idesc, 1);
                    umma_tf32_ts(tmem_base, a_hi + kk * 8, dbh, idesc, 1);
                }
                umma_commit(bar_empty(s));
            }
            umma_commit(bar_tmem);
        }
    } else {
        // 8 converter warps: the first four also split the A tile (one row per thread, TMEM quarter = warp % 4),
        // all eight share the B tile
        const int t = threadIdx.x - 64;          // 0..255
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const bool a_warp = warp < 6;
        const int row = q * 32 + lane;           // row of the A tile this thread splits
        constexpr int VECB = Cfg::B_BYTES / 16;
        for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
            const int rs = it % Cfg::R, s = it % Cfg::C;
            mbar_wait(bar_full(rs), (uint32_t)(it / Cfg::R) & 1u);
            mbar_wait(bar_empty(s), ((uint32_t)(it / Cfg::C) & 1u) ^ 1u);      // CONV slot + TMEM A slot free again
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint8_t* stage = base_ptr + rs * Cfg::RAW_BYTES;
            uint8_t* conv = base_ptr + Cfg::CONV0 + s * Cfg::CONV_BYTES;
            // A: this thread's 128-byte row (16-byte chunk c of row r sits at chunk c ^ (r & 7): 128B swizzle)
            if (a_warp) {
                uint32_t hi[32], lo[32];
                const uint8_t* arow = stage + row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
                    const float h0 = to_tf32(x.x), h1 = to_tf32(x.y), h2 = to_tf32(x.z), h3 = to_tf32(x.w);
                    hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1);
                    hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
                    lo[4 * c] = __float_as_uint(lo_tf32(x.x - h0)); lo[4 * c + 1] = __float_as_uint(lo_tf32(x.y - h1));
                    lo[4 * c + 2] = __float_as_uint(lo_tf32(x.z - h2)); lo[4 * c + 3] = __float_as_uint(lo_tf32(x.w - h3));
                }
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::TMEM_A0 + s * 64);
                tmem_st32(ta, hi);
                tmem_st32(ta + 32u, lo);
            }
            // B: position-preserving split of the raw tile into the CONV slot (hi | lo)
            const float4* braw = reinterpret_cast<const float4*>(stage + Cfg::A_BYTES);
            float4* bh = reinterpret_cast<float4*>(conv);
            float4* bl = reinterpret_cast<float4*>(conv + Cfg::B_BYTES);
#pragma unroll 4
            for (int i = t; i < VECB; i += TA_CONV_THREADS) {
                const float4 x = braw[i];
                float4 h, l;
                h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
                l.x = lo_tf32(x.x - h.x); l.y = lo_tf32(x.y - h.y); l.z = lo_tf32(x.z - h.z); l.w = lo_tf32(x.w - h.w);
                bh[i] = h;
                bl[i] = l;
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_ready(s));
            mbar_arrive(bar_rawfree(rs));
        }
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int orow = m0 + q * 32 + lane;
        float* Cb = C + (int64_t)blockIdx.z * sC;
        const bool vec_ok = ((ldc & 3) == 0) && ((sC & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        // two warps per TMEM quarter: they take alternate 32-column slabs of the accumulator
#pragma unroll 1
        for (int j = a_warp ? 0 : 1; j < BN / 32; j += 2) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
            if (orow < m) {
                float* crow = Cb + (int64_t)orow * ldc;
                const int c0 = n0 + j * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const int c = c0 + e;
                    if (c >= n) break;
                    float o0 = alpha * __uint_as_float(v[e]), o1 = alpha * __uint_as_float(v[e + 1]);
                    float o2 = alpha * __uint_as_float(v[e + 2]), o3 = alpha * __uint_as_float(v[e + 3]);
                    if (vec_ok && c + 3 < n) {
                        float4* dst = reinterpret_cast<float4*>(crow + c);
                        if (beta != 0.f) {
                            const float4 old = *dst;
                            o0 = fmaf(beta, old.x, o0); o1 = fmaf(beta, old.y, o1);
                            o2 = fmaf(beta, old.z, o2); o3 = fmaf(beta, old.w, o3);
                        }
                        *dst = make_float4(o0, o1, o2, o3);
                    } else {
                        const float o[4] = {o0, o1, o2, o3};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c + u < n) crow[c + u] = (beta != 0.f) ? fmaf(beta, crow[c + u], o[u]) : o[u];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant of the TMEM-A kernel: one CTA per SM walks a static list of 128 x 128 output tiles; the FP32
// accumulator is double-buffered in tensor memory (2 x 128 columns) and a dedicated group of four epilogue warps drains
// tile j (tcgen05.ld -> alpha / beta -> global) while the TMA producer, the converter warps and the MMA thread are
// already running the K loop of tile j + 1.  The one-tile-per-CTA kernels above pay barrier / TMEM set-up, the pipeline
// fill and the whole epilogue once per tile with nothing overlapped -- with K = 512 (potrf / trsm updates) that is as long
// as the K loop itself (profiles/r1b: 35 % tensor-pipe activity at K = 512 against 69 % at K = 4096).
// Converter warps work in two groups of four that take alternate K steps, so each group has two MMA periods
// (2 x 768 cycles at BN = 128) for the load -> split -> tcgen05.st / st.shared -> fence chain of its step.
// ------------------------------------------------------------------------------------------------
struct PeCfg {
    static constexpr int BN = 128;
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;               // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 4;                  // 16 KB
    static constexpr int RAW_BYTES = A_BYTES + B_BYTES;
    static constexpr int CONV_BYTES = 2 * B_BYTES;
    static constexpr int R = 4;
    static constexpr int C = 3;
    static constexpr int CONV0 = R * RAW_BYTES;
    static constexpr int BARS0 = CONV0 + C * CONV_BYTES;
    static constexpr int TMEM_A0 = 2 * BN;                          // accumulators in columns [0, 256); A slots after
    static constexpr int SMEM = BARS0 + 1024 + 256;
    static constexpr int THREADS = 448;                             // producer, MMA, 8 converter, 4 epilogue warps
    static constexpr int GROUP_THREADS = 128;                       // one converter group
    static_assert(2 * BN + C * 64 <= 512, "tensor memory budget");
    static_assert(SMEM <= 232448, "shared memory budget");
};

template <bool B_MN>
__global__ void __launch_bounds__(PeCfg::THREADS, 1)
gemm_tc_pe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ C, int64_t ldc, int64_t sC, int m, int n, int k, float alpha, float beta,
                  int tri, int batchA, int batchB, int tiles_m, int tiles_n, int S, int snake) {
    using Cfg = PeCfg;
    constexpr int BN = Cfg::BN;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + Cfg::BARS0;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_rawfree = [&](int s) { return bars + 8u * (Cfg::R + s); };
    auto bar_ready = [&](int s) { return bars + 8u * (2 * Cfg::R + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * Cfg::R + Cfg::C + s); };
    auto bar_accfull = [&](int s) { return bars + 8u * (2 * Cfg::R + 2 * Cfg::C + s); };
    auto bar_accfree = [&](int s) { return bars + 8u * (2 * Cfg::R + 2 * Cfg::C + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * Cfg::R + 2 * Cfg::C + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::BARS0 + 8 * (2 * Cfg::R + 2 * Cfg::C + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 32) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::R; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_rawfree(s), Cfg::GROUP_THREADS);
        }
        for (int s = 0; s < Cfg::C; ++s) {
            mbar_init(bar_ready(s), Cfg::GROUP_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_accfull(s), 1);
            mbar_init(bar_accfree(s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int tiles_per_batch = tiles_m * tiles_n;
    const int total_tiles = tiles_per_batch * S;
    // every role walks the same tile list; `tile_of` decodes tile t and returns false for tiles that are skipped
    // Tile schedule: position p of the list (rows of tiles ordered by DEcreasing K-loop length when A is triangular: with
    // tri & 2 the K loop of row block tm has tm + 1 blocks) is dealt to the CTAs in snake order (round i runs forwards for
    // even i, backwards for odd i), which balances the per-CTA sums of K steps to within one tile (plain round-robin on
    // 8 x 41 tiles of a 1024 solve leaves the slowest CTA with 14 units against a mean of 10).
    const int rounds = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_at = [&](int i) -> int {
        const int b = (snake && (i & 1)) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
        const int p = i * (int)gridDim.x + b;
        return p < total_tiles ? p : -1;
    };
    auto tile_of = [&](int t, int& z, int& m0, int& n0, int& kb_begin, int& kb_end) -> bool {
        if (t < 0) return false;
        z = t / tiles_per_batch;
        const int r = t - z * tiles_per_batch;
        int tm = r / tiles_n;
        const int tn = r - tm * tiles_n;
        if (snake && (tri & 2) && !(tri & 4)) tm = tiles_m - 1 - tm;          // heaviest rows first
        m0 = tm * TC_BM;
        n0 = tn * BN;
        if ((tri & 1) && n0 > m0 + TC_BM - 1) return false;
        kb_begin = (tri & 4) ? m0 / TC_BK : 0;
        const int k_end = (tri & 2) ? min(k, m0 + TC_BM) : k;
        kb_end = (k_end + TC_BK - 1) / TC_BK;
        return kb_end > kb_begin;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;
            for (int ri = 0; ri < rounds; ++ri) {
                const int t = tile_at(ri);
                int z, m0, n0, kb0, kb1;
                if (!tile_of(t, z, m0, n0, kb0, kb1)) continue;
                const int zA = batchA ? z : 0, zB = batchB ? z : 0;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % Cfg::R;
                    mbar_wait(bar_rawfree(s), ((uint32_t)(it / Cfg::R) & 1u) ^ 1u);
                    const uint32_t st = base + s * Cfg::RAW_BYTES;
                    mbar_expect_tx(bar_full(s), Cfg::A_BYTES + Cfg::B_BYTES);
                    tma_load_3d(st, &tmA, bar_full(s), kb * TC_BK, m0, zA);
                    if (!B_MN) {
                        tma_load_3d(st + Cfg::A_BYTES, &tmB, bar_full(s), kb * TC_BK, n0, zB);
                    } else {
#pragma unroll
                        for (int c = 0; c < BN / 32; ++c)
                            tma_load_3d(st + Cfg::A_BYTES + c * (TC_BK * 128), &tmB, bar_full(s), n0 + c * 32, kb * TC_BK, zB);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int it = 0, j = 0;
            for (int ri = 0; ri < rounds; ++ri) {
                const int t = tile_at(ri);
                int z, m0, n0, kb0, kb1;
                if (!tile_of(t, z, m0, n0, kb0, kb1)) continue;
                const int as = j & 1;
                mbar_wait(bar_accfree(as), ((uint32_t)(j >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % Cfg::C;
                    mbar_wait(bar_ready(s), (uint32_t)(it / Cfg::C) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_hi = base + Cfg::CONV0 + s * Cfg::CONV_BYTES, b_lo = b_hi + Cfg::B_BYTES;
                    const uint32_t a_hi = tmem_base + (uint32_t)(Cfg::TMEM_A0 + s * 64), a_lo = a_hi + 32u;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        uint64_t dbh, dbl;
                        if (!B_MN) {
                            dbh = smem_desc(b_hi + kk * 32, 16, 1024);
                            dbl = smem_desc(b_lo + kk * 32, 16, 1024);
                        } else {
                            dbh = smem_desc(b_hi + kk * 1024, TC_BK * 128, 512, 1);
                            dbl = smem_desc(b_lo + kk * 1024, TC_BK * 128, 512, 1);
                        }
                        umma_tf32_ts(tmem_d, a_lo + kk * 8, dbh, idesc, (kb != kb0 || kk != 0) ? 1u : 0u);
                        umma_tf32_ts(tmem_d, a_hi + kk * 8, dbl, idesc, 1);
                        umma_tf32_ts(tmem_d, a_hi + kk * 8, dbh, idesc, 1);
                    }
                    umma_commit(bar_empty(s));
                }
                umma_commit(bar_accfull(as));
                ++j;
            }
        }
    } else if (warp < 10) {
        // ------------------------------------------------------------------ converters: two groups, alternate K steps
        const int grp = (warp - 2) >> 2;                       // 0: warps 2..5, 1: warps 6..9
        const int tg = threadIdx.x - 64 - grp * Cfg::GROUP_THREADS;   // 0..127 within the group
        const int q = warp & 3;                                // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                         // row of the A tile this thread splits
        constexpr int VECB = Cfg::B_BYTES / 16;
        int it = 0;
        for (int ri = 0; ri < rounds; ++ri) {
            const int t = tile_at(ri);
            int z, m0, n0, kb0, kb1;
            if (!tile_of(t, z, m0, n0, kb0, kb1)) continue;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                if ((it & 1) != grp) continue;
                const int rs = it % Cfg::R, s = it % Cfg::C;
                mbar_wait(bar_full(rs), (uint32_t)(it / Cfg::R) & 1u);
                mbar_wait(bar_empty(s), ((uint32_t)(it / Cfg::C) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint8_t* stage = base_ptr + rs * Cfg::RAW_BYTES;
                uint8_t* conv = base_ptr + Cfg::CONV0 + s * Cfg::CONV_BYTES;
                {
                    uint32_t hi[32], lo[32];
                    const uint8_t* arow = stage + row * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
                        const float h0 = to_tf32(x.x), h1 = to_tf32(x.y), h2 = to_tf32(x.z), h3 = to_tf32(x.w);
                        hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1);
                        hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
                        lo[4 * c] = __float_as_uint(lo_tf32(x.x - h0)); lo[4 * c + 1] = __float_as_uint(lo_tf32(x.y - h1));
                        lo[4 * c + 2] = __float_as_uint(lo_tf32(x.z - h2)); lo[4 * c + 3] = __float_as_uint(lo_tf32(x.w - h3));
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::TMEM_A0 + s * 64);
                    tmem_st32(ta, hi);
                    tmem_st32(ta + 32u, lo);
                }
                const float4* braw = reinterpret_cast<const float4*>(stage + Cfg::A_BYTES);
                float4* bh = reinterpret_cast<float4*>(conv);
                float4* bl = reinterpret_cast<float4*>(conv + Cfg::B_BYTES);
#pragma unroll 4
                for (int i = tg; i < VECB; i += Cfg::GROUP_THREADS) {
                    const float4 x = braw[i];
                    float4 h, l;
                    h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
                    l.x = lo_tf32(x.x - h.x); l.y = lo_tf32(x.y - h.y); l.z = lo_tf32(x.z - h.z); l.w = lo_tf32(x.w - h.w);
                    bh[i] = h;
                    bl[i] = l;
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(bar_ready(s));
                mbar_arrive(bar_rawfree(rs));
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps (10..13)
        const int q = warp & 3;
        const bool vec_ok = ((ldc & 3) == 0) && ((sC & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        int j = 0;
        for (int ri = 0; ri < rounds; ++ri) {
            const int t = tile_at(ri);
            int z, m0, n0, kb0, kb1;
            if (!tile_of(t, z, m0, n0, kb0, kb1)) continue;
            const int as = j & 1;
            mbar_wait(bar_accfull(as), (uint32_t)(j >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int orow = m0 + q * 32 + lane;
            float* Cb = C + (int64_t)z * sC;
#pragma unroll 1
            for (int jj = 0; jj < BN / 32; ++jj) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + jj * 32), v);
                if (orow < m) {
                    float* crow = Cb + (int64_t)orow * ldc;
                    const int c0 = n0 + jj * 32;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        const int c = c0 + e;
                        if (c >= n) break;
                        float o0 = alpha * __uint_as_float(v[e]), o1 = alpha * __uint_as_float(v[e + 1]);
                        float o2 = alpha * __uint_as_float(v[e + 2]), o3 = alpha * __uint_as_float(v[e + 3]);
                        if (vec_ok && c + 3 < n) {
                            float4* dst = reinterpret_cast<float4*>(crow + c);
                            if (beta != 0.f) {
                                const float4 old = *dst;
                                o0 = fmaf(beta, old.x, o0); o1 = fmaf(beta, old.y, o1);
                                o2 = fmaf(beta, old.z, o2); o3 = fmaf(beta, old.w, o3);
                            }
                            *dst = make_float4(o0, o1, o2, o3);
                        } else {
                            const float o[4] = {o0, o1, o2, o3};
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (c + u < n) crow[c + u] = (beta != 0.f) ? fmaf(beta, crow[c + u], o[u]) : o[u];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_accfree(as));
            ++j;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side (tensor-map encoder: tc_common.cuh)
// ------------------------------------------------------------------------------------------------

template <int BN, bool B_MN>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, float* C, int64_t ldc, int64_t sC, int m, int n,
                     int k, float alpha, float beta, int S, int tri, int batchA, int batchB, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    auto kern = gemm_tc_kernel<BN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
            return (int)cudaGetLastError();
        attr_set = true;
    }
    dim3 grid(cdiv(n, BN), cdiv(m, TC_BM), S);
    cudaError_t e = launch_pdl(kern, grid, dim3(TC_THREADS), (size_t)Cfg::SMEM, st, tmA, tmB, C, ldc, sC, m, n, k, alpha, beta,
                               tri, batchA, batchB);
    if (e != cudaSuccess) return (int)e;
    return after_launch();
}

template <int BN, bool B_MN>
static int launch_tc_ta(const CUtensorMap& tmA, const CUtensorMap& tmB, float* C, int64_t ldc, int64_t sC, int m, int n,
                        int k, float alpha, float beta, int S, int tri, int batchA, int batchB, cudaStream_t st) {
    using Cfg = TaCfg<BN>;
    auto kern = gemm_tc_ta_kernel<BN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
            return (int)cudaGetLastError();
        attr_set = true;
    }
    dim3 grid(cdiv(n, BN), cdiv(m, TC_BM), S);
    kern<<<grid, TA_THREADS, Cfg::SMEM, st>>>(tmA, tmB, C, ldc, sC, m, n, k, alpha, beta, tri, batchA, batchB);
    return after_launch();
}

template <bool B_MN>
static int launch_tc_pe(const CUtensorMap& tmA, const CUtensorMap& tmB, float* C, int64_t ldc, int64_t sC, int m, int n,
                        int k, float alpha, float beta, int S, int tri, int batchA, int batchB, cudaStream_t st) {
    auto kern = gemm_tc_pe_kernel<B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PeCfg::SMEM) != cudaSuccess)
            return (int)cudaGetLastError();
        attr_set = true;
    }
    const int tiles_m = cdiv(m, TC_BM), tiles_n = cdiv(n, PeCfg::BN);
    const int64_t total = (int64_t)tiles_m * tiles_n * S;
    const int grid = (int)std::min<int64_t>(total, kNumSMs);
    static const int snake = [] { const char* e = getenv("MXF_GEMM_PE_SNAKE"); return e ? atoi(e) : 1; }();
    kern<<<grid, PeCfg::THREADS, PeCfg::SMEM, st>>>(tmA, tmB, C, ldc, sC, m, n, k, alpha, beta, tri, batchA, batchB, tiles_m,
                                                    tiles_n, S, (snake && (tri & 6)) ? 1 : 0);
    return after_launch();
}

static int pe_mode() {      // 0 = off, 1 = on for problems with more than one wave of 128 x 128 tiles and 256 <= K <= 2048 (default)
    static int v = [] { const char* e = getenv("MXF_GEMM_PE"); return e ? atoi(e) : 1; }();
    return v;
}

static int ta_mode() {      // 0 = off, 1 = on for long K loops (default)
    static int v = [] { const char* e = getenv("MXF_GEMM_TA"); return e ? atoi(e) : 1; }();
    return v;
}

// Returns MXF_ENOTIMPL when the problem does not fit the tensor-core path (caller falls back to the FMA kernel).
int gemm_tc_f32(int transA, int transB, int m, int n, int k, double alpha, const float* A, int64_t lda, int64_t sA,
                const float* B, int64_t ldb, int64_t sB, double beta, float* C, int64_t ldc, int64_t sC, int S, int tri,
                int wide, cudaStream_t st) {
    if (!tc_enabled() || transA || m <= 0 || n <= 0 || k <= 0 || S <= 0) return MXF_ENOTIMPL;
    if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return MXF_ENOTIMPL;
    if (S > 1 && (((sA != 0) && (sA & 3)) || ((sB != 0) && (sB & 3)))) return MXF_ENOTIMPL;
    const int batchA = (S > 1 && sA != 0) ? 1 : 0, batchB = (S > 1 && sB != 0) ? 1 : 0;
    // tile width: 64-wide tiles when 128-wide ones would leave most SMs idle
    const int64_t tiles128 = (int64_t)cdiv(n, 128) * cdiv(m, TC_BM) * S;
    static const int bn64_max = [] { const char* e = getenv("MXF_GEMM_BN64_MAX"); return e ? atoi(e) : 100; }();
    const bool bn64 = !wide && tiles128 < bn64_max;   // `wide`: C aliases A (in-place panel), one CTA must own full rows
    // 256-wide tiles halve the A-operand traffic per flop (the kernel is shared-memory-bandwidth bound): worth it once
    // they still fill the machine
    static const int bn256_min = [] { const char* e = getenv("MXF_GEMM_BN256_MIN"); return e ? atoi(e) : 120; }();
    const bool bn256 = !wide && !bn64 && (int64_t)cdiv(n, 256) * cdiv(m, TC_BM) * S >= bn256_min && n >= 256;
    const int BN = bn64 ? 64 : (bn256 ? 256 : 128);
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, A, m, k, lda, sA, batchA ? S : 1, TC_BK, TC_BM)) return MXF_ENOTIMPL;
    const bool b_mn = (transB == 0);
    if (!b_mn) {
        if (!make_map(&tmB, B, n, k, ldb, sB, batchB ? S : 1, TC_BK, BN)) return MXF_ENOTIMPL;
    } else {
        if (!make_map(&tmB, B, k, n, ldb, sB, batchB ? S : 1, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return MXF_ENOTIMPL;
    }
    const float al = (float)alpha, be = (float)beta;
    // persistent kernel (128-wide tiles, epilogue overlapped with the next tile's K loop): pays when a CTA gets several
    // tiles and the K loop is short enough for the per-tile overheads to matter
    static const int pe_kmax = [] { const char* e = getenv("MXF_GEMM_PE_KMAX"); return e ? atoi(e) : 2048; }();
    static const int pe_min_tiles = [] { const char* e = getenv("MXF_GEMM_PE_MIN_TILES"); return e ? atoi(e) : kNumSMs + 12; }();
    if (pe_mode() && !wide && k >= 256 && k <= pe_kmax && n >= 128 &&
        (int64_t)cdiv(n, 128) * cdiv(m, TC_BM) * S >= pe_min_tiles) {
        CUtensorMap tmBp;
        bool ok;
        if (!b_mn) ok = make_map(&tmBp, B, n, k, ldb, sB, batchB ? S : 1, TC_BK, PeCfg::BN);
        else ok = make_map(&tmBp, B, k, n, ldb, sB, batchB ? S : 1, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (ok)
            return b_mn ? launch_tc_pe<true>(tmA, tmBp, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                        : launch_tc_pe<false>(tmA, tmBp, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
    }
    if (ta_mode() && k >= 256) {
        if (bn256)
            return b_mn ? launch_tc_ta<256, true>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                        : launch_tc_ta<256, false>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
        if (bn64)
            return b_mn ? launch_tc_ta<64, true>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                        : launch_tc_ta<64, false>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
        return b_mn ? launch_tc_ta<128, true>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                    : launch_tc_ta<128, false>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
    }
    if (bn256) {
        return b_mn ? launch_tc<256, true>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                    : launch_tc<256, false>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
    }
    if (bn64) {
        return b_mn ? launch_tc<64, true>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                    : launch_tc<64, false>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
    }
    return b_mn ? launch_tc<128, true>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st)
                : launch_tc<128, false>(tmA, tmB, C, ldc, sC, m, n, k, al, be, S, tri, batchA, batchB, st);
}

}  // namespace mxf
