// Two-shot all-reduce of the gradient bucket over NVLink peer memory, as ONE kernel launch that can be captured in the
// step's CUDA graph (the data-parallel exchange of SURVEY.md section 8(e); the reference has no counterpart: it is
// single-device, the MXNet KVStore would play this role).
//
// Every rank owns a "symmetric" buffer (same size on every GPU, mapped into every peer's address space: allocated by
// torch.distributed._symmetric_memory -- plumbing -- the data path is this kernel):
//     [ n elements of gradient | flags ]
// 1. barrier: every rank's bucket is complete (its producers ran earlier on the same stream);
// 2. reduce-scatter by loads: rank r sums slice r of all W buckets, in rank order (so the result does not depend on which
//    rank computed it), scales it, and
// 3. all-gather by stores: writes the reduced slice into every rank's buffer (its own included);
// 4. barrier: all slices have landed everywhere.
// Per GPU (W - 1)/W of the bucket crosses NVLink once in each direction: 4.3 MB at W = 8 is ~10 us of wire time against the
// ~85 us of a NCCL call issued between the graph replay and the Adam launch -- and, being a plain kernel, it is captured with
// the step, so the data-parallel step is one graph replay like the single-GPU one.
//
// Flags: one 32-bit word per (barrier, block, source rank) in the DESTINATION rank's buffer, toggled 0 -> 1 by the source
// (atom.cas.release.sys) and 1 -> 0 by the owner (atom.cas.acquire.sys): self-resetting, so a graph replay needs no host
// reset.  Block b of every rank only talks to block b of the peers; all blocks are co-resident (grid <= 64).  A spin that
// exceeds `timeout_ns` sets *err and gives up (a dead peer becomes an error code, not a hung GPU).
#include "common.cuh"

namespace mxf {

constexpr int AR_MAX_WORLD = 16;
constexpr int AR_MAX_BLOCKS = 64;
constexpr int AR_THREADS = 512;

struct ArPeers {
    void* buf[AR_MAX_WORLD];
};

__device__ __forceinline__ unsigned long long ar_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}

// All threads of the block call.  Thread t < world raises flag (which, block, rank) on peer t and waits for flag
// (which, block, t) in its own buffer.
__device__ __forceinline__ bool ar_barrier(const ArPeers& p, int rank, int world, size_t flag_byte_off, int which,
                                           unsigned long long timeout_ns, int* err) {
    __threadfence_system();
    __syncthreads();
    __shared__ int ok_sh;
    if (threadIdx.x == 0) ok_sh = 1;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int t = threadIdx.x;
        const size_t slot = ((size_t)which * AR_MAX_BLOCKS + blockIdx.x) * AR_MAX_WORLD;
        uint32_t* theirs = reinterpret_cast<uint32_t*>(static_cast<char*>(p.buf[t]) + flag_byte_off) + slot + rank;
        uint32_t* mine = reinterpret_cast<uint32_t*>(static_cast<char*>(p.buf[rank]) + flag_byte_off) + slot + t;
        const unsigned long long t0 = ar_now();
        bool ok = true;
        while (cas_release_sys(theirs, 0u, 1u) != 0u)
            if (ar_now() - t0 > timeout_ns) { ok = false; break; }
        while (ok && cas_acquire_sys(mine, 1u, 0u) != 1u)
            if (ar_now() - t0 > timeout_ns) { ok = false; break; }
        if (!ok) { ok_sh = 0; if (err) atomicExch(err, 1); }
    }
    __syncthreads();
    return ok_sh != 0;
}

template <typename V> __device__ __forceinline__ V ld_peer(const V* p);
template <> __device__ __forceinline__ float4 ld_peer<float4>(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <> __device__ __forceinline__ double2 ld_peer<double2>(const double2* p) {
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void vacc(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vacc(double2& a, const double2& b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vscale(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }
__device__ __forceinline__ void vscale(double2& a, double s) { a.x *= s; a.y *= s; }

// nv: number of 16-byte vectors of the bucket (the buffer is padded to a multiple of 16 bytes)
template <typename V, typename S, int W>
__global__ void __launch_bounds__(AR_THREADS)
allreduce_two_shot_kernel(const ArPeers p, int rank, int world_rt, int64_t nv, S scale, size_t flag_byte_off,
                          unsigned long long timeout_ns, int* err) {
    const int world = W > 0 ? W : world_rt;
    if (!ar_barrier(p, rank, world, flag_byte_off, 0, timeout_ns, err)) return;
    const int64_t chunk = (nv + world - 1) / world;
    const int64_t lo = (int64_t)rank * chunk, hi = min(nv, lo + chunk);
    for (int64_t i = lo + (int64_t)blockIdx.x * AR_THREADS + threadIdx.x; i < hi; i += (int64_t)gridDim.x * AR_THREADS) {
        V acc;
        if (W > 0) {
            V v[W > 0 ? W : 1];
#pragma unroll
            for (int r = 0; r < W; ++r) v[r] = ld_peer<V>(static_cast<const V*>(p.buf[r]) + i);      // all loads in flight
            acc = v[0];
#pragma unroll
            for (int r = 1; r < W; ++r) vacc(acc, v[r]);
        } else {
            acc = ld_peer<V>(static_cast<const V*>(p.buf[0]) + i);
            for (int r = 1; r < world; ++r) vacc(acc, ld_peer<V>(static_cast<const V*>(p.buf[r]) + i));
        }
        vscale(acc, scale);
#pragma unroll
        for (int r = 0; r < (W > 0 ? W : AR_MAX_WORLD); ++r)
            if (r < world) static_cast<V*>(p.buf[r])[i] = acc;
    }
    ar_barrier(p, rank, world, flag_byte_off, 1, timeout_ns, err);
}

}  // namespace mxf

// Bytes of flag space that must follow the (16-byte padded) bucket in every rank's symmetric buffer; zeroed once by the host
// before the first call (then a barrier across the ranks).
extern "C" size_t mxf_allreduce_p2p_flag_bytes(void) { return (size_t)2 * mxf::AR_MAX_BLOCKS * mxf::AR_MAX_WORLD * 4; }

// In-place sum (times `scale`) of the first n elements of the W symmetric buffers bufs[0..W-1] (device pointers valid in THIS
// process; bufs[rank] is the local one); flags at byte offset flag_byte_off (16-byte aligned, >= n * sizeof(element)) of every
// buffer.  Must be called by all ranks, in the same order.  err: device int set to 1 when a peer did not show up in time.
extern "C" int mxf_allreduce_p2p(int dtype, void* const* bufs, int rank, int world, int64_t n, double scale,
                                 size_t flag_byte_off, double timeout_s, int* err, void* stream) {
    using namespace mxf;
    if (world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world || n < 0 || !bufs) return MXF_EINVAL;
    if (dtype != MXF_F32 && dtype != MXF_F64) return MXF_EDTYPE;
    const size_t esz = dtype == MXF_F32 ? 4 : 8;
    if ((flag_byte_off & 15) || flag_byte_off < (((size_t)n * esz + 15) / 16) * 16) return MXF_EINVAL;
    ArPeers p;
    for (int r = 0; r < AR_MAX_WORLD; ++r) p.buf[r] = r < world ? bufs[r] : nullptr;
    for (int r = 0; r < world; ++r)
        if (!p.buf[r] || (reinterpret_cast<uintptr_t>(p.buf[r]) & 15)) return MXF_EINVAL;
    const int64_t nv = ((int64_t)n * (int64_t)esz + 15) / 16;
    const int64_t per_rank = (nv + world - 1) / world;
    int blocks = (int)std::min<int64_t>(AR_MAX_BLOCKS, std::max<int64_t>(1, (per_rank + AR_THREADS - 1) / AR_THREADS));
    const unsigned long long tns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MXF_AR_LAUNCH(V, S, W) \
    allreduce_two_shot_kernel<V, S, W><<<blocks, AR_THREADS, 0, st>>>(p, rank, world, nv, (S)scale, flag_byte_off, tns, err)
    if (dtype == MXF_F32) {
        if (world == 2) MXF_AR_LAUNCH(float4, float, 2);
        else if (world == 4) MXF_AR_LAUNCH(float4, float, 4);
        else if (world == 8) MXF_AR_LAUNCH(float4, float, 8);
        else MXF_AR_LAUNCH(float4, float, 0);
    } else {
        if (world == 2) MXF_AR_LAUNCH(double2, double, 2);
        else if (world == 4) MXF_AR_LAUNCH(double2, double, 4);
        else if (world == 8) MXF_AR_LAUNCH(double2, double, 8);
        else MXF_AR_LAUNCH(double2, double, 0);
    }
#undef MXF_AR_LAUNCH
    return after_launch();
}
