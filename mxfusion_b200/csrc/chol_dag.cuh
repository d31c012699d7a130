// Interface of the single-launch tile-dataflow Cholesky + inverse (chol_dag.cu), used by chol_packed.cu.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mxf {

constexpr int DG_B = 64;                   // tile edge
constexpr int DG_THREADS = 256;
constexpr int DG_MAXT = 16;                // tiles per side
constexpr int DG_MAXN = DG_B * DG_MAXT;    // 1024
constexpr int DG_SYNC_INTS = 2 + 2 * DG_MAXT * DG_MAXT + 2;   // ticket counter, abort flag, L-final and W-final flags

// mode 1: A (n x n, SPD, lower part read) -> L in place (strict upper zeroed), W = L^-1, optional W^T, L^T and the
// inverses of the 128 x 128 diagonal blocks (+ transposes); mode 0: A holds a lower factor, only the inverse half runs.
// All outputs live in the per-sample `pack` at the given element offsets (a negative oWT / oLT / oDinv skips that output);
// W and W^T have row stride ldw, L^T has row stride ldlt.  `oSync` is DG_SYNC_INTS ints of scratch (zeroed here).
// info (may be null): first non-positive pivot, 1-based, + info_base; written with atomicCAS(0 -> value).
int dag_launch(int mode, float* A, int64_t lda, int64_t sA, int n, float* pack, int64_t sP, int64_t oW, int64_t oWT,
               int64_t oLT, int64_t oDinv, int64_t oDinvT, int64_t oSync, int ldw, int ldlt, int* info, int info_base, int S,
               cudaStream_t st);
int dag_tickets(int T, int mode);
extern std::atomic<int> g_dag_ctas;        // CTAs per matrix of the persistent dataflow kernel (default: one per SM)
int dag_set_prof(long long* dev_ptr);      // debug: 16 stamps per diagonal ticket (clock64; [14], [15] globaltimer ns)

}  // namespace mxf
