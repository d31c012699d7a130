/*
 * mxf_b200.h -- C ABI of libmxf_b200.so: the sm_100a kernels behind the
 * MXFusion variational-inference / Gaussian-process hot path.
 *
 * The reference (amzn/MXFusion) has no native code: every entry point below
 * replaces an MXNet operator call that the reference issues through its `F`
 * namespace.  The reference-side binding is a ctypes stub
 * (INTEGRATION.md); the in-tree binding is mxfusion_b200/_lib.py.
 *
 * Conventions
 *   - row-major, innermost dimension contiguous, batch-major over the
 *     reference's leading sample axis S (components/variables/
 *     runtime_variable.py:20-50).  Batch strides are given in ELEMENTS; a batch
 *     stride of 0 broadcasts one matrix over all S samples
 *     (runtime_variable.py:102-118 `arrays_as_samples`, without the copy).
 *   - dtype: MXF_F32 (reference default, common/config.py:18) or MXF_F64
 *     (what the reference's tests use).
 *   - ownership: the caller owns every buffer, including workspaces.  The
 *     library never allocates, frees or synchronises; all work is enqueued on
 *     the cudaStream_t passed as `stream` (NULL = legacy default stream).
 *   - errors: 0 = ok; negative = invalid argument (MXF_EINVAL...); positive =
 *     the cudaError_t of the failed launch.  No exception or abort crosses the
 *     ABI.  A non-positive-definite potrf input is reported through the
 *     device-side `info` array (1-based index of the first bad pivot, 0 = ok).
 *   - re-entrant: no global mutable state.
 */
#ifndef MXF_B200_H
#define MXF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MXF_F32 0
#define MXF_F64 1

#define MXF_OK 0
#define MXF_EINVAL (-1)
#define MXF_EDTYPE (-2)
#define MXF_ENOTIMPL (-3)
#define MXF_EWORKSPACE (-4)

/* stationary kernel kinds: gp/kernels/rbf.py:71-72, matern.py:84-88,116-120,148-151 */
#define MXF_KERN_RBF 0
#define MXF_KERN_MATERN12 1
#define MXF_KERN_MATERN32 2
#define MXF_KERN_MATERN52 3

int mxf_version(void);
/* number of kernel launches enqueued by this library in this process so far */
uint64_t mxf_launch_count(void);

/* ---- covariance build ---------------------------------------------------
 * Replaces StationaryKernel._compute_R2 + RBF/Matern._compute_K
 * (gp/kernels/stationary.py:74-107, rbf.py:54-72, matern.py:67-151): the
 * scale -> squared distance (|a|^2+|b|^2-2a.b) -> kernel -> store chain in ONE
 * pass over the (S,N,N2) output.
 *   X (S,N,D), X2 (S,N2,D) or NULL (then X2 := X, N2 := N),
 *   ls (S,ls_len) with ls_len in {1, D}, var (S,1)  ->  out (S,N,N2), ldo = row stride of out.
 *   diag_add (S,1) or NULL, diag_add_const: out[i][i] += diag_add[s] + diag_add_const
 *   when X2 == NULL (the `+ eye*noise_var`, `+ eye*jitter` of gp_regression.py:55-60,
 *   svgp_regression.py:70-72 folded into the store). */
int mxf_kbuild_fwd(int kind, int dtype,
                   const void* X, const void* X2, const void* ls, int ls_len, const void* var,
                   const void* diag_add, double diag_add_const,
                   void* out, int64_t ldo,
                   int S, int N, int N2, int D,
                   int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar, int64_t sDiag, int64_t sOut,
                   void* stream);

/* Tuning knob of mxf_kbuild_fwd: f32 cross-covariances (X2 != NULL, D <= 16, N2 >= 128) with at least `min_elems`
 * output elements take the tcgen05 + TMA-store kernel (csrc/kbuild_tc.cuh: the gemm2 of stationary.py:102 on the tensor
 * pipe), smaller ones the streaming FMA kernel.  Returns the previous value; a negative argument only queries. */
long long mxf_kbuild_tc_threshold(long long min_elems);

/* Adjoint of mxf_kbuild_fwd (replaces MXNet autograd through the same ops).
 *   G (S,N,N2) with row stride ldg -> dX (S,N,D) or NULL, dX2 (S,N2,D) or NULL,
 *   dls (S,ls_len), dvar (S,1); every output is OVERWRITTEN (write, not add).
 *   When X2 == NULL the symmetric contribution of both arguments goes to dX.
 *   ws: workspace of mxf_kbuild_bwd_workspace_bytes() bytes. */
size_t mxf_kbuild_bwd_workspace_bytes(int dtype, int S, int N, int N2, int D);
int mxf_kbuild_bwd(int kind, int dtype,
                   const void* X, const void* X2, const void* ls, int ls_len, const void* var,
                   const void* G, int64_t ldg,
                   void* dX, void* dX2, void* dls, void* dvar,
                   int S, int N, int N2, int D,
                   int64_t sX, int64_t sX2, int64_t sLs, int64_t sVar, int64_t sG,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- dense linear algebra (replaces mx.nd.linalg.*) -----------------------
 * gemm2 / syrk / trmm: C = alpha * op(A) op(B) + beta * C   (svgp_regression.py:76,82,89,90)
 *   op(A) is m x k, op(B) is k x n, C is m x n; lda/ldb/ldc row strides; sA/sB/sC batch strides.
 *   tri: 0 = full C; 1 = only tiles touching the lower triangle of C are computed
 *   (syrk-type trailing update).  Elements strictly above those tiles are left untouched. */
int mxf_gemm(int dtype, int transA, int transB, int m, int n, int k,
             double alpha, const void* A, int64_t lda, int64_t sA,
             const void* B, int64_t ldb, int64_t sB,
             double beta, void* C, int64_t ldc, int64_t sC,
             int S, int tri, void* stream);

/* linalg.potrf (svgp_regression.py:83-84, gp_regression.py:61): in-place lower
 * Cholesky of the S n x n matrices at A (row stride lda, batch stride sA); the
 * strict upper triangle is zeroed (MXNet convention).  info: int32[S] on device. */
int mxf_potrf(int dtype, void* A, int64_t lda, int64_t sA, int S, int n, int* info, void* stream);

/* linalg.trsm, left side, lower triangular A (svgp_regression.py:85-87,92; gp_regression.py:66):
 *   B := alpha * A^-1 B (transpose=0) or alpha * A^-T B (transpose=1); B is n x nrhs. */
int mxf_trsm(int dtype, int transpose, int n, int nrhs, double alpha,
             const void* A, int64_t lda, int64_t sA,
             void* B, int64_t ldb, int64_t sB, int S, void* stream);

/* ---- tensor-core formulation of potrf / trsm --------------------------------------------------
 * The "pack" of a lower factor L (n x n), per sample, is  [ Dinv | DinvT | LT ]:  the inverses of the NB x NB diagonal
 * blocks of L (NB = mxf_tri_block(dtype): 128 for f32, 64 for f64; the last block is padded with the identity), their
 * transposes, and L^T (row stride (n+3)&~3).  It holds mxf_tri_pack_elems(dtype, n) elements per sample (a multiple
 * of 4).  mxf_potrf_packed factors A in place exactly like mxf_potrf AND emits the pack as a by-product (the panel
 * solve and the trailing update are GEMMs -- tcgen05 for f32); mxf_tri_pack builds the pack of an existing factor;
 * mxf_trsm_packed solves with it as a chain of GEMMs  X_k = Dinv_k B_k ; B_rest -= L_rest,k X_k.
 * sP = batch stride of the pack in elements (0 = one factor shared by all S right-hand sides). */
int mxf_tri_block(int dtype);
size_t mxf_tri_pack_elems(int dtype, int n);
int mxf_tri_pack(int dtype, const void* L, int64_t lda, int64_t sA, int S, int n, void* pack, void* stream);
int mxf_potrf_packed(int dtype, void* A, int64_t lda, int64_t sA, int S, int n, int* info, void* pack, void* stream);
int mxf_trsm_packed(int dtype, int transpose, int n, int nrhs, double alpha, const void* L, int64_t lda, int64_t sA,
                    const void* pack, int64_t sP, void* B, int64_t ldb, int64_t sB, int S, void* stream);

/* acc[0] = max(acc[0], max_i |info[i]|): folds the `info` of the factorisations of a step (first non-positive pivot,
 * 1-based; the reference surfaces a non-PD matrix as an MXNetError at its next synchronisation) into
 * one device int that the host reads every few steps and maps to InferenceError (common/exceptions.py:20). */
int mxf_info_max(int* acc, const int* info, int n, void* stream);

/* Element offsets of the pieces of a factor's pack (per sample): out[0..7] = {Dinv, DinvT, L^T, W, W^T, top, ldt, nq}.
 * W / W^T are nq blocks of top x top elements (row stride top) holding the explicit inverse of the factor's diagonal
 * blocks of size top (f32, n <= 1024: ONE block, W = L^-1 -- written by the single-launch tile-dataflow potrf). */
int mxf_tri_pack_layout(int dtype, int n, int64_t* out);

/* When n is a multiple of 2*NB the pack also holds hierarchically built inverses of larger diagonal blocks
 * ([[Wa,0],[-Wc Lca Wa, Wc]], doubling from NB up to mxf_tri_top_block(dtype, n) <= 512, env MXF_TRI_INV) and their
 * transposes.  mxf_trsm_packed_oop then solves in n/top block steps of one or two LARGE GEMMs each (the triangular
 * structure of the inverse blocks is used to skip the zero half of K): X = op(L)^-1 B, out of place; B is scratch.
 * Returns MXF_ENOTIMPL when the pack has no level above NB (use mxf_trsm_packed). */
int mxf_tri_top_block(int dtype, int n);
int mxf_trsm_packed_oop(int dtype, int transpose, int n, int nrhs, const void* L, int64_t lda, int64_t sA,
                        const void* pack, int64_t sP, void* B, int64_t ldb, int64_t sB, void* X, int64_t ldx, int64_t sX,
                        int S, void* stream);

/* Cholesky adjoint helper: out = phi(P) + phi(P)^T with phi = lower triangle, i.e. the
 * symmetric matrix whose lower triangle (diagonal included) is copied from P. */
int mxf_copy_ltu(int dtype, const void* P, int64_t ldp, int64_t sP,
                 void* out, int64_t ldo, int64_t sO, int S, int n, void* stream);
/* Same with the lower triangle taken from the SUM of `parts` matrices P + g * sPart (the split-K partial products of
 * Phi = A A^T, replacing F.sum + the mirror; svgp_regression.py:89-90 as folded in ops.py). */
int mxf_copy_ltu_sum(int dtype, const void* P, int64_t ldp, int64_t sPart, int parts,
                     void* out, int64_t ldo, int n, void* stream);
/* Strided batched 2-D copy dst[s][r][c] = src[s][r][c] (src == NULL: zero fill): the slice assignments around the
 * solves of svgp_regression.py:85-87 (F.concat / slice_axis in the reference). */
int mxf_copy2d(int dtype, const void* src, int64_t lds, int64_t sS, void* dst, int64_t ldd, int64_t sD,
               int S, int rows, int cols, void* stream);
/* out = alpha * (A + A^T)  (symmetrise; in-place allowed only if out != A is false -> use distinct buffers) */
int mxf_symmetrize(int dtype, double alpha, const void* A, int64_t lda, int64_t sA,
                   void* out, int64_t ldo, int64_t sO, int S, int n, void* stream);
/* out = tril(A) (mode 0, diagonal kept) / strictly-lower (mode 1); upper part zero-filled. */
int mxf_tril(int dtype, int mode, const void* A, int64_t lda, int64_t sA,
             void* out, int64_t ldo, int64_t sO, int S, int n, void* stream);
/* out (n x m) = A^T, A is m x n. */
int mxf_transpose(int dtype, const void* A, int64_t lda, int64_t sA,
                  void* out, int64_t ldo, int64_t sO, int S, int m, int n, void* stream);

/* ---- reductions (linalg.sumlogdiag, F.sum(F.square(.)), ...) ---------------
 * svgp_regression.py:94-107, gp_regression.py:67-70.  One output scalar per sample. */
#define MXF_RED_SUM 0      /* sum a                 */
#define MXF_RED_SUMSQ 1    /* sum a*a               */
#define MXF_RED_DOT 2      /* sum a*b               */
#define MXF_RED_SUMLOG 3   /* sum log(a)            */
#define MXF_RED_SUMSQDIFF 4 /* sum (a-b)^2          */
/* a, b: (S, rows, cols) with row strides lda/ldb; out[s] = scale * reduce.  b may be NULL. */
int mxf_reduce(int op, int dtype, const void* a, int64_t lda, int64_t sA,
               const void* b, int64_t ldb, int64_t sB,
               int S, int64_t rows, int64_t cols, double scale, void* out, void* stream);
/* linalg.sumlogdiag(|A|) for n x n matrices: out[s] = sum_i log|A[s][i][i]|. */
int mxf_sumlogdiag(int dtype, const void* A, int64_t lda, int64_t sA, int S, int n,
                   void* out, void* stream);
/* make_diagonal (util/customop.py:22-81) fused into its only consumer:
 * A[s][i][i] += d[s][i] (+ c).  Used for S = syrk(W) + make_diagonal(diag). */
int mxf_add_diag(int dtype, void* A, int64_t lda, int64_t sA, const void* d, int64_t sD,
                 double c, int S, int n, void* stream);
/* out[s][i] = A[s][i][i]  (backward of make_diagonal, customop.py:45-57). */
int mxf_get_diag(int dtype, const void* A, int64_t lda, int64_t sA, void* out, int64_t sO,
                 int S, int n, void* stream);

/* out[s] = a[s] * X[s] + b[s] * Y[s] elementwise over n elements per sample; a, b are DEVICE scalars
 * (one per sample; NULL = 1 for a, 0 for b and then Y may be NULL) so that coefficients that depend on
 * learnable parameters or on the upstream gradient never visit the host.  sX/sY/sO batch strides. */
int mxf_axpby_dev(int dtype, const void* a, const void* X, int64_t sX, const void* b, const void* Y, int64_t sY,
                  void* out, int64_t sO, int S, int64_t n, void* stream);
/* The same on strided (rows x cols) views: out[s][r][c] = a[s] X[s][r][c] + b[s] Y[s][r][c]; a / b == NULL: 1,
 * Y == NULL: no second term.  Combines blocks of the solve buffers of svgp_regression.py:85-92 where they lie. */
int mxf_axpby2d(int dtype, const void* a, const void* X, int64_t ldx, int64_t sX, const void* b, const void* Y,
                int64_t ldy, int64_t sY, void* out, int64_t ldo, int64_t sO, int S, int rows, int cols, void* stream);

/* Softplus parameter transform (components/variables/var_trans.py:63-91, applied to every constrained
 * parameter on every forward, inference_alg.py:79-80): y = log(1 + exp(x)) + offset, overflow-safe;
 * adjoint gx = gy * sigmoid(x). */
int mxf_softplus_fwd(int dtype, const void* x, double offset, void* y, int64_t n, void* stream);
int mxf_softplus_bwd(int dtype, const void* x, const void* gy, void* gx, int64_t n, void* stream);

/* ---- SVGP bound, adjoint assembly (analytic gradient of svgp_regression.py:43-109) ----
 * With Phi = A A^T (A = L^-1 Kuf), T = C C^T (C = L^-1 Ls), U = Phi T, mt = L^-1 mu (M x P),
 * v = A (Y - A^T mt) (M x P) and per-sample device coefficients coef[s][0..5] =
 * {gP/2, c1 = g*scale*P*beta/2, g/2, g*scale*beta/2, g*scale*P*beta, g*scale*beta} this writes the three
 * symmetric M x M matrices whose two-/one-sided solves with L give the gradients wrt Kuu, S and Kuf:
 *   E   = coef2 mt mt^T + coef0 (T - I) - c1 Phi + c1 (U + U^T) - coef3 (v mt^T + mt v^T)
 *   E_S = coef0 I + c1 Phi
 *   E_R = -coef4 (T - I) - coef5 mt mt^T
 * into out as [E | E_S | E_R] (row stride ldo >= 3M, batch stride sO; extra columns are left untouched so that more
 * right-hand sides can ride along in the same triangular solve). */
int mxf_svgp_bwd_assemble(int dtype, const void* Phi, const void* T, const void* U, const void* mt, const void* v,
                          const void* coef, void* out, int64_t ldo, int64_t sO, int S, int M, int P, void* stream);

/* Scalar head / tail of the SVGP bound (svgp_regression.py:94-108) on the per-sample reductions, one launch each way
 * instead of ~20 scalar operators: with nv = noise[s], kv = kvar[s], beta = 1/nv,
 *   Q     = -sumr2/2 - P B kv/2 - P (trPhiT - trPhi)/2
 *   logL  = scale (beta Q - B P (log 2pi + log nv)/2) + P (M/2 + sldLs - sldL) - P trT/2 - mm/2
 * (sumr2 = sum (Y - A^T mt)^2, trPhi = tr(A A^T), trT = tr(C C^T), trPhiT = tr(Phi T), mm = sum mt^2, sldL / sldLs =
 * sumlogdiag of the two factors).  Outputs logL, beta, Q (S each).  The adjoint head writes coef (S x 6, the layout
 * mxf_svgp_bwd_assemble reads), gsb = g scale beta, -gsb, d logL / d noise_var, the Kff_diag part of d / d variance, and the
 * per-sample constants -g and -1 that the remaining axpby launches take as device coefficients. */
int mxf_svgp_bound_fwd(int dtype, int S, int P, int B, int M, double scale, const void* sumr2, const void* trPhi,
                       const void* trT, const void* trPhiT, const void* mm, const void* sldL, const void* sldLs,
                       const void* noise, int64_t sNoise, const void* kvar, int64_t sKvar, void* logL, void* beta,
                       void* Q, void* stream);
int mxf_svgp_coef_bwd(int dtype, int S, int P, int B, double scale, const void* g, const void* beta, const void* Q,
                      void* coef, void* gsb, void* neg_gsb, void* dnoise, void* dkvar_diag, void* neg_g, void* minus_one,
                      void* stream);

/* ---- Normal distribution: MC-ELBO pieces (normal.py:52-92, factor_graph.py:223) ----
 * Fused log-density + sample-mean + sum:
 *   out[0] = scale * (1/S) * sum_{s,i} [ -0.5 log 2pi - 0.5 log v - (x-m)^2 / (2 v) ]
 * x (Sx,n), m (Sm,n), v (Sv,n) with S = max(Sx,Sm,Sv); an operand with batch
 * stride 0 is broadcast over samples.  out must be zero-initialised by the caller
 * (accumulated with one atomic per CTA). */
int mxf_normal_logpdf_sum(int dtype, const void* x, int64_t sX, const void* m, int64_t sM,
                          const void* v, int64_t sV, int S, int64_t n, double scale,
                          void* out, void* stream);
/* Adjoint: gx/gm/gv (any may be NULL) receive d out / d x,m,v times gout[0];
 * an operand broadcast over S gets the sum over samples.  Outputs are overwritten. */
int mxf_normal_logpdf_sum_bwd(int dtype, const void* x, int64_t sX, const void* m, int64_t sM,
                              const void* v, int64_t sV, int S, int64_t n, double scale,
                              const void* gout, void* gx, void* gm, void* gv, void* stream);
/* Multi-tensor form of the two calls above: `count` Normal factors in one launch each way (HOST arrays of DEVICE
 * pointers / per-entry sizes).  Entry t: x, m, v with sample strides sX / sM / sV (0 = shared by the S[t] samples), n[t]
 * elements per sample, scalar[t] bit 0 / 1 / 2 = x / m / v is ONE element per sample (a constant prior parameter is read in
 * place, not broadcast into a copy), scale[t].  out[0] += sum_t scale_t / S_t sum log N (caller zeroes out).
 * The adjoint writes gx[t] / gm[t] / gv[t] (NULL = not needed; scalar operands cannot receive a gradient here). */
int mxf_normal_logpdf_multi(int dtype, int count, const void* const* x, const void* const* m, const void* const* v,
                            const int64_t* sX, const int64_t* sM, const int64_t* sV, const int64_t* n, const int* S,
                            const int* scalar, const double* scale, void* out, void* stream);
int mxf_normal_logpdf_multi_bwd(int dtype, int count, const void* const* x, const void* const* m, const void* const* v,
                                const int64_t* sX, const int64_t* sM, const int64_t* sV, const int64_t* n, const int* S,
                                const int* scalar, const double* scale, const void* gout, void* const* gx,
                                void* const* gm, void* const* gv, void* stream);
/* Reparameterised draw (normal.py:89-92): w[s][i] = eps[s][i]*sqrt(v[i]) + m[i].
 * If eps == NULL a counter-based Philox4x32-10 standard-normal stream keyed by
 * (seed, offset) is generated in-kernel and, if eps_out != NULL, stored for the adjoint.
 * step_counter (device int32[1] or NULL): its value is added to the high word of the Philox counter at run time, so
 * a launch replayed from a CUDA graph draws fresh noise every optimiser step (pass the counter mxf_adam_step bumps). */
int mxf_normal_reparam(int dtype, const void* eps, const void* m, int64_t sM, const void* v, int64_t sV,
                       int S, int64_t n, uint64_t seed, uint64_t offset, const int* step_counter, void* w,
                       void* eps_out, void* stream);
/* Adjoint of the draw: gm = gw, gv = gw * eps / (2 sqrt(v)); an operand shared by the samples (sM / sV = 0) receives
 * the sum over samples; gm / gv may be NULL. */
int mxf_normal_reparam_bwd(int dtype, const void* gw, const void* eps, const void* v, int64_t sM, int64_t sV, int S,
                           int64_t n, void* gm, void* gv, void* stream);
/* Multi-tensor forms (HOST arrays of DEVICE pointers / sizes): `count` independent draws in one launch, entry t with its
 * own Philox offset[t] (same seed and step counter), and their adjoints in one launch. */
int mxf_normal_reparam_multi(int dtype, int count, const void* const* m, const void* const* v, const int64_t* sM,
                             const int64_t* sV, const int64_t* n, const int* S, uint64_t seed, const uint64_t* offset,
                             const int* step_counter, void* const* w, void* const* eps_out, void* stream);
int mxf_normal_reparam_multi_bwd(int dtype, int count, const void* const* gw, const void* const* eps,
                                 const void* const* v, const int64_t* sM, const int64_t* sV, const int64_t* n,
                                 const int* S, void* const* gm, void* const* gv, void* stream);

/* ---- optimiser (mx.gluon.Trainer('adam').step(batch_size), minibatch_loop.py:71-91) ----
 * One fused update over a flat parameter bucket:
 *   g *= rescale; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *   w -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)
 * step_count: device int32[1], incremented by the kernel (so the update can live in a CUDA graph). */
int mxf_adam_step(int dtype, void* w, const void* g, void* m, void* v, int64_t n,
                  double lr, double beta1, double beta2, double eps, double rescale,
                  int* step_count, void* stream);

/* mx.gluon.Trainer('sgd') (grad_based_inference.py:67 accepts any optimiser name): w -= lr * rescale * g, or with
 * momentum != 0 (mom: state bucket) mom = momentum * mom - lr * rescale * g; w += mom.  Bumps step_count like Adam. */
int mxf_sgd_step(int dtype, void* w, const void* g, void* mom, int64_t n, double lr, double momentum, double rescale,
                 int* step_count, void* stream);

/* ---- flat parameter bucket: multi-tensor transform and gradient gather (inference_alg.py:79-80 applies
 *      `var_trans[k].transform` per constrained parameter on every forward; gluon Trainer / autograd accumulate one
 *      gradient per parameter) ----
 * Segment t of the flat buffers is [off[t], off[t] + n[t]); kind[t] = 0 identity, 1 softplus (+ offset[t]).
 * mxf_params_transform: tflat[seg] = transform(flat[seg]) for every segment, one launch (per 24 segments).
 * mxf_params_pack_grads: gflat[seg] = grads[t] * d transform / d raw (grads[t]: HOST array of DEVICE pointers to the
 * gradient w.r.t. the transformed value; NULL = no gradient, the segment is zero-filled); overwrites gflat. */
int mxf_params_transform(int dtype, int count, const int64_t* off, const int64_t* n, const int* kind,
                         const double* offset, const void* flat, void* tflat, void* stream);
int mxf_params_pack_grads(int dtype, int count, const void* const* grads, const int64_t* off, const int64_t* n,
                          const int* kind, const void* flat, void* gflat, void* stream);

/* ---- minibatch gather (gluon DataLoader batchify, minibatch_loop.py:68-70,78) ----
 * out[r][:] = src[idx[off + r]][:] for r < rows; idx is int64 on device; bit-exact copy.
 * `off` is read from device memory (int64[1]) so the gather can be replayed inside a CUDA graph. */
int mxf_gather_rows(int dtype, const void* src, int64_t cols, const int64_t* idx, const int64_t* off,
                    int64_t rows, void* out, void* stream);

/* ---- dense-tanh network with sampled weights (FunctionEvaluation.eval over an MXFusionGluonFunction,
 *      function_evaluation.py:47-99, mxfusion_gluon_function.py:97-111,166-194; BASELINE config 4) ----
 *   out[s, r, :] = W_L^s tanh( ... tanh(W_1^s x[s|0, r, :] + b_1^s) ... ) + b_L^s     (Dense layout: W (out, in))
 * replaces the reference's Python loop over the S weight samples (one Gluon forward and one autograd tape each).
 * n_layers <= 4 dense layers, every width <= 64 (else MXF_ENOTIMPL); widths[0..n_layers] on the host.
 * W, b, dW, db: HOST arrays of n_layers DEVICE pointers; sW / sb: sample strides in elements (0 = the tensor is shared
 * by all S samples).  b (and db) may be NULL (no bias).  x: (S|1, B, widths[0]) with sample stride sx (0 = shared).
 * The adjoint recomputes the forward pass and ADDS into dW / db (caller zeroes them): dW_l^s = sum_r delta_l act_{l-1}^T. */
int mxf_mlp_tanh_fwd(int dtype, int n_layers, const int* widths, const void* x, int64_t sx,
                     const void* const* W, const int64_t* sW, const void* const* b, const int64_t* sb,
                     void* out, int S, int B, void* stream);
int mxf_mlp_tanh_bwd(int dtype, int n_layers, const int* widths, const void* x, int64_t sx,
                     const void* const* W, const int64_t* sW, const void* const* b, const int64_t* sb,
                     const void* gout, void* const* dW, void* const* db, int S, int B, void* stream);

/* Debug hooks (scripts/prof_dag.py, scripts/prof_diag.py): a device buffer of int64 that the potrf kernels fill with
 * clock64 / globaltimer stamps per phase; NULL switches the stamps off (the default).  Not part of the drop-in surface. */
int mxf_debug_set_prof(void* dev_ptr);
int mxf_debug_set_dag_prof(void* dev_ptr);

/* Tuning knob of the single-launch dataflow potrf (f32, n <= 1024): CTAs per matrix of the persistent kernel (default 148 =
 * one per SM, so that two factorisations issued on two streams are both fully resident).  Returns the previous value;
 * ctas <= 0 only queries. */
int mxf_potrf_dag_ctas(int ctas);

/* ---- data-parallel exchange --------------------------------------------------------------
 * The reference is single-device (its multi-device story would be the MXNet KVStore behind gluon.Trainer.step,
 * inference/grad_based_inference.py:67, batch_loop.py:58); here the averaged gradient of the ranks is ONE kernel over
 * NVLink peer memory (csrc/allreduce_p2p.cu) that is captured in the step's CUDA graph.
 *   bufs[0..world-1]: device pointers (valid in this process) of the ranks' symmetric buffers, laid out as
 *   [ n elements | padding to 16 bytes | mxf_allreduce_p2p_flag_bytes() bytes of flags, zeroed once ]; flags start at
 *   flag_byte_off.  In place: every buffer ends up holding scale * sum over ranks.  *err (device int) is set to 1 when
 *   a peer did not arrive within timeout_s (<= 0: 10 s). */
size_t mxf_allreduce_p2p_flag_bytes(void);
int mxf_allreduce_p2p(int dtype, void* const* bufs, int rank, int world, int64_t n, double scale,
                      size_t flag_byte_off, double timeout_s, int* err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MXF_B200_H */
