"""Differentiable operators of the hot path: torch.autograd.Function wrappers whose forward AND
backward are C-ABI kernel launches (mxfusion_b200._raw -> libmxf_b200.so).

These stand where the reference calls MXNet operators through `F` and leaves the adjoints to MXNet
autograd.  Two kinds of operator live here:

* primitives that mirror the MXNet operators one for one (`kernel_matrix`, `potrf`, `trsm`, `gemm2`,
  `syrk`, `sumlogdiag`, `make_diagonal`, `softplus`, the Normal pieces), so that a user-written
  `InferenceAlgorithm.compute` reads like the reference's;
* the two fused module bounds, `svgp_log_pdf` (svgp_regression.py:43-109) and `gp_log_pdf`
  (gp_regression.py:42-76), each one autograd node with an analytic, hand-derived gradient.

Every array carries the reference's leading sample axis (runtime_variable.py:20-50).
"""
import math

import torch

from . import _raw as R   # tests may monkeypatch `ops.R` with a stand-in to check the algebra without a GPU

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
_LOG2PI = math.log(2.0 * math.pi)


def _lead(*ts):
    return max(t.shape[0] for t in ts if t is not None)


def _expand(t, S):
    """Differentiable broadcast of the sample axis (arrays_as_samples, runtime_variable.py:102-118)."""
    if t is None or t.shape[0] == S:
        return t
    return t.expand((S,) + tuple(t.shape[1:]))


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
class _Softplus(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offset):
        ctx.save_for_backward(x)
        return R.softplus_fwd(x, offset)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return R.softplus_bwd(x, gy), None


def softplus(x, offset=0.0):
    """var_trans.py:75 (Activation softrelu) + offset."""
    return _Softplus.apply(x, offset)


class _KernelMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, X, X2, ls, var, diag_add, diag_const):
        ctx.kind = kind
        ctx.save_for_backward(X, X2, ls, var)
        ctx.has_diag = diag_add is not None
        return R.kbuild_fwd(kind, X, X2, ls, var, diag_add=diag_add, diag_const=diag_const)

    @staticmethod
    def backward(ctx, G):
        X, X2, ls, var = ctx.saved_tensors
        dX, dX2, dls, dvar = R.kbuild_bwd(ctx.kind, X, X2, ls, var, G,
                                          need_dX=ctx.needs_input_grad[1] or X2 is None,
                                          need_dX2=ctx.needs_input_grad[2])
        ddiag = None
        if ctx.has_diag and ctx.needs_input_grad[5]:
            ddiag = R.reduce(R.RED_SUM, R.get_diag(G).unsqueeze(1)).unsqueeze(1)
        return None, dX, dX2, dls, dvar, ddiag, None


def kernel_matrix(kind, X, X2, lengthscale, variance, diag_add=None, diag_const=0.0):
    """Stationary covariance matrix (stationary.py:74-107 + rbf.py:71-72 / matern.py:84-151), with the
    `+ eye * noise_var` / `+ eye * jitter` of the callers folded into the store."""
    S = _lead(X, X2, lengthscale, variance, diag_add)
    return _KernelMatrix.apply(kind, _expand(X, S), _expand(X2, S), _expand(lengthscale, S),
                               _expand(variance, S), _expand(diag_add, S), diag_const)


class _Potrf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        L, info, pack = R.potrf_packed_(A.clone())
        _note_info(info)
        ctx.save_for_backward(L, pack)
        ctx.mark_non_differentiable(info)
        return L, info

    @staticmethod
    def backward(ctx, Lbar, _):
        # Murray (2016): Abar = 1/2 L^-T (P + P^T) L^-1 with P = Phi(L^T Lbar), Phi = lower triangle with the
        # diagonal halved, so P + P^T is the symmetric matrix built from the lower triangle of L^T Lbar.
        L, pack = ctx.saved_tensors
        G = R.gemm(L, R.tril(Lbar), transA=True)
        Psym = R.copy_ltu(G)
        Psym = R.trsm_solve(L, pack, Psym, transpose=True)
        Y = R.trsm_solve(L, pack, R.transpose(Psym), transpose=True)
        return R.symmetrize(Y, 0.25)      # 1/2 * (Y + Y^T)/2: symmetric by construction, cleans rounding


def potrf(A, return_info=False):
    """linalg.potrf: lower Cholesky factor with the strict upper triangle zeroed."""
    L, info = _Potrf.apply(A)
    return (L, info) if return_info else L


class _Trsm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, B, transpose, alpha):
        pack = R.tri_pack(L)
        X = R.trsm_packed_(L, pack, B.clone(), transpose=transpose, alpha=alpha)
        ctx.transpose, ctx.alpha = transpose, alpha
        ctx.save_for_backward(L, X, pack)
        return X

    @staticmethod
    def backward(ctx, Xbar):
        L, X, pack = ctx.saved_tensors
        Bbar = R.trsm_packed_(L, pack, Xbar.clone(), transpose=not ctx.transpose, alpha=ctx.alpha)
        Lbar = None
        if ctx.needs_input_grad[0]:
            # X = alpha op(L)^-1 B:  Lbar = -(1/alpha) tril(Bbar X^T)  (no transpose),  -(1/alpha) tril(X Bbar^T)
            if not ctx.transpose:
                Lbar = R.tril(R.gemm(Bbar, X, transB=True, alpha=-1.0 / ctx.alpha))
            else:
                Lbar = R.tril(R.gemm(X, Bbar, transB=True, alpha=-1.0 / ctx.alpha))
            if L.shape[0] != Lbar.shape[0]:
                Lbar = Lbar.sum(dim=0, keepdim=True)
        return Lbar, Bbar, None, None


def trsm(L, B, transpose=False, alpha=1.0):
    """linalg.trsm (left, lower): alpha * op(L)^-1 B."""
    S = _lead(L, B)
    return _Trsm.apply(L, _expand(B, S), bool(transpose), float(alpha))


class _Gemm2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, ta, tb, alpha):
        ctx.ta, ctx.tb, ctx.alpha = ta, tb, alpha
        ctx.save_for_backward(A, B)
        return R.gemm(A, B, ta, tb, alpha=alpha)

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        ta, tb, al = ctx.ta, ctx.tb, ctx.alpha
        dA = dB = None
        if ctx.needs_input_grad[0]:
            dA = R.gemm(B, G, tb, True, alpha=al) if ta else R.gemm(G, B, False, not tb, alpha=al)
        if ctx.needs_input_grad[1]:
            dB = R.gemm(G, A, True, ta, alpha=al) if tb else R.gemm(A, G, not ta, False, alpha=al)
        return dA, dB, None, None, None


def gemm2(A, B, transpose_a=False, transpose_b=False, alpha=1.0):
    """linalg.gemm2: alpha * op(A) op(B)."""
    S = _lead(A, B)
    return _Gemm2.apply(_expand(A, S), _expand(B, S), bool(transpose_a), bool(transpose_b), float(alpha))


def syrk(A, transpose=False, alpha=1.0):
    """linalg.syrk: alpha * A A^T (A^T A when transpose)."""
    return gemm2(A, A, transpose, not transpose, alpha)


class _SumLogDiag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        ctx.save_for_backward(A)
        return R.sumlogdiag(A)

    @staticmethod
    def backward(ctx, g):
        (A,) = ctx.saved_tensors
        d = R.get_diag(A)
        out = torch.zeros_like(A)
        R.add_diag_(out, (g.unsqueeze(1) / d))
        return out


def sumlogdiag(A):
    """linalg.sumlogdiag."""
    return _SumLogDiag.apply(A)


class _MakeDiagonal(torch.autograd.Function):
    """util/customop.py:22-81 (`make_diagonal` CustomOp: forward embeds, backward extracts)."""

    @staticmethod
    def forward(ctx, v):
        out = torch.zeros(tuple(v.shape) + (v.shape[-1],), dtype=v.dtype, device=v.device)
        R.add_diag_(out, v)
        return out

    @staticmethod
    def backward(ctx, G):
        return R.get_diag(G)


def make_diagonal(v):
    return _MakeDiagonal.apply(v)


class _NormalLogPdfSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, m, v, scale):
        ctx.scale = scale
        ctx.save_for_backward(x, m, v)
        return R.normal_logpdf_sum(x, m, v, scale)

    @staticmethod
    def backward(ctx, g):
        x, m, v = ctx.saved_tensors
        gx, gm, gv = R.normal_logpdf_sum_bwd(x, m, v, g, ctx.scale, need=tuple(ctx.needs_input_grad[:3]))
        return gx, gm, gv, None


def normal_log_pdf_sum(x, mean, variance, scale=1.0):
    """scale * F.sum(F.mean(Normal.log_pdf, axis=0)) in one pass: normal.py:67-69 + factor_graph.py:223."""
    return _NormalLogPdfSum.apply(x, mean, variance, float(scale))


class _NormalLogPdfMulti(torch.autograd.Function):
    """All Normal factors of a graph walk in one launch each way; inputs x_0, m_0, v_0, x_1, ... (scales in ctx)."""

    @staticmethod
    def forward(ctx, scales, *xmv):
        entries = [(xmv[3 * t], xmv[3 * t + 1], xmv[3 * t + 2], scales[t]) for t in range(len(scales))]
        ctx.scales = scales
        ctx.save_for_backward(*xmv)
        return R.normal_logpdf_multi(entries)

    @staticmethod
    def backward(ctx, g):
        xmv, scales = ctx.saved_tensors, ctx.scales
        entries = [(xmv[3 * t], xmv[3 * t + 1], xmv[3 * t + 2], scales[t]) for t in range(len(scales))]
        needs = [tuple(ctx.needs_input_grad[1 + 3 * t + k] for k in range(3)) for t in range(len(scales))]
        grads = R.normal_logpdf_multi_bwd(entries, g, needs)
        out = [None]
        for row in grads:
            out += list(row)
        return tuple(out)


def normal_log_pdf_sum_multi(entries):
    """sum over the entries (x, mean, variance, scale) of scale * F.sum(F.mean(Normal.log_pdf, axis=0)): the Normal
    factors of one factor-graph walk in a single launch (and a single adjoint launch)."""
    flat = []
    for x, m, v, _ in entries:
        flat += [x, m, v]
    return _NormalLogPdfMulti.apply(tuple(float(e[3]) for e in entries), *flat)


class _NormalReparam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, v, S, eps, seed, offset, step_counter):
        w, e = R.normal_reparam(m, v, S, eps=eps, seed=seed, offset=offset, return_eps=True,
                                step_counter=step_counter)
        ctx.save_for_backward(e, v)
        ctx.mS, ctx.vS = m.shape[0], v.shape[0]
        return w

    @staticmethod
    def backward(ctx, gw):
        e, v = ctx.saved_tensors
        # gm = gw, gv = gw eps / (2 sqrt(v)), summed over the samples for shared operands: one launch
        gm, gv = R.normal_reparam_bwd(gw, e, v, ctx.mS, need=tuple(ctx.needs_input_grad[:2]))
        return gm, gv, None, None, None, None, None


def normal_draw(mean, variance, num_samples, eps=None, seed=0, offset=0, step_counter=None):
    """Reparameterised draw eps*sqrt(v)+mu (normal.py:89-92); eps injected or Philox-generated in-kernel.
    `step_counter` (device int32[1]) is mixed into the Philox counter at run time: a draw replayed from a CUDA graph
    sees fresh noise every optimiser step."""
    return _NormalReparam.apply(mean, variance, int(num_samples), eps, int(seed), int(offset), step_counter)


class _NormalReparamMulti(torch.autograd.Function):
    """Independent reparameterised draws of several Normal factors in one launch (and one adjoint launch).  Inputs
    m_0, v_0, m_1, v_1, ...; outputs w_0, w_1, ..."""

    @staticmethod
    def forward(ctx, meta, *mv):
        S_list, seed, offsets, step_counter = meta
        entries = [(mv[2 * t], mv[2 * t + 1], S_list[t]) for t in range(len(S_list))]
        outs = R.normal_reparam_multi(entries, seed, offsets, step_counter)
        ctx.mS = [mv[2 * t].shape[0] for t in range(len(S_list))]
        ctx.save_for_backward(*[o[1] for o in outs], *[mv[2 * t + 1] for t in range(len(S_list))])
        return tuple(o[0] for o in outs)

    @staticmethod
    def backward(ctx, *gws):
        T = len(ctx.mS)
        saved = ctx.saved_tensors
        eps, vs = saved[:T], saved[T:]
        gws = [torch.zeros_like(eps[t]) if gws[t] is None else gws[t] for t in range(T)]
        needs = [(ctx.needs_input_grad[1 + 2 * t], ctx.needs_input_grad[2 + 2 * t]) for t in range(T)]
        grads = R.normal_reparam_multi_bwd([(gws[t], eps[t], vs[t], ctx.mS[t]) for t in range(T)], needs)
        out = [None]
        for gm, gv in grads:
            out += [gm, gv]
        return tuple(out)


def normal_draw_multi(entries, seed, offsets, step_counter=None):
    """entries: [(mean, variance, num_samples)] -> [w_t]; in-kernel Philox streams (seed, offsets[t])."""
    mv = []
    for m, v, _ in entries:
        mv += [m, v]
    meta = (tuple(int(e[2]) for e in entries), int(seed), tuple(int(o) for o in offsets), step_counter)
    return list(_NormalReparamMulti.apply(meta, *mv))


# --------------------------------------------------------------------------------------------------
# fused SVGP bound
# --------------------------------------------------------------------------------------------------
class _SVGPLogPdf(torch.autograd.Function):
    """svgp_regression.py:61-109 (homoscedastic noise).  Forward follows the reference's whitened quantities
    (A = L^-1 Kuf, C = L^-1 Ls, mt = L^-1 mu); the two B x M products of :89-90 are replaced by
    Phi = A A^T and T = C C^T, since  sum((A^T C)^2) = tr(Phi T)  and  sum(A^2) = tr(Phi).
    Backward is the closed-form gradient -- no tape."""

    @staticmethod
    def forward(ctx, kind, jitter, scale, X, Y, Z, noise, mu, W, dvec, ls, kvar):
        S, B, P, M = X.shape[0], X.shape[1], Y.shape[2], Z.shape[1]
        dt, dev = X.dtype, X.device
        PP = (P + 3) & ~3                                       # keep every row stride a multiple of 4 elements
        # all right-hand sides of the solves with L ride in ONE buffer [Kuf | Ls | mu]: one GEMM, not three.  S = W W^T +
        # diag is formed and factored IN PLACE inside that buffer (the factorisation takes any row stride), so Ls never
        # has to be copied into it.  Padding columns stay uninitialised: a GEMM column only depends on its own column.
        RH = torch.empty((S, M, B + M + PP), dtype=dt, device=dev)
        Sm = RH[:, :, B:B + M]
        Kuu = torch.empty((S, M, M), dtype=dt, device=dev)
        pk, pks = R.new_pack(Kuu), R.new_pack(Sm)
        info = torch.empty((2, S), dtype=torch.int32, device=dev)
        Sinv_l = torch.empty((S, M, M), dtype=dt, device=dev)
        Sinv = torch.empty((S, M, M), dtype=dt, device=dev)
        side = _side_stream(dev)

        def s_branch():
            R.gemm(W, W, transB=True, beta=0.0, C=Sm, tri=1)    # :76 syrk (lower tiles) ...
            R.add_diag_(Sm, dvec)                               # ... + make_diagonal
            R.potrf_packed_(Sm, info[1], pks)                   # :84
        # S^-1 = Ls^-T Ls^-1 is only needed by the adjoint (d logdet S), but it depends on nothing else: it is
        # computed here, on the side stream, while the main stream factors Kuu and runs the solves with L
        if side is not None:                                    # the two factorisations are independent: overlap them
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            join_ls = torch.cuda.Event()
            with torch.cuda.stream(side):
                s_branch()
                sldLs = R.sumlogdiag(Sm)
                if R.pack_inverse(pks, Sm) is not None:
                    join_ls.record()                            # S^-1 only reads the pack: the solves need not wait for it
                    _sinv_chain(Sm, pks, Sinv_l, Sinv)
                else:
                    _sinv_chain(Sm, pks, Sinv_l, Sinv)          # reads Ls, which an in-place solve below overwrites
                    join_ls.record()
            side_k = _side_stream(dev, 1)
            side_k.wait_stream(cur)
            with torch.cuda.stream(side_k):                     # Kuf (:73) and mu: third stream, beside the factorisations
                R.kbuild_fwd(kind, Z, X, ls, kvar, out=RH[:, :, :B])
                R.copy2d_(RH[:, :, B + M:B + M + P], mu)
            R.kbuild_fwd(kind, Z, None, ls, kvar, diag_const=jitter, out=Kuu)    # :69-72
            R.potrf_packed_(Kuu, info[0], pk)                   # :83
            cur.wait_event(join_ls)
            cur.wait_stream(side_k)
        else:
            R.kbuild_fwd(kind, Z, None, ls, kvar, diag_const=jitter, out=Kuu)
            R.kbuild_fwd(kind, Z, X, ls, kvar, out=RH[:, :, :B])
            R.copy2d_(RH[:, :, B + M:B + M + P], mu)
            R.potrf_packed_(Kuu, info[0], pk)
            s_branch()
            _sinv_chain(Sm, pks, Sinv_l, Sinv)
        L, Ls = Kuu, Sm
        sldL = R.sumlogdiag(L)
        if side is None:
            sldLs = R.sumlogdiag(Ls)
        RH = R.trsm_solve(L, pk, RH)                            # :85-87  A = L^-1 Kuf, C = L^-1 Ls, mt = L^-1 mu
        A, C, mt = RH[:, :, :B], RH[:, :, B:B + M], RH[:, :, B + M:B + M + P]
        # T = C C^T only meets Phi in tr(Phi T): its chain runs on the side stream (behind S^-1) beside the Phi chain
        Tl = torch.empty((S, M, M), dtype=dt, device=dev)
        T = torch.empty((S, M, M), dtype=dt, device=dev)
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                R.copy_ltu(R.gemm(C, C, transB=True, beta=0.0, C=Tl, tri=1), out=T)
        else:
            R.copy_ltu(R.gemm(C, C, transB=True, beta=0.0, C=Tl, tri=1), out=T)
        # the data-fit residual (A^T mt and its reduction) is independent of Phi: third stream
        side2 = _side_stream(dev, 1)
        if side2 is not None:
            side2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side2):
                G1 = R.gemm(A, mt, transA=True)                 # :89  (S,B,P)
                sumr2 = R.reduce(R.RED_SUMSQDIFF, Y, G1)
                mm = R.reduce(R.RED_SUMSQ, mt)
        else:
            G1 = R.gemm(A, mt, transA=True)
            sumr2 = R.reduce(R.RED_SUMSQDIFF, Y, G1)
            mm = R.reduce(R.RED_SUMSQ, mt)
        G = _split_k(M, B) if S == 1 else 1
        if G > 1:
            # Phi = A A^T is M x M x B: 36-72 output tiles on 148 SMs with a 4096-deep K loop each -> split K into G slabs
            # (batched GEMM over strided views of the same buffer); the partial lower triangles are added and mirrored
            # by one launch
            Av = A.as_strided((G, M, B // G), (B // G, RH.stride(1), 1), A.storage_offset())
            parts = torch.empty((G, M, M), dtype=dt, device=dev)
            R.gemm(Av, Av, transB=True, beta=0.0, C=parts, tri=1)
            Phi = R.copy_ltu_sum(parts)
        else:
            Pl = torch.empty((S, M, M), dtype=dt, device=dev)
            Phi = R.copy_ltu(R.gemm(A, A, transB=True, beta=0.0, C=Pl, tri=1))
        trPhi = R.reduce(R.RED_SUMSQ, A)
        trT = R.reduce(R.RED_SUMSQ, C)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)                    # join: T (and S^-1, for the adjoint) are ready
            torch.cuda.current_stream().wait_stream(side2)                   # and the residual
        trPhiT = R.reduce(R.RED_DOT, Phi, T)
        # :94-108 on the reduced scalars in one launch (`KL_u` of the reference is minus the KL):
        #   Q = -sumr2/2 - P B kv/2 - P (tr(Phi T) - tr Phi)/2,  data = beta Q - B P (log 2pi + log nv)/2,
        #   logL = scale data + P (M/2 + sld(Ls) - sld(L)) - P tr(T)/2 - |mt|^2/2
        logL, beta, Q = R.svgp_bound_fwd(P, B, M, scale, sumr2, trPhi, trT, trPhiT, mm, sldL, sldLs, noise, kvar)
        ctx.kind, ctx.scale, ctx.dims = kind, scale, (S, B, P, M)
        ctx.save_for_backward(X, Y, Z, ls, kvar, W, L, Sinv, RH, Phi, T, G1, beta, Q, pk)
        ctx.info = info
        _note_info(info)
        return logL

    @staticmethod
    def backward(ctx, g):
        X, Y, Z, ls, kvar, W, L, Sinv, RH, Phi, T, G1, beta, Q, pk = ctx.saved_tensors
        S, B, P, M = ctx.dims
        dt, dev = RH.dtype, RH.device
        PP = (P + 3) & ~3
        A, mt_v = RH[:, :, :B], RH[:, :, B + M:B + M + P]
        mt = torch.empty((S, M, P), dtype=dt, device=dev)
        R.copy2d_(mt, mt_v)
        sc = ctx.scale
        need = ctx.needs_input_grad      # (kind, jitter, scale, X, Y, Z, noise, mu, W, dvec, ls, kvar)
        coef, gsb, neg_gsb, dnoise_s, dkvar_diag, neg_g, minus1 = R.svgp_coef_bwd(P, B, sc, g.contiguous(), beta, Q)
        side = _side_stream(dev)
        side2 = _side_stream(dev, 1)
        cur = torch.cuda.current_stream() if side is not None else None
        U = torch.empty((S, M, M), dtype=dt, device=dev)
        if side is not None:                                    # U = Phi T beside v = A (Y - A^T mt)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                R.gemm(Phi, T, C=U)
        else:
            R.gemm(Phi, T, C=U)
        v = R.gemm(A, Y)
        R.gemm(Phi, mt, alpha=-1.0, beta=1.0, C=v)              # v = A (Y - A^T mt)
        if side is not None:
            cur.wait_stream(side)
        # first solve with L^T: [E | E_S | E_R | mt]  (mt rides along: w = L^-T mt)
        E4 = torch.empty((S, M, 3 * M + PP), dtype=dt, device=dev)
        R.svgp_bwd_assemble(Phi, T, U, mt, v, coef, out=E4)
        R.copy2d_(E4[:, :, 3 * M:3 * M + P], mt)
        E4 = R.trsm_solve(L, pk, E4, transpose=True)
        # The Kuf branch (one big product + kernel adjoint) and the Kuu branch (second solve + kernel adjoint) are
        # independent from here on: fork the Kuf branch onto the side stream
        w = E4[:, :, 3 * M:3 * M + P]
        wg = R.axpby2d(gsb, w)
        dKuf = torch.empty((S, M, B), dtype=dt, device=dev)

        def kuf_branch():
            # Kuf adjoint: (L^-T E_R) A + g s beta (L^-T mt) Y^T, then through the kernel
            R.gemm(E4[:, :, 2 * M:3 * M], A, C=dKuf)
            R.gemm(wg, Y, transB=True, beta=1.0, C=dKuf)
            return R.kbuild_bwd(ctx.kind, Z, X, ls, kvar, dKuf, need_dX=True, need_dX2=need[3])
        if side is not None:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                dZ1, dX, dls1, dvar1 = kuf_branch()
        # second solve with L^T: [(L^-T E)^T | (L^-T E_S)^T | g (s beta v - mt)]  ->  [Kuu adjoint | L^-T E_S L^-1 | mu adjoint]
        F2 = torch.empty((S, M, 2 * M + PP), dtype=dt, device=dev)
        R.transpose(E4[:, :, :M], out=F2[:, :, :M])
        R.transpose(E4[:, :, M:2 * M], out=F2[:, :, M:2 * M])
        R.axpby2d(gsb, v, neg_g, mt, out=F2[:, :, 2 * M:2 * M + P])
        F2 = R.trsm_solve(L, pk, F2, transpose=True)
        dKuu = F2[:, :, :M]
        # S adjoint: g P/2 S^-1 - L^-T E_S L^-1 (S^-1 from the forward pass); W adjoint 2 Sbar W ; diag adjoint diag(Sbar).
        # Formed BEFORE the kernel adjoints: qU_cov_W's gradient is 93 % of the gradient bucket (M^2 of M^2 + M D + ...),
        # and a data-parallel run starts its all-reduce here, under the ~100 us of kernel-adjoint work that follows
        def s_adjoint():
            Sb = R.axpby2d(coef[:, 0], Sinv, minus1, F2[:, :, M:2 * M])
            dW_ = R.gemm(Sb, W, alpha=2.0)
            if _EARLY_REDUCE[0] is not None and need[8]:
                _EARLY_REDUCE[0](dW_)
            return dW_, R.get_diag(Sb)
        if side2 is not None:                                   # beside the Kuu kernel adjoint (third stream)
            side2.wait_stream(cur)
            with torch.cuda.stream(side2):
                dW, dd = s_adjoint()
        else:
            dW, dd = s_adjoint()
        dmu = torch.empty((S, M, P), dtype=dt, device=dev)
        R.copy2d_(dmu, F2[:, :, 2 * M:2 * M + P])
        dZ2, _, dls2, dvar2 = R.kbuild_bwd(ctx.kind, Z, None, ls, kvar, dKuu)
        if side is None:
            dZ1, dX, dls1, dvar1 = kuf_branch()
        dnoise = dnoise_s.unsqueeze(1)
        dY = R.axpby_dev(neg_gsb, Y, gsb, G1) if need[4] else None          # -g s beta (Y - A^T mt)
        if side is not None:
            cur.wait_stream(side)                                            # join the Kuf branch
            cur.wait_stream(side2)                                           # and the S adjoint
        dZ = R.axpby2d(None, dZ1, None, dZ2)
        dls = R.axpby2d(None, dls1.unsqueeze(0), None, dls2.unsqueeze(0)).squeeze(0)
        dkvar = R.axpby2d(None, dvar1.unsqueeze(0), None, dvar2.unsqueeze(0)).squeeze(0)
        dkvar = R.axpby2d(None, dkvar.unsqueeze(0), None, dkvar_diag.reshape(1, -1, 1)).squeeze(0)   # Kff_diag term (:100)
        return None, None, None, dX, dY, dZ, dnoise, dmu, dW, dd, dls, dkvar


def _sinv_chain(Ls, pks, Sinv_l, Sinv):
    """S^-1 = Ls^-T Ls^-1.  The factor's pack holds the explicit inverse W = Ls^-1 and its transpose (f32, M <= 1024: a
    by-product of the single-launch potrf): one lower-tiles product W^T W with the zero blocks of both triangular
    operands skipped, then the mirror.  Without the explicit inverse: a solve with the identity first.  Writes into
    preallocated buffers (created on the main stream, filled on the side stream)."""
    inv = R.pack_inverse(pks, Ls)
    if inv is not None:
        Wi, WiT = inv
        R.gemm(WiT, Wi, beta=0.0, C=Sinv_l, tri=5)              # lower tiles; A = W^T is upper triangular
    else:
        S, M = Ls.shape[0], Ls.shape[1]
        eye = torch.eye(M, dtype=Ls.dtype, device=Ls.device).unsqueeze(0).repeat(S, 1, 1)
        Li = R.trsm_solve(Ls, pks, eye)                         # Ls^-1
        R.gemm(R.transpose(Li), Li, beta=0.0, C=Sinv_l, tri=1)
    R.copy_ltu(Sinv_l, out=Sinv)


# Device-side record of the factorisations' `info` (first non-positive pivot, 1-based; 0 = fine): every fused bound adds
# max(info) of its potrf calls into one int32 per device, without a host synchronisation.  The training loops read it
# every `INFO_CHECK_EVERY` steps / at the end of a run and raise InferenceError -- the reference surfaces a non-PD matrix
# as an MXNetError at its next synchronisation (svgp_regression.py:70-72, gp_regression.py:58-60).
_INFO_ACC = {}


def info_accumulator(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    acc = _INFO_ACC.get(key)
    if acc is None:
        acc = torch.zeros((1,), dtype=torch.int32, device=device)
        _INFO_ACC[key] = acc
    return acc


def _note_info(info):
    if info.device.type != 'cuda' or not hasattr(R, 'note_info_'):
        return
    R.note_info_(info_accumulator(info.device), info)


def check_factorisations(device, what="a Cholesky factorisation"):
    """Raises InferenceError if any potrf on `device` met a non-positive pivot since the last check (one D2H read)."""
    from .common.exceptions import InferenceError
    if device.type != 'cuda':
        return
    acc = info_accumulator(device)
    bad = int(acc.item())
    if bad != 0:
        acc.zero_()
        raise InferenceError("%s failed: the matrix is not positive definite (first non-positive pivot %d). Increase "
                             "`jitter` on the module's log-pdf algorithm (svgp_regression.py:70-72, "
                             "gp_regression.py:58-60) or check the kernel parameters." % (what, bad))


# Data-parallel training (inference/_stepper.py) registers a callback here: it is handed the gradient of the largest
# parameter as soon as it exists and starts its all-reduce on a communication stream.
_EARLY_REDUCE = [None]


def set_early_reduce(fn):
    _EARLY_REDUCE[0] = fn


_SIDE_STREAMS = {}


def _side_stream(device, idx=0):
    """Extra streams per device for independent kernel chains (capturable fork/join)."""
    if device.type != 'cuda':
        return None
    key = (device.index if device.index is not None else torch.cuda.current_device(), idx)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[key] = s
    return s


def svgp_log_pdf(kind, X, Y, Z, noise_var, qU_mean, qU_cov_W, qU_cov_diag, lengthscale, variance,
                 jitter=0.0, log_pdf_scaling=1.0, mean=None, return_info=False):
    """SVGPRegressionLogPdf.compute (svgp_regression.py:43-109) as one autograd node -> logL (S,)."""
    if noise_var.dim() != 2 or noise_var.shape[-1] != 1:
        raise NotImplementedError("heteroscedastic noise_var of shape (N, P) (svgp_regression.py:61-67) is not "
                                  "implemented on the fused path")
    if mean is not None:
        Y = Y - mean                                            # :78-80
    S = _lead(X, Y, Z, noise_var, qU_mean, qU_cov_W, qU_cov_diag, lengthscale, variance)
    args = [_expand(t, S).contiguous() for t in
            (X, Y, Z, noise_var, qU_mean, qU_cov_W, qU_cov_diag, lengthscale, variance)]
    return _SVGPLogPdf.apply(kind, float(jitter), float(log_pdf_scaling), *args)


# --------------------------------------------------------------------------------------------------
# fused exact-GP marginal likelihood
# --------------------------------------------------------------------------------------------------
class _GPLogPdf(torch.autograd.Function):
    """gp_regression.py:55-70.  Returns (logL (S,), L, LinvY); L and LinvY are the cached posterior
    quantities of :72-75 and carry no gradient."""

    @staticmethod
    def forward(ctx, kind, jitter, X, Y, noise, ls, kvar):
        S, N, P = X.shape[0], X.shape[1], Y.shape[2]
        K = R.kbuild_fwd(kind, X, None, ls, kvar, diag_add=noise, diag_const=jitter)   # :55-60
        L, info, pk = R.potrf_packed_(K)                                                # :61
        _note_info(info)
        LinvY = R.trsm_packed_(L, pk, Y.clone())                                        # :66
        logdet_l = R.sumlogdiag(L)                                                      # :67 (diag(L) > 0)
        ss = R.reduce(R.RED_SUMSQ, LinvY)
        logL = -logdet_l * P - 0.5 * ss - (0.5 * N * P) * _LOG2PI                       # :68-70
        ctx.kind, ctx.dims = kind, (S, N, P)
        ctx.save_for_backward(X, ls, kvar, L, LinvY, pk)
        ctx.mark_non_differentiable(L, LinvY)
        ctx.info = info
        return logL, L, LinvY

    @staticmethod
    def backward(ctx, g, _gL, _gLY):
        X, ls, kvar, L, LinvY, pk = ctx.saved_tensors
        S, N, P = ctx.dims
        g = g.contiguous()
        # Kbar = g (1/2 a a^T - P/2 K^-1),  a = K^-1 Y = L^-T LinvY ;  Ybar = -g a
        a = R.trsm_packed_(L, pk, LinvY.clone(), transpose=True)
        eye = torch.eye(N, dtype=X.dtype, device=X.device).unsqueeze(0).expand(S, N, N).contiguous()
        Kinv = R.trsm_solve(L, pk, R.trsm_solve(L, pk, eye), transpose=True)
        aaT = R.gemm(a, a, transB=True)
        Kbar = R.axpby_dev(0.5 * g, aaT, (-0.5 * P) * g, Kinv)
        dX, _, dls, dvar = R.kbuild_bwd(ctx.kind, X, None, ls, kvar, Kbar)
        dnoise = R.reduce(R.RED_SUM, R.get_diag(Kbar).unsqueeze(1)).unsqueeze(1)
        dY = R.axpby_dev(-g, a) if ctx.needs_input_grad[3] else None
        return None, None, dX, dY, dnoise, dls, dvar


def gp_log_pdf(kind, X, Y, noise_var, lengthscale, variance, jitter=0.0, mean=None):
    """GPRegressionLogPdf.compute (gp_regression.py:42-76) -> (logL (S,), L, LinvY)."""
    if mean is not None:
        Y = Y - mean
    S = _lead(X, Y, noise_var, lengthscale, variance)
    args = [_expand(t, S).contiguous() for t in (X, Y, noise_var, lengthscale, variance)]
    return _GPLogPdf.apply(kind, float(jitter), *args)


# --------------------------------------------------------------------------------------------------
# dense-tanh network with sampled weights (BNN function evaluation)
# --------------------------------------------------------------------------------------------------
class _MlpTanh(torch.autograd.Function):
    """function_evaluation.py:72-96 for a Dense/tanh stack: all S weight samples in one forward launch and one adjoint
    launch (csrc/mlp.cu).  Inputs: x, then W_1, b_1, ..., W_L, b_L (b_l may be None)."""

    @staticmethod
    def forward(ctx, n_layers, x, *wb):
        Ws, bs = list(wb[0::2]), list(wb[1::2])
        ctx.n_layers = n_layers
        ctx.has_b = [b is not None for b in bs]
        ctx.save_for_backward(x, *Ws, *[b for b in bs if b is not None])
        return R.mlp_tanh_fwd(x, Ws, bs)

    @staticmethod
    def backward(ctx, gout):
        L = ctx.n_layers
        saved = ctx.saved_tensors
        x, Ws, rest = saved[0], list(saved[1:1 + L]), list(saved[1 + L:])
        bs = [rest.pop(0) if h else None for h in ctx.has_b]
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("mlp_tanh: gradient with respect to the network input is not implemented")
        dWs, dbs = R.mlp_tanh_bwd(x, Ws, bs, gout)
        out = [None, None]
        for dW, db in zip(dWs, dbs):
            out += [dW, db]
        return tuple(out)


def mlp_tanh(x, weights, biases):
    """x (S|1,B,in); weights[l] (S|1,out_l,in_l) (torch.nn.Linear layout); biases[l] (S|1,out_l) or None."""
    wb = []
    for W, b in zip(weights, biases):
        wb += [W, b]
    return _MlpTanh.apply(len(weights), x, *wb)


# --------------------------------------------------------------------------------------------------
# variational sparse GP (Titsias collapsed bound): streamed whitened statistics
# --------------------------------------------------------------------------------------------------
STATS_KEEP_BYTES = 16 << 30     # keep L^-1 K(Z, X) for the adjoint when it is at most this large
STATS_CHUNK_ROWS = 28672     # upper bound on the rows of X per streamed block: K(Z, X_c) is M x 28k (~115 MB at M = 1024)


def _stats_chunk(M, S):
    """Rows per streamed block: the block's GEMMs tile its columns 256 wide, 4 row tiles per 512-row solve step, so a
    column-tile count that is a multiple of 37 makes every GEMM of the block a whole number of waves on 148 SMs
    (28672 rows = 448 CTAs = 3.03 waves; 28392 rows = 444 CTAs = 3 waves); also divisible into the syrk's K slabs."""
    G = _syrk_splits(M) if S == 1 else 1
    tiles = max(37, (STATS_CHUNK_ROWS // 256) // 37 * 37)
    return (tiles * 256) // (4 * G) * (4 * G)


def _split_k(M, K):
    """K slabs for an M x M x K product with lower tiles only: the largest count <= _syrk_splits(M) that divides K into
    slabs whose length is a multiple of 4 elements (16-byte TMA rows)."""
    for G in range(_syrk_splits(M), 1, -1):
        if K % (4 * G) == 0:
            return G
    return 1


def _syrk_splits(M):
    """K-axis slabs for the M x M syrk of a streamed block so that (lower 128 x 256 tiles) x slabs fills the 148 SMs in
    one wave."""
    tiles = sum(i // 2 + 1 for i in range((M + 127) // 128))
    return max(1, min(8, 148 // tiles))


class _WhitenedStats(torch.autograd.Function):
    """Phi = sum_c A_c A_c^T (S,M,M) and b = sum_c A_c Y_c (S,M,P) with A_c = L^-1 K(Z, X_c), streamed over blocks
    of rows of X so that neither K(Z,X) nor L^-1 K(Z,X) (M x N, 4.1 GB each at N=1e6, M=1024) is ever resident:
    these are the only places where sparsegp_regression.py:77-100 touches the N axis besides sum(Y^2).
    The adjoint re-streams X (recomputing K(Z, X_c) is cheaper than keeping it); the factor's adjoint has the closed
    form  Lbar = -tril(L^-T (G Phi + bbar b^T)),  G = Phibar + Phibar^T."""

    @staticmethod
    def forward(ctx, kind, chunk, X, Y, Z, ls, var, L):
        S, N, P = X.shape[0], X.shape[1], Y.shape[2]
        M = Z.shape[1]
        pack = R.tri_pack(L)
        Phi = torch.zeros((S, M, M), dtype=X.dtype, device=X.device)
        b = torch.zeros((S, M, P), dtype=X.dtype, device=X.device)
        G = _syrk_splits(M) if S == 1 else 1
        parts = torch.zeros((G, M, M), dtype=X.dtype, device=X.device) if G > 1 else None
        # 180 GB of HBM: when the whitened matrix L^-1 K(Z, X) fits a budget (4.1 GB at N=1e6, M=1024) its blocks are kept
        # for the adjoint, which then skips the K-build + solve recomputation of every block
        keep = any(ctx.needs_input_grad) and S * M * N * X.element_size() <= STATS_KEEP_BYTES
        kept = [] if keep else None
        for c0 in range(0, N, chunk):
            Xc, Yc = X[:, c0:c0 + chunk], Y[:, c0:c0 + chunk]
            A = R.trsm_solve(L, pack, R.kbuild_fwd(kind, Z, Xc, ls, var))       # :77, :81 on this block
            if kept is not None:
                kept.append(A)
            Bc = A.shape[2]
            if G > 1 and Bc % (4 * G) == 0 and A.is_contiguous():
                # :84 syrk(LinvKuf): M x M x Bc has too few output tiles for 148 SMs -> split the K axis into G slabs
                # (a batched GEMM over strided views of the same buffer), then add the G partial lower triangles
                Av = A.as_strided((G, M, Bc // G), (Bc // G, Bc, 1))
                R.gemm(Av, Av, transB=True, beta=0.0, C=parts, tri=True)
                Phi += parts.sum(dim=0, keepdim=True)
            else:
                R.gemm(A, A, transB=True, beta=1.0, C=Phi, tri=True)            # lower tiles only
            R.gemm(A, Yc.contiguous(), beta=1.0, C=b)                           # :90 gemm2(LinvKuf, Y)
        Phi = R.copy_ltu(Phi)
        ctx.kind, ctx.chunk = kind, chunk
        ctx.kept = kept
        ctx.save_for_backward(X, Y, Z, ls, var, L, pack, Phi, b)
        return Phi, b

    @staticmethod
    def backward(ctx, gPhi, gb):
        X, Y, Z, ls, var, L, pack, Phi, b = ctx.saved_tensors
        need = ctx.needs_input_grad      # (kind, chunk, X, Y, Z, ls, var, L)
        N, chunk = X.shape[1], ctx.chunk
        G = R.symmetrize(gPhi.contiguous(), 1.0)                                # Phibar + Phibar^T
        gb = gb.contiguous()
        dX = torch.empty_like(X) if need[2] else None
        dY = torch.empty_like(Y) if need[3] else None
        dZ = dls = dvar = None
        # the solve with L^T is applied once to the M x M / M x P factors instead of to every M x B_c block
        Gp = R.trsm_solve(L, pack, G.clone(), transpose=True)
        gbp = R.trsm_solve(L, pack, gb.clone(), transpose=True)
        for c0 in range(0, N, chunk):
            Xc, Yc = X[:, c0:c0 + chunk], Y[:, c0:c0 + chunk].contiguous()
            if ctx.kept is not None:
                A = ctx.kept[c0 // chunk]
            else:
                A = R.trsm_solve(L, pack, R.kbuild_fwd(kind=ctx.kind, X=Z, X2=Xc, ls=ls, var=var))
            Kbar = R.gemm(Gp, A)                                                # L^-T (G A + bbar Y^T), re-associated:
            R.gemm(gbp, Yc, transB=True, beta=1.0, C=Kbar)                      # (L^-T G) A + (L^-T bbar) Y^T
            dZc, dXc, dlsc, dvarc = R.kbuild_bwd(ctx.kind, Z, Xc, ls, var, Kbar, need_dX=True, need_dX2=need[2])
            dZ = dZc if dZ is None else dZ + dZc
            dls = dlsc if dls is None else dls + dlsc
            dvar = dvarc if dvar is None else dvar + dvarc
            if need[2]:
                dX[:, c0:c0 + chunk] = dXc
            if need[3]:
                dY[:, c0:c0 + chunk] = R.gemm(A, gb, transA=True)
        dL = None
        if need[7]:
            Hm = R.gemm(G, Phi)
            R.gemm(gb, b, transB=True, beta=1.0, C=Hm)
            dL = R.tril(R.trsm_solve(L, pack, Hm, transpose=True))
            dL = -dL
        return None, None, dX, dY, dZ, dls, dvar, dL


def whitened_stats(kind, X, Y, Z, lengthscale, variance, L, chunk=None):
    S = _lead(X, Y, Z, lengthscale, variance, L)
    args = [_expand(t, S).contiguous() for t in (X, Y, Z, lengthscale, variance, L)]
    return _WhitenedStats.apply(kind, int(chunk or _stats_chunk(Z.shape[-2], S)), *args)


def sparsegp_log_pdf(kind, X, Y, Z, noise_var, lengthscale, variance, jitter=0.0, mean=None, chunk=None):
    """SparseGPRegressionLogPdf.compute (sparsegp_regression.py:42-108) -> (logL (S,), wv, L, LA); wv, L, LA are the
    posterior quantities cached for prediction (:101-106) and carry no gradient.  The N axis is consumed by the
    streamed statistics above; everything after them is M x M (two potrf, one solve, reductions).  As in the
    reference, `log_pdf_scaling` does not enter this bound."""
    if noise_var.dim() != 2 or noise_var.shape[-1] != 1:
        raise NotImplementedError("noise_var must have shape (S, 1)")
    if mean is not None:
        Y = Y - mean                                                            # :87-89
    N, D, M = X.shape[-2], Y.shape[-1], Z.shape[-2]
    Kuu = kernel_matrix(kind, Z, None, lengthscale, variance, diag_const=jitter)    # :72-75
    L, info = potrf(Kuu, return_info=True)                                      # :80
    Phi, b = whitened_stats(kind, X, Y, Z, lengthscale, variance, L, chunk=chunk)
    nv = noise_var.unsqueeze(-1)                                                # (S,1,1)
    eye = torch.eye(M, dtype=Phi.dtype, device=Phi.device).unsqueeze(0)
    LA, infoA = potrf(eye + Phi / nv, return_info=True)                         # :83-85
    c = trsm(LA, b)                                                             # :90
    n1 = noise_var[:, 0]
    logL = -D * sumlogdiag(LA)                                                  # :92
    logL = logL - (torch.sum(torch.square(Y), dim=(-1, -2)) / n1 + (N * D) * (_LOG2PI + torch.log(n1))) / 2   # :93-94
    logL = logL + torch.sum(torch.square(c), dim=(-1, -2)) / (2 * torch.square(n1))                           # :95-97
    logL = logL - (D * N) * variance[:, 0] / (2 * n1)                           # :98  (Kdiag = variance)
    logL = logL + D * torch.sum(torch.diagonal(Phi, dim1=-2, dim2=-1), dim=-1) / (2. * n1)    # :99-100
    with torch.no_grad():                                                       # :101-106
        wv = trsm(L, trsm(LA, c, transpose=True), transpose=True) / nv
    return logL, wv, L.detach(), LA.detach(), (info, infoA)
