"""Thin tensor-level wrappers over the C ABI (include/mxf_b200.h).

Every function takes CUDA tensors that carry the reference's leading sample
axis S (components/variables/runtime_variable.py:20-50), enqueues kernels on
the current torch stream and returns torch tensors.  No autograd here (see
ops.py) and no fallback: a CPU tensor or a missing library raises.
"""
import torch

from . import _lib
from ._lib import lib, ptr, stream_ptr, check, dtype_code, require_cuda, launch_count  # noqa: F401

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
RED_SUM, RED_SUMSQ, RED_DOT, RED_SUMLOG, RED_SUMSQDIFF = 0, 1, 2, 3, 4


def _bstride(t, S):
    """Batch stride in elements; 0 broadcasts a single sample over S."""
    return 0 if t.shape[0] == 1 and S > 1 else (t.stride(0) if t.shape[0] > 1 else 0)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def kbuild_fwd(kind, X, X2, ls, var, diag_add=None, diag_const=0.0, out=None):
    """K (S,N,N2) for X (Sx,N,D), X2 (Sx2,N2,D) or None, ls (Sl,1|D), var (Sv,1)."""
    require_cuda(X, X2, ls, var, diag_add)
    X, ls, var = _c(X), _c(ls), _c(var)
    X2 = None if X2 is None else _c(X2)
    diag_add = None if diag_add is None else _c(diag_add)
    S = max(X.shape[0], ls.shape[0], var.shape[0], 1 if X2 is None else X2.shape[0],
            1 if diag_add is None else diag_add.shape[0])
    N, D = X.shape[1], X.shape[2]
    N2 = N if X2 is None else X2.shape[1]
    if out is None:
        out = torch.empty((S, N, N2), dtype=X.dtype, device=X.device)
    check(lib().mxf_kbuild_fwd(kind, dtype_code(X), ptr(X), ptr(X2), ptr(ls), ls.shape[-1], ptr(var),
                               ptr(diag_add), float(diag_const), ptr(out), out.stride(1),
                               S, N, N2, D, _bstride(X, S), 0 if X2 is None else _bstride(X2, S),
                               _bstride(ls, S), _bstride(var, S),
                               0 if diag_add is None else _bstride(diag_add, S), out.stride(0),
                               stream_ptr()), 'mxf_kbuild_fwd')
    return out


def kbuild_tc_threshold(min_elems=-1):
    """Output size (elements) from which f32 cross-covariances with D <= 16 take the tcgen05 + TMA-store kernel; returns
    the previous value (negative argument: query only)."""
    return int(lib().mxf_kbuild_tc_threshold(int(min_elems)))


def kbuild_bwd(kind, X, X2, ls, var, G, need_dX=True, need_dX2=True):
    """Adjoint of kbuild_fwd.  Returns (dX, dX2, dls, dvar), each with S leading."""
    require_cuda(X, X2, ls, var, G)
    X, ls, var = _c(X), _c(ls), _c(var)
    X2 = None if X2 is None else _c(X2)
    if G.stride(2) != 1:
        G = G.contiguous()
    S, N, N2 = G.shape
    D = X.shape[2]
    # operands broadcast over S are expanded: the adjoint then sums over S on the host side
    def ex(t):
        return t if t.shape[0] == S else t.expand((S,) + t.shape[1:]).contiguous()
    Xe, lse, vare = ex(X), ex(ls), ex(var)
    X2e = None if X2 is None else ex(X2)
    dX = torch.empty_like(Xe) if need_dX else None
    dX2 = torch.empty_like(X2e) if (need_dX2 and X2e is not None) else None
    dls = torch.empty_like(lse)
    dvar = torch.empty_like(vare)
    nbytes = lib().mxf_kbuild_bwd_workspace_bytes(dtype_code(X), S, N, N2, D)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=X.device)
    check(lib().mxf_kbuild_bwd(kind, dtype_code(X), ptr(Xe), ptr(X2e), ptr(lse), lse.shape[-1], ptr(vare),
                               ptr(G), G.stride(1), ptr(dX), ptr(dX2), ptr(dls), ptr(dvar),
                               S, N, N2, D, Xe.stride(0), 0 if X2e is None else X2e.stride(0),
                               lse.stride(0), vare.stride(0), G.stride(0), ptr(ws), nbytes, stream_ptr()),
          'mxf_kbuild_bwd')

    def red(g, t):
        if g is None:
            return None
        return g if t.shape[0] == S else g.sum(dim=0, keepdim=True)
    return red(dX, X), (None if X2 is None else red(dX2, X2)), red(dls, ls), red(dvar, var)


def gemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None, tri=False):
    """C = alpha op(A) op(B) + beta C, batched over the leading axis (stride-0 broadcast allowed)."""
    require_cuda(A, B, C)
    if A.stride(2) != 1:
        A = A.contiguous()
    if B.stride(2) != 1:
        B = B.contiguous()
    S = max(A.shape[0], B.shape[0], 1 if C is None else C.shape[0])
    m = A.shape[2] if transA else A.shape[1]
    k = A.shape[1] if transA else A.shape[2]
    n = B.shape[1] if transB else B.shape[2]
    kb = B.shape[2] if transB else B.shape[1]
    if k != kb:
        raise _lib.MXFusionB200Error("gemm: inner dimensions differ (%d vs %d)" % (k, kb))
    if C is None:
        C = torch.empty((S, m, n), dtype=A.dtype, device=A.device)
        if tri:
            C.zero_()
        beta = 0.0
    check(lib().mxf_gemm(dtype_code(A), int(transA), int(transB), m, n, k, float(alpha),
                         ptr(A), A.stride(1), _bstride(A, S), ptr(B), B.stride(1), _bstride(B, S),
                         float(beta), ptr(C), C.stride(1), C.stride(0) if C.shape[0] > 1 else 0,
                         S, int(tri), stream_ptr()), 'mxf_gemm')
    return C


def potrf_(A, info=None):
    """In-place lower Cholesky of A (S,n,n).  Returns (A, info int32[S])."""
    require_cuda(A)
    S, n, _ = A.shape
    if info is None:
        info = torch.empty((S,), dtype=torch.int32, device=A.device)
    check(lib().mxf_potrf(dtype_code(A), ptr(A), A.stride(1), A.stride(0), S, n, ptr(info), stream_ptr()),
          'mxf_potrf')
    return A, info


def trsm_(L, B, transpose=False, alpha=1.0):
    """In-place B := alpha op(L)^-1 B; L (Sl,n,n) lower, B (S,n,nrhs)."""
    require_cuda(L, B)
    S, n, nrhs = B.shape
    check(lib().mxf_trsm(dtype_code(B), int(transpose), n, nrhs, float(alpha), ptr(L), L.stride(1),
                         _bstride(L, S), ptr(B), B.stride(1), B.stride(0), S, stream_ptr()), 'mxf_trsm')
    return B


def _new_pack(A):
    S, n, _ = A.shape
    elems = lib().mxf_tri_pack_elems(dtype_code(A), n)
    return torch.empty((S, max(int(elems), 1)), dtype=A.dtype, device=A.device)


def new_pack(A):
    return _new_pack(A)


def potrf_packed_(A, info=None, pack=None):
    """In-place lower Cholesky of A (S,n,n) on the GEMM-based path; returns (A, info, pack) where `pack` holds the
    inverted diagonal blocks and L^T that turn later solves with this factor into GEMMs (mxf_trsm_packed)."""
    require_cuda(A)
    S, n, _ = A.shape
    if info is None:
        info = torch.empty((S,), dtype=torch.int32, device=A.device)
    if pack is None:
        pack = _new_pack(A)
    check(lib().mxf_potrf_packed(dtype_code(A), ptr(A), A.stride(1), A.stride(0), S, n, ptr(info), ptr(pack),
                                 stream_ptr()), 'mxf_potrf_packed')
    return A, info, pack


def tri_pack(L):
    """Pack of an existing lower factor L (S,n,n)."""
    require_cuda(L)
    if L.stride(2) != 1:
        L = L.contiguous()
    S, n, _ = L.shape
    pack = _new_pack(L)
    check(lib().mxf_tri_pack(dtype_code(L), ptr(L), L.stride(1), L.stride(0), S, n, ptr(pack), stream_ptr()),
          'mxf_tri_pack')
    return pack


def trsm_packed_(L, pack, B, transpose=False, alpha=1.0):
    """In-place B := alpha op(L)^-1 B as a chain of GEMMs; L (Sl,n,n), pack (Sl,*), B (S,n,nrhs)."""
    require_cuda(L, pack, B)
    S, n, nrhs = B.shape
    check(lib().mxf_trsm_packed(dtype_code(B), int(transpose), n, nrhs, float(alpha), ptr(L), L.stride(1),
                                _bstride(L, S), ptr(pack), pack.stride(0) if pack.shape[0] > 1 else 0,
                                ptr(B), B.stride(1), B.stride(0), S, stream_ptr()), 'mxf_trsm_packed')
    return B


def trsm_solve(L, pack, B, transpose=False):
    """X = op(L)^-1 B with the factor's pack.  Uses the large-block out-of-place chain when the pack has one (B is then
    scratch and a NEW tensor is returned), else the in-place NB-block chain (returns B)."""
    require_cuda(L, pack, B)
    S, n, nrhs = B.shape
    code = dtype_code(B)
    if nrhs > 8 and lib().mxf_tri_top_block(code, n) > lib().mxf_tri_block(code):
        X = torch.empty_like(B)
        check(lib().mxf_trsm_packed_oop(code, int(transpose), n, nrhs, ptr(L), L.stride(1), _bstride(L, S), ptr(pack),
                                        pack.stride(0) if pack.shape[0] > 1 else 0, ptr(B), B.stride(1), B.stride(0),
                                        ptr(X), X.stride(1), X.stride(0), S, stream_ptr()), 'mxf_trsm_packed_oop')
        return X
    return trsm_packed_(L, pack, B, transpose=transpose)


def _sq(fn, name, A, *extra, out=None):
    require_cuda(A, out)
    S, n, _ = A.shape
    if out is None:
        out = torch.empty((S, n, n), dtype=A.dtype, device=A.device)
    check(fn(dtype_code(A), *extra, ptr(A), A.stride(1), A.stride(0), ptr(out), out.stride(1), out.stride(0),
             S, n, stream_ptr()), name)
    return out


def copy_ltu(P, out=None):
    return _sq(lib().mxf_copy_ltu, 'mxf_copy_ltu', P, out=out)


def copy_ltu_sum(parts, out=None):
    """parts (G, n, n): lower triangles of G partial products -> (1, n, n) symmetric sum (one launch)."""
    require_cuda(parts, out)
    G, n, _ = parts.shape
    if out is None:
        out = torch.empty((1, n, n), dtype=parts.dtype, device=parts.device)
    check(lib().mxf_copy_ltu_sum(dtype_code(parts), ptr(parts), parts.stride(1), parts.stride(0), G, ptr(out), out.stride(1),
                                 n, stream_ptr()), 'mxf_copy_ltu_sum')
    return out


def copy2d_(dst, src=None):
    """dst[s, r, c] = src[s, r, c] for strided (S, rows, cols) views with unit innermost stride; src None: zero fill."""
    require_cuda(dst, src)
    S, rows, cols = dst.shape
    if dst.stride(2) != 1 or (src is not None and (src.stride(2) != 1 or tuple(src.shape) != tuple(dst.shape))):
        raise _lib.MXFusionB200Error("copy2d_: operands must be (S, rows, cols) views with unit innermost stride")
    check(lib().mxf_copy2d(dtype_code(dst), ptr(src), 0 if src is None else src.stride(1),
                           0 if src is None else src.stride(0), ptr(dst), dst.stride(1), dst.stride(0), S, rows, cols,
                           stream_ptr()), 'mxf_copy2d')
    return dst


_PACK_LAYOUTS = {}


def pack_layout(dtype_c, n):
    """{Dinv, DinvT, LT, W, WT} element offsets + top, ldt, nq of a factor's pack (mxf_tri_pack_layout)."""
    key = (dtype_c, n)
    lay = _PACK_LAYOUTS.get(key)
    if lay is None:
        import ctypes
        buf = (ctypes.c_int64 * 8)()
        check(lib().mxf_tri_pack_layout(dtype_c, n, ctypes.cast(buf, ctypes.c_void_p)), 'mxf_tri_pack_layout')
        lay = dict(zip(('dinv', 'dinvT', 'lt', 'w', 'wT', 'top', 'ldt', 'nq'), [int(v) for v in buf]))
        _PACK_LAYOUTS[key] = lay
    return lay


def pack_inverse(pack, like):
    """(W, WT): the factor's explicit inverse L^-1 and its transpose as (S, n, n) views into `pack` (row stride top), or
    None when the pack holds no single full-size inverse block (f64 factors, n > 1024)."""
    S, n, _ = like.shape
    lay = pack_layout(dtype_code(like), n)
    if lay['nq'] != 1 or lay['top'] < n:
        return None
    top = lay['top']
    W = pack.as_strided((pack.shape[0], n, n), (pack.stride(0), top, 1), pack.storage_offset() + lay['w'])
    WT = pack.as_strided((pack.shape[0], n, n), (pack.stride(0), top, 1), pack.storage_offset() + lay['wT'])
    return W, WT


def note_info_(acc, info):
    """acc[0] = max(acc[0], max |info|) on the device (no synchronisation)."""
    require_cuda(acc, info)
    info = _c(info)
    check(lib().mxf_info_max(ptr(acc), ptr(info), info.numel(), stream_ptr()), 'mxf_info_max')
    return acc


def symmetrize(A, alpha=1.0):
    return _sq(lib().mxf_symmetrize, 'mxf_symmetrize', A, float(alpha))


def tril(A, strict=False):
    return _sq(lib().mxf_tril, 'mxf_tril', A, 1 if strict else 0)


def transpose(A, out=None):
    require_cuda(A, out)
    if A.stride(2) != 1:
        A = A.contiguous()
    S, m, n = A.shape
    if out is None:
        out = torch.empty((S, n, m), dtype=A.dtype, device=A.device)
    check(lib().mxf_transpose(dtype_code(A), ptr(A), A.stride(1), A.stride(0), ptr(out), out.stride(1),
                              out.stride(0), S, m, n, stream_ptr()), 'mxf_transpose')
    return out


def reduce(op, a, b=None, scale=1.0):
    """out[s] = scale * sum over the trailing two axes of f(a, b); a (S,rows,cols)."""
    require_cuda(a, b)
    if a.stride(2) != 1:
        a = a.contiguous()
    if b is not None and b.stride(2) != 1:
        b = b.contiguous()
    S = a.shape[0] if b is None else max(a.shape[0], b.shape[0])
    rows, cols = a.shape[1], a.shape[2]
    out = torch.empty((S,), dtype=a.dtype, device=a.device)
    check(lib().mxf_reduce(op, dtype_code(a), ptr(a), a.stride(1), _bstride(a, S),
                           ptr(b), 0 if b is None else b.stride(1), 0 if b is None else _bstride(b, S),
                           S, rows, cols, float(scale), ptr(out), stream_ptr()), 'mxf_reduce')
    return out


def sumlogdiag(A):
    require_cuda(A)
    S, n, _ = A.shape
    out = torch.empty((S,), dtype=A.dtype, device=A.device)
    check(lib().mxf_sumlogdiag(dtype_code(A), ptr(A), A.stride(1), A.stride(0), S, n, ptr(out), stream_ptr()),
          'mxf_sumlogdiag')
    return out


def add_diag_(A, d=None, c=0.0):
    require_cuda(A, d)
    S, n, _ = A.shape
    d = None if d is None else _c(d)
    check(lib().mxf_add_diag(dtype_code(A), ptr(A), A.stride(1), A.stride(0), ptr(d),
                             0 if d is None else _bstride(d, S), float(c), S, n, stream_ptr()), 'mxf_add_diag')
    return A


def get_diag(A):
    require_cuda(A)
    S, n, _ = A.shape
    out = torch.empty((S, n), dtype=A.dtype, device=A.device)
    check(lib().mxf_get_diag(dtype_code(A), ptr(A), A.stride(1), A.stride(0), ptr(out), out.stride(0), S, n,
                             stream_ptr()), 'mxf_get_diag')
    return out


def _flat(t):
    """(S', ...) -> contiguous (S', n) view."""
    t = _c(t)
    return t.reshape(t.shape[0], -1)


def normal_logpdf_sum(x, m, v, scale=1.0):
    """scale * sum(mean_S(log N(x | m, v))) -> tensor of shape (1,)."""
    require_cuda(x, m, v)
    x, m, v = _flat(x), _flat(m), _flat(v)
    S = max(x.shape[0], m.shape[0], v.shape[0])
    n = x.shape[1]
    out = torch.zeros((1,), dtype=x.dtype, device=x.device)
    check(lib().mxf_normal_logpdf_sum(dtype_code(x), ptr(x), _bstride(x, S), ptr(m), _bstride(m, S),
                                      ptr(v), _bstride(v, S), S, n, float(scale), ptr(out), stream_ptr()),
          'mxf_normal_logpdf_sum')
    return out


def normal_logpdf_sum_bwd(x, m, v, gout, scale=1.0, need=(True, True, True)):
    require_cuda(x, m, v, gout)
    xs, ms, vs = x.shape, m.shape, v.shape
    x, m, v = _flat(x), _flat(m), _flat(v)
    S = max(x.shape[0], m.shape[0], v.shape[0])
    n = x.shape[1]
    gx = torch.empty_like(x) if need[0] else None
    gm = torch.empty_like(m) if need[1] else None
    gv = torch.empty_like(v) if need[2] else None
    gout = _c(gout).reshape(-1)
    check(lib().mxf_normal_logpdf_sum_bwd(dtype_code(x), ptr(x), _bstride(x, S), ptr(m), _bstride(m, S),
                                          ptr(v), _bstride(v, S), S, n, float(scale), ptr(gout),
                                          ptr(gx), ptr(gm), ptr(gv), stream_ptr()), 'mxf_normal_logpdf_sum_bwd')
    return (None if gx is None else gx.reshape(xs), None if gm is None else gm.reshape(ms),
            None if gv is None else gv.reshape(vs))


def normal_reparam(m, v, S, eps=None, seed=0, offset=0, return_eps=False, step_counter=None):
    """w (S, *shape) = eps*sqrt(v)+m; m, v carry a leading axis of 1 or S."""
    require_cuda(m, v, eps)
    shape = m.shape[1:]
    mf, vf = _flat(m), _flat(v)
    n = mf.shape[1]
    w = torch.empty((S, n), dtype=m.dtype, device=m.device)
    eps_out = None
    if eps is not None:
        eps = _c(eps).reshape(S, n)
    elif return_eps:
        eps_out = torch.empty_like(w)
    check(lib().mxf_normal_reparam(dtype_code(m), ptr(eps), ptr(mf), _bstride(mf, S), ptr(vf), _bstride(vf, S),
                                   S, n, int(seed), int(offset), ptr(step_counter), ptr(w), ptr(eps_out), stream_ptr()),
          'mxf_normal_reparam')
    w = w.reshape((S,) + tuple(shape))
    if return_eps:
        e = eps if eps is not None else eps_out
        return w, e.reshape((S,) + tuple(shape))
    return w


def adam_step_(w, g, m, v, step_count, lr, beta1=0.9, beta2=0.999, eps=1e-8, rescale=1.0):
    require_cuda(w, g, m, v, step_count)
    check(lib().mxf_adam_step(dtype_code(w), ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), float(lr), float(beta1),
                              float(beta2), float(eps), float(rescale), ptr(step_count), stream_ptr()),
          'mxf_adam_step')


def sgd_step_(w, g, mom, step_count, lr, momentum=0.0, rescale=1.0):
    """mx.optimizer.SGD on the flat bucket (mom: state bucket, only read when momentum != 0)."""
    require_cuda(w, g, mom, step_count)
    check(lib().mxf_sgd_step(dtype_code(w), ptr(w), ptr(g), ptr(mom), w.numel(), float(lr), float(momentum), float(rescale),
                             ptr(step_count), stream_ptr()), 'mxf_sgd_step')


def gather_rows(src, idx, off, rows, out=None):
    require_cuda(src, idx, off)
    src = _c(src)
    cols = src.shape[1]
    if out is None:
        out = torch.empty((rows, cols), dtype=src.dtype, device=src.device)
    check(lib().mxf_gather_rows(dtype_code(src), ptr(src), cols, ptr(idx), ptr(off), rows, ptr(out), stream_ptr()),
          'mxf_gather_rows')
    return out


def axpby_dev(a, X, b=None, Y=None, out=None):
    """out[s] = a[s]*X[s] + b[s]*Y[s]; a, b device tensors (S,) or None; X, Y (S|1, ...)."""
    require_cuda(a, X, b, Y)
    X = _c(X)
    Y = None if Y is None else _c(Y)
    S = max(X.shape[0], 1 if Y is None else Y.shape[0], 1 if a is None else a.numel(),
            1 if b is None else b.numel())
    n = X[0].numel()
    if out is None:
        out = torch.empty((S,) + tuple(X.shape[1:]), dtype=X.dtype, device=X.device)

    def coef(c):
        if c is None:
            return None
        c = _c(c).reshape(-1)
        return c if c.numel() == S else c.expand(S).contiguous()
    a, b = coef(a), coef(b)
    check(lib().mxf_axpby_dev(dtype_code(X), ptr(a), ptr(X), _bstride(X, S) if X.shape[0] > 1 else 0,
                              ptr(b), ptr(Y), 0 if Y is None else (n if Y.shape[0] > 1 else 0),
                              ptr(out), n, S, n, stream_ptr()), 'mxf_axpby_dev')
    return out


def axpby2d(a, X, b=None, Y=None, out=None):
    """out = a[s] X + b[s] Y on (S, rows, cols) views with unit innermost stride (any row / batch strides); a, b: device
    tensors (S,) or None (= 1)."""
    require_cuda(a, X, b, Y, out)
    S, rows, cols = X.shape
    if out is None:
        out = torch.empty((S, rows, cols), dtype=X.dtype, device=X.device)
    for t in (X, Y, out):
        if t is not None and (t.stride(2) != 1 or tuple(t.shape) != (S, rows, cols)):
            raise _lib.MXFusionB200Error("axpby2d: operands must be (S, rows, cols) views with unit innermost stride")

    def coef(c):
        if c is None:
            return None
        c = _c(c.reshape(-1))
        if c.numel() != S:
            raise _lib.MXFusionB200Error("axpby2d: one coefficient per sample expected")
        return c
    a, b = coef(a), coef(b)
    check(lib().mxf_axpby2d(dtype_code(X), ptr(a), ptr(X), X.stride(1), X.stride(0), ptr(b), ptr(Y),
                            0 if Y is None else Y.stride(1), 0 if Y is None else Y.stride(0), ptr(out), out.stride(1),
                            out.stride(0), S, rows, cols, stream_ptr()), 'mxf_axpby2d')
    return out


def softplus_fwd(x, offset=0.0):
    require_cuda(x)
    x = _c(x)
    y = torch.empty_like(x)
    check(lib().mxf_softplus_fwd(dtype_code(x), ptr(x), float(offset), ptr(y), x.numel(), stream_ptr()),
          'mxf_softplus_fwd')
    return y


def softplus_bwd(x, gy):
    require_cuda(x, gy)
    x, gy = _c(x), _c(gy)
    gx = torch.empty_like(x)
    check(lib().mxf_softplus_bwd(dtype_code(x), ptr(x), ptr(gy), ptr(gx), x.numel(), stream_ptr()),
          'mxf_softplus_bwd')
    return gx


def svgp_bwd_assemble(Phi, T, U, mt, v, coef, out=None):
    """[E | E_S | E_R] into out[:, :, :3M] (out may be wider: extra right-hand-side columns are left untouched)."""
    require_cuda(Phi, T, U, mt, v, coef, out)
    S, M, _ = Phi.shape
    P = mt.shape[2]
    if out is None:
        out = torch.empty((S, M, 3 * M), dtype=Phi.dtype, device=Phi.device)
    check(lib().mxf_svgp_bwd_assemble(dtype_code(Phi), ptr(_c(Phi)), ptr(_c(T)), ptr(_c(U)), ptr(_c(mt)), ptr(_c(v)),
                                      ptr(_c(coef)), ptr(out), out.stride(1), out.stride(0), S, M, P, stream_ptr()),
          'mxf_svgp_bwd_assemble')
    return out


def _mlp_pack(x, Ws, bs):
    import ctypes
    L = len(Ws)
    S = max([x.shape[0]] + [w.shape[0] for w in Ws] + [b.shape[0] for b in bs if b is not None])
    widths = (ctypes.c_int * (L + 1))(*([Ws[0].shape[2]] + [w.shape[1] for w in Ws]))
    for l in range(L):
        if Ws[l].shape[2] != widths[l]:
            raise _lib.MXFusionB200Error("mlp_tanh: layer %d expects %d inputs, got %d" % (l, Ws[l].shape[2], widths[l]))
    Wp = (ctypes.c_void_p * L)(*[w.data_ptr() for w in Ws])
    sW = (ctypes.c_int64 * L)(*[_bstride(w, S) for w in Ws])
    has_b = all(b is not None for b in bs)
    bp = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs]) if has_b else None
    sb = (ctypes.c_int64 * L)(*[_bstride(b, S) for b in bs]) if has_b else None
    return L, S, widths, Wp, sW, bp, sb


def mlp_tanh_fwd(x, Ws, bs):
    """out (S,B,width_L) of the dense-tanh stack; x (S|1,B,w0), Ws[l] (S|1,out,in), bs[l] (S|1,out) or None."""
    require_cuda(x, *Ws, *[b for b in bs if b is not None])
    x = _c(x)
    Ws = [_c(w) for w in Ws]
    bs = [None if b is None else _c(b) for b in bs]
    L, S, widths, Wp, sW, bp, sb = _mlp_pack(x, Ws, bs)
    B = x.shape[1]
    out = torch.empty((S, B, widths[L]), dtype=x.dtype, device=x.device)
    check(lib().mxf_mlp_tanh_fwd(dtype_code(x), L, widths, ptr(x), _bstride(x, S), Wp, sW, bp, sb, ptr(out), S, B,
                                 stream_ptr()), 'mxf_mlp_tanh_fwd')
    return out


def mlp_tanh_bwd(x, Ws, bs, gout):
    """Gradients (dWs, dbs) of sum(out * gout), shaped like Ws / bs (a tensor shared by the samples gets the sum)."""
    import ctypes
    require_cuda(x, gout, *Ws)
    x, gout = _c(x), _c(gout)
    Ws = [_c(w) for w in Ws]
    bs = [None if b is None else _c(b) for b in bs]
    L, S, widths, Wp, sW, bp, sb = _mlp_pack(x, Ws, bs)
    B = x.shape[1]
    dWs = [torch.zeros_like(w) for w in Ws]
    dbs = [None if b is None else torch.zeros_like(b) for b in bs]
    dWp = (ctypes.c_void_p * L)(*[w.data_ptr() for w in dWs])
    dbp = (ctypes.c_void_p * L)(*[b.data_ptr() for b in dbs]) if bp is not None else None
    check(lib().mxf_mlp_tanh_bwd(dtype_code(x), L, widths, ptr(x), _bstride(x, S), Wp, sW, bp, sb, ptr(gout), dWp, dbp,
                                 S, B, stream_ptr()), 'mxf_mlp_tanh_bwd')
    return dWs, dbs


def svgp_bound_fwd(P, B, M, scale, sumr2, trPhi, trT, trPhiT, mm, sldL, sldLs, noise, kvar):
    """(logL, beta, Q), each (S,), from the per-sample reductions; noise / kvar are (S|1, 1)."""
    require_cuda(sumr2, trPhi, trT, trPhiT, mm, sldL, sldLs, noise, kvar)
    S = sumr2.shape[0]
    out = torch.empty((3, S), dtype=sumr2.dtype, device=sumr2.device)
    check(lib().mxf_svgp_bound_fwd(dtype_code(sumr2), S, int(P), int(B), int(M), float(scale), ptr(_c(sumr2)),
                                   ptr(_c(trPhi)), ptr(_c(trT)), ptr(_c(trPhiT)), ptr(_c(mm)), ptr(_c(sldL)),
                                   ptr(_c(sldLs)), ptr(noise), _bstride(noise, S), ptr(kvar), _bstride(kvar, S),
                                   ptr(out[0]), ptr(out[1]), ptr(out[2]), stream_ptr()), 'mxf_svgp_bound_fwd')
    return out[0], out[1], out[2]


def svgp_coef_bwd(P, B, scale, g, beta, Q):
    """(coef (S,6), gsb, -gsb, dnoise, dkvar_diag, -g, -1) for the adjoint of the SVGP bound."""
    require_cuda(g, beta, Q)
    S = g.shape[0]
    coef = torch.empty((S, 6), dtype=g.dtype, device=g.device)
    out = torch.empty((6, S), dtype=g.dtype, device=g.device)
    check(lib().mxf_svgp_coef_bwd(dtype_code(g), S, int(P), int(B), float(scale), ptr(_c(g)), ptr(_c(beta)), ptr(_c(Q)),
                                  ptr(coef), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]), ptr(out[4]),
                                  ptr(out[5]), stream_ptr()), 'mxf_svgp_coef_bwd')
    return coef, out[0], out[1], out[2], out[3], out[4], out[5]


def params_transform(flat, tflat, offs, sizes, kinds, offsets):
    """tflat[seg] = transform(flat[seg]) for every (offset, size, kind, softplus offset) segment: one launch."""
    import ctypes
    require_cuda(flat, tflat)
    n = len(offs)
    check(lib().mxf_params_transform(dtype_code(flat), n, (ctypes.c_int64 * n)(*offs), (ctypes.c_int64 * n)(*sizes),
                                     (ctypes.c_int * n)(*kinds), (ctypes.c_double * n)(*offsets), ptr(flat), ptr(tflat),
                                     stream_ptr()), 'mxf_params_transform')


def params_pack_grads(flat, gflat, grads, offs, sizes, kinds):
    """gflat[seg] = grads[t] * d transform / d raw (None: zeros) for every segment: one launch."""
    import ctypes
    require_cuda(flat, gflat, *[g for g in grads if g is not None])
    n = len(offs)
    gs = [None if g is None else _c(g) for g in grads]
    gp = (ctypes.c_void_p * n)(*[None if g is None else g.data_ptr() for g in gs])
    check(lib().mxf_params_pack_grads(dtype_code(flat), n, gp, (ctypes.c_int64 * n)(*offs), (ctypes.c_int64 * n)(*sizes),
                                      (ctypes.c_int * n)(*kinds), ptr(flat), ptr(gflat), stream_ptr()),
          'mxf_params_pack_grads')


def normal_reparam_bwd(gw, eps, v, m_samples, need=(True, True)):
    """(gm, gv) of the reparameterised draw; `m_samples` = leading size of the mean (1: shared, gradients are summed)."""
    require_cuda(gw, eps, v)
    S = gw.shape[0]
    gwf, ef, vf = _flat(_c(gw)), _flat(_c(eps)), _flat(_c(v))
    n = gwf.shape[1]
    shape = tuple(gw.shape[1:])
    gm = torch.empty(((m_samples if m_samples > 1 else 1), n), dtype=gw.dtype, device=gw.device) if need[0] else None
    gv = torch.empty((vf.shape[0], n), dtype=gw.dtype, device=gw.device) if need[1] else None
    check(lib().mxf_normal_reparam_bwd(dtype_code(gw), ptr(gwf), ptr(ef), ptr(vf), n if m_samples > 1 else 0,
                                       _bstride(vf, S), S, n, ptr(gm), ptr(gv), stream_ptr()), 'mxf_normal_reparam_bwd')
    return (None if gm is None else gm.reshape((gm.shape[0],) + shape),
            None if gv is None else gv.reshape((gv.shape[0],) + shape))


def _nl_tables(entries):
    import ctypes
    T = len(entries)
    xs, ms, vs, sX, sM, sV, ns, Ss, flags, scales = [], [], [], [], [], [], [], [], [], []
    for x, m, v, scale in entries:
        require_cuda(x, m, v)
        x, m, v = _flat(x), _flat(m), _flat(v)
        S = max(x.shape[0], m.shape[0], v.shape[0])
        n = max(x.shape[1], m.shape[1], v.shape[1])
        fl = 0
        for bit, t in ((1, x), (2, m), (4, v)):
            if t.shape[1] != n:
                if t.shape[1] != 1:
                    raise _lib.MXFusionB200Error("normal_logpdf_multi: operands must be full-size or one element per sample")
                fl |= bit
        xs.append(x); ms.append(m); vs.append(v)
        sX.append(_bstride(x, S)); sM.append(_bstride(m, S)); sV.append(_bstride(v, S))
        ns.append(n); Ss.append(S); flags.append(fl); scales.append(float(scale))
    arr = lambda ct, vals: (ct * T)(*vals)
    c = dict(x=arr(ctypes.c_void_p, [t.data_ptr() for t in xs]), m=arr(ctypes.c_void_p, [t.data_ptr() for t in ms]),
             v=arr(ctypes.c_void_p, [t.data_ptr() for t in vs]), sX=arr(ctypes.c_int64, sX), sM=arr(ctypes.c_int64, sM),
             sV=arr(ctypes.c_int64, sV), n=arr(ctypes.c_int64, ns), S=arr(ctypes.c_int, Ss), fl=arr(ctypes.c_int, flags),
             scale=arr(ctypes.c_double, scales))
    return T, xs, ms, vs, flags, c


def normal_logpdf_multi(entries):
    """sum_t scale_t * sum(mean_S(log N(x_t | m_t, v_t))) for a list of (x, m, v, scale): ONE launch -> tensor (1,)."""
    T, xs, ms, vs, flags, c = _nl_tables(entries)
    out = torch.zeros((1,), dtype=xs[0].dtype, device=xs[0].device)
    check(lib().mxf_normal_logpdf_multi(dtype_code(xs[0]), T, c['x'], c['m'], c['v'], c['sX'], c['sM'], c['sV'], c['n'],
                                        c['S'], c['fl'], c['scale'], ptr(out), stream_ptr()), 'mxf_normal_logpdf_multi')
    return out


def normal_logpdf_multi_bwd(entries, gout, needs):
    """Gradients [(gx, gm, gv)] per entry (None where `needs[t][k]` is False), shaped like the operands: ONE launch."""
    import ctypes
    T, xs, ms, vs, flags, c = _nl_tables(entries)
    gout = _c(gout).reshape(-1)
    grads, ptrs = [], ([], [], [])
    for t, (x, m, v, _) in enumerate(entries):
        row = []
        for k, (flat, orig, bit) in enumerate(((xs[t], x, 1), (ms[t], m, 2), (vs[t], v, 4))):
            if needs[t][k]:
                if flags[t] & bit:
                    raise _lib.MXFusionB200Error("normal_logpdf_multi_bwd: a one-element operand cannot receive a gradient")
                g = torch.empty_like(flat)
                row.append(g.reshape(orig.shape))
                ptrs[k].append(g.data_ptr())
            else:
                row.append(None)
                ptrs[k].append(None)
        grads.append(tuple(row))
    arr = lambda vals: (ctypes.c_void_p * T)(*vals)
    check(lib().mxf_normal_logpdf_multi_bwd(dtype_code(xs[0]), T, c['x'], c['m'], c['v'], c['sX'], c['sM'], c['sV'], c['n'],
                                            c['S'], c['fl'], c['scale'], ptr(gout), arr(ptrs[0]), arr(ptrs[1]),
                                            arr(ptrs[2]), stream_ptr()), 'mxf_normal_logpdf_multi_bwd')
    return grads


def normal_reparam_multi(entries, seed, offsets, step_counter=None):
    """Draws for a list of (m, v, S): w_t (S, *shape) = eps * sqrt(v) + m with in-kernel Philox streams (seed, offsets[t]);
    ONE launch.  Returns [(w, eps)]."""
    import ctypes
    T = len(entries)
    ms, vs, outs, sM, sV, ns, Ss = [], [], [], [], [], [], []
    for m, v, S in entries:
        require_cuda(m, v)
        shape = tuple(m.shape[1:])
        mf, vf = _flat(m), _flat(v)
        n = mf.shape[1]
        w = torch.empty((S, n), dtype=m.dtype, device=m.device)
        e = torch.empty_like(w)
        ms.append(mf); vs.append(vf); outs.append((w, e, shape))
        sM.append(_bstride(mf, S)); sV.append(_bstride(vf, S)); ns.append(n); Ss.append(int(S))
    vp = lambda ts: (ctypes.c_void_p * T)(*[t.data_ptr() for t in ts])
    check(lib().mxf_normal_reparam_multi(dtype_code(ms[0]), T, vp(ms), vp(vs), (ctypes.c_int64 * T)(*sM),
                                         (ctypes.c_int64 * T)(*sV), (ctypes.c_int64 * T)(*ns), (ctypes.c_int * T)(*Ss),
                                         int(seed), (ctypes.c_uint64 * T)(*[int(o) for o in offsets]), ptr(step_counter),
                                         vp([o[0] for o in outs]), vp([o[1] for o in outs]), stream_ptr()),
          'mxf_normal_reparam_multi')
    return [(w.reshape((w.shape[0],) + shape), e.reshape((e.shape[0],) + shape)) for w, e, shape in outs]


def normal_reparam_multi_bwd(entries, needs):
    """Adjoints for a list of (gw, eps, v, m_samples): [(gm, gv)] (None where not needed); ONE launch."""
    import ctypes
    T = len(entries)
    gws, es, vs, sM, sV, ns, Ss, res, pm, pv = [], [], [], [], [], [], [], [], [], []
    for (gw, e, v, m_samples), need in zip(entries, needs):
        require_cuda(gw, e, v)
        S = gw.shape[0]
        shape = tuple(gw.shape[1:])
        gwf, ef, vf = _flat(gw), _flat(e), _flat(v)
        n = gwf.shape[1]
        gm = torch.empty(((m_samples if m_samples > 1 else 1), n), dtype=gw.dtype, device=gw.device) if need[0] else None
        gv = torch.empty((vf.shape[0], n), dtype=gw.dtype, device=gw.device) if need[1] else None
        gws.append(gwf); es.append(ef); vs.append(vf)
        sM.append(n if m_samples > 1 else 0); sV.append(_bstride(vf, S)); ns.append(n); Ss.append(S)
        pm.append(None if gm is None else gm.data_ptr()); pv.append(None if gv is None else gv.data_ptr())
        res.append((None if gm is None else gm.reshape((gm.shape[0],) + shape),
                    None if gv is None else gv.reshape((gv.shape[0],) + shape)))
    vp = lambda ts: (ctypes.c_void_p * T)(*[t.data_ptr() for t in ts])
    check(lib().mxf_normal_reparam_multi_bwd(dtype_code(gws[0]), T, vp(gws), vp(es), vp(vs), (ctypes.c_int64 * T)(*sM),
                                             (ctypes.c_int64 * T)(*sV), (ctypes.c_int64 * T)(*ns), (ctypes.c_int * T)(*Ss),
                                             (ctypes.c_void_p * T)(*pm), (ctypes.c_void_p * T)(*pv), stream_ptr()),
          'mxf_normal_reparam_multi_bwd')
    return res
