"""CPU stand-in for `mxfusion_b200._raw` -- TEST INFRASTRUCTURE ONLY.

The product's operators (mxfusion_b200/ops.py) call the CUDA library through `mxfusion_b200._raw`
and raise on CPU tensors.  To check the HOST-SIDE algebra (hand-derived adjoints, module/graph/loop
logic) in the CPU-only test tier, tests monkeypatch `ops.R` with this module, which offers the same
function signatures on plain torch CPU tensors.  Nothing under mxfusion_b200/ imports it.
"""
import math

import torch

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
RED_SUM, RED_SUMSQ, RED_DOT, RED_SUMLOG, RED_SUMSQDIFF = 0, 1, 2, 3, 4
launches = 0


def launch_count():
    return 0


def _k(kind, X, X2, ls, var):
    from oracle import torch_ref
    return torch_ref.K(kind, X, ls, var, X2)


def kbuild_fwd(kind, X, X2, ls, var, diag_add=None, diag_const=0.0, out=None):
    K = _k(kind, X, X2, ls, var)
    if X2 is None:
        n = K.shape[-1]
        d = torch.full((K.shape[0], 1), float(diag_const), dtype=K.dtype)
        if diag_add is not None:
            d = d + diag_add
        K = K + torch.eye(n, dtype=K.dtype).unsqueeze(0) * d.unsqueeze(-1)
    if out is not None:
        out.copy_(K)
        return out
    return K.contiguous()


def kbuild_bwd(kind, X, X2, ls, var, G, need_dX=True, need_dX2=True):
    with torch.enable_grad():
        Xr = X.detach().clone().requires_grad_()
        X2r = None if X2 is None else X2.detach().clone().requires_grad_()
        lsr = ls.detach().clone().requires_grad_()
        vr = var.detach().clone().requires_grad_()
        K = _k(kind, Xr, X2r, lsr, vr)
        (K * G).sum().backward()
    return Xr.grad, (None if X2 is None else X2r.grad), lsr.grad, vr.grad


def gemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None, tri=False):
    a = A.transpose(-1, -2) if transA else A
    b = B.transpose(-1, -2) if transB else B
    prod = alpha * torch.matmul(a, b)
    if C is None:
        return torch.tril(prod).contiguous() if tri else prod.contiguous()
    acc = prod if beta == 0.0 else prod + beta * C      # like the kernels, beta == 0 never reads C (it may be uninitialised)
    if tri:
        C.copy_(torch.where(torch.ones_like(C, dtype=torch.bool).tril(), acc, C))
    else:
        C.copy_(acc)
    return C


def potrf_(A, info=None):
    L, inf = torch.linalg.cholesky_ex(torch.tril(A) + torch.tril(A, -1).transpose(-1, -2))
    A.copy_(L)
    return A, inf.to(torch.int32)


def new_pack(A):
    return torch.zeros((A.shape[0], 1), dtype=A.dtype)


def potrf_packed_(A, info=None, pack=None):
    A, inf = potrf_(A, info)
    return A, inf, new_pack(A) if pack is None else pack


def tri_pack(L):
    return torch.zeros((L.shape[0], 1), dtype=L.dtype)


def trsm_packed_(L, pack, B, transpose=False, alpha=1.0):
    return trsm_(L, B, transpose=transpose, alpha=alpha)


def trsm_solve(L, pack, B, transpose=False):
    return trsm_(L, B.clone(), transpose=transpose)


def trsm_(L, B, transpose=False, alpha=1.0):
    Lt = torch.tril(L)
    if transpose:
        X = torch.linalg.solve_triangular(Lt.transpose(-1, -2), B, upper=True)
    else:
        X = torch.linalg.solve_triangular(Lt, B, upper=False)
    B.copy_(alpha * X)
    return B


def copy_ltu(P, out=None):
    r = (torch.tril(P) + torch.tril(P, -1).transpose(-1, -2)).contiguous()
    if out is not None:
        out.copy_(r)
        return out
    return r


def symmetrize(A, alpha=1.0):
    return (alpha * (A + A.transpose(-1, -2))).contiguous()


def tril(A, strict=False):
    return torch.tril(A, -1 if strict else 0).contiguous()


def transpose(A, out=None):
    t = A.transpose(-1, -2)
    if out is not None:
        out.copy_(t)
        return out
    return t.contiguous()


def reduce(op, a, b=None, scale=1.0):
    if op == RED_SUM:
        t = a
    elif op == RED_SUMSQ:
        t = a * a
    elif op == RED_DOT:
        t = a * b
    elif op == RED_SUMLOG:
        t = torch.log(a)
    else:
        t = (a - b) ** 2
    return scale * t.sum(dim=(1, 2))


def sumlogdiag(A):
    return torch.log(torch.abs(torch.diagonal(A, dim1=-2, dim2=-1))).sum(-1)


def add_diag_(A, d=None, c=0.0):
    n = A.shape[-1]
    add = torch.full((A.shape[0], n), float(c), dtype=A.dtype)
    if d is not None:
        add = add + d
    A.diagonal(dim1=-2, dim2=-1).add_(add)
    return A


def get_diag(A):
    return torch.diagonal(A, dim1=-2, dim2=-1).contiguous()


def _lp(x, m, v):
    return -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(v) - (x - m) ** 2 / (2 * v)


def normal_logpdf_sum(x, m, v, scale=1.0):
    return (scale * torch.sum(torch.mean(_lp(x, m, v), dim=0))).reshape(1)


def normal_logpdf_sum_bwd(x, m, v, gout, scale=1.0, need=(True, True, True)):
    with torch.enable_grad():
        xr, mr, vr = [t.detach().clone().requires_grad_() for t in (x, m, v)]
        (normal_logpdf_sum(xr, mr, vr, scale) * gout.reshape(-1)[0]).sum().backward()
    return (xr.grad if need[0] else None, mr.grad if need[1] else None, vr.grad if need[2] else None)


def normal_reparam(m, v, S, eps=None, seed=0, offset=0, return_eps=False, step_counter=None):
    shape = (S,) + tuple(m.shape[1:])
    if eps is None:
        g = torch.Generator().manual_seed(int(seed) * 1000003 + int(offset))
        eps = torch.randn(shape, generator=g, dtype=m.dtype)
    w = eps * torch.sqrt(v) + m
    return (w, eps) if return_eps else w


def adam_step_(w, g, m, v, step_count, lr, beta1=0.9, beta2=0.999, eps=1e-8, rescale=1.0):
    t = int(step_count.item()) + 1
    gg = g * rescale
    m.mul_(beta1).add_(gg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    w.sub_(lr_t * m / (v.sqrt() + eps))
    step_count += 1


def sgd_step_(w, g, mom, step_count, lr, momentum=0.0, rescale=1.0):
    if momentum != 0.0:
        mom.mul_(momentum).sub_(g, alpha=lr * rescale)
        w.add_(mom)
    else:
        w.sub_(g, alpha=lr * rescale)
    step_count += 1


def gather_rows(src, idx, off, rows, out=None):
    o = int(off.item()) if off is not None else 0
    r = src[idx[o:o + rows]]
    if out is not None:
        out.copy_(r)
        return out
    return r


def axpby_dev(a, X, b=None, Y=None, out=None):
    def cv(c):
        return c.reshape((-1,) + (1,) * (X.dim() - 1))
    r = X if a is None else cv(a) * X
    if Y is not None:
        r = r + cv(b) * Y
    if out is not None:
        out.copy_(r)
        return out
    return r.contiguous()


def softplus_fwd(x, offset=0.0):
    return torch.nn.functional.softplus(x) + offset


def softplus_bwd(x, gy):
    return gy * torch.sigmoid(x)


def svgp_bwd_assemble(Phi, T, U, mt, v, coef, out=None):
    M = Phi.shape[-1]
    I = torch.eye(M, dtype=Phi.dtype).unsqueeze(0)
    c = [coef[:, i].reshape(-1, 1, 1) for i in range(6)]
    mm = mt @ mt.transpose(-1, -2)
    vm = v @ mt.transpose(-1, -2)
    E = c[2] * mm + c[0] * (T - I) - c[1] * Phi + c[1] * (U + U.transpose(-1, -2)) - c[3] * (vm + vm.transpose(-1, -2))
    ES = c[0] * I + c[1] * Phi
    ER = -c[4] * (T - I) - c[5] * mm
    r = torch.cat([E, ES, ER], dim=-1).contiguous()
    if out is not None:
        out[:, :, :3 * M].copy_(r)
        return out
    return r


def _mlp(x, Ws, bs):
    h = x
    for l, (W, b) in enumerate(zip(Ws, bs)):
        h = torch.matmul(h, W.transpose(-1, -2))
        if b is not None:
            h = h + b.unsqueeze(-2)
        if l + 1 < len(Ws):
            h = torch.tanh(h)
    return h


def mlp_tanh_fwd(x, Ws, bs):
    return _mlp(x, Ws, bs).contiguous()


def mlp_tanh_bwd(x, Ws, bs, gout):
    with torch.enable_grad():
        Wr = [w.detach().clone().requires_grad_() for w in Ws]
        br = [None if b is None else b.detach().clone().requires_grad_() for b in bs]
        (_mlp(x.detach(), Wr, br) * gout).sum().backward()
    return [w.grad for w in Wr], [None if b is None else b.grad for b in br]


def svgp_bound_fwd(P, B, M, scale, sumr2, trPhi, trT, trPhiT, mm, sldL, sldLs, noise, kvar):
    nv, kv = noise[:, 0], kvar[:, 0]
    beta = 1.0 / nv
    Q = -0.5 * sumr2 - (0.5 * P * B) * kv - (0.5 * P) * (trPhiT - trPhi)
    data = beta * Q - (0.5 * B * P) * (math.log(2.0 * math.pi) + torch.log(nv))
    neg_kl = P * (0.5 * M + sldLs - sldL) - (0.5 * P) * trT - 0.5 * mm
    S = sumr2.shape[0]
    return (scale * data + neg_kl).expand(S).contiguous(), beta.expand(S).contiguous(), Q.expand(S).contiguous()


def svgp_coef_bwd(P, B, scale, g, beta, Q):
    gsb = g * (scale * beta)
    coef = torch.stack([g * (0.5 * P), gsb * (0.5 * P), 0.5 * g, 0.5 * gsb, gsb * P, gsb], dim=1).contiguous()
    return (coef, gsb, -gsb, g * scale * (-beta * beta * Q - (0.5 * B * P) * beta), -gsb * (0.5 * P * B), -g,
            torch.full_like(g, -1.0))


def params_transform(flat, tflat, offs, sizes, kinds, offsets):
    for o, n, k, c in zip(offs, sizes, kinds, offsets):
        x = flat[o:o + n]
        tflat[o:o + n] = torch.nn.functional.softplus(x) + c if k == 1 else x


def params_pack_grads(flat, gflat, grads, offs, sizes, kinds):
    for g, o, n, k in zip(grads, offs, sizes, kinds):
        if g is None:
            gflat[o:o + n] = 0
        else:
            gv = g.reshape(-1)
            gflat[o:o + n] = gv * torch.sigmoid(flat[o:o + n]) if k == 1 else gv


def normal_reparam_bwd(gw, eps, v, m_samples, need=(True, True)):
    gm = gw if m_samples == gw.shape[0] else gw.sum(dim=0, keepdim=True)
    gv = gw * eps * (0.5 / torch.sqrt(v))
    if v.shape[0] != gv.shape[0]:
        gv = gv.sum(dim=0, keepdim=True)
    return (gm if need[0] else None), (gv if need[1] else None)


def normal_logpdf_multi(entries):
    out = torch.zeros((1,), dtype=entries[0][0].dtype)
    for x, m, v, scale in entries:
        S = max(x.shape[0], m.shape[0], v.shape[0])
        lp = -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(v) - (x - m) ** 2 / (2 * v)
        lp = lp.expand((S,) + tuple(lp.shape[1:]))
        out = out + scale * lp.sum() / S
    return out


def normal_logpdf_multi_bwd(entries, gout, needs):
    res = []
    for (x, m, v, scale), need in zip(entries, needs):
        with torch.enable_grad():
            xr, mr, vr = (t.detach().clone().requires_grad_() for t in (x, m, v))
            S = max(x.shape[0], m.shape[0], v.shape[0])
            lp = -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(vr) - (xr - mr) ** 2 / (2 * vr)
            lp = lp.expand((S,) + tuple(lp.shape[1:]))
            (scale * lp.sum() / S * gout.reshape(-1)[0]).backward()
        res.append(tuple(t.grad if n else None for t, n in zip((xr, mr, vr), need)))
    return res


def normal_reparam_multi(entries, seed, offsets, step_counter=None):
    out = []
    for (m, v, S), off in zip(entries, offsets):
        g = torch.Generator().manual_seed(int(seed) * 1000003 + int(off))
        eps = torch.randn((S,) + tuple(m.shape[1:]), generator=g, dtype=m.dtype)
        out.append((eps * torch.sqrt(v) + m, eps))
    return out


def normal_reparam_multi_bwd(entries, needs):
    return [normal_reparam_bwd(gw, e, v, ms, need=need) for (gw, e, v, ms), need in zip(entries, needs)]


def copy_ltu_sum(parts, out=None):
    low = torch.tril(parts.sum(dim=0, keepdim=True))
    r = low + torch.tril(low, -1).transpose(-1, -2)
    if out is not None:
        out.copy_(r)
        return out
    return r.contiguous()


def copy2d_(dst, src=None):
    if src is None:
        dst.zero_()
    else:
        dst.copy_(src)
    return dst


def axpby2d(a, X, b=None, Y=None, out=None):
    def co(c):
        return 1.0 if c is None else c.reshape(-1, 1, 1)
    r = co(a) * X
    if Y is not None:
        r = r + co(b) * Y
    if out is not None:
        out.copy_(r)
        return out
    return r.contiguous()


def pack_inverse(pack, like):
    return None
