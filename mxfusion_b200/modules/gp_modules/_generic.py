"""The three GP bounds written over the differentiable primitives (ops.potrf / trsm / gemm2 / syrk / sumlogdiag /
make_diagonal, all CUDA launches with hand-written adjoints), operation for operation as the reference writes them.
Used when the kernel is not a single stationary kernel (Add / Multiply / Linear / Bias / White): the fused
operators of ops.py (`svgp_log_pdf`, `gp_log_pdf`, `sparsegp_log_pdf`) build K(X, X2) inside their own kernels and
therefore only take RBF / Matern."""
import math

import torch

from ... import ops

_LOG2PI = math.log(2.0 * math.pi)


def _eye(M, like):
    return torch.eye(M, dtype=like.dtype, device=like.device).unsqueeze(0)


def svgp_log_pdf(F, kern, kern_params, X, Y, Z, noise_var, mu, S_W, S_diag, jitter, log_pdf_scaling, mean=None):
    """svgp_regression.py:61-109 (homoscedastic noise)."""
    D, M = Y.shape[-1], Z.shape[-2]
    if noise_var.dim() == 2:                                             # :61-62 (heteroscedastic noise has ndim 3)
        noise_var = noise_var.unsqueeze(-2)
    if noise_var.shape[-1] == 1:                                         # :64-67
        beta_sum = D * torch.sum(1 / noise_var, dim=-1)
    else:
        beta_sum = torch.sum(1 / noise_var, dim=-1)
    Kuu = kern.K(F, Z, **kern_params)                                    # :69
    if jitter > 0.:
        Kuu = Kuu + _eye(M, Z) * jitter                                  # :70-72
    Kuf = kern.K(F, Z, X, **kern_params)                                 # :73
    Kff_diag = kern.Kdiag(F, X, **kern_params)                           # :74
    S = ops.syrk(S_W) + ops.make_diagonal(S_diag)                        # :76
    if mean is not None:
        Y = Y - mean                                                     # :78-80
    psi1Y = ops.gemm2(Kuf, Y / noise_var, False, False)                  # :82
    L = ops.potrf(Kuu)                                                   # :83
    Ls = ops.potrf(S)                                                    # :84
    LinvLs = ops.trsm(L, Ls)                                             # :85
    Linvmu = ops.trsm(L, mu)                                             # :86
    LinvKuf = ops.trsm(L, Kuf)                                           # :87
    KfuKuuInvmu = ops.gemm2(LinvKuf, Linvmu, True, False)                # :89
    KfuKuuInvLs = ops.gemm2(LinvKuf, LinvLs, True, False)                # :90
    LinvKufY = ops.trsm(L, psi1Y)                                        # :92
    KL_u = (M / 2. + ops.sumlogdiag(Ls)) * D - ops.sumlogdiag(L) * D \
        - torch.sum(torch.square(LinvLs), dim=(-1, -2)) / 2. * D \
        - torch.sum(torch.square(Linvmu), dim=(-1, -2)) / 2.            # :94-96
    logL = -torch.sum(torch.square(Y) / noise_var + _LOG2PI + torch.log(noise_var), dim=(-1, -2)) / 2.   # :98-99
    logL = logL - torch.sum(Kff_diag * beta_sum, dim=-1) / 2.                                            # :100
    logL = logL - torch.sum(torch.square(KfuKuuInvmu) / noise_var, dim=(-1, -2)) / 2.
    logL = logL - torch.sum(torch.square(KfuKuuInvLs) * beta_sum.unsqueeze(-1), dim=(-1, -2)) / 2.
    logL = logL + torch.sum(torch.square(LinvKuf) * beta_sum.unsqueeze(-2), dim=(-1, -2)) / 2.
    logL = logL + torch.sum(Linvmu * LinvKufY, dim=(-1, -2))                                             # :107
    return log_pdf_scaling * logL + KL_u                                                                 # :108


def gp_log_pdf(F, kern, kern_params, X, Y, noise_var, jitter, mean=None):
    """gp_regression.py:55-70 -> (logL, L, LinvY)."""
    D, N = Y.shape[-1], X.shape[-2]
    K = kern.K(F, X, **kern_params) + _eye(N, X) * noise_var.unsqueeze(-2)          # :55-57
    if jitter > 0.:
        K = K + _eye(N, X) * jitter                                                  # :58-60
    L = ops.potrf(K)                                                                 # :61
    if mean is not None:
        Y = Y - mean
    LinvY = ops.trsm(L, Y)                                                           # :66
    logdet_l = ops.sumlogdiag(L)                                                     # :67
    tmp = torch.sum(torch.square(LinvY) + _LOG2PI, dim=(-1, -2))                     # :68
    return -logdet_l * D - tmp / 2, L.detach(), LinvY.detach()                       # :70


def sparsegp_log_pdf(F, kern, kern_params, X, Y, Z, noise_var, jitter, mean=None):
    """sparsegp_regression.py:62-106, materialising K(Z, X) as the reference does -> (logL, wv, L, LA)."""
    D, M = Y.shape[-1], Z.shape[-2]
    noise_var_m = noise_var.unsqueeze(-2)                                            # :70
    Kuu = kern.K(F, Z, **kern_params)                                                # :72
    if jitter > 0.:
        Kuu = Kuu + _eye(M, Z) * jitter
    Kuf = kern.K(F, Z, X, **kern_params)                                             # :77
    Kff_diag = kern.Kdiag(F, X, **kern_params)                                       # :78
    L = ops.potrf(Kuu)                                                               # :80
    LinvKuf = ops.trsm(L, Kuf)                                                       # :81
    A = _eye(M, Z) + ops.syrk(LinvKuf) / noise_var_m                                 # :83-84
    LA = ops.potrf(A)                                                                # :85
    if mean is not None:
        Y = Y - mean
    LAInvLinvKufY = ops.trsm(LA, ops.gemm2(LinvKuf, Y))                              # :90
    logL = -D * ops.sumlogdiag(LA)                                                   # :92
    logL = logL - torch.sum(torch.square(Y) / noise_var_m + _LOG2PI + torch.log(noise_var_m), dim=(-1, -2)) / 2
    logL = logL + torch.sum(torch.square(LAInvLinvKufY) / (2 * torch.square(noise_var_m)), dim=(-1, -2))
    logL = logL - D * torch.sum(Kff_diag / (2 * noise_var), dim=-1)                  # :98
    logL = logL + D * torch.sum(torch.square(LinvKuf) / (2. * noise_var_m), dim=(-1, -2))   # :99-100
    with torch.no_grad():                                                            # :101-106
        wv = ops.trsm(L, ops.trsm(LA, LAInvLinvKufY, transpose=True), transpose=True) / noise_var_m
    return logL, wv, L.detach(), LA.detach()
