"""SVGP (Hensman et al. 2013) evidence lower bound restated in NumPy
(test infrastructure).

``mxfusion/modules/gp_modules/svgp_regression.py:43-109`` operation for
operation, plus an independent dense formulation used to pin it.
Semantics that are easy to lose: the reference's
``KL_u`` variable is MINUS the KL; ``log_pdf_scaling`` multiplies only the data
term (``svgp_regression.py:108``).
"""
import numpy as np
from . import kernels, linalg


def svgp_log_pdf(kind, X, Y, Z, noise_var, mu, S_W, S_diag, lengthscale, variance,
                 jitter=0.0, log_pdf_scaling=1.0, mean=None, return_parts=False):
    """All inputs carry the leading sample axis: X (S,B,Din), Y (S,B,P),
    Z (S,M,Din), noise_var (S,1), mu (S,M,P), S_W (S,M,M), S_diag (S,M),
    lengthscale (S,1|Din), variance (S,1).  Returns logL (S,)."""
    D = Y.shape[-1]
    M = Z.shape[-2]
    if noise_var.ndim == 2:                                  # :61-62
        noise_var = noise_var[..., None, :]
    if noise_var.shape[-1] == 1:                             # :64-67
        beta_sum = D * np.sum(1 / noise_var, axis=-1)
    else:
        beta_sum = np.sum(1 / noise_var, axis=-1)

    Kuu = kernels.K(kind, Z, lengthscale, variance)          # :69
    if jitter > 0.:
        Kuu = Kuu + np.eye(M, dtype=Z.dtype)[None] * jitter  # :70-72
    Kuf = kernels.K(kind, Z, lengthscale, variance, X)       # :73
    Kff_diag = kernels.Kdiag(X, variance)                    # :74

    S = linalg.syrk(S_W) + linalg.make_diagonal(S_diag)      # :76
    if mean is not None:
        Y = Y - mean                                         # :78-80

    psi1Y = linalg.gemm2(Kuf, Y / noise_var, False, False)   # :82
    L = linalg.potrf(Kuu)                                    # :83
    Ls = linalg.potrf(S)                                     # :84
    LinvLs = linalg.trsm(L, Ls)                              # :85
    Linvmu = linalg.trsm(L, mu)                              # :86
    LinvKuf = linalg.trsm(L, Kuf)                            # :87

    KfuKuuInvmu = linalg.gemm2(LinvKuf, Linvmu, True, False)   # :89
    KfuKuuInvLs = linalg.gemm2(LinvKuf, LinvLs, True, False)   # :90
    LinvKufY = linalg.trsm(L, psi1Y)                           # :92

    KL_u = (M / 2. + linalg.sumlogdiag(Ls)) * D - linalg.sumlogdiag(L) * D \
        - np.sum(np.sum(np.square(LinvLs), axis=-1), axis=-1) / 2. * D \
        - np.sum(np.sum(np.square(Linvmu), axis=-1), axis=-1) / 2.      # :94-96

    logL = -np.sum(np.sum(np.square(Y) / noise_var + np.log(2. * np.pi) +
                          np.log(noise_var), axis=-1), axis=-1) / 2.    # :98-99
    logL = logL - np.sum(Kff_diag * beta_sum, axis=-1) / 2.             # :100
    logL = logL - np.sum(np.sum(np.square(KfuKuuInvmu) / noise_var, axis=-1), axis=-1) / 2.
    logL = logL - np.sum(np.sum(np.square(KfuKuuInvLs) * beta_sum[..., None], axis=-1), axis=-1) / 2.
    logL = logL + np.sum(np.sum(np.square(LinvKuf) * beta_sum[..., None, :], axis=-1), axis=-1) / 2.
    logL = logL + np.sum(np.sum(Linvmu * LinvKufY, axis=-1), axis=-1)   # :107
    data_term = logL
    logL = log_pdf_scaling * logL + KL_u                                # :108
    if return_parts:
        return logL, data_term, -KL_u
    return logL


def svgp_elbo_independent(kind, X, Y, Z, noise_var, mu, S_W, S_diag, lengthscale, variance,
                          jitter=0.0, log_pdf_scaling=1.0):
    """Independent formulation (unbatched inputs, homoscedastic noise): dense
    inverses, marginals of q(f), expected Gaussian log-likelihood and the closed
    form KL(q(u)||p(u)).  Returns (elbo, data_term, kl)."""
    M = Z.shape[0]
    P = Y.shape[1]
    s2 = float(noise_var[0])
    Kuu = kernels.K_direct(kind, Z[None], lengthscale[None], variance[None])[0] + jitter * np.eye(M)
    Kuf = kernels.K_direct(kind, Z[None], lengthscale[None], variance[None], X[None])[0]
    kff = np.full(X.shape[0], float(variance[0]))
    S = S_W @ S_W.T + np.diag(S_diag)
    Kinv = np.linalg.inv(Kuu)
    A = Kinv @ Kuf                                  # (M,B)
    m = A.T @ mu                                    # (B,P)
    v = kff - np.sum(Kuf * A, axis=0) + np.sum(A * (S @ A), axis=0)
    ell = np.sum(-0.5 * np.log(2 * np.pi * s2) - 0.5 * ((Y - m) ** 2 + v[:, None]) / s2)
    _, ld_k = np.linalg.slogdet(Kuu)
    _, ld_s = np.linalg.slogdet(S)
    kl = 0.5 * (P * np.trace(Kinv @ S) + np.sum(mu * (Kinv @ mu)) - P * M + P * ld_k - P * ld_s)
    return log_pdf_scaling * ell - kl, ell, kl


def svgp_predict(kind, Xt, Z, noise_var, mu, S_W, S_diag, lengthscale, variance,
                 jitter=0.0, noise_free=True, diagonal_variance=True, mean=None):
    """svgp_regression.py:145-182 (mean/variance prediction), batched over S."""
    M = Z.shape[-2]
    N = Xt.shape[-2]
    S = linalg.syrk(S_W) + linalg.make_diagonal(S_diag)
    Kuu = kernels.K(kind, Z, lengthscale, variance)
    if jitter > 0.:
        Kuu = Kuu + np.eye(M, dtype=Z.dtype) * jitter
    L = linalg.potrf(Kuu)
    Ls = linalg.potrf(S)
    LinvLs = linalg.trsm(L, Ls)
    Linvmu = linalg.trsm(L, mu)
    LinvSLinvT = linalg.syrk(LinvLs)
    wv = linalg.trsm(L, Linvmu, transpose=True)
    Kxt = kernels.K(kind, Z, lengthscale, variance, Xt)
    mean_f = linalg.gemm2(Kxt, wv, True, False)
    if mean is not None:
        mean_f = mean_f + mean
    LinvKxt = linalg.trsm(L, Kxt)
    tmp = linalg.gemm2(LinvSLinvT, LinvKxt)
    if diagonal_variance:
        var = kernels.Kdiag(Xt, variance) - np.sum(np.square(LinvKxt), axis=-2) + \
            np.sum(tmp * LinvKxt, axis=-2)
        var = var[..., None]
        if not noise_free:
            var = var + noise_var
    else:
        var = kernels.K(kind, Xt, lengthscale, variance) - linalg.syrk(LinvKxt, True) + \
            linalg.gemm2(LinvKxt, tmp, True, False)
        var = var[..., None]
        if not noise_free:
            var = var + np.eye(N, dtype=Xt.dtype).reshape(1, N, N, 1) * noise_var[..., None, :]
    return mean_f, var
