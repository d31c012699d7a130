"""Variational sparse GP regression (Titsias 2009 collapsed bound) restated in NumPy
(test infrastructure).

``mxfusion/modules/gp_modules/sparsegp_regression.py:42-108`` (bound) and
``:119-171`` (mean / variance prediction), operation for operation, plus an
independent dense formulation (log N(Y | 0, Qff + s2 I) - tr(Kff - Qff) / (2 s2))
used to pin it.  As in the reference, ``log_pdf_scaling`` is NOT applied by this
bound, and ``wv``, ``L``, ``LA`` are the quantities cached for prediction
(``:101-106``).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import numpy as np
from . import kernels, linalg


def sparsegp_log_pdf(kind, X, Y, Z, noise_var, lengthscale, variance, jitter=0.0, mean=None,
                     return_cache=False):
    """X (S,N,Din), Y (S,N,P), Z (S,M,Din), noise_var (S,1), lengthscale (S,1|Din),
    variance (S,1).  Returns logL (S,) [, (wv, L, LA)]."""
    D = Y.shape[-1]
    M = Z.shape[-2]
    noise_var_m = noise_var[..., None, :]                               # :70
    Kuu = kernels.K(kind, Z, lengthscale, variance)                      # :72
    if jitter > 0.:
        Kuu = Kuu + np.eye(M, dtype=Z.dtype)[None] * jitter              # :73-75
    Kuf = kernels.K(kind, Z, lengthscale, variance, X)                   # :77
    Kff_diag = kernels.Kdiag(X, variance)                                # :78
    L = linalg.potrf(Kuu)                                                # :80
    LinvKuf = linalg.trsm(L, Kuf)                                        # :81
    A = np.eye(M, dtype=Z.dtype)[None] + linalg.syrk(LinvKuf) / noise_var_m   # :83-84
    LA = linalg.potrf(A)                                                 # :85
    if mean is not None:
        Y = Y - mean                                                     # :87-89
    LAInvLinvKufY = linalg.trsm(LA, linalg.gemm2(LinvKuf, Y))            # :90
    logL = -D * linalg.sumlogdiag(LA)                                    # :92
    logL = logL - np.sum(np.sum(np.square(Y) / noise_var_m + np.log(2. * np.pi) +
                                np.log(noise_var_m), axis=-1), axis=-1) / 2   # :93-94
    logL = logL + np.sum(np.sum(np.square(LAInvLinvKufY) / (2 * np.square(noise_var_m)), axis=-1), axis=-1)  # :95-97
    logL = logL - D * np.sum(Kff_diag / (2 * noise_var), axis=-1)        # :98
    logL = logL + D * np.sum(np.sum(np.square(LinvKuf) / (2. * noise_var_m), axis=-1), axis=-1)  # :99-100
    if return_cache:
        wv = linalg.trsm(L, linalg.trsm(LA, LAInvLinvKufY, transpose=True), transpose=True) / noise_var_m  # :102-105
        return logL, (wv, L, LA)
    return logL


def sparsegp_bound_independent(kind, X, Y, Z, noise_var, lengthscale, variance, jitter=0.0):
    """Independent formulation (unbatched inputs): sum_p log N(y_p | 0, Qff + s2 I) - P tr(Kff - Qff) / (2 s2)
    with Qff = Kfu Kuu^-1 Kuf formed densely."""
    N, P = Y.shape
    M = Z.shape[0]
    s2 = float(noise_var[0])
    Kuu = kernels.K_direct(kind, Z[None], lengthscale[None], variance[None])[0] + jitter * np.eye(M)
    Kuf = kernels.K_direct(kind, Z[None], lengthscale[None], variance[None], X[None])[0]
    Qff = Kuf.T @ np.linalg.solve(Kuu, Kuf)
    C = Qff + s2 * np.eye(N)
    _, ld = np.linalg.slogdet(C)
    quad = np.sum(Y * np.linalg.solve(C, Y))
    ll = -0.5 * P * (N * np.log(2 * np.pi) + ld) - 0.5 * quad
    return ll - P * (N * float(variance[0]) - np.trace(Qff)) / (2 * s2)


def sparsegp_predict(kind, Xt, Z, wv, L, LA, noise_var, lengthscale, variance,
                     noise_free=True, diagonal_variance=True, mean=None):
    """sparsegp_regression.py:131-171 (mean / variance prediction), batched over S."""
    Kxt = kernels.K(kind, Z, lengthscale, variance, Xt)
    mu = linalg.gemm2(Kxt, wv, True, False)
    if mean is not None:
        mu = mu + mean
    LinvKxt = linalg.trsm(L, Kxt)
    LAinvLinvKxt = linalg.trsm(LA, LinvKxt)
    if diagonal_variance:
        var = kernels.Kdiag(Xt, variance) - np.sum(np.square(LinvKxt), axis=-2) + \
            np.sum(np.square(LAinvLinvKxt), axis=-2)
        if not noise_free:
            var = var + noise_var
    else:
        N = Xt.shape[-2]
        var = kernels.K(kind, Xt, lengthscale, variance) - linalg.syrk(LinvKxt, True) + \
            linalg.syrk(LAinvLinvKxt, True)
        if not noise_free:
            var = var + np.eye(N, dtype=Xt.dtype)[None] * noise_var[..., None, :]
    return mu, var
