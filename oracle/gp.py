"""Exact GP regression marginal likelihood restated in NumPy (test infrastructure).

``mxfusion/modules/gp_modules/gp_regression.py:42-76``.
"""
import numpy as np
from . import kernels, linalg


def gp_log_pdf(kind, X, Y, noise_var, lengthscale, variance, jitter=0.0, mean=None):
    """Returns (logL (S,), L (S,N,N), LinvY (S,N,P)).  gp_regression.py:55-70.
    Note: log_pdf_scaling is *not* applied here (gp_regression.py:70)."""
    N = X.shape[-2]
    D = Y.shape[-1]
    K = kernels.K(kind, X, lengthscale, variance) + \
        np.eye(N, dtype=X.dtype)[None] * noise_var[..., None, :]
    if jitter > 0.:
        K = K + np.eye(N, dtype=X.dtype)[None] * jitter
    L = linalg.potrf(K)
    if mean is not None:
        Y = Y - mean
    LinvY = linalg.trsm(L, Y)
    logdet_l = linalg.sumlogdiag(np.abs(L))
    tmp = np.sum((np.square(LinvY) + np.log(2. * np.pi)).reshape(Y.shape[0], -1), axis=-1)
    logL = -logdet_l * D - tmp / 2
    return logL, L, LinvY


def gp_log_pdf_independent(kind, X, Y, noise_var, lengthscale, variance, jitter=0.0):
    """Independent formulation: sum over output columns of a dense
    multivariate-normal log-density (scipy).  Unbatched inputs."""
    from scipy.stats import multivariate_normal
    K = kernels.K_direct(kind, X[None], lengthscale[None], variance[None])[0]
    C = K + (float(noise_var[0]) + jitter) * np.eye(X.shape[0])
    return sum(multivariate_normal.logpdf(Y[:, d], mean=np.zeros(X.shape[0]), cov=C)
               for d in range(Y.shape[1]))


def gp_predict(kind, Xt, X_cond, L, LinvY, noise_var, lengthscale, variance,
               noise_free=True, diagonal_variance=True, mean=None):
    """gp_regression.py:158-196 (mean/variance prediction)."""
    Kxt = kernels.K(kind, X_cond, lengthscale, variance, Xt)
    LinvKxt = linalg.trsm(L, Kxt)
    mu = linalg.gemm2(LinvKxt, LinvY, True, False)
    if mean is not None:
        mu = mu + mean
    if diagonal_variance:
        var = kernels.Kdiag(Xt, variance) - np.sum(np.square(LinvKxt), axis=-2)
        if not noise_free:
            var = var + noise_var
    else:
        N = Xt.shape[-2]
        var = kernels.K(kind, Xt, lengthscale, variance) - linalg.syrk(LinvKxt, True)
        if not noise_free:
            var = var + np.eye(N, dtype=Xt.dtype)[None] * noise_var[..., None, :]
    return mu, var
