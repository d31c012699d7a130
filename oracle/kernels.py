"""Stationary covariance functions restated in NumPy (test infrastructure).

Follows ``mxfusion/components/distributions/gp/kernels/stationary.py:74-124``,
``rbf.py:54-72`` and ``matern.py:67-151`` operation for operation, including
the expanded form  r2 = |a|^2 + |b|^2 - 2 a.b  (so r2 may come out slightly
negative / non-zero on the diagonal), the ``clip(r2, 1e-14, inf)`` before the
square root in the Matern family, and Matern-5/2's use of the *unclipped* r2 in
its 5/3 r2 term (``matern.py:84-87``).

Layout: X (S,N,D), X2 (S,N2,D), lengthscale (S,1) or (S,D), variance (S,1)
-> K (S,N,N2); Kdiag (S,N).
"""
import numpy as np

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
KIND_BY_NAME = {'rbf': RBF, 'matern12': MATERN12, 'matern32': MATERN32, 'matern52': MATERN52}


def r2(X, lengthscale, X2=None):
    """stationary.py:90-107."""
    ls = lengthscale[..., None, :]
    if X2 is None:
        xsc = X / ls
        amat = (xsc @ np.swapaxes(xsc, -1, -2)) * -2
        dg = np.sum(np.square(xsc), axis=-1)
        amat = amat + dg[..., :, None]
        amat = amat + dg[..., None, :]
    else:
        x1 = X / ls
        x2 = X2 / ls
        amat = (x1 @ np.swapaxes(x2, -1, -2)) * -2
        amat = amat + np.sum(np.square(x1), axis=-1, keepdims=True)
        amat = amat + np.sum(np.square(x2), axis=-1)[..., None, :]
    return amat


def K(kind, X, lengthscale, variance, X2=None):
    """rbf.py:71-72, matern.py:84-88, 116-120, 148-151."""
    R2 = r2(X, lengthscale, X2)
    var = variance[..., None]  # (S,1,1)
    if kind == RBF:
        return np.exp(R2 / -2) * var
    R = np.sqrt(np.clip(R2, 1e-14, np.inf))
    if kind == MATERN52:
        return (1 + np.sqrt(5) * R + 5 / 3. * R2) * np.exp(-np.sqrt(5) * R) * var
    if kind == MATERN32:
        return (1 + np.sqrt(3) * R) * np.exp(-np.sqrt(3) * R) * var
    if kind == MATERN12:
        return np.exp(-R) * var
    raise ValueError(kind)


def Kdiag(X, variance):
    """stationary.py:123-124: zeros(S,N) + variance."""
    return np.zeros(X.shape[:-1], dtype=X.dtype) + variance


def K_direct(kind, X, lengthscale, variance, X2=None):
    """Independent formulation (explicit differences, closed-form Matern in r)
    used only to cross-check ``K``; not a restatement of the reference."""
    X2 = X if X2 is None else X2
    ls = lengthscale[..., None, None, :]
    diff = (X[..., :, None, :] - X2[..., None, :, :]) / ls
    r = np.sqrt(np.sum(diff * diff, axis=-1))
    var = variance[..., None]
    if kind == RBF:
        return var * np.exp(-0.5 * r * r)
    if kind == MATERN52:
        return var * (1 + np.sqrt(5) * r + 5. / 3. * r * r) * np.exp(-np.sqrt(5) * r)
    if kind == MATERN32:
        return var * (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r)
    return var * np.exp(-r)
