"""MXNet ``linalg`` operator semantics restated in NumPy (test infrastructure).

MXNet's source is not under /root/reference; semantics follow the MXNet 1.x
operator documentation for ``mx.nd.linalg.*``.  All operators are batched over
every leading axis, exactly as the reference uses them with its sample axis
(X is (S,N,D), K is (S,N,N)).

Call sites in the reference: ``modules/gp_modules/svgp_regression.py:76-92``,
``modules/gp_modules/gp_regression.py:61-67``,
``components/distributions/gp/kernels/stationary.py:94,102``.
"""
import numpy as np
import scipy.linalg as sla


def _batched(fn, *arrays):
    lead = arrays[0].shape[:-2]
    flat = [a.reshape((-1,) + a.shape[-2:]) for a in arrays]
    out = [fn(*[f[i] for f in flat]) for i in range(flat[0].shape[0])]
    out = np.stack(out, axis=0)
    return out.reshape(lead + out.shape[1:])


def potrf(A):
    """linalg.potrf: lower Cholesky factor, upper triangle zero."""
    return _batched(lambda a: np.linalg.cholesky(a), A)


def trsm(A, B, transpose=False, rightside=False, lower=True, alpha=1.0):
    """linalg.trsm: alpha * op(A)^-1 B  (or B op(A)^-1 when rightside)."""
    def one(a, b):
        if rightside:
            # X op(A) = alpha B  <=>  op(A)^T X^T = alpha B^T
            xt = sla.solve_triangular(a, b.T, lower=lower, trans='N' if transpose else 'T')
            return alpha * xt.T
        return alpha * sla.solve_triangular(a, b, lower=lower, trans='T' if transpose else 'N')
    return _batched(one, A, B)


def trmm(A, B, transpose=False, rightside=False, lower=True, alpha=1.0):
    """linalg.trmm: alpha * op(A) B with A triangular."""
    tri = np.tril(A) if lower else np.triu(A)
    opA = np.swapaxes(tri, -1, -2) if transpose else tri
    return alpha * (B @ opA if rightside else opA @ B)


def gemm2(A, B, transpose_a=False, transpose_b=False, alpha=1.0):
    """linalg.gemm2: alpha * op(A) op(B)."""
    a = np.swapaxes(A, -1, -2) if transpose_a else A
    b = np.swapaxes(B, -1, -2) if transpose_b else B
    return alpha * (a @ b)


def syrk(A, transpose=False, alpha=1.0):
    """linalg.syrk: alpha * A A^T  (A^T A when transpose)."""
    At = np.swapaxes(A, -1, -2)
    return alpha * (At @ A if transpose else A @ At)


def sumlogdiag(A):
    """linalg.sumlogdiag: sum of log of the diagonal, per matrix."""
    return np.sum(np.log(np.diagonal(A, axis1=-2, axis2=-1)), axis=-1)


def make_diagonal(a):
    """(…, n) -> (…, n, n) diagonal embed.  util/customop.py:26-43."""
    n = a.shape[-1]
    return np.eye(n, dtype=a.dtype) * a[..., None]


def make_diagonal_backward(b_grad):
    """Gradient of make_diagonal: extract the diagonal.  util/customop.py:45-57."""
    return np.diagonal(b_grad, axis1=-2, axis2=-1).copy()
