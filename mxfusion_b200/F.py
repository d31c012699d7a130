"""The `F` namespace handed to `InferenceAlgorithm.compute(F, variables)`.

In the reference `F` is ``mxnet.ndarray`` or ``mxnet.symbol``.  Here it exposes, under the MXNet
operator names the hot path uses, the differentiable CUDA operators of
mxfusion_b200.ops; shape plumbing (expand_dims, reshape, broadcast) goes to torch views."""
import torch

from . import ops as _ops


class linalg(object):
    potrf = staticmethod(_ops.potrf)
    trsm = staticmethod(_ops.trsm)
    gemm2 = staticmethod(_ops.gemm2)
    syrk = staticmethod(_ops.syrk)
    sumlogdiag = staticmethod(_ops.sumlogdiag)

    @staticmethod
    def trmm(A, B, transpose=False, rightside=False, alpha=1.0):
        """linalg.trmm: alpha * op(tril(A)) B (B op(tril(A)) when rightside)."""
        if rightside:
            return _ops.gemm2(B, A, False, transpose, alpha)
        return _ops.gemm2(A, B, transpose, False, alpha)


def expand_dims(a, axis):
    return a.unsqueeze(axis)


def reshape(a, shape):
    return a.reshape(shape)


def sum(a, axis=None):
    return torch.sum(a) if axis is None else torch.sum(a, dim=axis)


def mean(a, axis=None):
    return torch.mean(a) if axis is None else torch.mean(a, dim=axis)


square, log, exp, sqrt, abs = torch.square, torch.log, torch.exp, torch.sqrt, torch.abs
broadcast_add, broadcast_mul = torch.add, torch.mul
broadcast_minus = broadcast_sub = torch.sub
broadcast_div = torch.div


def eye(N, dtype=None, ctx=None):
    return torch.eye(N, dtype=dtype, device=ctx)


def Custom(x, op_type=None, **kw):
    """`F.Custom(x, op_type='make_diagonal')` (util/customop.py:22-81)."""
    if op_type == 'make_diagonal':
        return _ops.make_diagonal(x)
    raise NotImplementedError(op_type)
