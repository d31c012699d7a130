"""NDArray = a torch.Tensor subclass carrying the handful of MXNet-only methods the reference calls."""
import numpy as np
import torch

from ..context import cpu

_DT = {'float32': torch.float32, 'float64': torch.float64, 'int32': torch.int32, 'int64': torch.int64,
       np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64,
       np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64, np.dtype('float32'): torch.float32,
       np.dtype('float64'): torch.float64, None: None}


def _dt(d):
    if isinstance(d, torch.dtype):
        return d
    return _DT[d]


class NDArray(torch.Tensor):
    @staticmethod
    def __new__(cls, data, *a, **k):
        return torch.Tensor._make_subclass(cls, data, data.requires_grad)

    def asnumpy(self):
        return self.detach().cpu().numpy()

    def asscalar(self):
        return self.detach().reshape(-1)[0].item()

    @property
    def context(self):
        return cpu()

    def as_in_context(self, ctx):
        return self

    def copyto(self, other):
        return self.clone()

    def astype(self, dtype):
        return _wrap(torch.Tensor(self).to(_dt(dtype)))

    def reshape(self, *shape, **kw):
        if 'shape' in kw:
            shape = kw['shape']
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return _wrap(torch.Tensor.reshape(self, tuple(int(s) for s in shape)))

    def expand_dims(self, axis):
        return _wrap(torch.Tensor.unsqueeze(self, axis))

    @property
    def shape(self):
        return tuple(torch.Tensor.size(self))

    @property
    def dtype_np(self):
        return np.float64 if torch.Tensor(self).dtype == torch.float64 else np.float32

    def wait_to_read(self):
        pass

    def attach_grad(self):
        self.requires_grad_(True)


def _wrap(t):
    if isinstance(t, NDArray):
        return t
    if isinstance(t, torch.Tensor):
        return t.as_subclass(NDArray)
    return t


def _raw(t):
    return t


def array(a, dtype=None, ctx=None):
    if isinstance(a, torch.Tensor):
        t = a.detach().clone()
    else:
        t = torch.as_tensor(np.array(a))
    d = _dt(dtype) if dtype is not None else (torch.float32 if not t.is_floating_point() or True else t.dtype)
    if dtype is None:
        d = torch.float32          # mx.nd.array defaults to float32
    return _wrap(t.to(d))


def zeros(shape, dtype=None, ctx=None, **kw):
    if isinstance(shape, int):
        shape = (shape,)
    return _wrap(torch.zeros(tuple(shape), dtype=_dt(dtype) or torch.float32))


def ones(shape, dtype=None, ctx=None, **kw):
    if isinstance(shape, int):
        shape = (shape,)
    return _wrap(torch.ones(tuple(shape), dtype=_dt(dtype) or torch.float32))


def zeros_like(a):
    return _wrap(torch.zeros_like(a))


def ones_like(a):
    return _wrap(torch.ones_like(a))


def eye(N, M=0, k=0, dtype=None, ctx=None):
    return _wrap(torch.eye(int(N), dtype=_dt(dtype) or torch.float32))


def arange(start, stop=None, step=1.0, dtype=None, ctx=None):
    return _wrap(torch.arange(start, stop, step).to(_dt(dtype) or torch.float32))


def _axis(axis):
    if axis is None or axis == ():
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else int(axis)


def sum(a, axis=None, keepdims=False, **kw):
    ax = _axis(axis)
    return _wrap(torch.sum(a) if ax is None else torch.sum(a, dim=ax, keepdim=keepdims))


def mean(a, axis=None, keepdims=False, **kw):
    ax = _axis(axis)
    return _wrap(torch.mean(a) if ax is None else torch.mean(a, dim=ax, keepdim=keepdims))


def prod(a, axis=None, keepdims=False):
    return _wrap(torch.prod(a) if axis is None else torch.prod(a, dim=int(axis), keepdim=keepdims))


def expand_dims(a, axis):
    return _wrap(torch.unsqueeze(a, axis))


def _mx_reshape_shape(src, shape):
    """MXNet reshape special codes 0 (copy), -1 (infer); the reference uses only those."""
    out = []
    for i, s in enumerate(shape):
        out.append(src[i] if s == 0 else int(s))
    return tuple(out)


def reshape(a, shape=None, **kw):
    return _wrap(torch.reshape(a, _mx_reshape_shape(tuple(a.shape), tuple(shape))))


def square(a):
    return _wrap(torch.square(a))


def sqrt(a):
    return _wrap(torch.sqrt(a))


def log(a):
    return _wrap(torch.log(a))


def exp(a):
    return _wrap(torch.exp(a))


def expm1(a):
    return _wrap(torch.expm1(a))


def abs(a):
    return _wrap(torch.abs(a))


def sign(a):
    return _wrap(torch.sign(a))


def gammaln(a):
    return _wrap(torch.lgamma(a))


def clip(a, a_min, a_max):
    return _wrap(torch.clamp(a, min=a_min, max=None if a_max == np.inf else a_max))


def stop_gradient(a):
    return _wrap(a.detach())


def transpose(a, axes=None):
    return _wrap(a.permute(*axes) if axes else a.t())


def dot(a, b):
    return _wrap(torch.matmul(a, b))


def concat(*arrays, dim=1):
    return _wrap(torch.cat(list(arrays), dim=dim))


def broadcast_to(a, shape, out=None):
    shape = tuple(a.shape[i] if s == 0 else s for i, s in enumerate(shape))
    r = a.expand(shape)
    if out is not None:
        with torch.no_grad():
            torch.Tensor(out).copy_(torch.Tensor(r))
        return out
    return _wrap(r)


def one_hot(indices, depth, dtype=None, **kw):
    return _wrap(torch.nn.functional.one_hot(torch.Tensor(indices).long(), int(depth)).to(_dt(dtype) or torch.float32))


def broadcast_axis(a, axis, size):
    shape = list(a.shape)
    axes = axis if isinstance(axis, (tuple, list)) else [axis]
    sizes = size if isinstance(size, (tuple, list)) else [size]
    for ax, sz in zip(axes, sizes):
        shape[ax] = sz
    return _wrap(a.expand(tuple(shape)))


def broadcast_add(a, b):
    return _wrap(torch.add(a, b))


def broadcast_sub(a, b):
    return _wrap(torch.sub(a, b))


broadcast_minus = broadcast_sub


def broadcast_mul(a, b):
    return _wrap(torch.mul(a, b))


def broadcast_div(a, b):
    return _wrap(torch.div(a, b))


def broadcast_power(a, b):
    return _wrap(torch.pow(a, b))


add, subtract, multiply, divide, power = broadcast_add, broadcast_sub, broadcast_mul, broadcast_div, broadcast_power


def Activation(a, act_type=None):
    if act_type == 'softrelu':
        return _wrap(torch.nn.functional.softplus(a))
    if act_type == 'sigmoid':
        return _wrap(torch.sigmoid(a))
    if act_type == 'tanh':
        return _wrap(torch.tanh(a))
    if act_type == 'relu':
        return _wrap(torch.relu(a))
    raise NotImplementedError(act_type)


def Custom(x, op_type=None, **kwargs):
    from ..operator import custom
    return custom(x, op_type=op_type, **kwargs)


class _Linalg(object):
    """MXNet 1.x linalg operator semantics (see oracle/linalg.py for the same restatement in NumPy)."""

    @staticmethod
    def potrf(A):
        return _wrap(torch.linalg.cholesky(A))

    @staticmethod
    def trsm(A, B, transpose=False, rightside=False, lower=True, alpha=1.0):
        T = torch.tril(A) if lower else torch.triu(A)
        if rightside:
            # X op(A) = alpha B
            opA = T.transpose(-1, -2) if transpose else T
            X = torch.linalg.solve_triangular(opA, B, upper=(not lower) != transpose, left=False)
        else:
            opA = T.transpose(-1, -2) if transpose else T
            X = torch.linalg.solve_triangular(opA, B, upper=(not lower) != transpose, left=True)
        return _wrap(alpha * X)

    @staticmethod
    def trmm(A, B, transpose=False, rightside=False, lower=True, alpha=1.0):
        T = torch.tril(A) if lower else torch.triu(A)
        opA = T.transpose(-1, -2) if transpose else T
        return _wrap(alpha * (torch.matmul(B, opA) if rightside else torch.matmul(opA, B)))

    @staticmethod
    def gemm2(A, B, transpose_a=False, transpose_b=False, alpha=1.0):
        a = A.transpose(-1, -2) if transpose_a else A
        b = B.transpose(-1, -2) if transpose_b else B
        return _wrap(alpha * torch.matmul(a, b))

    @staticmethod
    def gemm(A, B, C, transpose_a=False, transpose_b=False, alpha=1.0, beta=1.0):
        return _wrap(_Linalg.gemm2(A, B, transpose_a, transpose_b, alpha) + beta * C)

    @staticmethod
    def syrk(A, transpose=False, alpha=1.0):
        At = A.transpose(-1, -2)
        return _wrap(alpha * (torch.matmul(At, A) if transpose else torch.matmul(A, At)))

    @staticmethod
    def sumlogdiag(A):
        return _wrap(torch.sum(torch.log(torch.diagonal(A, dim1=-2, dim2=-1)), dim=-1))

    @staticmethod
    def potri(A):
        L = torch.tril(A)
        return _wrap(torch.cholesky_inverse(L))


linalg = _Linalg()


class _Random(object):
    @staticmethod
    def normal(loc=0, scale=1, shape=None, dtype=None, ctx=None, out=None):
        return _wrap(torch.randn(tuple(shape), dtype=_dt(dtype) or torch.float32) * scale + loc)

    @staticmethod
    def uniform(low=0, high=1, shape=None, dtype=None, ctx=None, out=None):
        return _wrap(torch.rand(tuple(shape), dtype=_dt(dtype) or torch.float32) * (high - low) + low)


random = _Random()
