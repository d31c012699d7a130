"""Stationary kernels (mxfusion/components/distributions/gp/kernels/stationary.py:22-132).

The reference builds K through syrk/gemm2 + five broadcast/elementwise passes + exp; here
`_compute_K` is ONE fused kernel launch (ops.kernel_matrix -> mxf_kbuild_fwd) and its adjoint one
more (mxf_kbuild_bwd)."""
import torch

from .kernel import NativeKernel
from ....variables.variable import Variable
from ....variables.var_trans import PositiveTransformation
from ..... import ops


class StationaryKernel(NativeKernel):
    KIND = None

    def __init__(self, input_dim, ARD=False, variance=1., lengthscale=1., name='stationary', active_dims=None,
                 dtype=None, ctx=None):
        super(StationaryKernel, self).__init__(input_dim=input_dim, name=name, active_dims=active_dims,
                                               dtype=dtype, ctx=ctx)
        self.ARD = ARD
        if not isinstance(variance, Variable):
            variance = Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=variance)
        if not isinstance(lengthscale, Variable):
            lengthscale = Variable(shape=(input_dim if ARD else 1,), transformation=PositiveTransformation(),
                                   initial_value=lengthscale)
        self.variance = variance
        self.lengthscale = lengthscale

    def _compute_K(self, F, X, lengthscale, variance, X2=None):
        """stationary.py:74-107 + the subclass's _compute_K, fused."""
        return ops.kernel_matrix(self.KIND, X, X2, lengthscale, variance)

    def _compute_Kdiag(self, F, X, lengthscale, variance):
        """stationary.py:109-124: zeros(S,N) + variance."""
        return torch.zeros(X.shape[:-1], dtype=X.dtype, device=X.device) + variance


class RBF(StationaryKernel):
    """rbf.py:21-72."""
    broadcastable = True
    KIND = ops.RBF

    def __init__(self, input_dim, ARD=False, variance=1., lengthscale=1., name='rbf', active_dims=None, dtype=None,
                 ctx=None):
        super(RBF, self).__init__(input_dim=input_dim, ARD=ARD, variance=variance, lengthscale=lengthscale,
                                  name=name, active_dims=active_dims, dtype=dtype, ctx=ctx)


class Matern(StationaryKernel):
    """matern.py:21-57."""
    broadcastable = True

    def __init__(self, input_dim, order, ARD=False, variance=1., lengthscale=1., name='matern', active_dims=None,
                 dtype=None, ctx=None):
        super(Matern, self).__init__(input_dim=input_dim, ARD=ARD, variance=variance, lengthscale=lengthscale,
                                     name=name, active_dims=active_dims, dtype=dtype, ctx=ctx)
        self.order = order


class Matern52(Matern):
    """matern.py:59-88."""
    KIND = ops.MATERN52

    def __init__(self, input_dim, ARD=False, variance=1., lengthscale=1., name='matern52', active_dims=None,
                 dtype=None, ctx=None):
        super(Matern52, self).__init__(input_dim=input_dim, order=2, ARD=ARD, variance=variance,
                                       lengthscale=lengthscale, name=name, active_dims=active_dims, dtype=dtype,
                                       ctx=ctx)


class Matern32(Matern):
    """matern.py:91-120."""
    KIND = ops.MATERN32

    def __init__(self, input_dim, ARD=False, variance=1., lengthscale=1., name='matern32', active_dims=None,
                 dtype=None, ctx=None):
        super(Matern32, self).__init__(input_dim=input_dim, order=1, ARD=ARD, variance=variance,
                                       lengthscale=lengthscale, name=name, active_dims=active_dims, dtype=dtype,
                                       ctx=ctx)


class Matern12(Matern):
    """matern.py:123-151."""
    KIND = ops.MATERN12

    def __init__(self, input_dim, ARD=False, variance=1., lengthscale=1., name='matern12', active_dims=None,
                 dtype=None, ctx=None):
        super(Matern12, self).__init__(input_dim=input_dim, order=0, ARD=ARD, variance=variance,
                                       lengthscale=lengthscale, name=name, active_dims=active_dims, dtype=dtype,
                                       ctx=ctx)
