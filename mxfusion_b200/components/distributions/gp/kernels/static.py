"""Bias and White kernels (mxfusion/components/distributions/gp/kernels/static.py:22-164): constant / diagonal
covariances, pure broadcasts of the variance parameter (no arithmetic over N x N beyond the fill)."""
import torch

from .kernel import NativeKernel
from ....variables.variable import Variable
from ....variables.var_trans import PositiveTransformation


class _Static(NativeKernel):
    broadcastable = True

    def __init__(self, input_dim, variance=1., name='static', active_dims=None, dtype=None, ctx=None):
        super(_Static, self).__init__(input_dim=input_dim, name=name, active_dims=active_dims, dtype=dtype, ctx=ctx)
        if not isinstance(variance, Variable):
            variance = Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=variance)
        self.variance = variance

    def _compute_Kdiag(self, F, X, variance):
        """static.py:72-86 / :149-164: the variance broadcast to (S, N)."""
        return variance.expand(variance.shape[0], X.shape[-2]) + torch.zeros(X.shape[:-1], dtype=X.dtype, device=X.device)


class Bias(_Static):
    """k(x, y) = s2 (static.py:22-86)."""

    def __init__(self, input_dim, variance=1., name='bias', active_dims=None, dtype=None, ctx=None):
        super(Bias, self).__init__(input_dim=input_dim, variance=variance, name=name, active_dims=active_dims,
                                   dtype=dtype, ctx=ctx)

    def _compute_K(self, F, X, variance, X2=None):
        """static.py:51-70."""
        if X2 is None:
            X2 = X
        S = max(X.shape[0], X2.shape[0], variance.shape[0])
        return variance.unsqueeze(-1) + torch.zeros((S, X.shape[-2], X2.shape[-2]), dtype=X.dtype, device=X.device)


class White(_Static):
    """K = s2 I for X2 = None, zeros otherwise (static.py:89-164)."""

    def __init__(self, input_dim, variance=1., name='white', active_dims=None, dtype=None, ctx=None):
        super(White, self).__init__(input_dim=input_dim, variance=variance, name=name, active_dims=active_dims,
                                    dtype=dtype, ctx=ctx)

    def _compute_K(self, F, X, variance, X2=None):
        """static.py:118-147."""
        N = X.shape[-2]
        if X2 is None:
            eye = torch.eye(N, dtype=X.dtype, device=X.device).unsqueeze(0)
            return eye * variance.unsqueeze(-1)
        S = max(X.shape[0], X2.shape[0])
        return torch.zeros((S, N, X2.shape[-2]), dtype=X.dtype, device=X.device)
