"""Linear kernel k(x, y) = sum_i s2_i x_i y_i (mxfusion/components/distributions/gp/kernels/linear.py:22-111).
The product is a tensor-core GEMM (ops.syrk / ops.gemm2 -> mxf_gemm); the per-dimension scaling is an elementwise
pass over X (N x D), not over the covariance matrix."""
import torch

from .kernel import NativeKernel
from ....variables.variable import Variable
from ....variables.var_trans import PositiveTransformation
from ..... import ops


class Linear(NativeKernel):
    broadcastable = True

    def __init__(self, input_dim, ARD=False, variances=1., name='linear', active_dims=None, dtype=None, ctx=None):
        super(Linear, self).__init__(input_dim=input_dim, name=name, active_dims=active_dims, dtype=dtype, ctx=ctx)
        self.ARD = ARD
        if not isinstance(variances, Variable):
            variances = Variable(shape=(input_dim if ARD else 1,), transformation=PositiveTransformation(),
                                 initial_value=variances)
        self.variances = variances

    def _compute_K(self, F, X, variances, X2=None):
        """linear.py:57-85."""
        if self.ARD:
            var_sqrt = torch.sqrt(variances).unsqueeze(-2)
            xsc = X * var_sqrt
            if X2 is None:
                return ops.syrk(xsc)
            return ops.gemm2(xsc, X2 * var_sqrt, False, True)
        A = ops.syrk(X) if X2 is None else ops.gemm2(X, X2, False, True)
        return A * variances.unsqueeze(-1)

    def _compute_Kdiag(self, F, X, variances):
        """linear.py:87-101."""
        return torch.sum(torch.square(X) * variances.unsqueeze(-2), dim=-1)
