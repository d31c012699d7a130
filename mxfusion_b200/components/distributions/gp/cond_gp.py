"""Conditional Gaussian process p(Y | X_c, Y_c, X) (mxfusion/components/distributions/gp/cond_gp.py:25-234).

Reference behaviour kept for parity (cond_gp.py:170): the whitened residual is summed over the OUTPUT axis before it
is squared (`F.sum(trsm(L, Y - mean), axis=-1)`), which differs from the product of per-output densities when the
output dimension is larger than one (SURVEY 8f rank 3 flags it)."""
import math

import torch

from ..distribution import Distribution
from ...variables.variable import Variable
from .... import ops

_LOG2PI = math.log(2.0 * math.pi)


class ConditionalGaussianProcess(Distribution):
    def __init__(self, X, X_cond, Y_cond, kernel, mean=None, mean_cond=None, rand_gen=None, dtype=None, ctx=None):
        if (mean is None) and (mean_cond is not None):
            raise ValueError("A mean function for the conditional values must be given together with `mean`.")
        inputs = [('X', X), ('X_cond', X_cond), ('Y_cond', Y_cond)] + [(k, v) for k, v in kernel.parameters.items()]
        self._has_mean = mean is not None
        self._has_mean_cond = mean_cond is not None
        if mean is not None:
            inputs.append(('mean', mean))
        if mean_cond is not None:
            inputs.append(('mean_cond', mean_cond))
        super(ConditionalGaussianProcess, self).__init__(inputs=inputs, outputs=None,
                                                         input_names=[k for k, _ in inputs],
                                                         output_names=['random_variable'], rand_gen=rand_gen,
                                                         dtype=dtype, ctx=ctx)
        self.kernel = kernel

    @staticmethod
    def define_variable(X, X_cond, Y_cond, kernel, shape=None, mean=None, mean_cond=None, rand_gen=None, dtype=None,
                        ctx=None):
        gp = ConditionalGaussianProcess(X=X, X_cond=X_cond, Y_cond=Y_cond, kernel=kernel, mean=mean,
                                        mean_cond=mean_cond, rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        gp.set_outputs([Variable(value=None, shape=X.shape[:-1] + (1,) if shape is None else shape)])
        return gp.random_variable

    def _moments(self, F, X, X_cond, Y_cond, kernel_params):
        """Shared by log-pdf and sampling (cond_gp.py:150-168 / :201-213): (L of the conditional covariance, mean)."""
        mean_cond = kernel_params.pop('mean_cond', None) if self._has_mean_cond else None
        K = self.kernel.K(F, X, **kernel_params)
        Kc = self.kernel.K(F, X_cond, X, **kernel_params)
        Kcc = self.kernel.K(F, X_cond, **kernel_params)
        Lcc = ops.potrf(Kcc)
        LccInvKc = ops.trsm(Lcc, Kc)
        cov = K - ops.syrk(LccInvKc, transpose=True)
        L = ops.potrf(cov)
        if mean_cond is not None:
            Y_cond = Y_cond - mean_cond
        LccInvY = ops.trsm(Lcc, Y_cond)
        rv_mean = ops.gemm2(LccInvKc, LccInvY, True, False)
        return L, rv_mean

    def log_pdf_impl(self, X, X_cond, Y_cond, random_variable, F=None, **kernel_params):
        """cond_gp.py:124-174."""
        mean = kernel_params.pop('mean', None) if self._has_mean else None
        D = random_variable.shape[-1]
        L, rv_mean = self._moments(F, X, X_cond, Y_cond, kernel_params)
        if mean is not None:
            random_variable = random_variable - mean
        LinvY = torch.sum(ops.trsm(L, random_variable - rv_mean), dim=-1)          # :170 (sum over outputs first)
        logdet_l = ops.sumlogdiag(L)
        return (-logdet_l * D - torch.sum(torch.square(LinvY) + _LOG2PI, dim=-1) / 2) * self.log_pdf_scaling

    def draw_samples_impl(self, X, X_cond, Y_cond, rv_shape, num_samples=1, F=None, **kernel_params):
        """cond_gp.py:176-223."""
        mean = kernel_params.pop('mean', None) if self._has_mean else None
        L, rv_mean = self._moments(F, X, X_cond, Y_cond, kernel_params)
        out_shape = (num_samples,) + tuple(rv_shape)
        die = self._rand_gen.sample_normal(shape=out_shape, dtype=self.dtype, ctx=X.device)
        rv = ops.gemm2(L, die) + rv_mean
        if mean is not None:
            rv = rv + mean
        return rv

    def replicate_self(self, attribute_map=None):
        rep = super(ConditionalGaussianProcess, self).replicate_self(attribute_map)
        rep._has_mean, rep._has_mean_cond = self._has_mean, self._has_mean_cond
        rep.kernel = self.kernel.replicate_self(attribute_map)
        return rep
