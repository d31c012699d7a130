"""Gaussian-process prior distribution (mxfusion/components/distributions/gp/gp.py:25-162): log-pdf and sampling of
Y ~ N(g(X), K(X, X)) through the CUDA primitives (kernel build, potrf, trsm, sumlogdiag; trmm as a GEMM)."""
import math

import torch

from ..distribution import Distribution
from ...variables.variable import Variable
from .... import ops

_LOG2PI = math.log(2.0 * math.pi)


class GaussianProcess(Distribution):
    def __init__(self, X, kernel, mean=None, rand_gen=None, dtype=None, ctx=None):
        inputs = [('X', X)] + [(k, v) for k, v in kernel.parameters.items()]
        self._has_mean = mean is not None
        if mean is not None:
            inputs.append(('mean', mean))
        super(GaussianProcess, self).__init__(inputs=inputs, outputs=None, input_names=[k for k, _ in inputs],
                                              output_names=['random_variable'], rand_gen=rand_gen, dtype=dtype,
                                              ctx=ctx)
        self.kernel = kernel

    @property
    def has_mean(self):
        return self._has_mean

    @staticmethod
    def define_variable(X, kernel, shape=None, mean=None, rand_gen=None, dtype=None, ctx=None):
        gp = GaussianProcess(X=X, kernel=kernel, mean=mean, rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        gp.set_outputs([Variable(value=None, shape=X.shape[:-1] + (1,) if shape is None else shape)])
        return gp.random_variable

    def log_pdf_impl(self, X, random_variable, F=None, **kernel_params):
        """gp.py:95-121."""
        mean = kernel_params.pop('mean', None) if self._has_mean else None
        D = random_variable.shape[-1]
        K = self.kernel.K(F, X, **kernel_params)
        L = ops.potrf(K)
        if mean is not None:
            random_variable = random_variable - mean
        LinvY = ops.trsm(L, random_variable)
        logdet_l = ops.sumlogdiag(L)                      # diag(L) > 0: F.abs is a no-op on a Cholesky factor
        return (-logdet_l * D - torch.sum(torch.square(LinvY) + _LOG2PI, dim=(-1, -2)) / 2) * self.log_pdf_scaling

    def draw_samples_impl(self, X, rv_shape, num_samples=1, F=None, **kernel_params):
        """gp.py:123-153: L die (+ mean), die ~ N(0, 1) of shape (num_samples,) + rv_shape."""
        mean = kernel_params.pop('mean', None) if self._has_mean else None
        K = self.kernel.K(F, X, **kernel_params)
        L = ops.potrf(K)
        out_shape = (num_samples,) + tuple(rv_shape)
        die = self._rand_gen.sample_normal(shape=out_shape, dtype=self.dtype, ctx=X.device)
        rv = ops.gemm2(L, die)                            # trmm: L is lower triangular with a zeroed upper part
        if mean is not None:
            rv = rv + mean
        return rv

    def replicate_self(self, attribute_map=None):
        rep = super(GaussianProcess, self).replicate_self(attribute_map)
        rep._has_mean = self._has_mean
        rep.kernel = self.kernel.replicate_self(attribute_map)
        return rep
