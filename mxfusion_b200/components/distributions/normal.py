"""Univariate Normal (mxfusion/components/distributions/normal.py:26-116), the distribution behind the
mean-field MC-ELBO (BASELINE config 4)."""
import math

import torch

from .distribution import Distribution
from ..variables.variable import Variable
from ..variables.runtime_variable import arrays_as_samples
from ... import ops


def _to_rv_shape(t, rv_shape):
    """Broadcast everything after the sample axis to the random variable's shape (NumPy rules: trailing dimensions are
    aligned, e.g. a (S, 1) mean against a (N, 1) variable)."""
    rv_shape = tuple(rv_shape)
    if tuple(t.shape[1:]) == rv_shape:
        return t
    t = t.reshape((t.shape[0],) + (1,) * (len(rv_shape) - (t.dim() - 1)) + tuple(t.shape[1:]))
    return t.expand((t.shape[0],) + rv_shape)


class Normal(Distribution):
    def __init__(self, mean, variance, rand_gen=None, dtype=None, ctx=None):
        inputs = [('mean', mean), ('variance', variance)]
        super(Normal, self).__init__(inputs=inputs, outputs=None, input_names=['mean', 'variance'],
                                     output_names=['random_variable'], rand_gen=rand_gen, dtype=dtype, ctx=ctx)

    def log_pdf_impl(self, mean, variance, random_variable, F=None):
        """normal.py:52-70, elementwise (used when the per-element values are needed)."""
        logvar = math.log(2 * math.pi) / -2 + torch.log(variance) / -2
        return (logvar + torch.square(random_variable - mean) / (-2 * variance)) * self.log_pdf_scaling

    def log_pdf_sum(self, F, variables):
        """normal.py:67-69 + factor_graph.py:223 fused: one pass, one scalar."""
        kw = self.fetch_runtime_inputs(variables)
        kw.update(self.fetch_runtime_outputs(variables))
        x, m, v = kw['random_variable'], kw['mean'], kw['variance']
        n = max(x[0].numel(), m[0].numel(), v[0].numel())
        if not (x[0].numel() == m[0].numel() == v[0].numel() == n):
            shape = tuple(torch.broadcast_shapes(x.shape[1:], m.shape[1:], v.shape[1:]))

            def bc(t):      # broadcast everything after the sample axis (NumPy rules: align trailing dimensions)
                t = t.reshape((t.shape[0],) + (1,) * (len(shape) - (t.dim() - 1)) + tuple(t.shape[1:]))
                return t.expand((t.shape[0],) + shape)
            x, m, v = bc(x), bc(m), bc(v)
        return ops.normal_log_pdf_sum(x, m, v, self.log_pdf_scaling).reshape(())

    def log_pdf_operands(self, F, variables):
        """(x, mean, variance, scale) for the batched evaluation of all Normal factors of a graph walk
        (FactorGraph.log_pdf -> ops.normal_log_pdf_sum_multi), or None when this factor needs the general path: every
        operand must be full-size or a single element per sample, and a single-element operand must not need a gradient."""
        kw = self.fetch_runtime_inputs(variables)
        kw.update(self.fetch_runtime_outputs(variables))
        x, m, v = kw['random_variable'], kw['mean'], kw['variance']
        n = max(x[0].numel(), m[0].numel(), v[0].numel())
        S = max(x.shape[0], m.shape[0], v.shape[0])
        for t in (x, m, v):
            k = t[0].numel()
            if t.shape[0] not in (1, S) or (k != n and (k != 1 or t.requires_grad)):
                return None
        return x, m, v, self.log_pdf_scaling

    def draw_operands(self, F, variables, num_samples):
        """(mean, variance, num_samples, (seed, offset)) for the batched draw of independent Normal factors
        (FactorGraph.draw_samples -> ops.normal_draw_multi), or None when the generator injects fixed noise."""
        gen = self._rand_gen
        if not getattr(gen, 'in_kernel', False):
            return None
        from ..variables.runtime_variable import arrays_as_samples
        kw = arrays_as_samples(F, self.fetch_runtime_inputs(variables))
        mean, variance = kw['mean'], kw['variance']
        rv_shape = self._realized_shape(variables)
        return _to_rv_shape(mean, rv_shape), _to_rv_shape(variance, rv_shape), num_samples, gen.next_stream()

    def draw_samples_impl(self, mean, variance, rv_shape, num_samples=1, F=None):
        """normal.py:72-92: eps * sqrt(variance) + mean with eps ~ N(0,1) of shape (S,) + rv_shape."""
        full = (num_samples,) + tuple(rv_shape)
        mean, variance = _to_rv_shape(mean, rv_shape), _to_rv_shape(variance, rv_shape)
        gen = self._rand_gen
        if getattr(gen, 'in_kernel', False):
            from .random_gen import step_counter
            seed, offset = gen.next_stream()
            return ops.normal_draw(mean, variance, num_samples, seed=seed, offset=offset,
                                   step_counter=step_counter(mean.device))
        eps = gen.sample_normal(shape=full, dtype=self.dtype, ctx=self.ctx)
        return ops.normal_draw(mean, variance, num_samples, eps=eps)

    @staticmethod
    def define_variable(mean=0., variance=1., shape=None, rand_gen=None, dtype=None, ctx=None):
        normal = Normal(mean=mean, variance=variance, rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        normal._generate_outputs(shape=shape)
        return normal.random_variable

    def _generate_outputs(self, shape):
        self.set_outputs([Variable(value=None, shape=shape if shape is not None else (1,))])
        # Variable(value=None) made a plain parameter node; attaching it as our output makes it a RANDVAR
