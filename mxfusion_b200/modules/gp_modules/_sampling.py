"""Default draw-samples algorithms of the GP modules.

* `GPRegressionSampling` restates gp_regression.py:78-135: Y = chol(K(X, X) + noise_var I) eps (+ mean).
* `InducingGPSampling` is what the reference gets by attaching `ForwardSamplingAlgorithm` to the inner graph of the
  sparse modules (svgp_regression.py:353-381, 398-403; sparsegp_regression.py:340-377): the topological walk
  U ~ GP(Z) (gp.py:123-153) -> F ~ GP(X | Z, U) (cond_gp.py:176-223) -> Y ~ N(F, noise_var) (normal.py:72-92), one
  `sample_normal` call per variable in that order.  The inner graphs of this repo do not hold U and F as nodes (the
  bounds never evaluate them), so the walk is written out here on the same distribution code (`draw_samples_impl` of
  GaussianProcess / ConditionalGaussianProcess / Normal)."""
import torch

from ...inference.inference_alg import SamplingAlgorithm
from ...components.variables.runtime_variable import arrays_as_samples
from ...components.distributions.random_gen import MXNetRandomGenerator
from ...components.distributions.normal import Normal
from ...components.distributions.gp import GaussianProcess, ConditionalGaussianProcess
from ...components.variables.variable import Variable
from ... import ops


def _gen(alg):
    return alg._rand_gen if alg._rand_gen is not None else MXNetRandomGenerator


def _y_shape(model, X):
    """(N, P): the modules define Y with the rows of X and a fixed output dimension (gp_regression.py:320-327)."""
    return (X.shape[-2], int(model.Y.shape[-1]))


class GPRegressionSampling(SamplingAlgorithm):
    def __init__(self, model, observed, num_samples=1, target_variables=None, rand_gen=None, dtype=None):
        super(GPRegressionSampling, self).__init__(model=model, observed=observed, num_samples=num_samples,
                                                   target_variables=target_variables)
        self._rand_gen = rand_gen
        self._dtype = dtype

    def compute(self, F, variables):
        X = variables[self.model.X]
        noise_var = variables[self.model.noise_var]
        kern = self.model.kernel
        kern_params = kern.fetch_parameters(variables)
        X, noise_var, kern_params = arrays_as_samples(F, [X, noise_var, kern_params])
        N = X.shape[-2]
        K = kern.K(F, X, **kern_params) + \
            torch.eye(N, dtype=X.dtype, device=X.device).unsqueeze(0) * noise_var.unsqueeze(-2)      # :112-114
        L = ops.potrf(K)
        Y_shape = _y_shape(self.model, X)
        out_shape = (self.num_samples,) + tuple(Y_shape)
        die = _gen(self).sample_normal(shape=out_shape, dtype=self._dtype, ctx=X.device)
        y = ops.gemm2(L.expand((self.num_samples,) + tuple(L.shape[1:])), die)                     # trmm(L, die)
        if getattr(self.model, 'has_mean', False):
            y = y + variables[self.model.mean]
        samples = {self.model.Y.uuid: y}
        if self.target_variables:
            return tuple(samples[v] for v in self.target_variables)
        return samples


class InducingGPSampling(SamplingAlgorithm):
    def __init__(self, model, observed, num_samples=1, target_variables=None, rand_gen=None, dtype=None):
        super(InducingGPSampling, self).__init__(model=model, observed=observed, num_samples=num_samples,
                                                 target_variables=target_variables)
        self._rand_gen = rand_gen
        self._dtype = dtype

    def compute(self, F, variables):
        m = self.model
        X, Z = variables[m.X], variables[m.inducing_inputs]
        noise_var = variables[m.noise_var]
        kern = m.kernel
        kern_params = kern.fetch_parameters(variables)
        X, Z, noise_var, kern_params = arrays_as_samples(F, [X, Z, noise_var, kern_params])
        Y_shape = _y_shape(m, X)
        S, gen, dt = self.num_samples, _gen(self), self._dtype
        # the three factors of the reference's inner graph, stand-alone (only their draw_samples_impl is used)
        gp_u = GaussianProcess(X=Variable(shape=Z.shape[1:]), kernel=kern, rand_gen=gen, dtype=dt)
        U = gp_u.draw_samples_impl(X=Z, rv_shape=(Z.shape[-2], Y_shape[-1]), num_samples=S, F=F, **kern_params)
        extra = {}
        has_mean = getattr(m, 'has_mean', False)
        cgp = ConditionalGaussianProcess(X=Variable(shape=X.shape[1:]), X_cond=Variable(shape=Z.shape[1:]),
                                         Y_cond=Variable(shape=U.shape[1:]), kernel=kern,
                                         mean=Variable(shape=Y_shape) if has_mean else None, rand_gen=gen, dtype=dt)
        if has_mean:
            extra['mean'] = variables[m.mean]
        Xs, Zs, kps = arrays_as_samples(F, [X, Z, kern_params])
        Xs, Zs, U, kps = _match_samples([Xs, Zs, U, kps], S)
        Fv = cgp.draw_samples_impl(X=Xs, X_cond=Zs, Y_cond=U, rv_shape=Y_shape, num_samples=S, F=F, **kps, **extra)
        normal = Normal(mean=Variable(shape=Y_shape), variance=Variable(shape=Y_shape), rand_gen=gen, dtype=dt,
                        ctx=X.device)
        nv = noise_var.reshape((noise_var.shape[0],) + (1,) * (len(Y_shape) - 1) + (-1,))
        Y = normal.draw_samples_impl(mean=Fv, variance=nv.expand((nv.shape[0],) + Y_shape), rv_shape=Y_shape,
                                     num_samples=S, F=F)
        samples = {m.Y.uuid: Y}
        if self.target_variables:
            return tuple(samples[v] for v in self.target_variables)
        return samples


def _match_samples(items, S):
    """Broadcast leading sample axes of size 1 to S (U carries S samples once it has been drawn)."""
    def ex(t):
        if isinstance(t, dict):
            return {k: ex(v) for k, v in t.items()}
        return t.expand((S,) + tuple(t.shape[1:])) if t.shape[0] == 1 and S > 1 else t
    return [ex(t) for t in items]
