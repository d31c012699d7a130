"""Stochastic variational sparse GP regression module (Hensman et al. 2013)
(mxfusion/modules/gp_modules/svgp_regression.py:30-457)."""
import numpy as np
import torch

from ..module import Module
from ...models import Model, Posterior
from ...components.variables.variable import Variable
from ...components.variables.var_trans import PositiveTransformation
from ...inference.variational import VariationalInference
from ...inference.inference_alg import SamplingAlgorithm
from ... import ops


def _active(F, kern, *arrays):
    if kern.active_dims is None:
        return arrays
    from ...components.distributions.gp.kernels.kernel import slice_axis
    return tuple(slice_axis(F, a, -1, kern.active_dims) for a in arrays)


class SVGPRegressionLogPdf(VariationalInference):
    """The SVGP evidence lower bound with the analytic KL(q(u)||p(u)) (svgp_regression.py:30-109), as ONE
    fused operator (ops.svgp_log_pdf): kernel build, syrk + make_diagonal, two potrf, the triangular
    solves and every reduction are CUDA launches of libmxf_b200.so, and the gradient is analytic."""

    def __init__(self, model, posterior, observed, jitter=0.):
        super(SVGPRegressionLogPdf, self).__init__(model=model, posterior=posterior, observed=observed)
        self.log_pdf_scaling = 1
        self.jitter = jitter

    def compute(self, F, variables):
        X = variables[self.model.X]
        Y = variables[self.model.Y]
        Z = variables[self.model.inducing_inputs]
        noise_var = variables[self.model.noise_var]
        mu = variables[self.posterior.qU_mean]
        S_W = variables[self.posterior.qU_cov_W]
        S_diag = variables[self.posterior.qU_cov_diag]
        kern = self.model.kernel
        mean = variables[self.model.mean] if self.model.has_mean else None
        if getattr(kern, 'KIND', None) is None or noise_var.dim() == 3:     # combination kernels, or per-point noise
            # (svgp_regression.py:61-67: noise_var of shape (N, 1|P)): primitive-by-primitive
            from . import _generic
            return _generic.svgp_log_pdf(F, kern, kern.fetch_parameters(variables), X, Y, Z, noise_var, mu, S_W,
                                         S_diag, self.jitter, self.log_pdf_scaling, mean=mean)
        kp = kern._strip(kern.fetch_parameters(variables))
        X, Z = _active(F, kern, X, Z)
        return ops.svgp_log_pdf(kern.KIND, X, Y, Z, noise_var, mu, S_W, S_diag, kp['lengthscale'], kp['variance'],
                                jitter=self.jitter, log_pdf_scaling=self.log_pdf_scaling, mean=mean)


class SVGPRegressionMeanVariancePrediction(SamplingAlgorithm):
    """svgp_regression.py:112-189."""

    def __init__(self, model, posterior, observed, noise_free=True, diagonal_variance=True, jitter=0.):
        super(SVGPRegressionMeanVariancePrediction, self).__init__(model=model, observed=observed,
                                                                   extra_graphs=[posterior])
        self.jitter = jitter
        self.noise_free = noise_free
        self.diagonal_variance = diagonal_variance

    def compute(self, F, variables):
        outcomes = {self.model.Y.uuid: _svgp_predict_moments(self, F, variables)}
        if self.target_variables:
            return tuple(outcomes[v] for v in self.target_variables)
        return outcomes


def _svgp_predict_moments(self, F, variables, squeeze_var=False):
    """Predictive mean and (diagonal or full) covariance (svgp_regression.py:145-182 / :216-266)."""
    X = variables[self.model.X]
    N = X.shape[-2]
    Z = variables[self.model.inducing_inputs]
    noise_var = variables[self.model.noise_var]
    mu = variables[self.graphs[1].qU_mean]
    S_W = variables[self.graphs[1].qU_cov_W]
    S_diag = variables[self.graphs[1].qU_cov_diag]
    kern = self.model.kernel
    kern_params = kern.fetch_parameters(variables)
    kp = kern._strip(kern_params) if getattr(kern, 'KIND', None) is not None else None
    S = ops.syrk(S_W) + ops.make_diagonal(S_diag)
    if getattr(kern, 'KIND', None) is None:
        Kuu = kern.K(F, Z, **kern_params)
        if self.jitter > 0.:
            Kuu = Kuu + torch.eye(Z.shape[-2], dtype=Z.dtype, device=Z.device).unsqueeze(0) * self.jitter
    else:
        (Zs,) = _active(F, kern, Z)
        Kuu = ops.kernel_matrix(kern.KIND, Zs, None, kp['lengthscale'], kp['variance'], diag_const=self.jitter)
    L = ops.potrf(Kuu)
    Ls = ops.potrf(S)
    LinvLs = ops.trsm(L, Ls)
    Linvmu = ops.trsm(L, mu)
    LinvSLinvT = ops.syrk(LinvLs)
    wv = ops.trsm(L, Linvmu, transpose=True)
    Kxt = kern.K(F, Z, X, **kern_params)
    mean_f = ops.gemm2(Kxt, wv, True, False)
    if self.model.has_mean:
        mean_f = mean_f + variables[self.model.mean]
    LinvKxt = ops.trsm(L, Kxt)
    tmp = ops.gemm2(LinvSLinvT, LinvKxt)
    if self.diagonal_variance:
        var = kern.Kdiag(F, X, **kern_params) - torch.sum(torch.square(LinvKxt), dim=-2) + \
            torch.sum(tmp * LinvKxt, dim=-2)
        var = var.unsqueeze(-1)
        if not self.noise_free:
            var = var + noise_var
    else:
        var = kern.K(F, X, **kern_params) - ops.syrk(LinvKxt, True) + ops.gemm2(LinvKxt, tmp, True, False)
        var = var.unsqueeze(-1)
        if not self.noise_free:
            var = var + torch.eye(N, dtype=X.dtype, device=X.device).reshape(1, N, N, 1) * \
                noise_var.unsqueeze(-2)
    if squeeze_var:
        var = var.squeeze(-1)
    return mean_f, var


class SVGPRegressionSamplingPrediction(SamplingAlgorithm):
    """svgp_regression.py:192-280: draws from the predictive distribution."""

    def __init__(self, model, posterior, observed, rand_gen=None, noise_free=True, diagonal_variance=True, jitter=0.):
        super(SVGPRegressionSamplingPrediction, self).__init__(model=model, observed=observed,
                                                               extra_graphs=[posterior])
        self.noise_free = noise_free
        self.diagonal_variance = diagonal_variance
        self._rand_gen = rand_gen
        self.jitter = jitter

    def compute(self, F, variables):
        from .gp_regression import _draw_from_moments
        mu, var = _svgp_predict_moments(self, F, variables, squeeze_var=True)
        jitter, self.jitter = self.jitter, 0.          # :232-234 use the jitter on Kuu only; :267 adds none to cov
        try:
            samples = _draw_from_moments(self, mu, var)
        finally:
            self.jitter = jitter
        outcomes = {self.model.Y.uuid: samples}
        if self.target_variables:
            return tuple(outcomes[v] for v in self.target_variables)
        return outcomes


class SVGPRegression(Module):
    """`m.Y = SVGPRegression.define_variable(X=m.X, kernel=k, noise_var=m.noise_var, shape=(m.N, 1),
    num_inducing=M)`; hidden parameters: inducing_inputs (M,D), qU_mean (M,P), qU_cov_W (M,M),
    qU_cov_diag (M,) softplus (svgp_regression.py:349-381)."""

    def __init__(self, X, kernel, noise_var, inducing_inputs=None, num_inducing=10, mean=None, rand_gen=None,
                 dtype=None, ctx=None):
        if not isinstance(X, Variable):
            X = Variable(value=X)
        if not isinstance(noise_var, Variable):
            noise_var = Variable(value=noise_var)
        if inducing_inputs is None:
            inducing_inputs = Variable(shape=(num_inducing, kernel.input_dim),
                                       initial_value=np.random.randn(num_inducing, kernel.input_dim))   # :318-320
        inputs = [('X', X), ('inducing_inputs', inducing_inputs), ('noise_var', noise_var)]
        self._has_mean = mean is not None
        if mean is not None:
            inputs.append(('mean', mean))
        super(SVGPRegression, self).__init__(inputs=inputs, outputs=None, input_names=[k for k, _ in inputs],
                                             output_names=['random_variable'], rand_gen=rand_gen, dtype=dtype,
                                             ctx=ctx)
        self.kernel = kernel

    def _generate_outputs(self, output_shapes=None):
        shape = output_shapes['random_variable']
        if shape is None:
            shape = self.X.shape[:-1] + (1,)
        self.set_outputs([Variable(shape=shape)])

    def _build_module_graphs(self):
        Y = self.random_variable
        graph = Model(name='sparsegp_regression')
        graph.X = self.X.replicate_self()
        graph.inducing_inputs = self.inducing_inputs.replicate_self()
        M = self.inducing_inputs.shape[0]
        graph.noise_var = self.noise_var.replicate_self()
        graph.__dict__['has_mean'] = self._has_mean
        if self._has_mean:
            graph.mean = self.mean.replicate_self()
        graph.Y = Y.replicate_self()
        graph.__dict__['kernel'] = self.kernel
        for name, var in self.kernel.parameters.items():
            graph.add_component(var, name)
        post = Posterior(graph)
        post.qU_cov_diag = Variable(shape=(M,), transformation=PositiveTransformation())
        post.qU_cov_W = Variable(shape=(M, M))
        post.qU_mean = Variable(shape=(M, Y.shape[-1]))
        return graph, [post]

    def _attach_default_inference_algorithms(self):
        observed = [v for _, v in self.inputs] + [v for _, v in self.outputs]
        self.attach_log_pdf_algorithms(targets=self.output_names, conditionals=self.input_names,
                                       algorithm=SVGPRegressionLogPdf(self._module_graph, self._extra_graphs[0],
                                                                      observed), alg_name='svgp_log_pdf')
        observed = [v for _, v in self.inputs]
        from ._sampling import InducingGPSampling
        self.attach_draw_samples_algorithms(targets=self.output_names, conditionals=self.input_names,
                                            algorithm=InducingGPSampling(self._module_graph, observed, rand_gen=self._rand_gen,
                                                         dtype=self.dtype), alg_name='svgp_sampling')
        self.attach_prediction_algorithms(targets=self.output_names, conditionals=self.input_names,
                                          algorithm=SVGPRegressionMeanVariancePrediction(
                                              self._module_graph, self._extra_graphs[0], observed),
                                          alg_name='svgp_predict')

    @staticmethod
    def define_variable(X, kernel, noise_var, shape=None, inducing_inputs=None, num_inducing=10, mean=None,
                        rand_gen=None, dtype=None, ctx=None):
        gp = SVGPRegression(X=X, kernel=kernel, noise_var=noise_var, inducing_inputs=inducing_inputs,
                            num_inducing=num_inducing, mean=mean, rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        gp._generate_outputs({'random_variable': shape})
        return gp.random_variable
