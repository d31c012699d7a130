"""Variational sparse GP regression module (Titsias 2009)
(mxfusion/modules/gp_modules/sparsegp_regression.py:30-424)."""
import torch

from ..module import Module
from ...models import Model, Posterior
from ...components.variables.variable import Variable
from ...components.variables.runtime_variable import arrays_as_samples
from ...components.distributions.random_gen import MXNetRandomGenerator
from ...inference.variational import VariationalInference
from ...inference.inference_alg import SamplingAlgorithm
from ... import ops


def _active(F, kern, *arrays):
    if kern.active_dims is None:
        return arrays
    from ...components.distributions.gp.kernels.kernel import slice_axis
    return tuple(slice_axis(F, a, -1, kern.active_dims) for a in arrays)


class SparseGPRegressionLogPdf(VariationalInference):
    """The collapsed variational lower bound (sparsegp_regression.py:30-108) through ops.sparsegp_log_pdf: the N axis
    is consumed by streamed whitened statistics (K(Z, X_c) -> L^-1 K -> syrk / gemm2 per block of rows, never
    materialising the M x N matrices), the rest is M x M.  As in the reference, `log_pdf_scaling` is accepted but not
    applied, and wv, L, LA are published into the posterior graph for prediction (:101-106)."""

    def __init__(self, model, posterior, observed, jitter=0.):
        super(SparseGPRegressionLogPdf, self).__init__(model=model, posterior=posterior, observed=observed)
        self.log_pdf_scaling = 1
        self.jitter = jitter
        self.chunk_rows = None          # rows of X per streamed block (None: ops.STATS_CHUNK_ROWS)

    def compute(self, F, variables):
        X = variables[self.model.X]
        Y = variables[self.model.Y]
        Z = variables[self.model.inducing_inputs]
        noise_var = variables[self.model.noise_var]
        kern = self.model.kernel
        mean = variables[self.model.mean] if self.model.has_mean else None
        if getattr(kern, 'KIND', None) is None:          # Add / Multiply / Linear / static kernels: materialising path
            from . import _generic
            logL, wv, L, LA = _generic.sparsegp_log_pdf(F, kern, kern.fetch_parameters(variables), X, Y, Z, noise_var,
                                                        self.jitter, mean=mean)
            with torch.no_grad():
                self.set_parameter(variables, self.posterior.wv, wv[0])
                self.set_parameter(variables, self.posterior.L, L[0])
                self.set_parameter(variables, self.posterior.LA, LA[0])
            return logL
        kp = kern._strip(kern.fetch_parameters(variables))
        X, Z = _active(F, kern, X, Z)
        logL, wv, L, LA, _ = ops.sparsegp_log_pdf(kern.KIND, X, Y, Z, noise_var, kp['lengthscale'], kp['variance'],
                                                  jitter=self.jitter, mean=mean, chunk=self.chunk_rows)
        with torch.no_grad():
            self.set_parameter(variables, self.posterior.wv, wv[0])
            self.set_parameter(variables, self.posterior.L, L[0])
            self.set_parameter(variables, self.posterior.LA, LA[0])
        return logL


def _predict_moments(alg, F, variables):
    """Shared by the two prediction algorithms (sparsegp_regression.py:131-165 / :200-234)."""
    X = variables[alg.model.X]
    Z = variables[alg.model.inducing_inputs]
    noise_var = variables[alg.model.noise_var]
    L = variables[alg.graphs[1].L]
    LA = variables[alg.graphs[1].LA]
    wv = variables[alg.graphs[1].wv]
    kern = alg.model.kernel
    kern_params = kern.fetch_parameters(variables)
    X, Z, noise_var, L, LA, wv, kern_params = arrays_as_samples(F, [X, Z, noise_var, L, LA, wv, kern_params])
    N = X.shape[-2]
    Kxt = kern.K(F, Z, X, **kern_params)
    mu = ops.gemm2(Kxt, wv, True, False)
    if alg.model.has_mean:
        mu = mu + variables[alg.model.mean]
    LinvKxt = ops.trsm(L, Kxt)
    LAinvLinvKxt = ops.trsm(LA, LinvKxt)
    if alg.diagonal_variance:
        var = kern.Kdiag(F, X, **kern_params) - torch.sum(torch.square(LinvKxt), dim=-2) + \
            torch.sum(torch.square(LAinvLinvKxt), dim=-2)
        if not alg.noise_free:
            var = var + noise_var
    else:
        var = kern.K(F, X, **kern_params) - ops.syrk(LinvKxt, True) + ops.syrk(LAinvLinvKxt, True)
        if not alg.noise_free:
            var = var + torch.eye(N, dtype=X.dtype, device=X.device).unsqueeze(0) * noise_var.unsqueeze(-2)
    return mu, var


class SparseGPRegressionMeanVariancePrediction(SamplingAlgorithm):
    """sparsegp_regression.py:111-171."""

    def __init__(self, model, posterior, observed, target_variables=None, noise_free=True, diagonal_variance=True):
        super(SparseGPRegressionMeanVariancePrediction, self).__init__(model=model, observed=observed,
                                                                       extra_graphs=[posterior])
        self.noise_free = noise_free
        self.diagonal_variance = diagonal_variance

    def compute(self, F, variables):
        outcomes = {self.model.Y.uuid: _predict_moments(self, F, variables)}
        if self.target_variables:
            return tuple(outcomes[v] for v in self.target_variables)
        return outcomes


class SparseGPRegressionSamplingPrediction(SamplingAlgorithm):
    """sparsegp_regression.py:174-255: draws from the predictive distribution (independent per point with
    `diagonal_variance`, else through the Cholesky factor of the full predictive covariance)."""

    def __init__(self, model, posterior, observed, rand_gen=None, noise_free=True, diagonal_variance=True, jitter=0.):
        super(SparseGPRegressionSamplingPrediction, self).__init__(model=model, observed=observed,
                                                                   extra_graphs=[posterior])
        self.noise_free = noise_free
        self.diagonal_variance = diagonal_variance
        self._rand_gen = MXNetRandomGenerator if rand_gen is None else rand_gen
        self.jitter = jitter

    def compute(self, F, variables):
        mu, var = _predict_moments(self, F, variables)
        out_shape = (self.num_samples,) + tuple(mu.shape[1:])
        die = self._rand_gen.sample_normal(shape=out_shape, dtype=mu.dtype, ctx=mu.device)
        if self.diagonal_variance:
            samples = mu + die * torch.sqrt(var.unsqueeze(-1))
        else:
            cov = var
            if self.jitter > 0.:
                n = cov.shape[-1]
                cov = cov + torch.eye(n, dtype=cov.dtype, device=cov.device) * self.jitter
            Lc = ops.potrf(cov)
            samples = mu + ops.gemm2(Lc.expand((self.num_samples,) + tuple(Lc.shape[1:])), die)
        outcomes = {self.model.Y.uuid: samples}
        if self.target_variables:
            return tuple(outcomes[v] for v in self.target_variables)
        return outcomes


class SparseGPRegression(Module):
    """`m.Y = SparseGPRegression.define_variable(X=m.X, kernel=k, noise_var=m.noise_var, shape=(m.N, 1),
    num_inducing=M)`; hidden parameter: inducing_inputs (M,D); posterior placeholders L, LA (M,M), wv (M,P)
    (sparsegp_regression.py:258-424)."""

    def __init__(self, X, kernel, noise_var, inducing_inputs=None, num_inducing=10, mean=None, rand_gen=None,
                 dtype=None, ctx=None):
        if not isinstance(X, Variable):
            X = Variable(value=X)
        if not isinstance(noise_var, Variable):
            noise_var = Variable(value=noise_var)
        if inducing_inputs is None:
            inducing_inputs = Variable(shape=(num_inducing, kernel.input_dim))
        inputs = [('X', X), ('inducing_inputs', inducing_inputs), ('noise_var', noise_var)]
        self._has_mean = mean is not None
        if mean is not None:
            inputs.append(('mean', mean))
        super(SparseGPRegression, self).__init__(inputs=inputs, outputs=None, input_names=[k for k, _ in inputs],
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
        post = Posterior(graph)          # place holders of intermediate results, used for prediction (:342-348)
        post.L = Variable(shape=(M, M))
        post.LA = Variable(shape=(M, M))
        post.wv = Variable(shape=(M, Y.shape[-1]))
        return graph, [post]

    def _attach_default_inference_algorithms(self):
        observed = [v for _, v in self.inputs] + [v for _, v in self.outputs]
        self.attach_log_pdf_algorithms(targets=self.output_names, conditionals=self.input_names,
                                       algorithm=SparseGPRegressionLogPdf(self._module_graph, self._extra_graphs[0],
                                                                          observed), alg_name='sgp_log_pdf')
        observed = [v for _, v in self.inputs]
        from ._sampling import InducingGPSampling
        self.attach_draw_samples_algorithms(targets=self.output_names, conditionals=self.input_names,
                                            algorithm=InducingGPSampling(self._module_graph, observed, rand_gen=self._rand_gen,
                                                         dtype=self.dtype), alg_name='sgp_sampling')
        self.attach_prediction_algorithms(targets=self.output_names, conditionals=self.input_names,
                                          algorithm=SparseGPRegressionMeanVariancePrediction(
                                              self._module_graph, self._extra_graphs[0], observed),
                                          alg_name='sgp_predict')

    @staticmethod
    def define_variable(X, kernel, noise_var, shape=None, inducing_inputs=None, num_inducing=10, mean=None,
                        rand_gen=None, dtype=None, ctx=None):
        gp = SparseGPRegression(X=X, kernel=kernel, noise_var=noise_var, inducing_inputs=inducing_inputs,
                                num_inducing=num_inducing, mean=mean, rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        gp._generate_outputs({'random_variable': shape})
        return gp.random_variable
