"""Exact Gaussian-process regression module (mxfusion/modules/gp_modules/gp_regression.py:31-428)."""
import torch

from ..module import Module
from ...models import Model, Posterior
from ...components.variables.variable import Variable
from ...components.variables.runtime_variable import arrays_as_samples
from ...inference.variational import VariationalInference
from ...inference.inference_alg import SamplingAlgorithm
from ... import ops


class GPRegressionLogPdf(VariationalInference):
    """log p(Y | X) of an exact GP (gp_regression.py:31-76): K + (noise + jitter) I -> potrf -> trsm ->
    sumlogdiag, as ONE fused operator (ops.gp_log_pdf) with an analytic adjoint.  As in the reference,
    `log_pdf_scaling` is NOT applied (:70) and X, L, L^-1 Y are published into the posterior graph for
    prediction (:72-75)."""

    def __init__(self, model, posterior, observed, jitter=0.):
        super(GPRegressionLogPdf, self).__init__(model=model, posterior=posterior, observed=observed)
        self.log_pdf_scaling = 1
        self.jitter = jitter

    def compute(self, F, variables):
        X = variables[self.model.X]
        Y = variables[self.model.Y]
        noise_var = variables[self.model.noise_var]
        kern = self.model.kernel
        mean = variables[self.model.mean] if self.model.has_mean else None
        if getattr(kern, 'KIND', None) is None:          # Add / Multiply / Linear / static kernels
            from . import _generic
            logL, L, LinvY = _generic.gp_log_pdf(F, kern, kern.fetch_parameters(variables), X, Y, noise_var,
                                                 self.jitter, mean=mean)
            with torch.no_grad():
                self.set_parameter(variables, self.posterior.X, X[0])
                self.set_parameter(variables, self.posterior.L, L[0])
                self.set_parameter(variables, self.posterior.LinvY, LinvY[0])
            return logL
        kp = kern._strip(kern.fetch_parameters(variables))
        Xk = X
        if kern.active_dims is not None:
            from ...components.distributions.gp.kernels.kernel import slice_axis
            Xk = slice_axis(F, X, -1, kern.active_dims)
        logL, L, LinvY = ops.gp_log_pdf(kern.KIND, Xk, Y, noise_var, kp['lengthscale'], kp['variance'],
                                        jitter=self.jitter, mean=mean)
        with torch.no_grad():
            self.set_parameter(variables, self.posterior.X, X[0])
            self.set_parameter(variables, self.posterior.L, L[0])
            self.set_parameter(variables, self.posterior.LinvY, LinvY[0])
        return logL


class GPRegressionMeanVariancePrediction(SamplingAlgorithm):
    """gp_regression.py:138-196."""

    def __init__(self, model, posterior, observed, noise_free=True, diagonal_variance=True):
        super(GPRegressionMeanVariancePrediction, self).__init__(model=model, observed=observed,
                                                                 extra_graphs=[posterior])
        self.noise_free = noise_free
        self.diagonal_variance = diagonal_variance

    def compute(self, F, variables):
        outcomes = {self.model.Y.uuid: _gp_predict_moments(self, F, variables)}
        if self.target_variables:
            return tuple(outcomes[v] for v in self.target_variables)
        return outcomes


def _gp_predict_moments(alg, F, variables):
    """Predictive mean and (diagonal or full) covariance (gp_regression.py:158-190 / :225-259)."""
    X = variables[alg.model.X]
    N = X.shape[-2]
    noise_var = variables[alg.model.noise_var]
    X_cond = variables[alg.graphs[1].X]
    L = variables[alg.graphs[1].L]
    LinvY = variables[alg.graphs[1].LinvY]
    kern = alg.model.kernel
    kern_params = kern.fetch_parameters(variables)
    X, noise_var, X_cond, L, LinvY, kern_params = arrays_as_samples(
        F, [X, noise_var, X_cond, L, LinvY, kern_params])
    Kxt = kern.K(F, X_cond, X, **kern_params)
    LinvKxt = ops.trsm(L, Kxt)
    mu = ops.gemm2(LinvKxt, LinvY, True, False)
    if alg.model.has_mean:
        mu = mu + variables[alg.model.mean]
    if alg.diagonal_variance:
        var = kern.Kdiag(F, X, **kern_params) - torch.sum(torch.square(LinvKxt), dim=-2)
        if not alg.noise_free:
            var = var + noise_var
    else:
        var = kern.K(F, X, **kern_params) - ops.syrk(LinvKxt, True)
        if not alg.noise_free:
            var = var + torch.eye(N, dtype=X.dtype, device=X.device).unsqueeze(0) * noise_var.unsqueeze(-2)
    return mu, var


def _draw_from_moments(alg, mu, var):
    """Samples from N(mu, var) (independent per point, or through the Cholesky factor of the full covariance):
    gp_regression.py:246-268."""
    from ...components.distributions.random_gen import MXNetRandomGenerator
    gen = alg._rand_gen if alg._rand_gen is not None else MXNetRandomGenerator
    out_shape = (alg.num_samples,) + tuple(mu.shape[1:])
    die = gen.sample_normal(shape=out_shape, dtype=mu.dtype, ctx=mu.device)
    if alg.diagonal_variance:
        return mu + die * torch.sqrt(var.unsqueeze(-1))
    cov = var
    if getattr(alg, 'jitter', 0.) > 0.:
        n = cov.shape[-1]
        cov = cov + torch.eye(n, dtype=cov.dtype, device=cov.device) * alg.jitter
    Lc = ops.potrf(cov)
    return mu + ops.gemm2(Lc.expand((alg.num_samples,) + tuple(Lc.shape[1:])), die)


class GPRegressionSamplingPrediction(SamplingAlgorithm):
    """gp_regression.py:199-275: draws from the predictive distribution."""

    def __init__(self, model, posterior, observed, rand_gen=None, noise_free=True, diagonal_variance=True, jitter=0.):
        super(GPRegressionSamplingPrediction, self).__init__(model=model, observed=observed,
                                                             extra_graphs=[posterior])
        self.noise_free = noise_free
        self.diagonal_variance = diagonal_variance
        self._rand_gen = rand_gen
        self.jitter = jitter

    def compute(self, F, variables):
        mu, var = _gp_predict_moments(self, F, variables)
        outcomes = {self.model.Y.uuid: _draw_from_moments(self, mu, var)}
        if self.target_variables:
            return tuple(outcomes[v] for v in self.target_variables)
        return outcomes


class GPRegression(Module):
    """`m.Y = GPRegression.define_variable(X=m.X, kernel=k, noise_var=m.noise_var, shape=(m.N, 1))`."""

    def __init__(self, X, kernel, noise_var, mean=None, rand_gen=None, dtype=None, ctx=None):
        if not isinstance(X, Variable):
            X = Variable(value=X)
        if not isinstance(noise_var, Variable):
            noise_var = Variable(value=noise_var)
        inputs = [('X', X), ('noise_var', noise_var)]
        self._has_mean = mean is not None
        if mean is not None:
            inputs.append(('mean', mean))
        super(GPRegression, self).__init__(inputs=inputs, outputs=None, input_names=[k for k, _ in inputs],
                                           output_names=['random_variable'], rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        self.kernel = kernel

    def _generate_outputs(self, output_shapes):
        shape = output_shapes['random_variable']
        if shape is None:
            shape = self.X.shape[:-1] + (1,)
        self.set_outputs([Variable(shape=shape)])

    def _build_module_graphs(self):
        Y = self.random_variable
        graph = Model(name='gp_regression')
        graph.X = self.X.replicate_self()
        graph.noise_var = self.noise_var.replicate_self()
        graph.__dict__['has_mean'] = self._has_mean
        if self._has_mean:
            graph.mean = self.mean.replicate_self()
        graph.Y = Y.replicate_self()
        graph.__dict__['kernel'] = self.kernel
        for name, var in self.kernel.parameters.items():
            graph.add_component(var, name)
        post = Posterior(graph)                      # stores what prediction needs (gp_regression.py:352-356)
        post.L = Variable(shape=graph.X.shape[:-1] + graph.X.shape[-2:-1])
        post.LinvY = Variable(shape=graph.X.shape[:-1] + graph.Y.shape[-1:])
        post.X = Variable(shape=graph.X.shape)
        return graph, [post]

    def _attach_default_inference_algorithms(self):
        observed = [v for _, v in self.inputs] + [v for _, v in self.outputs]
        self.attach_log_pdf_algorithms(targets=self.output_names, conditionals=self.input_names,
                                       algorithm=GPRegressionLogPdf(self._module_graph, self._extra_graphs[0],
                                                                    observed), alg_name='gp_log_pdf')
        observed = [v for _, v in self.inputs]
        from ._sampling import GPRegressionSampling
        self.attach_draw_samples_algorithms(targets=self.output_names, conditionals=self.input_names,
                                            algorithm=GPRegressionSampling(self._module_graph, observed, rand_gen=self._rand_gen,
                                                         dtype=self.dtype), alg_name='gp_sampling')
        self.attach_prediction_algorithms(targets=self.output_names, conditionals=self.input_names,
                                          algorithm=GPRegressionMeanVariancePrediction(
                                              self._module_graph, self._extra_graphs[0], observed),
                                          alg_name='gp_predict')

    @staticmethod
    def define_variable(X, kernel, noise_var, shape=None, mean=None, rand_gen=None, dtype=None, ctx=None):
        gp = GPRegression(X=X, kernel=kernel, noise_var=noise_var, mean=mean, rand_gen=rand_gen, dtype=dtype,
                          ctx=ctx)
        gp._generate_outputs({'random_variable': shape})
        return gp.random_variable
