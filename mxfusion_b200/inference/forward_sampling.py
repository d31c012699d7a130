"""Forward sampling (mxfusion/inference/forward_sampling.py:24-95): draw samples of the model's variables by walking the
factor graph from the observed variables (FactorGraph.draw_samples)."""
from .inference_alg import SamplingAlgorithm
from .inference import TransferInference
from ..components.variables.variable import Variable


class ForwardSamplingAlgorithm(SamplingAlgorithm):
    """forward_sampling.py:24-56."""

    def compute(self, F, variables):
        return self.model.draw_samples(F=F, variables=variables, targets=self.target_variables,
                                       num_samples=self.num_samples)


class ForwardSampling(TransferInference):
    """forward_sampling.py:59-95: forward sampling with parameters transferred from a finished inference."""

    def __init__(self, num_samples, model, observed, var_tie, infr_params, target_variables=None, hybridize=False,
                 constants=None, dtype=None, context=None):
        if target_variables is not None:
            target_variables = [v.uuid for v in target_variables if isinstance(v, Variable)]
        infr = ForwardSamplingAlgorithm(num_samples=num_samples, model=model, observed=observed,
                                        target_variables=target_variables)
        super(ForwardSampling, self).__init__(inference_algorithm=infr, var_tie=var_tie, infr_params=infr_params,
                                              constants=constants, hybridize=hybridize, dtype=dtype, context=context)


class VariationalPosteriorForwardSamplingAlgorithm(SamplingAlgorithm):
    """Forward sampling with the latent variables drawn from the variational posterior instead of their priors.

    The reference builds a merged graph for this (`merge_posterior_into_model`, forward_sampling.py:99-116: clone the
    model, replace every latent variable's prior factor by its posterior factor) and forward-samples it.  The same draws
    come out of two walks with no graph surgery: sample the posterior graph, then forward-sample the model with those
    values already present -- `FactorGraph.draw_samples` skips a distribution whose outputs are known."""

    def __init__(self, model, posterior, observed, num_samples=1, target_variables=None):
        super(VariationalPosteriorForwardSamplingAlgorithm, self).__init__(
            model=model, observed=observed, num_samples=num_samples, target_variables=target_variables,
            extra_graphs=[posterior])

    def compute(self, F, variables):
        drawn = self.graphs[1].draw_samples(F=F, variables=variables, num_samples=self.num_samples)
        variables.update(drawn)
        rest = self.model.draw_samples(F=F, variables=variables, num_samples=self.num_samples)
        samples = dict(drawn)
        samples.update(rest)
        if self.target_variables:
            return tuple(samples[t.uuid if isinstance(t, Variable) else t] for t in self.target_variables)
        return samples


class VariationalPosteriorForwardSampling(TransferInference):
    """forward_sampling.py:119-157: `VariationalPosteriorForwardSampling(num_samples, observed, inherited_inference,
    target_variables)`; the model, the posterior and the learned parameters are taken from `inherited_inference`."""

    def __init__(self, num_samples, observed, inherited_inference, target_variables=None, hybridize=False,
                 constants=None, dtype=None, context=None):
        from .variational import StochasticVariationalInference
        from .map import MAP
        from ..common.exceptions import InferenceError
        alg = inherited_inference.inference_algorithm
        if not isinstance(alg, (StochasticVariationalInference, MAP)):
            raise InferenceError('inherited_inference needs to be a subclass of SVIInference or SVIMiniBatchInference.')
        if target_variables is not None:
            target_variables = [v.uuid for v in target_variables if isinstance(v, Variable)]
        infr = VariationalPosteriorForwardSamplingAlgorithm(model=alg.model, posterior=alg.posterior, observed=observed,
                                                            num_samples=num_samples, target_variables=target_variables)
        super(VariationalPosteriorForwardSampling, self).__init__(
            inference_algorithm=infr, var_tie={}, infr_params=inherited_inference.params, constants=constants,
            hybridize=hybridize, dtype=dtype, context=context)
