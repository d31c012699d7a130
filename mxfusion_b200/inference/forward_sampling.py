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
