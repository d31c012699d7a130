"""Prediction through the modules' registered prediction algorithms
(mxfusion/inference/prediction.py:22-85): walk the model; functions are evaluated, unobserved distributions are
sampled, modules are asked to `predict` their outputs."""
from .inference_alg import SamplingAlgorithm
from ..components.functions.mxfusion_function import FunctionEvaluation
from ..components.distributions.distribution import Distribution
from ..components.model_component import ModelComponent
from ..modules.module import Module
from ..common.exceptions import InferenceError


class ModulePredictionAlgorithm(SamplingAlgorithm):
    def compute(self, F, variables):
        outcomes = {}
        for f in self.model.ordered_factors:
            out_uuid = [v.uuid for _, v in f.outputs]
            if isinstance(f, FunctionEvaluation):
                res = f.eval(F=F, variables=variables, always_return_tuple=True)
            elif isinstance(f, Distribution):
                known = [u in variables for u in out_uuid]
                if all(known):
                    continue
                if any(known):
                    raise InferenceError("Part of the outputs of the distribution " + f.__class__.__name__ +
                                         " has been observed!")
                res = f.draw_samples(F=F, num_samples=self.num_samples, variables=variables, always_return_tuple=True)
            elif isinstance(f, Module):
                res = f.predict(F=F, variables=variables, targets=out_uuid, num_samples=self.num_samples)
            else:
                continue
            for v, uuid in zip(res, out_uuid):
                variables[uuid] = v
                outcomes[uuid] = v
        if self.target_variables:
            return tuple(outcomes[t.uuid if isinstance(t, ModelComponent) else t] for t in self.target_variables)
        return outcomes
