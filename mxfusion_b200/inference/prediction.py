"""Prediction through a module's registered prediction algorithm (mxfusion/inference/prediction.py:22-85)."""
from .inference_alg import SamplingAlgorithm
from ..components.variables.variable import VariableType
from ..modules.module import Module
from ..common.exceptions import InferenceError


class ModulePredictionAlgorithm(SamplingAlgorithm):
    def compute(self, F, variables):
        outcomes = {}
        for f in self.model.ordered_factors:
            if isinstance(f, Module):
                targets = [v.uuid for _, v in f.outputs if v.uuid not in variables]
                if not targets:
                    continue
                known = {k: v for k, v in variables.items()}
                res = f.predict(F=F, variables=known, num_samples=self.num_samples, targets=targets)
                for uuid, r in zip(targets, res):
                    outcomes[uuid] = r
            else:
                missing = [v for _, v in f.outputs if v.uuid not in variables]
                if missing and f.is_probabilistic:
                    raise InferenceError("ModulePredictionAlgorithm can only predict the outputs of modules; " +
                                         str(f) + " has unobserved outputs.")
        if self.target_variables:
            return tuple(outcomes[v.uuid if hasattr(v, 'uuid') else v] for v in self.target_variables)
        return outcomes
