"""Inference algorithms and the executor that runs them
(mxfusion/inference/inference_alg.py:25-293).

`ObjectiveBlock` is the reference's Gluon HybridBlock wrapper; here it is a plain callable doing the
same pre-processing -- variable ties, parameter transforms (softplus kernel), leading sample axis,
constants -- before calling `InferenceAlgorithm.compute(F, variables)`, then writing back any
``SET_`` parameters (:60-90)."""
from abc import ABC, abstractmethod

import torch

from ..common.constants import SET_PARAMETER_PREFIX
from ..components.variables.variable import Variable, VariableType
from ..components.variables.runtime_variable import add_sample_dimension_to_arrays
from ..components.model_component import ModelComponent


def variables_to_UUID(variables):
    return [v.uuid if isinstance(v, ModelComponent) else v for v in variables]


class ObjectiveBlock(object):
    def __init__(self, infr_method, constants, data_def, var_trans, var_ties, excluded, params=None):
        self._infr_method = infr_method
        # The LIVE dictionary of the parameters (inference_alg.py:41 of the reference keeps it too): the minibatch loop
        # updates the shape constants (m.N = batch size) after the executor exists.  Constants given as arrays are moved
        # next to the parameters (same device / dtype) on first use and cached; shape constants stay ints.
        self._constants_live = constants
        self._constants_cache = {}
        self._data_def = data_def
        self._var_trans = var_trans
        self._var_ties = var_ties
        self._infr_params = params
        self._excluded = excluded

    @property
    def _constants(self):
        params = self._infr_params
        dev, dt = params.mxnet_context, params.flat.dtype if params.flat is not None else None
        out = {}
        for k, v in self._constants_live.items():
            if isinstance(v, torch.Tensor):
                hit = self._constants_cache.get(k)
                if hit is None or hit[0] is not v:
                    hit = (v, v.to(device=dev, dtype=dt if v.is_floating_point() and dt is not None else v.dtype))
                    self._constants_cache[k] = hit
                out[k] = hit[1]
            else:
                out[k] = v
        return out

    def __call__(self, x, *args):
        """`x` is the reference's dummy first argument (``mx.nd.zeros(1)``); ignored."""
        from .. import F
        fused = getattr(self, 'pretransformed', None)       # set by the training step: uuids whose transformed value
        params = self._infr_params                           # is already in `tflat` (one launch for all of them)
        if fused is not None:
            kw = {name: (params.leaf(name) if name in params._params else p.tensor)
                  for name, p in params.param_dict.items() if name not in self._excluded}
        else:
            kw = {name: p.tensor for name, p in params.param_dict.items() if name not in self._excluded}
        for to_uuid, from_uuid in self._var_ties.items():
            kw[to_uuid] = kw[from_uuid]
        data = {k: v for k, v in zip(self._data_def, args)}
        variables = add_sample_dimension_to_arrays(F, data)
        for k, t in self._var_trans.items():
            if k in kw and not (fused is not None and k in fused):
                kw[k] = t.transform(kw[k], F=F)
        add_sample_dimension_to_arrays(F, kw, out=variables)
        add_sample_dimension_to_arrays(F, self._constants, out=variables)
        obj = self._infr_method.compute(F=F, variables=variables)
        with torch.no_grad():
            for k, v in list(variables.items()):
                if isinstance(k, str) and k.startswith(SET_PARAMETER_PREFIX):
                    self._infr_params[v[0]] = v[1]
        return obj

    def hybridize(self):        # accepted and ignored: there is no symbolic mode
        pass

    def initialize(self, ctx=None):
        pass


class InferenceAlgorithm(ABC):
    def __init__(self, model, observed, extra_graphs=None):
        self._model_graph = model
        self._extra_graphs = extra_graphs if extra_graphs is not None else []
        self._graphs = [model] + self._extra_graphs
        self._observed = set(observed)
        self._observed_uuid = variables_to_UUID(observed)
        self._observed_names = [v.name for v in observed]

    @property
    def observed_variables(self):
        return self._observed

    @property
    def observed_variable_UUIDs(self):
        return self._observed_uuid

    @property
    def observed_variable_names(self):
        return self._observed_names

    @property
    def model(self):
        return self._model_graph

    @property
    def graphs(self):
        return self._graphs

    def prepare_executor(self, rv_scaling=None):
        """inference_alg.py:165-190."""
        excluded, var_trans = set(), {}
        rv_scaling = {} if rv_scaling is None else rv_scaling
        for g in self._graphs:
            for v in g.variables.values():
                if v.type == VariableType.PARAMETER and v.transformation is not None:
                    var_trans[v.uuid] = v.transformation
                if v.type == VariableType.RANDVAR:
                    v.factor.log_pdf_scaling = rv_scaling.get(v.uuid, 1)
        return var_trans, excluded

    def create_executor(self, data_def, params, var_ties, rv_scaling=None):
        """inference_alg.py:192-219."""
        var_trans, excluded = self.prepare_executor(rv_scaling=rv_scaling)
        for m in self.model.modules.values():
            vt, ex = m.prepare_executor(rv_scaling=rv_scaling)
            var_trans.update(vt)
            excluded = excluded.union(ex)
        return ObjectiveBlock(infr_method=self, params=params, constants=params.constants, data_def=data_def,
                              var_trans=var_trans, var_ties=var_ties, excluded=excluded)

    @abstractmethod
    def compute(self, F, variables):
        raise NotImplementedError

    def set_parameter(self, variables, target_variable, target_value):
        """inference_alg.py:236-251: publish a value computed by the algorithm as a parameter."""
        variables[target_variable.uuid] = target_value
        variables[SET_PARAMETER_PREFIX + target_variable.uuid] = (target_variable, target_value)


class SamplingAlgorithm(InferenceAlgorithm):
    def __init__(self, model, observed, num_samples=1, target_variables=None, extra_graphs=None):
        super(SamplingAlgorithm, self).__init__(model=model, observed=observed, extra_graphs=extra_graphs)
        self.num_samples = num_samples
        self.target_variables = target_variables
