"""Factor nodes: a relation among variables -- a distribution, a function evaluation or a module
(mxfusion/components/factor.py:46-263).  Inputs and outputs are reachable as attributes by edge name."""
import numpy as np
import torch

from .model_component import ModelComponent
from .variables.variable import Variable
from ..common.exceptions import ModelSpecificationError
from ..common.config import torch_dtype


def _define_variable_from_constant(v):
    if isinstance(v, Variable):
        return v
    if isinstance(v, (int, float)):
        return Variable(value=torch.tensor([v], dtype=torch_dtype()))
    if isinstance(v, (torch.Tensor, np.ndarray)):
        return Variable(value=v)
    raise ModelSpecificationError('The inputs/outputs of a factor can only be a int, float, array or Variable, '
                                  'but get ' + str(v) + '.')


class Factor(ModelComponent):
    is_probabilistic = False      # True for distributions and modules (their outputs are random variables)

    def __init__(self, inputs, outputs, input_names, output_names):
        super(Factor, self).__init__()
        self._input_names = list(input_names) if input_names is not None else []
        self._output_names = list(output_names) if output_names is not None else []
        inputs = [(k, _define_variable_from_constant(v)) for k, v in inputs] if inputs is not None else []
        outputs = [(k, _define_variable_from_constant(v)) for k, v in outputs] if outputs is not None else []
        both = set(v.uuid for _, v in inputs) & set(v.uuid for _, v in outputs)
        if both:
            raise RuntimeError("The inputs and outputs variables of " + type(self).__name__ +
                               " have name conflict: " + str(both) + ".")
        self.predecessors = inputs
        self.successors = outputs

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        d = self.__dict__
        if name in d.get('_input_names', ()):
            for n, node in d.get('_in', ()):
                if n == name:
                    return node
        if name in d.get('_output_names', ()):
            for n, node in d.get('_out', ()):
                if n == name:
                    return node
        raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, name))

    @property
    def inputs(self):
        by = dict(self._in)
        return [(n, by[n]) for n in self._input_names if n in by]

    @inputs.setter
    def inputs(self, inputs):
        self.predecessors = inputs

    @property
    def outputs(self):
        by = dict(self._out)
        return [(n, by[n]) for n in self._output_names if n in by]

    @outputs.setter
    def outputs(self, outputs):
        self.successors = outputs

    @property
    def input_names(self):
        return self._input_names

    @property
    def output_names(self):
        return self._output_names

    def set_outputs(self, variables):
        variables = [variables] if not isinstance(variables, (list, tuple)) else variables
        self.successors = [(n, v) for n, v in zip(self._output_names, variables)]

    def set_single_input(self, key, value):
        self.predecessors = [(k, value) if k == key else (k, v) for k, v in self.inputs]

    def fetch_runtime_inputs(self, params):
        return {n: params[v.uuid] for n, v in self.inputs}

    def fetch_runtime_outputs(self, params):
        return {n: params[v.uuid] for n, v in self.outputs}

    def __repr__(self):
        return type(self).__name__ + '(' + ', '.join(str(n) + '=' + str(v) for n, v in self.inputs) + ')'
