"""Variable nodes (mxfusion/components/variables/variable.py:24-265): constants, parameters, random
variables (produced by a Distribution or Module) and function outputs."""
from enum import Enum

import numpy as np
import torch

from ..model_component import ModelComponent
from ...common.config import torch_dtype
from ...common.exceptions import ModelSpecificationError


class VariableType(Enum):
    CONSTANT = 0
    PARAMETER = 1
    RANDVAR = 2
    FUNCVAR = 3


def _as_tensor(value):
    if isinstance(value, torch.Tensor):
        return value
    if isinstance(value, np.ndarray):
        return torch.as_tensor(value, dtype=torch_dtype())
    if isinstance(value, (int, float)):
        return torch.tensor([value], dtype=torch_dtype())
    raise ModelSpecificationError("Variable type {} not supported".format(type(value)))


class Variable(ModelComponent):
    def __init__(self, value=None, shape=None, transformation=None, isInherited=False, initial_value=None):
        super(Variable, self).__init__()
        if shape is not None and not isinstance(shape, tuple):
            raise AssertionError("Shape is expected to be a tuple or None")
        self.shape = shape
        self.attributes = [s for s in shape if isinstance(s, Variable)] if shape is not None else []
        self.isInherited = isInherited
        self.inherited_name = None
        self._transformation = transformation
        self._value = None
        self.isConstant = False
        if initial_value is not None:
            initial_value = _as_tensor(initial_value)
        self._initial_value = initial_value
        from ..factor import Factor
        if isinstance(value, Factor):
            if transformation is not None:
                raise NotImplementedError('Constraints on random variables / function outputs are not supported!')
            if shape is None and not value.is_probabilistic:
                raise ModelSpecificationError("The shape argument was not given when defining a variable as the "
                                              "outcome of a function evaluation.")
            value.set_outputs(self)
        elif value is None:
            if self.shape is None:
                self.shape = (1,)
        else:
            self.isConstant = True
            if isinstance(value, (int, float)):
                self._value = value
                self.shape = (1,)
            else:
                value = _as_tensor(value)
                if self.shape is None:
                    self.shape = tuple(value.shape)
                if tuple(self.shape) != tuple(value.shape):
                    raise ModelSpecificationError(
                        "Shape mismatch in Variable creation. The array shape " + str(tuple(value.shape)) +
                        " does not match with the shape argument " + str(self.shape) + ".")
                self._value = value

    # typing --------------------------------------------------------------------------------------
    @property
    def factor(self):
        return self._in[0][1] if self._in else None

    @property
    def type(self):
        f = self.factor
        if f is None:
            return VariableType.CONSTANT if self.isConstant else VariableType.PARAMETER
        return VariableType.RANDVAR if f.is_probabilistic else VariableType.FUNCVAR

    @property
    def constant(self):
        if self.type == VariableType.CONSTANT:
            return self._value
        raise ModelSpecificationError("The constant property is not accessible for variable with the type " +
                                      str(self.type) + ".")

    def get_constant(self):
        return self.constant

    @property
    def transformation(self):
        return self._transformation

    @property
    def initial_value(self):
        return self._initial_value

    @property
    def initial_value_before_transformation(self):
        if self._transformation is None or self._initial_value is None:
            return self._initial_value
        return self._transformation.inverseTransform(self._initial_value)

    def set_prior(self, distribution):
        self.assign_factor(distribution)

    def assign_factor(self, factor):
        factor.set_outputs(self)

    def replicate_self(self, attribute_map=None):
        """Same uuid, no edges: the node as seen from another graph."""
        shape = self.shape
        if attribute_map is not None and shape is not None:
            shape = tuple(attribute_map.get(s, s) if isinstance(s, Variable) else s for s in shape)
        v = Variable.__new__(Variable)
        ModelComponent.__init__(v)
        v.shape = shape
        v.attributes = [s for s in shape if isinstance(s, Variable)] if shape is not None else []
        v.isInherited, v.inherited_name = self.isInherited, self.inherited_name
        v._transformation = self._transformation
        v._value = self._value if self.type == VariableType.CONSTANT else None
        v.isConstant = self.type == VariableType.CONSTANT
        v._initial_value = self._initial_value
        v._uuid = self._uuid
        v.name = self.name
        return v

    def __repr__(self):
        s = "Variable"
        if self.name is not None:
            s += " {}".format(self.name)
        return s + " ({})".format(self.uuid[:5])
