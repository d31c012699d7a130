"""Parameter transformations (mxfusion/components/variables/var_trans.py:24-147).  `transform` runs
on every forward for every constrained parameter (inference_alg.py:79-80) through the softplus
kernel; `inverseTransform` runs once at initialisation / `params[var] = value` on tiny arrays."""
import torch

from ... import ops


class VariableTransformation(object):
    def transform(self, var, F=None, dtype=None):
        raise NotImplementedError

    def inverseTransform(self, out_var, F=None, dtype=None):
        raise NotImplementedError


class Softplus(VariableTransformation):
    """y = log(1 + exp(x)) + offset  (var_trans.py:63-91)."""

    def __init__(self, offset):
        self._offset = offset

    def transform(self, var, F=None, dtype=None):
        return ops.softplus(var, self._offset)

    def inverseTransform(self, out_var, F=None, dtype=None):
        return torch.log(torch.expm1(out_var - self._offset))          # var_trans.py:91


class PositiveTransformation(Softplus):
    def __init__(self):
        super(PositiveTransformation, self).__init__(offset=0.)


class Logistic(VariableTransformation):
    """lower + (upper-lower) * sigmoid(x)  (var_trans.py:105-147); off the GP hot path, torch elementwise."""

    def __init__(self, lower, upper):
        if lower >= upper:
            raise ValueError('The lower bound is above the upper bound')
        self._lower, self._upper = lower, upper
        self._difference = upper - lower

    def transform(self, var, F=None, dtype=None):
        return self._lower + self._difference * torch.sigmoid(var)

    def inverseTransform(self, out_var, F=None, dtype=None):
        c = torch.clamp(out_var, self._lower + 1e-10, self._upper - 1e-10)
        return torch.log((c - self._lower) / (self._upper - c))
