"""Softplus parameter transform restated in NumPy (test infrastructure).

``mxfusion/components/variables/var_trans.py:63-91``:
``transform`` = Activation(softrelu) + offset = log(1 + exp(x)) + offset,
``inverseTransform`` = log(expm1(y - offset)).
"""
import numpy as np


def softplus(x, offset=0.0):
    """var_trans.py:75 (softrelu)."""
    return np.logaddexp(0.0, x) + offset


def softplus_inverse(y, offset=0.0):
    """var_trans.py:91."""
    return np.log(np.expm1(y - offset))
