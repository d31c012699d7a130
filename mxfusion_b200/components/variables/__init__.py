from .variable import Variable, VariableType  # noqa: F401
from .var_trans import Softplus, PositiveTransformation, Logistic, VariableTransformation  # noqa: F401
from .runtime_variable import (add_sample_dimension, add_sample_dimension_to_arrays, expectation,  # noqa: F401
                               is_sampled_array, get_num_samples, as_samples, arrays_as_samples)
