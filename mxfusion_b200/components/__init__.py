from .model_component import ModelComponent  # noqa: F401
from .factor import Factor  # noqa: F401
from .variables import Variable, VariableType, PositiveTransformation, Softplus, Logistic  # noqa: F401
