from .mxfusion_function import MXFusionFunction, FunctionEvaluation  # noqa: F401
from .torch_function import MXFusionTorchFunction, MXFusionGluonFunction  # noqa: F401
