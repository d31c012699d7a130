from .parameter import Parameter, ParameterDict  # noqa: F401
from .block import Block, HybridBlock  # noqa: F401
from .trainer import Trainer  # noqa: F401
from . import data, nn  # noqa: F401
