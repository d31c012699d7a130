from .module import Module  # noqa: F401
from .gp_modules import GPRegression, SVGPRegression, SparseGPRegression  # noqa: F401
