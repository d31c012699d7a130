from .kernel import Kernel, NativeKernel  # noqa: F401
from .stationary import StationaryKernel, RBF, Matern, Matern12, Matern32, Matern52  # noqa: F401
