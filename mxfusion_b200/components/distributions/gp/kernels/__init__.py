from .kernel import Kernel, NativeKernel, CombinationKernel  # noqa: F401
from .stationary import StationaryKernel, RBF, Matern, Matern12, Matern32, Matern52  # noqa: F401
from .linear import Linear  # noqa: F401
from .static import Bias, White  # noqa: F401
from .add_kernel import AddKernel  # noqa: F401
from .multiply_kernel import MultiplyKernel  # noqa: F401
