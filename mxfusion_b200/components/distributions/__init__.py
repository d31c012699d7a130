from .distribution import Distribution  # noqa: F401
from .normal import Normal  # noqa: F401
from .pointmass import PointMass  # noqa: F401
from .random_gen import MXNetRandomGenerator, MockMXNetRandomGenerator  # noqa: F401
from .gp.kernels import RBF, Matern12, Matern32, Matern52  # noqa: F401
from .gp import GaussianProcess, ConditionalGaussianProcess  # noqa: F401
