from .ndarray import *  # noqa: F401,F403
from .ndarray import NDArray, array, linalg, random  # noqa: F401
from . import ndarray  # noqa: F401
