"""mxfusion_b200: B200-native (sm_100a) implementation of MXFusion's variational-inference /
Gaussian-process hot path behind the reference's own Model / Module / Inference API.

The arithmetic lives in libmxf_b200.so (hand-written CUDA, C ABI in include/mxf_b200.h); this
package is the host-side mirror of the reference interface for that path (same module layout as
`mxfusion`: components / models / modules / inference).  There is no CPU path.
"""
__version__ = '0.1.0'

from .common import config  # noqa: F401
from .components import Variable  # noqa: F401
from .models import Model, Posterior  # noqa: F401
from . import components, models, modules, inference, ops, F  # noqa: F401
