"""A minimal stand-in for the `mxnet` package -- TEST INFRASTRUCTURE ONLY, used by tests/golden/make_golden.py.

`import mxnet` is impossible in this image, yet the reference (amzn/MXFusion, pure Python) can
only be executed on top of it.  This package provides just enough of the MXNet 1.x NDArray / Gluon / autograd
surface, on torch CPU tensors (float64 by default, so torch autograd supplies the gradients MXNet autograd would),
for the reference's OWN source under /root/reference to run its GP / SVGP / Normal / inference-loop code paths.
Operator semantics follow the MXNet 1.x operator documentation (the same restatement as oracle/linalg.py).
What this pins: everything the reference does in Python around the operators -- graph walks, sample-axis handling,
parameter transforms, log_pdf_scaling, the loop's batching / rescaling -- which is exactly the part a restatement
can get wrong.  It never ships to the GPU box and nothing outside tests/golden/ imports it.
"""
from . import ndarray, ndarray as nd, symbol, symbol as sym, autograd, initializer, initializer as init, gluon, operator, context, random  # noqa: E401,F401
from .context import cpu, gpu, Context, current_context  # noqa: F401
from .base import MXNetError  # noqa: F401

__version__ = '1.3.0-standin'
