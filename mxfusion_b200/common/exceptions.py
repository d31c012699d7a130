"""Exception types of the public surface; names follow the reference so user code's `except` clauses keep working
(mxfusion/common/exceptions.py:15-25).  All derive from one package base class."""


class MXFusionError(Exception):
    """Base of every error this package raises on purpose."""


class ModelSpecificationError(MXFusionError):
    """The model / posterior / kernel definition is inconsistent (wrong factor, shape, or combination)."""


class InferenceError(MXFusionError):
    """An inference algorithm was asked for something the graph cannot provide (partially observed outputs, a
    non-positive-definite covariance reported by potrf's `info`, an unsupported optimiser, ...)."""


class SerializationError(MXFusionError):
    """A saved inference archive does not match the graphs it is loaded into."""
