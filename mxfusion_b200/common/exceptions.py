"""Exception classes of the reference's public surface (mxfusion/common/exceptions.py:15-25)."""


class ModelSpecificationError(Exception):
    pass


class InferenceError(Exception):
    pass


class SerializationError(Exception):
    pass
