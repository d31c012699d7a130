from .symbol import Symbol  # noqa: F401


def array(*a, **k):
    raise NotImplementedError("symbolic mode does not exist in the stand-in")
