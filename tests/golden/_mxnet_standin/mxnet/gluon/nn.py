from .block import Block, HybridBlock  # noqa: F401


class HybridSequential(HybridBlock):
    pass


class Dense(HybridBlock):
    pass
