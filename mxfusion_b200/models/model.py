"""mxfusion/models/model.py:21-48."""
from .factor_graph import FactorGraph


class Model(FactorGraph):
    """The model definition: ``m = Model(); m.x = Variable(...); m.y = Normal.define_variable(...)``."""
    pass
