"""mxfusion/models/posterior.py:20-66: a factor graph over replicas (same uuid) of a model's variables."""
from .factor_graph import FactorGraph
from ..components.model_component import ModelComponent


class Posterior(FactorGraph):
    def __init__(self, model, name=None, verbose=False):
        super(Posterior, self).__init__(name=name, verbose=verbose)
        self.__dict__['_model'] = model

    def _replica(self, node, name):
        rep = node.replicate_self()
        if name is not None:
            setattr(self, name, rep)
        else:
            rep.graph = self
            self._anchors[rep.uuid] = rep
        return rep

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        model = self.__dict__.get('_model')
        if model is not None and name in model.__dict__ and isinstance(model.__dict__[name], ModelComponent):
            return self._replica(model.__dict__[name], name)
        raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, name))

    def __getitem__(self, item):
        key = item.uuid if isinstance(item, ModelComponent) else item
        comps = self.components
        if key in comps:
            return comps[key]
        if key in self._model:
            node = self._model[key]
            return self._replica(node, node.name)
        raise AttributeError("'%s' object has no item '%s'" % (type(self).__name__, item))
