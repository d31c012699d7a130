"""Graph nodes.  Identity is the UUID: two Python objects with the same uuid are the same node in
different graphs (a model variable and its posterior replica), and a node compares equal to its uuid
string so that ``variables[model.X]`` finds the entry stored under ``model.X.uuid``
(mxfusion/components/model_component.py:53-57)."""
from uuid import uuid4


class ModelComponent(object):
    def __init__(self):
        self.name = None
        self._uuid = str(uuid4()).replace('-', '_')
        self.attributes = []
        self._in = []      # [(edge name, component)] incoming
        self._out = []     # [(edge name, component)] outgoing
        self.graph = None  # owning FactorGraph (informational)

    @property
    def uuid(self):
        return self._uuid

    def __hash__(self):
        return hash(self._uuid)

    def __eq__(self, other):
        if isinstance(other, ModelComponent):
            return self._uuid == other._uuid
        if isinstance(other, str):
            return self._uuid == other
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __repr__(self):
        return self._uuid

    # edges ---------------------------------------------------------------------------------------
    @property
    def predecessors(self):
        return list(self._in)

    @predecessors.setter
    def predecessors(self, edges):
        for _, node in self._in:
            node._out = [(n, c) for n, c in node._out if c is not self]
        self._in = list(edges)
        for name, node in self._in:
            node._out.append((name, self))

    @property
    def successors(self):
        return list(self._out)

    @successors.setter
    def successors(self, edges):
        for _, node in self._out:
            node._in = [(n, c) for n, c in node._in if c is not self]
        self._out = list(edges)
        for name, node in self._out:
            node._in.append((name, self))

    def as_json(self):
        return {'uuid': self._uuid, 'name': self.name, 'type': type(self).__name__,
                'attributes': [a.uuid for a in self.attributes]}
