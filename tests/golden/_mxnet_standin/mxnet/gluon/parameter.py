from collections import OrderedDict

import numpy as np
import torch

from .. import initializer
from ..ndarray.ndarray import _wrap, _dt


class Parameter(object):
    def __init__(self, name, shape=None, dtype='float32', init=None, allow_deferred_init=False, grad_req='write'):
        self.name, self.shape, self.dtype, self.init, self.grad_req = name, shape, dtype, init, grad_req
        self._data = None

    def initialize(self, init=None, ctx=None, default_init=None, force_reinit=False):
        if self._data is not None and not force_reinit:
            return
        dt = _dt(self.dtype) or torch.float32
        ini = self.init
        if isinstance(ini, initializer.Constant):
            v = ini.value
            v = torch.as_tensor(np.asarray(v.detach() if isinstance(v, torch.Tensor) else v), dtype=dt)
            t = v.expand(tuple(self.shape)).clone() if tuple(v.shape) != tuple(self.shape) else v.clone()
        else:
            # Gluon's default for `init=None` is Uniform(0.07) (drawn from the NumPy global generator here)
            t = torch.as_tensor(np.random.uniform(-0.07, 0.07, size=tuple(self.shape)), dtype=dt)
        self._data = t.requires_grad_(self.grad_req != 'null')

    def data(self, ctx=None):
        return _wrap(self._data)

    def list_data(self):
        return [self.data()]

    def grad(self, ctx=None):
        return _wrap(self._data.grad)

    def set_data(self, value):
        v = torch.as_tensor(np.asarray(value.detach() if isinstance(value, torch.Tensor) else value))
        if self._data is None:
            self._data = v.to(_dt(self.dtype) or torch.float32).clone().requires_grad_(True)
            self.shape = tuple(v.shape)
        else:
            with torch.no_grad():
                self._data.data = v.to(self._data.dtype).reshape(self._data.shape).clone()

    def zero_grad(self):
        if self._data is not None and self._data.grad is not None:
            self._data.grad = None

    def _reduce(self):
        return self.data()

    def _load_init(self, data, ctx):
        self.set_data(data)


class ParameterDict(object):
    def __init__(self, prefix='', shared=None):
        self._params = OrderedDict()
        self.prefix = prefix

    def get(self, name, **kwargs):
        if name not in self._params:
            self._params[name] = Parameter(name, **kwargs)
        return self._params[name]

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        for p in self._params.values():
            p.initialize(ctx=ctx)

    def update(self, other):
        items = other.items() if hasattr(other, 'items') else other
        for k, v in items:
            self._params[k] = v

    def items(self):
        return self._params.items()

    def keys(self):
        return self._params.keys()

    def values(self):
        return self._params.values()

    def __getitem__(self, k):
        return self._params[k]

    def __contains__(self, k):
        return k in self._params

    def __iter__(self):
        return iter(self._params)

    def __len__(self):
        return len(self._params)

    def zero_grad(self):
        for p in self._params.values():
            p.zero_grad()
