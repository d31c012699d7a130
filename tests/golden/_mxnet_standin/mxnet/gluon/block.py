from .parameter import Parameter, ParameterDict
from .. import ndarray as nd


class Block(object):
    def __init__(self, prefix=None, params=None):
        self.__dict__['_reg_params'] = {}
        self._params = params if params is not None else ParameterDict()
        self.prefix = prefix or ''

    def __setattr__(self, name, value):
        if isinstance(value, Parameter):
            self._reg_params[name] = value
        object.__setattr__(self, name, value)

    def collect_params(self, select=None):
        d = ParameterDict()
        d.update(self._params.items())
        d.update(self._reg_params.items())
        return d

    @property
    def params(self):
        return self._params

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        self.collect_params().initialize(ctx=ctx)

    def hybridize(self, active=True, **kw):
        pass

    def __call__(self, *args):
        return self.forward(*args)

    def forward(self, *args):
        raise NotImplementedError


class HybridBlock(Block):
    def forward(self, x, *args):
        params = {k: v.data() for k, v in self._reg_params.items()}
        return self.hybrid_forward(nd, x, *args, **params)
