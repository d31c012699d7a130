"""mx.operator.CustomOp / CustomOpProp / register + the machinery F.Custom needs."""
import torch

_REGISTRY = {}


class CustomOp(object):
    def assign(self, dst, req, src):
        if req == 'null':
            return
        with torch.no_grad():
            if req in ('write', 'inplace'):
                torch.Tensor(dst).copy_(torch.Tensor(src))
            elif req == 'add':
                torch.Tensor(dst).add_(torch.Tensor(src))


class CustomOpProp(object):
    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad

    def list_arguments(self):
        return ['data']

    def list_outputs(self):
        return ['output']

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[0]], []

    def infer_type(self, in_type):
        return in_type, [in_type[0]] * len(self.list_outputs()), []


def register(name):
    def deco(cls):
        _REGISTRY[name] = cls
        return cls
    return deco


def custom(x, op_type=None, **kwargs):
    from .ndarray.ndarray import NDArray, _wrap
    kwargs.pop('name', None)          # `name=` is the symbol name in MXNet, not an operator argument
    prop = _REGISTRY[op_type](**{k: str(v) for k, v in kwargs.items()})
    _, out_shapes, _ = prop.infer_shape([list(x.shape)])
    op = prop.create_operator(None, [tuple(x.shape)], [x.dtype])

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp):
            out = _wrap(torch.zeros(tuple(out_shapes[0]), dtype=inp.dtype))
            with torch.no_grad():
                op.forward(True, ['write'], [_wrap(inp.detach())], [out], [])
            ctx.save_for_backward(inp, torch.Tensor(out))
            return torch.Tensor(out).clone()

        @staticmethod
        def backward(ctx, g):
            inp, out = ctx.saved_tensors
            gin = _wrap(torch.zeros_like(inp))
            with torch.no_grad():
                op.backward(['write'], [_wrap(g)], [_wrap(inp)], [_wrap(out)], [gin], [])
            return torch.Tensor(gin).clone()
    return _wrap(_Fn.apply(torch.Tensor(x) if isinstance(x, NDArray) else x))
