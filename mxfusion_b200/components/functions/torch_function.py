"""A torch.nn.Module used as a deterministic function inside a model: the role MXFusionGluonFunction
plays for Gluon blocks (mxfusion/components/functions/mxfusion_gluon_function.py:25-212).

Module parameters become model Variables `fn.parameters[<prefix><name>]` that can be given priors
(`m.f.parameters['fc1_weight']`-style access via `fn.parameters`).  With sampled weights the reference
loops over samples in Python (function_evaluation.py:72-96); here the S sample networks run as one
batched call (`torch.func.functional_call` under `vmap`)."""
import torch
from torch.func import functional_call, vmap

from .mxfusion_function import MXFusionFunction
from ..variables.variable import Variable
from ...common.exceptions import ModelSpecificationError


class MXFusionTorchFunction(MXFusionFunction):
    def __init__(self, block, num_outputs, dtype=None, broadcastable=False, name=None):
        if not isinstance(block, torch.nn.Module):
            raise ModelSpecificationError("The block argument must be a torch.nn.Module (stands in for a Gluon block).")
        super(MXFusionTorchFunction, self).__init__(func_name=name or type(block).__name__.lower(), dtype=dtype,
                                                    broadcastable=broadcastable)
        self._block = block
        self.num_outputs = num_outputs
        self._params = {}
        for pname, p in block.named_parameters():
            key = pname.replace('.', '_')
            v = Variable(shape=tuple(p.shape), isInherited=True, initial_value=p.detach().clone())
            v.inherited_name = pname
            self._params[key] = v
        self._key_to_pname = {k: v.inherited_name for k, v in self._params.items()}

    @property
    def block(self):
        return self._block

    @property
    def parameters(self):
        return self._params

    @property
    def input_names(self):
        import inspect
        sig = inspect.signature(self._block.forward)
        return [p for p in sig.parameters]

    @property
    def output_names(self):
        return [self.name + "_output_" + str(i) for i in range(self.num_outputs)]

    def eval(self, F, **kw):
        inputs = [kw[n] for n in self.input_names if n in kw]
        weights = {self._key_to_pname[k]: kw[k] for k in self._params if k in kw}
        S = max([t.shape[0] for t in inputs] + [w.shape[0] for w in weights.values()] + [1])

        def one(ws, *xs):
            return functional_call(self._block, ws, xs)
        in_dims_w = {k: (0 if w.shape[0] == S and S > 1 else None) for k, w in weights.items()}
        ws = {k: (w if in_dims_w[k] == 0 else w[0]) for k, w in weights.items()}
        in_dims_x = tuple(0 if (x.shape[0] == S and S > 1) else None for x in inputs)
        xs = tuple(x if d == 0 else x[0] for x, d in zip(inputs, in_dims_x))
        if S == 1:
            out = one(ws, *xs)
            return out.unsqueeze(0) if not isinstance(out, (tuple, list)) else tuple(o.unsqueeze(0) for o in out)
        return vmap(one, in_dims=(in_dims_w,) + in_dims_x)(ws, *xs)


MXFusionGluonFunction = MXFusionTorchFunction     # reference name kept as an alias
