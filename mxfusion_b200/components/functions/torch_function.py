"""A torch.nn.Module used as a deterministic function inside a model: the role MXFusionGluonFunction
plays for Gluon blocks (mxfusion/components/functions/mxfusion_gluon_function.py:25-212).

Module parameters become model Variables `fn.parameters[<prefix><name>]` that can be given priors
(`m.f.parameters['fc1_weight']`-style access via `fn.parameters`).  With sampled weights the reference
loops over samples in Python (function_evaluation.py:72-96); here a Dense/tanh stack (the BNN notebooks' network) runs
all S sample networks in ONE fused CUDA launch forward and one backward (csrc/mlp.cu via ops.mlp_tanh); any other block
structure is evaluated as one batched `torch.func.functional_call` under `vmap` (library code, not claimed)."""
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
        self._dense_tanh = self._match_dense_tanh(block)

    @staticmethod
    def _match_dense_tanh(block):
        """Sequential(Linear, Tanh, Linear, ..., Linear) with <= 4 dense layers of width <= 64 (the reference's BNN
        notebooks, bnn_regression.ipynb cell 6): evaluated by the fused kernels of csrc/mlp.cu.  Returns the list of
        (weight key, bias key | None) per layer, or None when the block has any other structure."""
        if not isinstance(block, torch.nn.Sequential) or len(block) == 0 or len(block) % 2 == 0:
            return None
        layers = []
        for i, (name, mod) in enumerate(block.named_children()):
            if i % 2 == 0:
                if not isinstance(mod, torch.nn.Linear) or max(mod.in_features, mod.out_features) > 64:
                    return None
                layers.append((name + '_weight', name + '_bias' if mod.bias is not None else None))
            elif not isinstance(mod, torch.nn.Tanh):
                return None
        return layers if len(layers) <= 4 else None

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
        if self._dense_tanh is not None and len(inputs) == 1 and inputs[0].dim() == 3 and \
                not inputs[0].requires_grad and all(k in kw for pair in self._dense_tanh for k in pair if k):
            from ... import ops
            return ops.mlp_tanh(inputs[0], [kw[wk] for wk, _ in self._dense_tanh],
                                [kw[bk] if bk else None for _, bk in self._dense_tanh])

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
