"""Parameter store of an inference run (mxfusion/inference/inference_parameters.py:27-252).

The reference keeps one Gluon Parameter per variable in a ParameterDict keyed by UUID.  Here all
trainable parameters live in ONE flat device buffer (`flat`) with a matching flat gradient buffer
(`gflat`) and Adam moments; each variable's tensor / gradient is a view.  One fused Adam launch then
updates everything, and data-parallel training all-reduces one contiguous bucket over NCCL/NVLink."""
from collections import OrderedDict
import warnings

import numpy as np
import torch

from ..components.model_component import ModelComponent
from ..components.variables.variable import Variable, VariableType
from ..common.config import get_default_dtype, get_default_device, torch_dtype
from ..common.exceptions import ModelSpecificationError


def realize_shape(shape, constants):
    """mxfusion/util/inference.py:48-59."""
    out = []
    for s in shape:
        if isinstance(s, (int, np.integer)):
            out.append(int(s))
        elif isinstance(s, Variable):
            if s.type == VariableType.CONSTANT:
                out.append(int(s.get_constant()))
            else:
                out.append(int(constants[s.uuid]))
        else:
            raise ModelSpecificationError("The shape of a Variable should either an integer or a Variable, but "
                                          "encountered {}!".format(type(s)))
    return tuple(out)


def discover_shape_constants(data_shapes, graphs):
    """mxfusion/util/inference.py:62-87: bind symbolic dimensions (m.N) from the data shapes."""
    shape_constants, variables = {}, {}
    for g in graphs:
        variables.update(g.variables)
    for var_id, shape in data_shapes.items():
        def_shape = variables[var_id].shape
        for s1, s2 in zip(def_shape, shape):
            if isinstance(s1, (int, np.integer)):
                if s1 != s2:
                    raise ModelSpecificationError(
                        "Variable ({}) shape mismatch between expected and found! s1 : {} s2 : {}".format(
                            str(variables[var_id]), str(s1), str(s2)))
            elif isinstance(s1, Variable):
                shape_constants[s1] = s2
            else:
                raise ModelSpecificationError("The shape of a Variable should either an integer or a Variable, "
                                              "but encountered {}!".format(str(type(s1))))
    return shape_constants


class Parameter(object):
    """One entry of the store: views into the flat buffers.  `tensor` is the autograd leaf."""

    def __init__(self, name, shape, init):
        self.name, self.shape, self.init = name, tuple(shape), init
        self.offset = None
        self.tensor = None
        self.grad_req = 'write'

    def data(self, ctx=None):
        return self.tensor.detach()

    def set_data(self, value):
        # `.data` shares storage but not the autograd version counter: values published between a forward and
        # its backward (the SET_ mechanism) must not invalidate tensors saved for that backward
        self.tensor.data.copy_(torch.as_tensor(value, dtype=self.tensor.dtype).to(self.tensor.device)
                               .reshape(self.tensor.shape))

    def grad(self):
        return self.tensor.grad


class InferenceParameters(object):
    def __init__(self, constants=None, dtype=None, context=None):
        self.dtype = dtype if dtype is not None else get_default_dtype()
        self.mxnet_context = context if context is not None else get_default_device()
        self._constants = {}
        self._var_ties = {}
        if constants is not None:
            self.update_constants(constants)
        self._params = OrderedDict()
        self._external = OrderedDict()      # carried-over parameters (TransferInference), not trained here
        self.flat = self.gflat = self.adam_m = self.adam_v = self.adam_t = None

    # constants ---------------------------------------------------------------------------------------
    def update_constants(self, constants):
        self._constants.update({(k.uuid if isinstance(k, ModelComponent) else k): v for k, v in constants.items()})

    @property
    def constants(self):
        return self._constants

    @property
    def var_ties(self):
        return self._var_ties

    @property
    def param_dict(self):
        d = OrderedDict(self._external)
        d.update(self._params)
        return d

    # allocation --------------------------------------------------------------------------------------
    def declare(self, var, constants=None):
        """Register a parameter variable (inference_parameters.py:81-86)."""
        if var.uuid in self._params or var.uuid in self._external:
            return
        shape = realize_shape(var.shape, self._constants if constants is None else constants)
        init = var.initial_value_before_transformation if var.initial_value is not None else None
        self._params[var.uuid] = Parameter(var.uuid, shape, init)

    def initialize_params(self, graphs, observed_uuid):
        """inference_parameters.py:63-90."""
        self._params = OrderedDict()
        for g in graphs:
            for var in g.get_constants():
                self._constants[var.uuid] = var.constant
            excluded = set(self._constants.keys()).union(observed_uuid).union(self._external.keys())
            for var in g.get_parameters(excluded=excluded):
                self.declare(var)
            for m in g.modules.values():
                m.initialize_hidden_parameters(self, excluded, self._constants)
        self._allocate()

    def initialize_with_carryover_params(self, graphs, observed_uuid, var_ties, carryover_params):
        """inference_parameters.py:92-136: reuse the values learned by a previous inference."""
        var_uuid = set()
        for g in graphs:
            var_uuid |= set(g.variables.keys())
            for m in g.modules.values():
                var_uuid |= set(m.hidden_parameters)
        self._external = OrderedDict()
        for carry in carryover_params:
            for uuid, p in carry.param_dict.items():
                if uuid in var_uuid:
                    if uuid in self._external:
                        warnings.warn('The variable with UUID ' + uuid + ' exists in multiple carryover parameter sets.')
                    self._external[uuid] = p
        self.initialize_params(graphs, set(observed_uuid))

    def _allocate(self):
        dt, dev = torch_dtype(self.dtype), self.mxnet_context
        total = 0
        # large parameters (qU_cov_W: M^2 elements) go to the END of the bucket: a data-parallel step all-reduces them early,
        # from inside the backward pass, and the remainder is then ONE contiguous prefix (inference/_stepper.py)
        order = sorted(self._params.values(), key=lambda q: (int(np.prod(q.shape)) if len(q.shape) else 1) >= 65536)
        for p in order:
            # every view starts on a 128-byte boundary: TMA / vectorised kernels read parameters in place
            total = (total + 31) & ~31
            p.offset = total
            total += int(np.prod(p.shape)) if len(p.shape) else 1
        self.flat = torch.zeros((total,), dtype=dt, device=dev)
        self.gflat = torch.zeros((total,), dtype=dt, device=dev)
        self.adam_m = torch.zeros((total,), dtype=dt, device=dev)
        self.adam_v = torch.zeros((total,), dtype=dt, device=dev)
        self.adam_t = torch.zeros((1,), dtype=torch.int32, device=dev)
        gen = torch.Generator(device='cpu').manual_seed(0)
        for p in self._params.values():
            n = int(np.prod(p.shape)) if len(p.shape) else 1
            view = self.flat[p.offset:p.offset + n].view(p.shape)
            if p.init is not None:
                init = torch.as_tensor(p.init, dtype=dt)
                view.copy_(init.expand(p.shape) if init.numel() != n else init.reshape(p.shape))
            else:
                # Gluon's default initialiser for `init=None` is Uniform(0.07)
                view.copy_(((torch.rand(p.shape, generator=gen, dtype=torch.float64) * 2 - 1) * 0.07).to(dt))
            p.tensor = view.data.requires_grad_(True)          # own version counter, storage shared with `flat`
            p.tensor.grad = self.gflat[p.offset:p.offset + n].view(p.shape)

    def refresh_leaves(self):
        """Re-wrap every parameter view as a brand-new autograd leaf.  A leaf's AccumulateGrad node remembers the
        stream it was created on and stays alive while any earlier result (e.g. a loss the user still holds)
        references it; fresh leaves let a loop run -- and be graph-captured -- on its own stream."""
        for p in self._params.values():
            n = p.tensor.numel()
            view = self.flat[p.offset:p.offset + n].view(p.shape)
            p.tensor = view.data.requires_grad_(True)
            p.tensor.grad = self.gflat[p.offset:p.offset + n].view(p.shape)

    # fused transforms / gradient gather (training loops) ---------------------------------------------------
    def setup_fused_transforms(self, var_trans):
        """Prepare the one-launch parameter plumbing of the training step: `transform_all_()` writes softplus(flat) for
        every Softplus-constrained parameter into `tflat` (the executor then reads those values through the leaves
        `tleaf`, skipping the per-parameter transform launches of inference_alg.py:79-80), and `pack_grads_()` gathers
        all leaf gradients into `gflat` with the softplus chain rule applied (instead of one accumulation launch per
        parameter plus one softplus adjoint per constrained parameter)."""
        from ..components.variables.var_trans import Softplus
        self.tflat = torch.zeros_like(self.flat)
        self._segments = []
        for uuid, p in self._params.items():
            n = p.tensor.numel()
            t = var_trans.get(uuid)
            fused = isinstance(t, Softplus)
            p.tleaf = None
            if fused:
                p.tleaf = self.tflat[p.offset:p.offset + n].view(p.shape).data.requires_grad_(True)
            self._segments.append((uuid, p, p.offset, n, 1 if fused else 0, float(t._offset) if fused else 0.0))
        self._fused_uuids = set(u for u, _, _, _, k, _ in self._segments if k == 1)

    def leaf(self, uuid):
        p = self._params[uuid]
        return p.tleaf if getattr(p, 'tleaf', None) is not None else p.tensor

    def transform_all_(self):
        seg = [s for s in self._segments if s[4] == 1]
        if seg:
            from .. import ops
            ops.R.params_transform(self.flat, self.tflat, [s[2] for s in seg], [s[3] for s in seg], [1] * len(seg),
                                   [s[5] for s in seg])

    def clear_leaf_grads(self):
        for _, p, _, _, _, _ in self._segments:
            p.tensor.grad = None
            if p.tleaf is not None:
                p.tleaf.grad = None

    def pack_grads_(self, out=None, skip_offset=None):
        """`out`: a bucket other than `gflat` (the data-parallel step packs straight into peer-mapped memory);
        `skip_offset`: the segment starting there is left alone (its gradient was written -- and exchanged -- earlier)."""
        from .. import ops
        seg = [s for s in self._segments if skip_offset is None or s[2] != skip_offset]
        grads = [(p.tleaf if k == 1 else p.tensor).grad for _, p, _, _, k, _ in seg]
        ops.R.params_pack_grads(self.flat, self.gflat if out is None else out, grads, [s[2] for s in seg], [s[3] for s in seg],
                                [s[4] for s in seg])

    def fix_all(self):
        for p in self._params.values():
            p.grad_req = 'null'

    # access ------------------------------------------------------------------------------------------
    def _get(self, uuid):
        if uuid in self._params:
            return self._params[uuid]
        return self._external[uuid]

    def __getitem__(self, key, ctx=None):
        if not isinstance(key, Variable):
            raise KeyError("The access key of inference parameter needs to be Variable, but got " +
                           str(type(key)) + ".")
        val = self._get(key.uuid).data()
        if key.transformation is not None:
            with torch.no_grad():
                val = key.transformation.transform(val)
        return val

    def __setitem__(self, key, item):
        if not isinstance(key, Variable):
            raise KeyError("The access key of inference parameter needs to be Variable, but get " +
                           str(type(key)) + ".")
        if key.uuid not in self._params and key.uuid not in self._external:
            # parameters published by an algorithm (SET_ prefix) are created on first assignment
            t = torch.as_tensor(item).detach()
            p = Parameter(key.uuid, tuple(t.shape), None)
            p.tensor = t.clone()
            self._external[key.uuid] = p
            return
        p = self._get(key.uuid)
        item = torch.as_tensor(item, dtype=p.tensor.dtype).detach()
        if key.transformation is not None:
            item = key.transformation.inverseTransform(item)
        if tuple(item.shape) != tuple(p.tensor.shape) and item.numel() != p.tensor.numel():
            p.tensor = item.clone().to(p.tensor.device)       # published values may change shape (cached L, X)
        else:
            p.set_data(item)

    def __contains__(self, k):
        return k in self.__dict__

    def get_serializable(self):
        params = {k: p.data().cpu().numpy() for k, p in self.param_dict.items()}
        tensor_consts = {k: v.detach().cpu().numpy() for k, v in self._constants.items()
                         if isinstance(v, torch.Tensor)}
        other = {k: v for k, v in self._constants.items() if k not in tensor_consts}
        return params, tensor_consts, other
