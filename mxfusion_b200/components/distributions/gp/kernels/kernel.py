"""Kernel base classes (mxfusion/components/distributions/gp/kernels/kernel.py:25-373)."""
from copy import copy

import torch

from ....functions.mxfusion_function import MXFusionFunction
from ....variables.variable import Variable
from .....common.exceptions import ModelSpecificationError


def slice_axis(F, array, axis, indices):
    """active_dims gather along one axis (mxfusion/util/util.py:23-62); index work, bit-exact."""
    idx = torch.as_tensor(list(indices), dtype=torch.long, device=array.device)
    return torch.index_select(array, axis if axis >= 0 else array.dim() + axis, idx)


class Kernel(MXFusionFunction):
    broadcastable = False

    def __init__(self, input_dim, name, active_dims=None, dtype=None, ctx=None):
        self.__dict__['_parameter_names'] = []
        super(Kernel, self).__init__(func_name=name, dtype=dtype, broadcastable=self.broadcastable)
        self.input_dim = input_dim
        self.ctx = ctx
        self.active_dims = active_dims

    def __setattr__(self, name, value):
        if isinstance(value, Variable) and name not in self._parameter_names:
            self._parameter_names.append(name)
        super(Kernel, self).__setattr__(name, value)

    @property
    def local_parameters(self):
        return {getattr(self, n) for n in self._parameter_names}

    @property
    def parameters(self):
        raise NotImplementedError

    @property
    def input_names(self):
        return ['X', 'X2']

    @property
    def output_names(self):
        return ['covariance']

    def _strip(self, kernel_params):
        off = len(self.name) + 1
        return {k[off:]: v for k, v in kernel_params.items() if k.startswith(self.name + '_')}

    def K(self, F, X, X2=None, **kernel_params):
        """kernel.py:96-123."""
        params = self._strip(kernel_params)
        if self.active_dims is not None:
            X = slice_axis(F, X, -1, self.active_dims)
            if X2 is not None:
                X2 = slice_axis(F, X2, -1, self.active_dims)
        return self._compute_K(F=F, X=X, X2=X2, **params)

    def Kdiag(self, F, X, **kernel_params):
        """kernel.py:125-147."""
        params = self._strip(kernel_params)
        if self.active_dims is not None:
            X = slice_axis(F, X, -1, self.active_dims)
        return self._compute_Kdiag(F=F, X=X, **params)

    def _compute_K(self, F, X, X2=None, **kernel_params):
        raise NotImplementedError

    def _compute_Kdiag(self, F, X, **kernel_params):
        raise NotImplementedError

    def fetch_parameters(self, params):
        """kernel.py:232-245."""
        return {n: params[v.uuid] for n, v in self.parameters.items()}

    def eval(self, F, X, X2=None, **kernel_params):
        return self.K(F, X, X2, **kernel_params)

    def replicate_self(self, attribute_map=None):
        rep = copy(self)
        rep.__dict__['_parameter_names'] = []
        for n in self._parameter_names:
            setattr(rep, n, getattr(self, n).replicate_self(attribute_map))
        rep.active_dims = copy(self.active_dims)
        return rep

    def add(self, other, name='add'):
        """kernel.py:149-160."""
        if not isinstance(other, Kernel):
            raise ModelSpecificationError("Only a Gaussian Process Kernel can be added to a Gaussian Process Kernel.")
        from .add_kernel import AddKernel
        return AddKernel([self, other], name=name, ctx=self.ctx, dtype=self.dtype)

    def __add__(self, other):
        return self.add(other)

    def multiply(self, other, name='mul'):
        """kernel.py:168-181."""
        if not isinstance(other, Kernel):
            raise ModelSpecificationError(
                "Only a Gaussian Process Kernel can be multiplied with a Gaussian Process Kernel.")
        from .multiply_kernel import MultiplyKernel
        return MultiplyKernel([self, other], name=name, ctx=self.ctx, dtype=self.dtype)

    def __mul__(self, other):
        return self.multiply(other)


class NativeKernel(Kernel):
    @property
    def parameters(self):
        return {self.name + '_' + n: getattr(self, n) for n in self._parameter_names}

    @property
    def parameter_names(self):
        return [self.name + '_' + n for n in self._parameter_names]


def rename_duplicate_names(names):
    """['a', 'b', 'a', 'a'] -> [(2, 'a0'), (3, 'a1')] (mxfusion/util/util.py:65-99)."""
    import re
    all_names = set(names)
    if len(all_names) == len(names):
        return []
    cur_names, renames = set(), []
    prog = re.compile(r'^(.*)(\d+)$')
    for i, n in enumerate(names):
        if n in cur_names:
            res = prog.match(n)
            if res is None or len(res.groups()) == 0:
                prefix, count = n, 0
            else:
                prefix, count = res.groups()[0], int(res.groups()[1]) + 1
            while prefix + str(count) in all_names:
                count += 1
            renames.append((i, prefix + str(count)))
            all_names.add(prefix + str(count))
        else:
            cur_names.add(n)
    return renames


class CombinationKernel(Kernel):
    """Covariance computed by combining the covariance matrices of sub-kernels (kernel.py:309-373).  Parameter names
    nest: `<combination>_<sub-kernel>_<parameter>`."""

    def __init__(self, sub_kernels, name, dtype=None, ctx=None):
        input_dim = max([k.input_dim for k in sub_kernels])
        for i, n in rename_duplicate_names([k.name for k in sub_kernels]):
            sub_kernels[i].name = n
        super(CombinationKernel, self).__init__(input_dim=input_dim, name=name, dtype=dtype, ctx=ctx)
        self.__dict__['sub_kernels'] = sub_kernels
        for k in sub_kernels:
            self.__dict__[k.name] = k

    @property
    def local_parameters(self):
        out = set()
        for k in self.sub_kernels:
            out |= set(k.local_parameters)
        return out

    @property
    def parameters(self):
        p = {}
        for k in self.sub_kernels:
            p.update(k.parameters)
        return {self.name + '_' + k: v for k, v in p.items()}

    @property
    def parameter_names(self):
        names = []
        for k in self.sub_kernels:
            names.extend([self.name + '_' + n for n in k.parameter_names])
        return names

    def replicate_self(self, attribute_map=None):
        rep = copy(self)
        rep.__dict__['_parameter_names'] = []
        rep.__dict__['sub_kernels'] = [k.replicate_self(attribute_map) for k in self.sub_kernels]
        for k in rep.sub_kernels:
            rep.__dict__[k.name] = k
        rep.active_dims = copy(self.active_dims)
        return rep


class _FoldKernel(CombinationKernel):
    """Combination by folding a binary operator over the sub-kernels' covariance matrices / diagonals.  Each sub-kernel's
    matrix comes from its own CUDA path (stationary: mxf_kbuild_fwd, linear: mxf_gemm); the fold is one elementwise pass
    per operand.  `FLATTEN` names the classes whose sub-kernels are absorbed at construction."""
    OP = None
    FLATTEN = ()

    def __init__(self, sub_kernels, name, dtype=None, ctx=None):
        flat = []
        for k in sub_kernels:
            flat.extend(k.sub_kernels if isinstance(k, self.FLATTEN) else [k])
        super(_FoldKernel, self).__init__(sub_kernels=flat, name=name, dtype=dtype, ctx=ctx)

    def _fold(self, parts):
        from functools import reduce
        return reduce(type(self).OP, parts)

    def _compute_K(self, F, X, X2=None, **kernel_params):
        return self._fold([k.K(F=F, X=X, X2=X2, **kernel_params) for k in self.sub_kernels])

    def _compute_Kdiag(self, F, X, **kernel_params):
        return self._fold([k.Kdiag(F=F, X=X, **kernel_params) for k in self.sub_kernels])
