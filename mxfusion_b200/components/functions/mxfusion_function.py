"""Function factors (mxfusion/components/functions/mxfusion_function.py, function_evaluation.py):
an `MXFusionFunction` describes a deterministic map; calling it on Variables creates a
`FunctionEvaluation` factor whose outputs are FUNCVAR nodes evaluated during the graph walk."""
from ..factor import Factor
from ..variables.variable import Variable
from ..variables.runtime_variable import arrays_as_samples
from ...common.config import get_default_dtype
from ...common.exceptions import ModelSpecificationError


class FunctionEvaluation(Factor):
    is_probabilistic = False

    def __init__(self, inputs, outputs, input_names, output_names, broadcastable=False):
        self.broadcastable = broadcastable
        super(FunctionEvaluation, self).__init__(inputs=inputs, outputs=outputs, input_names=input_names,
                                                 output_names=output_names)

    def replicate_self(self, attribute_map=None):
        rep = self.__class__.__new__(self.__class__)
        Factor.__init__(rep, None, None, list(self._input_names), list(self._output_names))
        rep._uuid = self._uuid
        rep.broadcastable = self.broadcastable
        rep.__dict__.update({k: v for k, v in self.__dict__.items() if k not in rep.__dict__})
        return rep

    def eval(self, F, variables, always_return_tuple=False):
        """function_evaluation.py:47-99: gather inputs by UUID, broadcast the sample axis, evaluate."""
        kw = arrays_as_samples(F, self.fetch_runtime_inputs(variables))
        out = self.eval_impl(F=F, **kw)
        if always_return_tuple and not isinstance(out, (list, tuple)):
            out = (out,)
        return out

    def eval_impl(self, F, **kwargs):
        raise NotImplementedError


class _BoundEvaluation(FunctionEvaluation):
    def __init__(self, func, inputs, input_names, output_names):
        self._func = func
        super(_BoundEvaluation, self).__init__(inputs=inputs, outputs=None, input_names=input_names,
                                               output_names=output_names, broadcastable=func.broadcastable)

    def eval_impl(self, F, **kwargs):
        return self._func.eval(F, **kwargs)

    @property
    def function(self):
        return self._func

    @property
    def parameters(self):
        """The function's parameter variables as wired into this evaluation (gluon_func_eval.py:26-38)."""
        names = set(self._func.parameters.keys())
        return {n: v for n, v in self.inputs if n in names}


class MXFusionFunction(object):
    """mxfusion_function.py:21-149."""

    def __init__(self, func_name, dtype=None, broadcastable=False):
        self.broadcastable = broadcastable
        self._func_name = func_name
        self.dtype = get_default_dtype() if dtype is None else dtype

    @property
    def name(self):
        return self._func_name

    @name.setter
    def name(self, value):
        self._func_name = value

    @property
    def parameters(self):
        return {}

    @property
    def input_names(self):
        raise NotImplementedError

    @property
    def output_names(self):
        raise NotImplementedError

    def eval(self, F, **input_kws):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        names = list(self.input_names)
        given = dict(zip(names, args))
        for k, v in kwargs.items():
            if k in given:
                raise ModelSpecificationError("The input " + k + " is given twice.")
            given[k] = v
        inputs = [(k, given[k]) for k in names if k in given] + list(self.parameters.items())
        fe = _BoundEvaluation(self, inputs, [k for k, _ in inputs], list(self.output_names))
        outs = [Variable(value=None, shape=None if not hasattr(self, 'output_shapes') else self.output_shapes[i])
                for i, _ in enumerate(self.output_names)]
        fe.set_outputs(outs)
        return outs[0] if len(outs) == 1 else tuple(outs)
