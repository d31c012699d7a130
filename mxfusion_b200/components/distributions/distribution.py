"""Distribution factors (mxfusion/components/distributions/distribution.py:22-153)."""
from ..factor import Factor
from ..variables.variable import Variable
from ..variables.runtime_variable import arrays_as_samples
from .random_gen import MXNetRandomGenerator
from ...common.config import get_default_dtype


class Distribution(Factor):
    is_probabilistic = True

    def __init__(self, inputs, input_names, output_names, outputs=None, rand_gen=None, dtype=None, ctx=None):
        super(Distribution, self).__init__(inputs=inputs, outputs=outputs, input_names=input_names,
                                           output_names=output_names)
        self._rand_gen = MXNetRandomGenerator if rand_gen is None else rand_gen
        self.dtype = get_default_dtype() if dtype is None else dtype
        self.ctx = ctx
        self.log_pdf_scaling = 1

    @property
    def random_variable(self):
        return self.outputs[0][1]

    def replicate_self(self, attribute_map=None):
        rep = self.__class__.__new__(self.__class__)
        Factor.__init__(rep, None, None, list(self._input_names), list(self._output_names))
        rep._uuid = self._uuid
        rep._rand_gen, rep.dtype, rep.ctx, rep.log_pdf_scaling = self._rand_gen, self.dtype, self.ctx, 1
        return rep

    def log_pdf(self, F, variables, targets=None):
        """distribution.py:56-91: gather runtime inputs/outputs by UUID, broadcast samples, dispatch."""
        kw = self.fetch_runtime_inputs(variables)
        kw.update(self.fetch_runtime_outputs(variables))
        kw = arrays_as_samples(F, kw)
        return self.log_pdf_impl(F=F, **kw)

    def log_pdf_sum(self, F, variables):
        """F.sum(expectation(log_pdf)) (factor_graph.py:223); subclasses fuse it into one kernel."""
        import torch
        return torch.sum(torch.mean(self.log_pdf(F, variables), dim=0))

    def draw_samples(self, F, variables, num_samples=1, always_return_tuple=False, targets=None):
        kw = self.fetch_runtime_inputs(variables)
        kw = arrays_as_samples(F, kw)
        rv_shape = None
        out = self.draw_samples_impl(F=F, rv_shape=self._realized_shape(variables), num_samples=num_samples, **kw)
        if always_return_tuple and not isinstance(out, (list, tuple)):
            out = (out,)
        return out

    def _realized_shape(self, variables):
        """Output shape with symbolic dimensions replaced by the shape constants found in `variables`."""
        shape = []
        for s in self.random_variable.shape:
            if isinstance(s, Variable):
                s = variables[s.uuid] if s.uuid in variables else s.constant
            shape.append(int(s))
        return tuple(shape)

    def log_pdf_impl(self, F, **kwargs):
        raise NotImplementedError

    def draw_samples_impl(self, rv_shape, num_samples=1, F=None, **kwargs):
        raise NotImplementedError

    @staticmethod
    def define_variable(shape=None, rand_gen=None, dtype=None, ctx=None, **kwargs):
        raise NotImplementedError
