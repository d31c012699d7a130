"""Product of kernels (mxfusion/components/distributions/gp/kernels/multiply_kernel.py:19-87).

Reference behaviour kept for parity (multiply_kernel.py:33-42): the constructor absorbs the sub-kernels of ANY combination
kernel among its operands -- so `(a + b) * c` evaluates as `a * b * c` with parameters `mul_a_*`, `mul_b_*`, `mul_c_*` --
and its default name is 'add' (`Kernel.multiply` passes name='mul')."""
import operator

from .kernel import CombinationKernel, _FoldKernel


class MultiplyKernel(_FoldKernel):
    OP = staticmethod(operator.mul)
    FLATTEN = (CombinationKernel,)

    def __init__(self, sub_kernels, name='add', dtype=None, ctx=None):
        super(MultiplyKernel, self).__init__(sub_kernels, name=name, dtype=dtype, ctx=ctx)
