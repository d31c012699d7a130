"""Sum of kernels (mxfusion/components/distributions/gp/kernels/add_kernel.py:19-88): nested sums are flattened."""
import operator

from .kernel import _FoldKernel


class AddKernel(_FoldKernel):
    OP = staticmethod(operator.add)

    def __init__(self, sub_kernels, name='add', dtype=None, ctx=None):
        super(AddKernel, self).__init__(sub_kernels, name=name, dtype=dtype, ctx=ctx)


AddKernel.FLATTEN = (AddKernel,)
