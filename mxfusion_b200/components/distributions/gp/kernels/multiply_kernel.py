"""MultiplyKernel: covariance by multiplying the covariance matrices of a list of kernels
(mxfusion/components/distributions/gp/kernels/multiply_kernel.py:19-87).  Each sub-kernel's matrix comes from its own CUDA path
(stationary: mxf_kbuild_fwd, linear: mxf_gemm); the combination is one elementwise pass."""
from .kernel import CombinationKernel


class MultiplyKernel(CombinationKernel):
    """Reference behaviour kept for parity (multiply_kernel.py:33-42): the constructor flattens ANY combination kernel
    among its operands -- so `(a + b) * c` evaluates as `a * b * c` with parameters `mul_a_*`, `mul_b_*`, `mul_c_*` -- and
    its default name is 'add' (`Kernel.multiply` passes name='mul')."""

    def __init__(self, sub_kernels, name='add', dtype=None, ctx=None):
        kernels = []
        for k in sub_kernels:
            if isinstance(k, CombinationKernel):
                kernels.extend(k.sub_kernels)
            else:
                kernels.append(k)
        super(MultiplyKernel, self).__init__(sub_kernels=kernels, name=name, dtype=dtype, ctx=ctx)

    def _compute_K(self, F, X, X2=None, **kernel_params):
        K = self.sub_kernels[0].K(F=F, X=X, X2=X2, **kernel_params)
        for k in self.sub_kernels[1:]:
            K = K * k.K(F=F, X=X, X2=X2, **kernel_params)
        return K

    def _compute_Kdiag(self, F, X, **kernel_params):
        K = self.sub_kernels[0].Kdiag(F=F, X=X, **kernel_params)
        for k in self.sub_kernels[1:]:
            K = K * k.Kdiag(F=F, X=X, **kernel_params)
        return K
