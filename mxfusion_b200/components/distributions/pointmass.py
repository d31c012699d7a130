"""Point-mass 'distribution' used by MAP for latent variables (mxfusion/components/distributions/pointmass.py:22-77)."""
import torch

from .distribution import Distribution
from ..variables.variable import Variable


class PointMass(Distribution):
    def __init__(self, location, rand_gen=None, dtype=None, ctx=None):
        super(PointMass, self).__init__(inputs=[('location', location)], outputs=None, input_names=['location'],
                                        output_names=['random_variable'], rand_gen=rand_gen, dtype=dtype, ctx=ctx)

    def log_pdf_impl(self, location, random_variable, F=None):
        return torch.zeros((1,), dtype=location.dtype, device=location.device)      # pointmass.py:35-44

    def draw_samples_impl(self, location, rv_shape, num_samples=1, F=None):
        return location.expand((num_samples,) + tuple(location.shape[1:]))

    @staticmethod
    def define_variable(location, shape=None, rand_gen=None, dtype=None, ctx=None):
        p = PointMass(location=location, rand_gen=rand_gen, dtype=dtype, ctx=ctx)
        p.set_outputs([Variable(value=None, shape=shape if shape is not None else (1,))])
        return p.random_variable
