"""Global defaults, as mutable module attributes exactly like the reference's
mxfusion/common/config.py:18-52 (notebooks set ``config.DEFAULT_DTYPE = 'float64'``).

`torch.device` stands in for the MXNet context.  The default device is ``cuda:<LOCAL_RANK>``; there
is no CPU execution path in this package, so a CPU device is accepted for building graphs but every
operator raises when handed CPU tensors.
"""
import os

import numpy as np
import torch

DEFAULT_DTYPE = 'float32'
MXNET_DEFAULT_DEVICE = None     # name kept from the reference; holds a torch.device (or None = auto)

_DT = {'float32': torch.float32, 'float64': torch.float64, np.float32: torch.float32, np.float64: torch.float64,
       np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64,
       torch.float32: torch.float32, torch.float64: torch.float64}


def get_default_dtype():
    return DEFAULT_DTYPE


def torch_dtype(dtype=None):
    dtype = DEFAULT_DTYPE if dtype is None else dtype
    try:
        return _DT[dtype]
    except (KeyError, TypeError):
        raise ValueError("unsupported dtype %r (float32 / float64 only)" % (dtype,))


def get_default_device():
    if MXNET_DEFAULT_DEVICE is not None:
        return torch.device(MXNET_DEFAULT_DEVICE)
    if torch.cuda.is_available():
        return torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    return torch.device('cpu')
