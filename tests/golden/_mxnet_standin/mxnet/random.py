import numpy as _np
import torch as _torch


def seed(s):
    _torch.manual_seed(int(s))
    _np.random.seed(int(s))
