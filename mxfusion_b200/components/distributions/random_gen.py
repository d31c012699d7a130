"""Random-number facade (mxfusion/components/distributions/random_gen.py:23-219).  The default
generator draws standard normals with the counter-based Philox stream inside the reparameterisation
kernel, so no separate RNG pass touches HBM; a mock generator injects fixed draws, which is how the
reference's tests make sampling deterministic (mxfusion/util/testutils.py:58-93)."""
import itertools

import torch

_COUNTER = itertools.count(1)
_SEED = [0]
_RANK_SALT = [0]             # mixed into the seed of in-kernel draws: data-parallel ranks draw DIFFERENT noise
_STEP_COUNTER = [None]       # device int32[1] bumped once per optimiser step (InferenceParameters.adam_t)


def set_step_counter(t):
    """Registers the optimiser's device step counter: in-kernel draws mix it into their Philox counter at run time,
    so a draw captured in a CUDA graph (inference/_stepper.py) is different on every replayed step."""
    _STEP_COUNTER[0] = t


def step_counter(device=None):
    t = _STEP_COUNTER[0]
    if t is None or (device is not None and t.device != torch.device(device)):
        return None
    return t


def set_rank(rank):
    """Data-parallel runs (inference/_stepper.py): every rank keeps the same (seed, offset) sequence, so without a salt
    all ranks would draw identical Monte-Carlo noise and adding ranks would not reduce its variance."""
    _RANK_SALT[0] = int(rank) * 0x9E3779B1


def seed(value):
    """Seed the in-kernel Philox stream (the reference seeds mx.random, conftest.py:21-25)."""
    _SEED[0] = int(value)
    global _COUNTER
    _COUNTER = itertools.count(1)


class RandomGenerator(object):
    @staticmethod
    def sample_normal(loc=0, scale=1, shape=None, dtype=None, out=None, ctx=None):
        raise NotImplementedError


class MXNetRandomGenerator(RandomGenerator):
    """Name kept from the reference.  `next_stream()` hands the kernel a fresh (seed, offset) pair."""
    in_kernel = True

    @staticmethod
    def next_stream():
        return (_SEED[0] + _RANK_SALT[0]) & 0x7FFFFFFFFFFFFFFF, next(_COUNTER)

    @staticmethod
    def sample_normal(loc=0, scale=1, shape=None, dtype=None, out=None, ctx=None):
        from ... import ops
        from ...common.config import torch_dtype, get_default_device
        dev = ctx if ctx is not None else get_default_device()
        m = torch.zeros((1,) + tuple(shape[1:]), dtype=torch_dtype(dtype), device=dev)
        v = torch.ones_like(m)
        s, o = MXNetRandomGenerator.next_stream()
        return ops.normal_draw(m, v, shape[0], seed=s, offset=o) * scale + loc


class MockMXNetRandomGenerator(RandomGenerator):
    """Returns the given samples reshaped to the requested shape (testutils.py:58-93)."""
    in_kernel = False

    def __init__(self, samples):
        self._samples = samples

    def sample_normal(self, loc=0, scale=1, shape=None, dtype=None, out=None, ctx=None):
        return self._samples.reshape(shape)
