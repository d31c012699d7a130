"""Sample-axis helpers (mxfusion/components/variables/runtime_variable.py:20-118).  Every runtime
array carries a leading sample axis of size 1 (not sampled) or S."""
import torch


def add_sample_dimension(F, array):
    return array.unsqueeze(0)


def add_sample_dimension_to_arrays(F, arrays, out=None):
    out = {} if out is None else out
    for uuid, v in arrays.items():
        out[uuid] = v.unsqueeze(0) if isinstance(v, torch.Tensor) else v
    return out


def expectation(F, array):
    """Mean over the sample axis (runtime_variable.py:53-60).  With a single sample the mean is the array itself: a view,
    not a kernel launch (the reference launches `F.mean` regardless; every 2 us launch counts in a 1.4 ms step)."""
    if array.shape[0] == 1:
        return array[0]
    return torch.mean(array, dim=0)


def is_sampled_array(F, array):
    return array.shape[0] > 1


def get_num_samples(F, array):
    return array.shape[0]


def as_samples(F, array, num_samples):
    if array.shape[0] == num_samples:
        return array
    return array.expand((num_samples,) + tuple(array.shape[1:]))


def arrays_as_samples(F, arrays):
    """Broadcast the sample axis of a nested list/dict of arrays to the largest one
    (runtime_variable.py:102-118).  `expand` is a stride-0 view: the kernels read it in place."""
    def leaves(a):
        if isinstance(a, dict):
            for v in a.values():
                yield from leaves(v)
        elif isinstance(a, (list, tuple)):
            for v in a:
                yield from leaves(v)
        else:
            yield a
    S = max(a.shape[0] for a in leaves(arrays))

    def conv(a):
        if isinstance(a, dict):
            return {k: conv(v) for k, v in a.items()}
        if isinstance(a, (list, tuple)):
            return [conv(v) for v in a]
        return as_samples(F, a, S)
    return conv(arrays)
