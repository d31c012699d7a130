"""ArrayDataset + DataLoader(shuffle=True, last_batch='rollover'): RandomSampler shuffles np.arange(n) with
np.random.shuffle each epoch; the incomplete last batch is kept and prepended to the next epoch."""
import numpy as np
import torch

from ..ndarray.ndarray import _wrap


class ArrayDataset(object):
    def __init__(self, *arrays):
        self.arrays = arrays

    def __len__(self):
        return len(self.arrays[0])


class DataLoader(object):
    def __init__(self, dataset, batch_size=None, shuffle=False, last_batch='keep', **kw):
        self.ds, self.bs, self.shuffle, self.last_batch = dataset, batch_size, shuffle, last_batch
        self._prev = np.zeros((0,), dtype=np.int64)

    def __iter__(self):
        n = len(self.ds)
        idx = np.arange(n)
        if self.shuffle:
            np.random.shuffle(idx)
        if self.last_batch == 'rollover':
            idx = np.concatenate([self._prev, idx])
        nfull = len(idx) // self.bs
        for b in range(nfull):
            sel = torch.as_tensor(idx[b * self.bs:(b + 1) * self.bs])
            yield [_wrap(torch.Tensor(a)[sel]) for a in self.ds.arrays]
        rest = idx[nfull * self.bs:]
        if self.last_batch == 'rollover':
            self._prev = rest
        elif self.last_batch == 'keep' and len(rest):
            sel = torch.as_tensor(rest)
            yield [_wrap(torch.Tensor(a)[sel]) for a in self.ds.arrays]
