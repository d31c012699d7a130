"""The training-loop arithmetic restated in NumPy (test infrastructure).

``mxfusion/inference/minibatch_loop.py:65-92`` and ``batch_loop.py:46-60``:
shuffled minibatches with ``last_batch='rollover'``, Adam through
``mx.gluon.Trainer(...).step(batch_size)`` i.e. ``rescale_grad = 1/batch_size``.

MXNet/Gluon sources are not under /root/reference.  Restated from the MXNet 1.x
docs: ``gluon.data.RandomSampler`` shuffles ``np.arange(n)`` with
``np.random.shuffle`` once per epoch; ``BatchSampler(last_batch='rollover')``
keeps the short remainder and prepends it to the next epoch;
``mx.optimizer.Adam`` (beta1 .9, beta2 .999, eps 1e-8) applies the bias
correction to the learning rate:
``lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m/(sqrt(v)+eps)``.
"""
import numpy as np


class RolloverBatchSampler(object):
    """Integer work: must be reproduced bit-exactly by the product sampler."""

    def __init__(self, n, batch_size, rng):
        self.n, self.batch_size, self.rng = n, batch_size, rng
        self._prev = np.zeros((0,), dtype=np.int64)

    def epoch(self):
        """Yields the index batches of one epoch (minibatch_loop.py:68-70,78)."""
        idx = np.arange(self.n, dtype=np.int64)
        self.rng.shuffle(idx)
        idx = np.concatenate([self._prev, idx])
        nfull = idx.shape[0] // self.batch_size
        for b in range(nfull):
            yield idx[b * self.batch_size:(b + 1) * self.batch_size]
        self._prev = idx[nfull * self.batch_size:]


def adam_step(w, g, m, v, t, lr, rescale_grad=1.0, beta1=0.9, beta2=0.999, eps=1e-8):
    """One MXNet Adam update (t is 1-based).  Returns (w, m, v)."""
    g = g * rescale_grad
    m = beta1 * m + (1. - beta1) * g
    v = beta2 * v + (1. - beta2) * g * g
    lr_t = lr * np.sqrt(1. - beta2 ** t) / (1. - beta1 ** t)
    w = w - lr_t * m / (np.sqrt(v) + eps)
    return w, m, v
