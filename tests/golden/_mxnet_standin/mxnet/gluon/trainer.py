"""gluon.Trainer with MXNet's Adam: rescale_grad = 1/batch_size, bias correction folded into the learning rate
(mx.optimizer.Adam, MXNet 1.x docs) -- the same restatement as oracle/loop.py:adam_step."""
import math

import torch


class Trainer(object):
    def __init__(self, params, optimizer, optimizer_params=None, kvstore='device'):
        if optimizer != 'adam':
            raise NotImplementedError(optimizer)
        self._params = list(params.values()) if hasattr(params, 'values') else list(params)
        op = optimizer_params or {}
        self.lr = op.get('learning_rate', 0.001)
        self.b1, self.b2, self.eps = op.get('beta1', 0.9), op.get('beta2', 0.999), op.get('epsilon', 1e-8)
        self._state = {}
        self._t = {}

    def step(self, batch_size, ignore_stale_grad=False):
        with torch.no_grad():
            for i, p in enumerate(self._params):
                d = p._data
                if d is None or d.grad is None or p.grad_req == 'null':
                    continue
                g = d.grad / batch_size
                m, v = self._state.get(i, (torch.zeros_like(d), torch.zeros_like(d)))
                t = self._t.get(i, 0) + 1
                m = self.b1 * m + (1 - self.b1) * g
                v = self.b2 * v + (1 - self.b2) * g * g
                lr_t = self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
                d.data = d.data - lr_t * m / (torch.sqrt(v) + self.eps)
                self._state[i], self._t[i] = (m, v), t
                d.grad = None
