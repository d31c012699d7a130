"""Minibatch gradient loop (mxfusion/inference/minibatch_loop.py:21-95).

Semantics kept from the reference: `max_iter` counts EPOCHS (:75); batches come from a shuffled
sampler with ``last_batch='rollover'`` (:68-70); gradients are rescaled by 1/batch_size
(`trainer.step(batch_size=...)`, :90-91); `rv_scaling` re-weights the likelihood of the listed random
variables (:39).  B200-side: the data set is copied to HBM once and each step's rows are gathered on
the device by a kernel reading the epoch's permutation (`data_resident=True`, default), or -- when
`data_resident=False` -- gathered on the host into pinned memory and streamed host-to-device every
step.  The loss is accumulated on the device; the host reads it once per epoch (or per step when
verbose, as the reference always does, :92)."""
import numpy as np
import torch
import torch.distributed as dist

from .grad_loop import GradLoop
from ._stepper import Stepper
from .. import ops


class RolloverBatchSampler(object):
    """Index batches of one epoch after another: shuffle `arange(n)` with the NumPy global generator (what
    gluon's RandomSampler does), prepend the remainder kept from the previous epoch, emit the full batches,
    keep the new remainder (BatchSampler(last_batch='rollover')).  Integer work: bit-exact."""

    def __init__(self, n, batch_size, rng=None, shuffle=True):
        self.n, self.batch_size, self.shuffle = int(n), int(batch_size), shuffle
        self.rng = np.random if rng is None else rng
        self._prev = np.zeros((0,), dtype=np.int64)

    def epoch_indices(self):
        idx = np.arange(self.n, dtype=np.int64)
        if self.shuffle:
            self.rng.shuffle(idx)
        idx = np.concatenate([self._prev, idx])
        nfull = idx.shape[0] // self.batch_size
        self._prev = idx[nfull * self.batch_size:].copy()
        return idx[:nfull * self.batch_size], nfull


class MinibatchInferenceLoop(GradLoop):
    def __init__(self, batch_size=100, rv_scaling=None, data_resident=True, use_cuda_graph=True, rng=None):
        super(MinibatchInferenceLoop, self).__init__()
        self.batch_size = batch_size
        self.rv_scaling = {v.uuid: s for v, s in rv_scaling.items()} if rv_scaling is not None else rv_scaling
        self.data_resident = data_resident
        self.use_cuda_graph = use_cuda_graph
        self.rng = rng
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 0

    def run(self, infr_executor, data, param_dict, ctx, optimizer='adam', learning_rate=1e-3, max_iter=1000,
            verbose=False, update_shape_constants=None, max_steps=None, on_step=None):
        B = self.batch_size
        n = data[0].shape[0]
        sampler = RolloverBatchSampler(n, B, rng=self.rng)
        dev = torch.device(ctx)
        cols = [int(np.prod(d.shape[1:])) for d in data]
        if self.data_resident:
            src = [d.to(dev).reshape(n, c).contiguous() for d, c in zip(data, cols)]
            # two index buffers (+ pinned staging): the next epoch's permutation is shuffled on the host and uploaded while
            # the device still works through the steps queued for the current epoch
            idx_devs = [torch.empty((n + B,), dtype=torch.int64, device=dev) for _ in range(2)]
            idx_pin = [torch.empty((n + B,), dtype=torch.int64) for _ in range(2)]
            if dev.type == 'cuda':
                idx_pin = [t.pin_memory() for t in idx_pin]
            off = torch.zeros((1,), dtype=torch.int64, device=dev)
        else:
            src = [d.detach().cpu().reshape(n, c).contiguous() for d, c in zip(data, cols)]
            # ring of pinned staging buffers: a slot is refilled by the host only after the H2D copy that read it
            # has completed (event), so the host can run ahead of the device by RING-1 steps -- far enough (up to 64 steps,
            # at most 64 MB of pinned memory) that the host-side shuffle of the next epoch's permutation (tens of ms at
            # N = 1e6) is covered by steps already queued
            step_bytes = sum(B * c * d.element_size() for d, c in zip(src, cols))
            RING = int(max(4, min(64, (64 << 20) // max(step_bytes, 1))))
            pinned = [[torch.empty((B, c), dtype=d.dtype).pin_memory() if dev.type == 'cuda'
                       else torch.empty((B, c), dtype=d.dtype) for d, c in zip(src, cols)] for _ in range(RING)]
            copied = [torch.cuda.Event() if dev.type == 'cuda' else None for _ in range(RING)]
            self.h2d_bytes_per_step = sum(p.numel() * p.element_size() for p in pinned[0])
        example = [torch.empty((B,) + tuple(d.shape[1:]), dtype=d.dtype, device=dev) for d in data]
        if update_shape_constants is not None:
            update_shape_constants(example)
        stepper = Stepper(infr_executor, param_dict, optimizer, learning_rate, 1.0 / B, example,
                          use_cuda_graph=self.use_cuda_graph)
        flat_in = [s.reshape(B, c) for s, c in zip(stepper.static_in, cols)]
        loss_acc = torch.zeros((), dtype=example[0].dtype, device=dev)
        loss_host = torch.empty((1,), dtype=example[0].dtype)
        if dev.type == 'cuda':
            loss_host = loss_host.pin_memory()
        self.d2h_bytes_per_step = loss_host.element_size()
        epoch_losses, steps_done = [], 0

        def stage(e, idx):
            """Epoch e's indices -> the device (resident data) or a host tensor; stream-ordered, no host wait."""
            if self.data_resident:
                k = idx.shape[0]
                idx_pin[e & 1][:k].copy_(torch.from_numpy(idx))
                idx_devs[e & 1][:k].copy_(idx_pin[e & 1][:k], non_blocking=True)
                return idx_devs[e & 1]
            return torch.from_numpy(idx)

        idx, nfull = sampler.epoch_indices()
        staged = stage(0, idx)
        for e in range(max_iter):
            if self.data_resident:
                idx_dev = staged
            else:
                idx_t = staged
            loss_acc.zero_()
            for i in range(nfull):
                if self.data_resident:
                    off.fill_(i * B)
                    for s, dst in zip(src, flat_in):
                        ops.R.gather_rows(s, idx_dev, off, B, out=dst)
                    loss = stepper.step()
                else:
                    sel = idx_t[i * B:(i + 1) * B]
                    slot = steps_done % RING
                    if copied[slot] is not None and steps_done >= RING:
                        copied[slot].synchronize()
                    for s, p, dst in zip(src, pinned[slot], flat_in):
                        torch.index_select(s, 0, sel, out=p)
                        dst.copy_(p, non_blocking=True)
                    if copied[slot] is not None:
                        copied[slot].record()
                    loss = stepper.step()
                    loss_host.copy_(loss.reshape(1), non_blocking=True)       # the reference's asscalar()
                loss_acc += loss
                steps_done += 1
                if on_step is not None:
                    on_step(steps_done, loss)
                if verbose:
                    print('\repoch {} Iteration {} loss: {}\t\t\t'.format(e + 1, i + 1, float(loss)), end='')
                if max_steps is not None and steps_done >= max_steps:
                    break
            nfull_done = nfull
            last = (e + 1 >= max_iter) or (max_steps is not None and steps_done >= max_steps)
            if not last:
                # the reference shuffles at the start of the next epoch (DataLoader, minibatch_loop.py:68-70); same draws in
                # the same order here, only earlier on the host's clock: the device is still busy with this epoch's steps
                idx, nfull = sampler.epoch_indices()
                staged = stage(e + 1, idx)
            if verbose:
                print('epoch-loss: {} '.format(float(loss_acc) / max(nfull_done, 1)))
            epoch_losses.append(loss_acc / max(nfull_done, 1))
            # a non-positive-definite factorisation is recorded on the device by the bounds; surfaced once per epoch
            # (the reference gets an MXNetError at its per-step asscalar(), minibatch_loop.py:92)
            ops.check_factorisations(dev, "a Cholesky factorisation during epoch %d" % (e + 1))
            stepper.check_exchange()
            if max_steps is not None and steps_done >= max_steps:
                break
        self.last_stepper = stepper
        return epoch_losses
