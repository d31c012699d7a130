"""One optimisation step of the hot loop as a replayable unit.

The reference's loop body (minibatch_loop.py:81-92, batch_loop.py:52-60) is: record -> executor ->
backward -> Trainer.step(batch_size) -> loss.asscalar().  Here the forward+backward of a fixed-shape
step is captured once into a CUDA graph (about a hundred kernel launches replayed with one driver
call), the gradient bucket is all-reduced over NCCL when running data-parallel, and one fused Adam
launch updates the flat parameter buffer.  Nothing synchronises with the host."""
import torch
import torch.distributed as dist

from .. import ops
from ..common.exceptions import InferenceError


class Stepper(object):
    def __init__(self, infr_executor, params, optimizer, learning_rate, rescale_grad, example_batch,
                 use_cuda_graph=True, warmup_steps=2):
        if optimizer not in ('adam', 'sgd'):
            raise InferenceError("optimizer='adam' (the reference's default, grad_based_inference.py:67) and 'sgd' are "
                                 "implemented on the fused update path; got %r" % (optimizer,))
        self.optimizer = optimizer
        self.executor, self.params = infr_executor, params
        self.lr = float(learning_rate)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rescale = float(rescale_grad) / self.world          # all-reduce(SUM) / world = average over ranks
        params.refresh_leaves()
        # one-launch parameter transform on the way in, one-launch gradient gather (with the softplus chain rule) on the
        # way out; needs the executor's transformation table (ObjectiveBlock)
        self.fused = hasattr(infr_executor, '_var_trans') and hasattr(ops.R, 'params_pack_grads') and \
            len(getattr(infr_executor, '_var_ties', {})) == 0
        if self.fused:
            params.setup_fused_transforms(infr_executor._var_trans)
            infr_executor.pretransformed = params._fused_uuids
        from ..components.distributions import random_gen
        random_gen.set_step_counter(params.adam_t if params.adam_t.is_cuda else None)
        if self.world > 1:
            # replicas must start from the same point (e.g. SVGP's default inducing inputs are np.random.randn per process,
            # svgp_regression.py:318-320) and draw different Monte-Carlo noise
            for t in (params.flat, params.adam_m, params.adam_v, params.adam_t):
                dist.broadcast(t, src=0)
            random_gen.set_rank(dist.get_rank())
        self.static_in = [torch.empty_like(b) for b in example_batch]
        if self.static_in and self.static_in[0].is_cuda:
            ops.info_accumulator(self.static_in[0].device)        # created eagerly, never inside a graph capture
        self.loss = None
        self.graph = None
        self.use_graph = bool(use_cuda_graph) and self.static_in[0].is_cuda if self.static_in else False
        # forward+backward always runs on this side stream, so the autograd nodes that are created during the
        # eager warm-up steps live on the stream the capture will use
        self.stream = torch.cuda.Stream(device=self.static_in[0].device) if self.use_graph else None
        self.warmup_steps = warmup_steps
        # data parallel: the gradient of the largest parameter is all-reduced on `comm` while the backward pass continues;
        # with MXF_DP_INGRAPH=1 (opt-in) the collectives and the Adam update are captured in the step's CUDA graph as well
        import os
        self.comm = torch.cuda.Stream(device=self.static_in[0].device) if (self.world > 1 and self.static_in and
                                                                             self.static_in[0].is_cuda) else None
        # Gradient exchange: ONE kernel over NVLink peer memory (inference/_p2p.py, csrc/allreduce_p2p.cu) when every rank
        # could map its peers' buckets -- a plain launch, captured in the step's graph together with the Adam update, so the
        # data-parallel step is a single graph replay; NCCL (outside the graph) otherwise.
        self.p2p = None
        if self.world > 1 and self.fused and self.comm is not None:
            from ._p2p import try_peer_bucket
            self.p2p = try_peer_bucket(params.gflat.numel(), params.gflat.dtype, params.gflat.device)
        self.exchange = 'single' if self.world == 1 else ('p2p-kernel' if self.p2p is not None else 'nccl')
        self.in_graph = bool(self.use_graph) and self.world > 1 and (
            self.p2p is not None or os.environ.get('MXF_DP_INGRAPH', '0') == '1')
        # NCCL inside the graph is OFF by default: with torch 2.11 / NCCL 2.28 a dist.all_reduce issued from inside the captured backward pass hung
        # both ranks on a 2 x B200 box (capture or first replay), so the default data-parallel step
        # keeps the collective OUTSIDE the graph: replay -> all-reduce of the bucket -> fused Adam.  MXF_DP_EARLY=1 enables
        # the early reduce for eager (use_cuda_graph=False) steps.
        self.early_reduce_enabled = os.environ.get('MXF_DP_EARLY', '0') == '1' and self.p2p is None and \
            (self.in_graph or not self.use_graph)
        self._early, self._early_segment = None, None
        # Peer-memory exchange: the gradient of the LAST segment of the bucket (the largest parameters are laid out last:
        # qU_cov_W is 93 % of the SVGP bucket) is copied into the bucket and exchanged on `comm` as soon as the backward pass
        # has formed it; the kernel adjoints that follow hide its transfer AND the wait for the slowest rank.  Only the small
        # remainder is exchanged at the end of the step.  MXF_DP_EARLY=0 disables.
        self._peer_early_seg, self._peer_early_ok, self._peer_early_probe = None, False, None
        if self.p2p is not None and os.environ.get('MXF_DP_EARLY', '1') != '0':
            segs = [sg for sg in params._segments]
            total = params.gflat.numel()
            last = max(segs, key=lambda sg: sg[2]) if segs else None
            if last is not None and last[4] == 0 and 2 * last[3] >= total and \
                    sum(1 for sg in segs if sg[3] == last[3]) == 1 and (last[2] * params.gflat.element_size()) % 16 == 0:
                self._peer_early_seg, self._peer_early_ok = (last[2], last[3]), True
        self.n_calls = 0
        self.launches_per_step = None    # library kernels launched by one forward+backward (counted on an eager step)

    # ---- data-parallel plumbing ------------------------------------------------------------------------------------
    def _early_reduce(self, t):
        """Called from inside the backward pass (ops.set_early_reduce) with the gradient of the largest parameter: its
        all-reduce starts now, on the communication stream, and overlaps the rest of the backward pass."""
        if self._early is not None:
            return                                              # one tensor per step
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        self._early = (t.data_ptr(), t.numel())

    def _peer_early(self, t):
        """The peer-memory twin of _early_reduce: copy the freshly formed gradient into its bucket segment and start the
        exchange of that segment on `comm` (forked from the stream the gradient was formed on)."""
        if self._early is not None or not self._peer_early_ok or t.numel() != self._peer_early_seg[1]:
            return
        off, n = self._peer_early_seg
        self.p2p.buf[off:off + n].copy_(t.reshape(-1))
        if self._peer_early_probe is None and not torch.cuda.is_current_stream_capturing():
            self._peer_early_probe = t.detach().clone()       # first eager step: checked against the leaf gradient below
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self.p2p.all_reduce_range_(off, n, slot=1)
        self._early = (t.data_ptr(), t.numel())

    def _reduce_rest(self):
        """All-reduce of the gradient bucket, minus the segment that was reduced early."""
        p = self.params
        if self._early is None:
            dist.all_reduce(p.gflat, op=dist.ReduceOp.SUM)
            return
        seg = self._early_segment
        if seg is None:
            raise InferenceError("data-parallel step: the early-reduced gradient does not belong to a parameter segment")
        off, n = seg
        if off > 0:
            dist.all_reduce(p.gflat[:off], op=dist.ReduceOp.SUM)
        if off + n < p.gflat.numel():
            dist.all_reduce(p.gflat[off + n:], op=dist.ReduceOp.SUM)

    def _fwd_bwd(self):
        if self.fused:
            self.params.transform_all_()
            self.params.clear_leaf_grads()
            self._early, self._early_segment = None, None
            if self.world > 1 and self.comm is not None and self.early_reduce_enabled:
                ops.set_early_reduce(self._early_reduce)
            elif self.p2p is not None and self._peer_early_ok:
                ops.set_early_reduce(self._peer_early)
            try:
                loss, loss_for_gradient = self.executor(None, *self.static_in)
                loss_for_gradient.backward()
            finally:
                ops.set_early_reduce(None)
            if self.p2p is not None:
                self._pack_and_exchange_peer()
                return loss.detach().reshape(())
            if self._early is not None:
                torch.cuda.current_stream().wait_stream(self.comm)          # the reduced values are packed below
                for _, p, off, n, k, _ in self.params._segments:
                    g = (p.tleaf if k == 1 else p.tensor).grad
                    if g is not None and g.data_ptr() == self._early[0] and g.numel() == self._early[1]:
                        self._early_segment = (off, n)
            self.params.pack_grads_(out=None if self.p2p is None else self.p2p.grad)
            if self.world > 1 and self.in_graph:
                self._update()
            return loss.detach().reshape(())
        self.params.gflat.zero_()
        loss, loss_for_gradient = self.executor(None, *self.static_in)
        loss_for_gradient.backward()
        return loss.detach().reshape(())

    def _pack_and_exchange_peer(self):
        """End of a data-parallel step on the peer-memory path: pack the gradients into the symmetric bucket, exchange what
        has not been exchanged yet, update.  Inside the captured graph when there is one."""
        p, cur = self.params, torch.cuda.current_stream()
        early = self._early is not None
        if early and isinstance(self._peer_early_probe, torch.Tensor):
            # one-time check (first eager step): the tensor handed to the hook IS the parameter's whole gradient (a parameter
            # used by a second factor would receive further contributions after the hook ran)
            off, n = self._peer_early_seg
            leaf = [(q.tleaf if k == 1 else q.tensor).grad for _, q, o, _, k, _ in p._segments if o == off][0]
            self._peer_early_ok = leaf is not None and torch.equal(leaf.reshape(-1), self._peer_early_probe.reshape(-1))
            self._peer_early_probe = False
            if not self._peer_early_ok:
                cur.wait_stream(self.comm)                    # the early exchange must be over before the segment is rewritten
                early = False
        if early:
            off, n = self._peer_early_seg
            p.pack_grads_(out=self.p2p.grad, skip_offset=off)
            if off > 0:
                self.p2p.all_reduce_range_(0, off, slot=0)
            cur.wait_stream(self.comm)
        else:
            p.pack_grads_(out=self.p2p.grad)
            self.p2p.all_reduce_()
        self._apply(self.p2p.grad)

    def _update(self):
        p = self.params
        if self.p2p is not None:
            return                                            # done inside _fwd_bwd (_pack_and_exchange_peer)
        if self.world > 1:
            if self.fused:
                self._reduce_rest()
            else:
                dist.all_reduce(p.gflat, op=dist.ReduceOp.SUM)
        self._apply(p.gflat)

    def _apply(self, g):
        """One fused update of the flat parameter bucket (mx.gluon.Trainer(optimizer).step, minibatch_loop.py:91)."""
        p = self.params
        if self.optimizer == 'sgd':
            ops.R.sgd_step_(p.flat, g, None, p.adam_t, lr=self.lr, momentum=0.0, rescale=self.rescale)
        else:
            ops.R.adam_step_(p.flat, g, p.adam_m, p.adam_v, p.adam_t, lr=self.lr, rescale=self.rescale)

    def check_exchange(self):
        """Raises InferenceError when the peer-memory all-reduce gave up waiting for a rank (one D2H read)."""
        if self.p2p is not None:
            self.p2p.check()

    def step(self, batch=None):
        """Runs one step on `batch` (copied into the static inputs) and returns the device loss scalar."""
        if batch is not None:
            for dst, src in zip(self.static_in, batch):
                dst.copy_(src, non_blocking=True)
        self.n_calls += 1
        if not self.use_graph:
            c0 = ops.R.launch_count()
            self.loss = self._fwd_bwd()
            self.launches_per_step = ops.R.launch_count() - c0
        elif self.graph is None and self.n_calls <= self.warmup_steps:
            cur = torch.cuda.current_stream()                     # eager warm-up on the capture stream
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                c0 = ops.R.launch_count()
                self.loss = self._fwd_bwd()
                self.launches_per_step = ops.R.launch_count() - c0
            cur.wait_stream(self.stream)
        elif self.graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream):
                self._static_loss = self._fwd_bwd()
            self.graph = g
            g.replay()
            self.loss = self._static_loss
        else:
            self.graph.replay()
            self.loss = self._static_loss
        if not (self.world > 1 and self.in_graph and self.fused):
            self._update()
        return self.loss
