"""Full-batch gradient loop (mxfusion/inference/batch_loop.py:19-61)."""
from .grad_loop import GradLoop
from ._stepper import Stepper
from .. import ops


class BatchInferenceLoop(GradLoop):
    def __init__(self, use_cuda_graph=True):
        self.use_cuda_graph = use_cuda_graph

    def run(self, infr_executor, data, param_dict, ctx, optimizer='adam', learning_rate=1e-3, max_iter=1000,
            n_prints=10, verbose=False):
        """One step per iteration on the whole data; `Trainer.step(batch_size=1)` (batch_loop.py:60)."""
        data = [d.to(ctx) for d in data]
        stepper = Stepper(infr_executor, param_dict, optimizer, learning_rate, 1.0, data,
                          use_cuda_graph=self.use_cuda_graph)
        for dst, src in zip(stepper.static_in, data):
            dst.copy_(src)
        iter_step = max(max_iter // n_prints, 1)
        loss = None
        check_every = max(1, min(100, iter_step))
        dev = data[0].device if data else None
        for i in range(max_iter):
            loss = stepper.step()
            if dev is not None and ((i + 1) % check_every == 0 or i == max_iter - 1):
                # non-PD factorisations are recorded on the device; read back every few steps (no per-step sync)
                ops.check_factorisations(dev, "a Cholesky factorisation at iteration %d" % (i + 1))
                stepper.check_exchange()
            if verbose:
                print('\rIteration {} loss: {}\t\t\t\t'.format(i + 1, float(loss)), end='')
                if ((i + 1) % iter_step == 0 and i > 0) or i == max_iter - 1:
                    print()
        self.last_stepper = stepper
        return loss
