"""BASELINE config 4: mean-field BNN (Normal posterior), 1-50-50-1 tanh MLP (bnn_regression.ipynb cell 6: H=50), S=3
Monte-Carlo samples, StochasticVariationalInference MC-ELBO under a MinibatchInferenceLoop with B=4096, through the
public API.  The CPU leg restates the reference's step (per-sample Python loop over the network,
function_evaluation.py:72-96; Normal log-pdfs; autograd; Adam with grads / B) in torch on the host cores.

    python scripts/bench_bnn.py [steps]      -> one JSON line (also gpurun_out/bnn_bench.json)
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, B, H, S = 100000, 4096, 50, 3
STEPS = int(sys.argv[1]) if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1].isdigit() else 200


def data():
    g = torch.Generator().manual_seed(0)
    x = torch.rand((N, 1), generator=g) * 2 - 1
    y = torch.sin(3 * x) + 0.05 * torch.randn((N, 1), generator=g)
    return x, y


def gpu_leg(x, y, steps=None, warm=10, data_resident=True, wall=False):
    STEPS = steps if steps is not None else globals()['STEPS']
    import mxfusion_b200 as mf
    from mxfusion_b200 import _lib
    from mxfusion_b200.components.distributions import Normal
    from mxfusion_b200.components.functions import MXFusionGluonFunction
    from mxfusion_b200.inference import (GradBasedInference, StochasticVariationalInference,
                                         create_Gaussian_meanfield, MinibatchInferenceLoop)
    dev = torch.device('cuda:0')
    mf.config.DEFAULT_DTYPE = 'float32'
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(1, H), torch.nn.Tanh(), torch.nn.Linear(H, H), torch.nn.Tanh(),
                              torch.nn.Linear(H, 1))
    m = mf.Model()
    m.N = mf.Variable()
    m.f = MXFusionGluonFunction(net, num_outputs=1, broadcastable=False)
    m.x = mf.Variable(shape=(m.N, 1))
    m.v = mf.Variable(shape=(1,), transformation=mf.components.PositiveTransformation(), initial_value=0.01)
    m.r = m.f(m.x)
    for _, v in m.r.factor.parameters.items():
        v.set_prior(Normal(mean=torch.tensor([0.]), variance=torch.tensor([1.])))
    m.y = Normal.define_variable(mean=m.r, variance=m.v, shape=(m.N, 1))
    observed = [m.y, m.x]
    q = create_Gaussian_meanfield(model=m, observed=observed)
    alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=observed)
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.y: N / float(B)}, rng=np.random.RandomState(0),
                                  data_resident=data_resident)
    infr = GradBasedInference(inference_algorithm=alg, grad_loop=loop, context=dev)
    infr.initialize(y=(N, 1), x=(N, 1))
    for _, v in m.r.factor.parameters.items():
        infr.params[q[v].factor.mean] = v.initial_value
        infr.params[q[v].factor.variance] = torch.full(v.shape, 1e-6)
    st = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def on_step(k, loss):
        if k == 1:
            st['first'] = float(loss)
        if k == warm:
            torch.cuda.synchronize()
            st['t0'] = time.perf_counter()
            e0.record()
        elif k == warm + STEPS:
            e1.record()
            torch.cuda.synchronize()
            st['t1'] = time.perf_counter()
            st['last'] = float(loss)
    l0 = _lib.launch_count()
    infr.run(max_iter=1 + (warm + STEPS) * B // N + 1, learning_rate=1e-3, max_steps=warm + STEPS, on_step=on_step,
             y=y, x=x)
    ms = (1e3 * (st['t1'] - st['t0']) if wall else e0.elapsed_time(e1)) / STEPS
    return dict(ms_per_step=ms, iters_per_s=1e3 / ms, first_loss=st['first'], last_loss=st['last'],
                launches_per_step=getattr(loop.last_stepper, 'launches_per_step', None),
                eager_launches=_lib.launch_count() - l0, h2d_bytes_per_step=loop.h2d_bytes_per_step,
                d2h_bytes_per_step=loop.d2h_bytes_per_step)


def cpu_leg(x, y, budget_s=10.0):
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    shapes = [(H, 1), (H,), (H, H), (H,), (1, H), (1,)]
    mu = [torch.randn(s) * 0.3 for s in shapes]
    rho = [torch.full(s, math.log(math.expm1(1e-6))) for s in shapes]     # softplus^-1(1e-6)
    nv = torch.tensor([math.log(math.expm1(0.01))])
    leaves = [t.requires_grad_() for t in mu + rho + [nv]]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    sp = torch.nn.functional.softplus

    def logn(v, mean, var):
        return -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(var) - torch.square(v - mean) / (2 * var)

    def step(i):
        sel = torch.randint(0, N, (B,))
        xb, yb = x[sel], y[sel]
        opt.zero_grad()
        var = [sp(r) for r in rho]
        ws = [m_.unsqueeze(0) + torch.randn((S,) + tuple(m_.shape)) * torch.sqrt(v_.unsqueeze(0)) for m_, v_ in zip(mu, var)]
        outs = []
        for s in range(S):                                               # function_evaluation.py:80-93
            h = torch.tanh(xb @ ws[0][s].T + ws[1][s])
            h = torch.tanh(h @ ws[2][s].T + ws[3][s])
            outs.append((h @ ws[4][s].T + ws[5][s]).unsqueeze(0))
        f = torch.cat(outs, 0)
        logp = (N / float(B)) * logn(yb.unsqueeze(0), f, sp(nv)).mean(0).sum()
        for w in ws:
            logp = logp + logn(w, torch.zeros(()), torch.ones(())).mean(0).sum()
        logq = sum(logn(w, m_.unsqueeze(0), v_.unsqueeze(0)).mean(0).sum() for w, m_, v_ in zip(ws, mu, var))
        loss = -(logp - logq)
        loss.backward()
        for t in leaves:
            t.grad /= B
        opt.step()
    for i in range(3):
        step(i)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s:
        step(n)
        n += 1
    dt = time.perf_counter() - t0
    return dict(cpu_iters_per_s=n / dt, cpu_iters=n, cpu_cores=os.cpu_count())


def main():
    x, y = data()
    out = dict(workload='mean-field BNN 1-%d-%d-1 tanh, S=%d, SVI MC-ELBO, minibatch=%d, N=%d, f32' % (H, H, S, B, N))
    out.update(gpu_leg(x, y))
    if '--no-cpu' not in sys.argv:
        out.update(cpu_leg(x, y))
        out['speedup_vs_cpu'] = out['iters_per_s'] / out['cpu_iters_per_s']
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'bnn_bench.json'), 'w'))


if __name__ == '__main__':
    main()
