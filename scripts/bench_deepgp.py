"""BASELINE config 5: two-layer deep GP, N=100k, M=256 inducing points per layer, RBF, doubly-stochastic VI (S Monte-Carlo
samples of the hidden layer per step), data-sharded over the GPUs of one box like the SVGP configs (one NCCL all-reduce of
the flat gradient per step).  There is no reference implementation; the CPU leg is the same bound written
with torch CPU LAPACK/BLAS ops + autograd (oracle-style restatement of THIS repo's composition, labelled as such).
Unspecified by BASELINE.json and fixed here: D=8 inputs, hidden width 8 (identity mean), P=1, minibatch 2048 per GPU, S=4."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, D, H, M, B, S = 100000, 8, 8, 256, 2048, 4
LR, JITTER = 1e-2, 1e-4
METRIC = 'deep_gp_dsvi_iters_per_sec'


def data():
    g = torch.Generator().manual_seed(0)
    X = torch.rand((N, D), generator=g) * 6 - 3
    f = torch.sin(X).sum(1, keepdim=True) / math.sqrt(D)
    Y = torch.sign(f) * f.abs().sqrt() + 0.05 * torch.randn((N, 1), generator=g)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:M]
    return X, Y, X[perm].clone()


def build(Xs, Ys, Z0, device, world, data_resident=True):
    import mxfusion_b200 as mf
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import DeepGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    mf.config.DEFAULT_DTYPE = 'float32'
    np.random.seed(7)
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, D))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.Z0 = mf.Variable(shape=(M, D), initial_value=Z0)
    m.Z1 = mf.Variable(shape=(M, H), initial_value=Z0[:, :H].clone())
    m.Y = DeepGPRegression.define_variable(X=m.X, kernels=[RBF(D, name='rbf_l0', lengthscale=2.0), RBF(H, name='rbf_l1', lengthscale=2.0)],
                                           noise_var=m.noise_var, inducing_inputs=[m.Z0, m.Z1], shape=(m.N, 1))
    alg = m.Y.factor.dgp_log_pdf
    alg.jitter, alg.num_samples = JITTER, S
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / float(B)}, data_resident=data_resident,
                                  rng=np.random.RandomState(1234))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, dtype='float32',
                              context=device)
    infr.initialize(X=tuple(Xs.shape), Y=tuple(Ys.shape))
    post = m.Y.factor._extra_graphs[0]
    for l, out in ((0, H), (1, 1)):
        infr.params[post.qU_mean[l]] = torch.zeros((M, out))
        infr.params[post.qU_cov_W[l]] = torch.zeros((M, M))
        infr.params[post.qU_cov_diag[l]] = torch.full((M,), 1e-2 if l == 0 else 1.0)
    return infr, loop


def timed(infr, loop, Xs, Ys, steps, warmup, use_events, barrier):
    st = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def on_step(k, loss):
        if k == 1:
            st['first'] = loss.clone()
        if k == warmup:
            barrier()
            torch.cuda.synchronize()
            st['t0'] = time.perf_counter()
            e0.record()
        elif k == warmup + steps:
            e1.record()
            barrier()
            torch.cuda.synchronize()
            st['t1'] = time.perf_counter()
            st['last'] = float(loss)
    infr.run(X=Xs, Y=Ys, max_iter=2 + (warmup + steps) * B // Xs.shape[0], learning_rate=LR, max_steps=warmup + steps,
             on_step=on_step)
    secs = e0.elapsed_time(e1) / 1e3 if use_events else st['t1'] - st['t0']
    return secs, float(st['first']), st['last']


def cpu_leg(X, Y, Z0, budget_s=10.0):
    """The same bound (Cholesky formulation, svgp_regression.py:83-96 per layer) with torch CPU ops + autograd + Adam."""
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    inv = lambda v: torch.log(torch.expm1(torch.as_tensor(v, dtype=torch.float32)))
    leaves = dict(Z0=Z0.clone(), Z1=Z0[:, :H].clone(), ls0=inv([2.0]), v0=inv([1.0]), ls1=inv([2.0]), v1=inv([1.0]), nv=inv([0.01]),
                  m0=torch.zeros((M, H)), W0=torch.zeros((M, M)), d0=inv(torch.full((M,), 1e-2)),
                  m1=torch.zeros((M, 1)), W1=torch.zeros((M, M)), d1=inv(torch.ones((M,))))
    for t in leaves.values():
        t.requires_grad_()
    opt = torch_ref.AdamMX(list(leaves.values()), LR)
    sp = torch.nn.functional.softplus
    eye = torch.eye(M)

    def layer(h, Z, ls, var, m, W, d):
        Kuu = torch_ref.K(torch_ref.RBF, Z[None], ls[None], var[None])[0] + JITTER * eye
        L = torch.linalg.cholesky(Kuu)
        Ls = torch.linalg.cholesky(W @ W.T + torch.diag(d))
        Kuf = torch_ref.K(torch_ref.RBF, Z[None].expand(h.shape[0], -1, -1), ls[None], var[None], h)
        A = torch.linalg.solve_triangular(L, Kuf, upper=False)
        mt = torch.linalg.solve_triangular(L, m, upper=False)
        C = torch.linalg.solve_triangular(L, Ls, upper=False)
        mean = A.transpose(-1, -2) @ mt
        v = var[0] - (A * A).sum(-2) + ((C.T @ A) ** 2).sum(-2)
        P = m.shape[1]
        nkl = P * (0.5 * M + torch.log(torch.diagonal(Ls)).sum() - torch.log(torch.diagonal(L)).sum()) - 0.5 * P * (C * C).sum() \
            - 0.5 * (mt * mt).sum()
        return mean, v, nkl

    def step(i):
        sel = torch.randint(0, N, (B,))
        xb, yb = X[sel], Y[sel]
        mean, v, k0 = layer(xb[None], leaves['Z0'], sp(leaves['ls0']), sp(leaves['v0']), leaves['m0'], leaves['W0'], sp(leaves['d0']))
        h = mean + xb[None] + torch.sqrt(torch.clamp(v, min=0.))[..., None] * torch.randn((S, B, H))
        mean, v, k1 = layer(h, leaves['Z1'], sp(leaves['ls1']), sp(leaves['v1']), leaves['m1'], leaves['W1'], sp(leaves['d1']))
        nv = sp(leaves['nv'])[0]
        data = -0.5 * B * (math.log(2 * math.pi) + torch.log(nv)) - ((yb[None] - mean) ** 2).sum((-1, -2)) / (2 * nv) - v.sum(-1) / (2 * nv)
        loss = -((N / float(B)) * data + k0 + k1).mean()
        loss.backward()
        opt.step(B)
        return float(loss)
    for i in range(2):
        step(i)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s:
        step(n)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt, os.cpu_count()


def _line(args, world, secs, wall2, first, last, loop, loop2, clocks, pk_kind, steps, warmup):
    unit = 'minibatch iterations (B=%d rows per GPU, S=%d samples: 2-layer ELBO fwd + grad + Adam) per second, summed over GPUs' % (B, S)
    st = loop.last_stepper
    return {'metric': METRIC, 'value': world * steps / secs, 'unit': unit, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': 1e3 * secs / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'DeepGPRegression 2 layers, N=%d D=%d hidden=%d M=%d/layer RBF, doubly-stochastic VI S=%d, '
                                   'minibatch=%d per GPU, f32 (BASELINE c5; no reference implementation: parity unpinned by the '
                                   'reference, pinned by the one-layer SVGP reduction + an independent dense oracle)' % (N, D, H, M, S, B),
                       'parallelism': 'dp%d' % world, 'l2_flush': 'none'},
            'clocks': clocks,
            'e2e': {'value': world * steps / wall2, 'unit': unit, 'h2d_bytes_per_step': loop2.h2d_bytes_per_step,
                    'd2h_bytes_per_step': loop2.d2h_bytes_per_step, 'ms_per_step': 1e3 * wall2 / steps},
            'gpu_launches': int((st.launches_per_step or 0) * steps), 'launches_per_step': st.launches_per_step,
            'first_loss': first, 'final_loss': last, 'peaks': pk_kind, 'roofline': None}


def bench_line(args, device, pk_kind, ClockSampler):
    steps, warmup = args.steps, max(args.warmup, 3)
    X, Y, Z0 = data()
    infr, loop = build(X, Y, Z0, device, 1)
    sampler = ClockSampler(0)
    sampler.start()
    secs, first, last = timed(infr, loop, X, Y, steps, warmup, True, lambda: None)
    clocks = sampler.finish()
    infr2, loop2 = build(X, Y, Z0, device, 1, data_resident=False)
    wall2, _, _ = timed(infr2, loop2, X, Y, steps, warmup, False, lambda: None)
    line = _line(args, 1, secs, wall2, first, last, loop, loop2, clocks, pk_kind, steps, warmup)
    if not args.no_cpu_baseline:
        ips, n, dt, cores = cpu_leg(X, Y, Z0)
        line['cpu_baseline'] = {'value': ips, 'unit': line['unit'], 'cores': cores, 'kind': 'port',
                                'sample': '%d iterations in %.1f s (torch CPU f32 restatement of the same composition; the reference '
                                          'has no deep GP)' % (n, dt)}
    return line


def main_distributed(args, emit):
    import torch.distributed as dist
    import bench
    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=device)
    steps, warmup = args.steps, max(args.warmup, 3)
    X, Y, Z0 = data()
    shard = N // world
    Xs, Ys = X[rank * shard:(rank + 1) * shard], Y[rank * shard:(rank + 1) * shard]

    def mx(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    infr, loop = build(Xs, Ys, Z0, device, world)
    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    secs, first, last = timed(infr, loop, Xs, Ys, steps, warmup, True, dist.barrier)
    clocks = sampler.finish() if sampler else None
    secs = mx(secs)
    infr2, loop2 = build(Xs, Ys, Z0, device, world, data_resident=False)
    wall2, _, _ = timed(infr2, loop2, Xs, Ys, steps, warmup, False, dist.barrier)
    wall2 = mx(wall2)
    if rank == 0:
        emit(_line(args, world, secs, wall2, first, last, loop, loop2, clocks, bench.peaks()[1], steps, warmup))
    dist.destroy_process_group()
