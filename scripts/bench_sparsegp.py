"""Full-batch variational sparse GP (Titsias bound) at the headline size N=1e6, M=1024, D=8 ("the full-batch
variant (B=N) is reported separately because that is where K(X,Z) is HBM-sized"): one evaluation of the bound and its
gradient through ops.sparsegp_log_pdf (streamed whitened statistics), timed with CUDA events; the CPU leg runs the
op-for-op restatement (oracle/torch_ref.sparsegp_log_pdf) on a bounded row sample and scales linearly in N.

    python scripts/bench_sparsegp.py [N] [M] [chunk]      -> one JSON line (also gpurun_out/sparsegp_bench.json)
"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxfusion_b200 import ops, _lib  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith('--')]
N = int(_pos[0]) if len(_pos) > 0 else 1000000
M = int(_pos[1]) if len(_pos) > 1 else 1024
CHUNK = int(_pos[2]) if len(_pos) > 2 else None
D = 8
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
X = torch.rand((1, N, D), generator=g) * 6 - 3
Y = torch.sin(X).sum(-1, keepdim=True) / math.sqrt(D) + 0.05 * torch.randn((1, N, 1), generator=g)
perm = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:M]
Z = X[:, perm].clone()
par = dict(noise_var=torch.full((1, 1), 0.01), lengthscale=torch.ones((1, 1)), variance=torch.ones((1, 1)))


def run_gpu(steps=3, warm=1):
    t = dict(X=X.to(dev), Y=Y.to(dev), Z=Z.to(dev).requires_grad_(),
             **{k: v.to(dev).requires_grad_() for k, v in par.items()})
    res = {}
    for mode in ('fwd', 'fwd_bwd'):
        ts = []
        for i in range(warm + steps):
            for v in t.values():
                v.grad = None
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = _lib.launch_count()
            a.record()
            with torch.set_grad_enabled(mode == 'fwd_bwd'):
                logL = ops.sparsegp_log_pdf(ops.RBF, t['X'], t['Y'], t['Z'], t['noise_var'], t['lengthscale'],
                                            t['variance'], jitter=1e-4, chunk=CHUNK)[0]
                if mode == 'fwd_bwd':
                    (-logL.sum()).backward()
            b.record()
            torch.cuda.synchronize()
            if i >= warm:
                ts.append(a.elapsed_time(b))
            res[mode + '_launches'] = _lib.launch_count() - l0
        res[mode + '_ms'] = sorted(ts)[len(ts) // 2]
    res['logL'] = float(logL)
    res['peak_mem_gb'] = torch.cuda.max_memory_allocated() / 1e9
    return res


def run_cpu(n_sample=32768):
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    t = dict(X=X[:, :n_sample].clone(), Y=Y[:, :n_sample].clone(), Z=Z.clone().requires_grad_(),
             **{k: v.clone().requires_grad_() for k, v in par.items()})

    def one():
        logL = torch_ref.sparsegp_log_pdf(torch_ref.RBF, t['X'], t['Y'], t['Z'], t['noise_var'], t['lengthscale'],
                                          t['variance'], jitter=1e-4)
        (-logL.sum()).backward()
    one()
    t0 = time.perf_counter()
    one()
    dt = time.perf_counter() - t0
    return dict(cpu_sample_rows=n_sample, cpu_sample_s=dt, cpu_full_batch_s_extrapolated=dt * N / n_sample,
                cpu_cores=os.cpu_count())


out = dict(workload='SparseGPRegression full batch N=%d M=%d D=%d RBF f32, bound + gradient' % (N, M, D),
           chunk_rows=CHUNK or ops._stats_chunk(M, 1))
out.update(run_gpu())
# algorithmic work of the streamed statistics: K-build N*M, trsm M^2 N, syrk M^2 N (lower half), fwd; bwd ~ 3x
out['fwd_gflop'] = (2.0 * M * M * N) / 1e9
out['fwd_tflops'] = out['fwd_gflop'] / out['fwd_ms']
if '--no-cpu' not in sys.argv:
    out.update(run_cpu())
    out['speedup_vs_cpu_fwd_bwd'] = out['cpu_full_batch_s_extrapolated'] * 1e3 / out['fwd_bwd_ms']
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'sparsegp_bench.json'), 'w'))
