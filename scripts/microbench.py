"""Per-kernel micro-benchmarks on one B200 (CUDA events on the launching stream, warm-up, inputs
larger than L2 where the kernel is HBM-bound).  Writes gpurun_out/microbench.json.

    python scripts/microbench.py [kbuild] [potrf] [gemm] [trsm] [all]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxfusion_b200 import _raw  # noqa: E402

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
except Exception:
    pass
HBM = PEAKS.get('hbm_gbs', 6650.0)


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_kbuild(res):
    dev = torch.device('cuda:0')
    for kind, kname in [(_raw.RBF, 'rbf'), (_raw.MATERN52, 'matern52')]:
        for (N, M, D) in [(1000000, 1024, 8), (1000000, 1024, 16), (1 << 18, 1024, 8)]:
            g = torch.Generator(device='cpu').manual_seed(0)
            X = (torch.rand((1, N, D), generator=g) * 6 - 3).to(dev)
            Z = X[:, :M].clone()
            ls = torch.ones((1, 1), device=dev)
            var = torch.ones((1, 1), device=dev)
            for name, a, b in [('K(Z,X)', Z, X), ('K(X,Z)', X, Z)]:
                out = torch.empty((1, a.shape[1], b.shape[1]), device=dev)
                med, best = timeit(lambda: _raw.kbuild_fwd(kind, a, b, ls, var, out=out))
                nbytes = 4 * (N * M + N * D + M * D + D + 1)
                r = dict(kernel='kbuild_fwd', kind=kname, which=name, N=N, M=M, D=D, ms_median=med, ms_best=best,
                         alg_bytes=nbytes, gbs=nbytes / med / 1e6, frac_of_measured_hbm=nbytes / med / 1e6 / HBM)
                print(r, flush=True)
                res.append(r)
            del X, out


def bench_gemm(res):
    dev = torch.device('cuda:0')
    for (m, n, k) in [(1024, 4096, 1024), (4096, 4096, 4096), (1024, 1024, 4096), (8192, 8192, 128)]:
        A = torch.randn((1, m, k), device=dev)
        B = torch.randn((1, k, n), device=dev)
        C = torch.empty((1, m, n), device=dev)
        med, best = timeit(lambda: _raw.gemm(A, B, C=C))
        r = dict(kernel='gemm_f32', m=m, n=n, k=k, ms_median=med, tflops=2.0 * m * n * k / med / 1e9)
        print(r, flush=True)
        res.append(r)
        torch.backends.cuda.matmul.allow_tf32 = False
        med, best = timeit(lambda: torch.matmul(A, B, out=C))
        r = dict(kernel='torch_matmul_f32', m=m, n=n, k=k, ms_median=med, tflops=2.0 * m * n * k / med / 1e9)
        print(r, flush=True)
        res.append(r)


def bench_potrf(res):
    dev = torch.device('cuda:0')
    for n in [512, 1024, 2048, 4096, 8192]:
        W = torch.randn((n, n), device=dev)
        A0 = (W @ W.t() / n + torch.eye(n, device=dev)).unsqueeze(0)
        A = A0.clone(); pk = _raw.new_pack(A)

        def run():
            A.copy_(A0)
            _raw.potrf_packed_(A, pack=pk)
        med, best = timeit(run, iters=5)
        med_copy, _ = timeit(lambda: A.copy_(A0), iters=5)
        t = med - med_copy
        r = dict(kernel='potrf_f32', n=n, ms_median=t, tflops=n ** 3 / 3.0 / t / 1e9)
        print(r, flush=True)
        res.append(r)
        med, _ = timeit(lambda: torch.linalg.cholesky(A0), iters=5)
        r = dict(kernel='torch_cholesky_f32', n=n, ms_median=med, tflops=n ** 3 / 3.0 / med / 1e9)
        print(r, flush=True)
        res.append(r)


def bench_trsm(res):
    dev = torch.device('cuda:0')
    for (n, nrhs) in [(1024, 4096), (1024, 1024), (512, 2048)]:
        L = torch.tril(torch.randn((1, n, n), device=dev)) * 0.01 + torch.eye(n, device=dev)
        B0 = torch.randn((1, n, nrhs), device=dev)
        B = B0.clone()

        def run():
            B.copy_(B0)
            _raw.trsm_(L, B)
        med, _ = timeit(run)
        r = dict(kernel='trsm_f32', n=n, nrhs=nrhs, ms_median=med, tflops=n * n * nrhs / med / 1e9)
        print(r, flush=True)
        res.append(r)
        med, _ = timeit(lambda: torch.linalg.solve_triangular(L, B0, upper=False))
        r = dict(kernel='torch_trsm_f32', n=n, nrhs=nrhs, ms_median=med, tflops=n * n * nrhs / med / 1e9)
        print(r, flush=True)
        res.append(r)


def bench_normal(res):
    """MC-ELBO pieces (normal.py:52-92 + factor_graph.py:223) at a stress size (n >= 2^26)."""
    dev = torch.device('cuda:0')
    n, S = 1 << 26, 3
    m = torch.randn((1, n), device=dev)
    v = torch.rand((1, n), device=dev) + 0.5
    w = torch.empty((S, n), device=dev)
    # reparameterised draw with the Philox stream generated in-kernel: read 2n (mean, variance), write S*n
    med, best = timeit(lambda: _raw.normal_reparam(m, v, S, seed=1, offset=0), iters=8)
    nbytes = 4 * n * (2 + S)
    r = dict(kernel='normal_reparam(philox)', n=n, S=S, ms_median=med, alg_bytes=nbytes, gbs=nbytes / med / 1e6,
             frac_of_measured_hbm=nbytes / med / 1e6 / HBM)
    print(r, flush=True)
    res.append(r)
    x = _raw.normal_reparam(m, v, S, seed=1, offset=0)
    # fused log-density + sample mean + sum: read S*n (samples) + 2n (mean, variance), write 4 bytes
    med, best = timeit(lambda: _raw.normal_logpdf_sum(x, m, v), iters=8)
    nbytes = 4 * n * (2 + S)
    r = dict(kernel='normal_logpdf_sum', n=n, S=S, ms_median=med, alg_bytes=nbytes, gbs=nbytes / med / 1e6,
             frac_of_measured_hbm=nbytes / med / 1e6 / HBM)
    print(r, flush=True)
    res.append(r)
    g = torch.ones((1,), device=dev)
    med, best = timeit(lambda: _raw.normal_logpdf_sum_bwd(x, m, v, g, need=(False, True, True)), iters=8)
    nbytes = 4 * n * (2 + S + 2)
    r = dict(kernel='normal_logpdf_sum_bwd', n=n, S=S, ms_median=med, alg_bytes=nbytes, gbs=nbytes / med / 1e6,
             frac_of_measured_hbm=nbytes / med / 1e6 / HBM)
    print(r, flush=True)
    res.append(r)
    a = torch.randn((1, 8192, 8192), device=dev)
    med, best = timeit(lambda: _raw.reduce(_raw.RED_SUMSQ, a), iters=8)
    nbytes = 4 * a.numel()
    r = dict(kernel='reduce_sumsq', n=a.numel(), ms_median=med, alg_bytes=nbytes, gbs=nbytes / med / 1e6,
             frac_of_measured_hbm=nbytes / med / 1e6 / HBM)
    print(r, flush=True)
    res.append(r)


def bench_gemmshapes(res):
    """The GEMM shapes of the blocked potrf / trsm / streamed statistics (set MXF_GEMM_PE=0/1 to compare kernels)."""
    dev = torch.device('cuda:0')
    shapes = [(512, 28392, 512, False, False), (1024, 28392, 1024, False, False), (7680, 7680, 512, True, True),
              (4096, 4096, 512, True, True), (2048, 2048, 512, True, True), (7680, 512, 512, True, False),
              (1024, 4096, 1024, False, False), (4096, 4096, 1024, True, False), (4096, 4096, 2048, True, False)]
    for (m, n, k, tb, tri) in shapes:
        A = torch.randn((1, m, k), device=dev)
        B = torch.randn((1, n, k) if tb else (1, k, n), device=dev)
        C = torch.zeros((1, m, n), device=dev)
        med, best = timeit(lambda: _raw.gemm(A, B, False, tb, alpha=-1.0, beta=1.0, C=C, tri=tri))
        fl = 2.0 * m * n * k * (0.5 if tri else 1.0)
        r = dict(kernel='gemm_f32', pe=os.environ.get('MXF_GEMM_PE', '1'), m=m, n=n, k=k, transB=tb, tri=tri, ms_median=med,
                 tflops=fl / med / 1e9)
        print(r, flush=True)
        res.append(r)


def main():
    which = sys.argv[1:] or ['all']
    res = []
    print('device', torch.cuda.get_device_name(0), 'host cores', os.cpu_count(), flush=True)
    if 'kbuild' in which or 'all' in which:
        bench_kbuild(res)
    if 'gemm' in which or 'all' in which:
        bench_gemm(res)
    if 'gemmshapes' in which:
        bench_gemmshapes(res)
    if 'potrf' in which or 'all' in which:
        bench_potrf(res)
    if 'trsm' in which or 'all' in which:
        bench_trsm(res)
    if 'normal' in which or 'all' in which:
        bench_normal(res)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'microbench_%d.json' % int(time.time())), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()

