"""K(X, Z) at the BASELINE shapes (N = 1e6, M = 1024, D = 8 / 16): the streaming FMA kernel against the tcgen05 + TMA-store
kernel (csrc/kbuild_tc.cuh), CUDA events on the launching stream, output (4.1 GB) larger than L2.

    python scripts/bench_kbuild.py [--out gpurun_out/kbuild_bench.json] [--quick]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxfusion_b200 import _raw  # noqa: E402

try:
    HBM = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6536.0)
except Exception:
    HBM = 6536.0


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


def main():
    out_path = 'gpurun_out/kbuild_bench.json'
    if '--out' in sys.argv:
        out_path = sys.argv[sys.argv.index('--out') + 1]
    quick = '--quick' in sys.argv
    dev = torch.device('cuda:0')
    res = []
    kinds = [(_raw.RBF, 'rbf'), (_raw.MATERN52, 'matern52')]
    if not quick:
        kinds += [(_raw.MATERN32, 'matern32'), (_raw.MATERN12, 'matern12')]
    shapes = [(1000000, 1024, 16), (1000000, 1024, 8)]
    if not quick:
        shapes += [(1 << 18, 1024, 16), (4096, 1024, 16), (1000000, 512, 16)]
    for (N, M, D) in shapes:
        g = torch.Generator(device='cpu').manual_seed(0)
        X = (torch.rand((1, N, D), generator=g) * 6 - 3).to(dev)
        Z = X[:, :M].clone()
        ls = torch.ones((1, 1), device=dev)
        var = torch.ones((1, 1), device=dev)
        out = torch.empty((1, N, M), device=dev)
        for kind, kname in kinds:
            row = dict(kernel='kbuild_fwd K(X,Z)', kind=kname, N=N, M=M, D=D)
            nbytes = 4 * (N * M + N * D + M * D + D + 1)
            ref = None
            for path, thr in (('fma', 1 << 62), ('tc', 0)):
                old = _raw.kbuild_tc_threshold(thr)
                try:
                    med, best = timeit(lambda: _raw.kbuild_fwd(kind, X, Z, ls, var, out=out))
                finally:
                    _raw.kbuild_tc_threshold(old)
                row['ms_' + path] = med
                row['ms_best_' + path] = best
                row['gbs_' + path] = nbytes / med / 1e6
                row['frac_hbm_' + path] = nbytes / med / 1e6 / HBM
                if ref is None:
                    ref = out[0, :: max(1, N // 4096)].clone()
                else:
                    row['max_abs_diff_tc_vs_fma'] = float((out[0, :: max(1, N // 4096)] - ref).abs().max())
            row['alg_bytes'] = nbytes
            print(json.dumps(row), flush=True)
            res.append(row)
        del X, out
    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    json.dump(dict(hbm_peak_gbs=HBM, rows=res), open(out_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
