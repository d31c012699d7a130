"""potrf (+ pack: explicit inverse, L^T) timings on one B200, CUDA events around back-to-back launches after warm-up,
next to torch.linalg.cholesky (cuSOLVER/MAGMA) on the same matrices.  Writes gpurun_out/potrf_bench.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import _raw  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device('cuda:0')
    out = []
    g = torch.Generator(device='cpu').manual_seed(0)
    for n in [int(a) for a in sys.argv[1:]] or [256, 512, 1024, 2048, 4096, 8192]:
        Z = torch.rand((n, 8), generator=g) * 6 - 3
        r2 = torch.cdist(Z.double(), Z.double()) ** 2
        A = (torch.exp(-0.5 * r2) + 1e-3 * torch.eye(n, dtype=torch.float64)).float().to(dev)[None]
        work = A.clone()
        pack = _raw.new_pack(work)
        info = torch.zeros((1,), dtype=torch.int32, device=dev)

        def ours():
            work.copy_(A)
            _raw.potrf_packed_(work, info, pack)

        def copy_only():
            work.copy_(A)

        def lib():
            torch.linalg.cholesky(A)
        t_copy = timeit(copy_only)
        t_ours = timeit(ours) - t_copy
        t_lib = timeit(lib)
        # under a CUDA graph (what the training step sees)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ours()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            ours()
        t_graph = timeit(gr.replay) - t_copy
        Lref = torch.linalg.cholesky(A.double())
        err = float((work.double() - Lref).abs().max() / Lref.abs().max())
        row = {'n': n, 'ms_potrf_pack': t_ours, 'ms_potrf_pack_graph': t_graph, 'ms_torch_cholesky': t_lib,
               'tflops_fp32': n ** 3 / 3 / t_ours / 1e9, 'info': int(info.item()), 'max_rel_err_vs_f64': err}
        print(json.dumps(row), flush=True)
        out.append(row)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/potrf_bench.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
