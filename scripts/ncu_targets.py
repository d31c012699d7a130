"""Small single-purpose launches for ncu captures (one GPU, few kernels):
    python scripts/ncu_targets.py kbuild     # K(X,Z) N=1e6 M=1024 D=8 RBF, 3 launches
    python scripts/ncu_targets.py kbuild_tc [D] [kind]   # the same on the tcgen05 + TMA-store kernel (default D=16 Matern-5/2)
    python scripts/ncu_targets.py potrf N    # one GEMM-based potrf of an N x N f32 matrix (default 8192)
    python scripts/ncu_targets.py gemm       # 4096^3 NT / NN tcgen05 GEMMs
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import _raw as R
dev = torch.device('cuda:0')
what = sys.argv[1] if len(sys.argv) > 1 else 'kbuild'
if what == 'kbuild':
    N, M, D = 1000000, 1024, 8
    g = torch.Generator(device='cpu').manual_seed(0)
    X = (torch.rand((1, N, D), generator=g) * 6 - 3).to(dev)
    Z = X[:, :M].clone()
    ls = torch.ones((1, 1), device=dev); var = torch.ones((1, 1), device=dev)
    out = torch.empty((1, N, M), device=dev)
    for _ in range(3):
        R.kbuild_fwd(R.RBF, X, Z, ls, var, out=out)
elif what == 'kbuild_tc':
    # K(X,Z) N=1e6 M=1024 on the tcgen05 + TMA-store kernel: python scripts/ncu_targets.py kbuild_tc [D] [kind]
    N, M = 1000000, 1024
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    kind = int(sys.argv[3]) if len(sys.argv) > 3 else R.MATERN52
    g = torch.Generator(device='cpu').manual_seed(0)
    X = (torch.rand((1, N, D), generator=g) * 6 - 3).to(dev)
    Z = X[:, :M].clone()
    ls = torch.ones((1, 1), device=dev); var = torch.ones((1, 1), device=dev)
    out = torch.empty((1, N, M), device=dev)
    R.kbuild_tc_threshold(0)
    for _ in range(3):
        R.kbuild_fwd(kind, X, Z, ls, var, out=out)
elif what == 'potrf':
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    W = torch.randn((n, n), device=dev)
    A = (W @ W.t() / n + torch.eye(n, device=dev)).unsqueeze(0)
    del W
    R.potrf_packed_(A)
elif what == 'gemm':
    n = 4096
    A = torch.randn((1, n, n), device=dev); B = torch.randn((1, n, n), device=dev)
    for tb in (True, False):
        for _ in range(2):
            R.gemm(A, B, False, tb)
elif what == 'kbwd':
    M, B, D = 1024, 4096, 8
    g = torch.Generator(device='cpu').manual_seed(0)
    Z = (torch.rand((1, M, D), generator=g) * 6 - 3).to(dev)
    X = (torch.rand((1, B, D), generator=g) * 6 - 3).to(dev)
    G = torch.randn((1, M, B), generator=g).to(dev)
    ls = torch.ones((1, 1), device=dev); var = torch.ones((1, 1), device=dev)
    for _ in range(3):
        R.kbuild_bwd(R.RBF, Z, X, ls, var, G, need_dX=True, need_dX2=False)
torch.cuda.synchronize()
print('done', what)
