"""Times every raw operator of the fused SVGP step at the headline shapes (M=1024, B=4096, D=8, f32) with CUDA
events, eager launches, warm L2.  Writes gpurun_out/step_breakdown.json."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxfusion_b200 import _raw as R

dev = torch.device('cuda:0')
M, B, D, P = 1024, 4096, 8, 1
if len(sys.argv) > 1:
    M, B = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device='cpu').manual_seed(0)
X = (torch.rand((1, B, D), generator=g) * 6 - 3).to(dev)
Z = (torch.rand((1, M, D), generator=g) * 6 - 3).to(dev)
Y = torch.randn((1, B, P), generator=g).to(dev)
ls = torch.ones((1, 1), device=dev); var = torch.ones((1, 1), device=dev)
W = (torch.randn((1, M, M), generator=g) * 0.01).to(dev)
res = []


def t(name, fn, setup=None, iters=20):
    for _ in range(3):
        if setup: setup()
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if setup: setup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    res.append((name, ts[len(ts) // 2]))
    print('%-44s %9.1f us' % (name, ts[len(ts) // 2]), flush=True)


Kuu0 = R.kbuild_fwd(0, Z, None, ls, var, diag_const=1e-3)
Kuu = Kuu0.clone()
t('kbuild Kuu (M x M)', lambda: R.kbuild_fwd(0, Z, None, ls, var, diag_const=1e-3))
Kuf = R.kbuild_fwd(0, Z, X, ls, var)
t('kbuild Kuf (M x B)', lambda: R.kbuild_fwd(0, Z, X, ls, var, out=Kuf))
t('gemm tri W W^T (M,M,M)', lambda: R.gemm(W, W, transB=True, tri=True))
t('potrf_packed (M)', lambda: R.potrf_packed_(Kuu), setup=lambda: Kuu.copy_(Kuu0))
t('  (copy M x M)', lambda: Kuu.copy_(Kuu0))
L, info, pk = R.potrf_packed_(Kuu0.clone())
RHS_M = torch.randn((1, M, M), device=dev); RHS_B = torch.randn((1, M, B), device=dev)
RHS_3M = torch.randn((1, M, 3 * M), device=dev); RHS_2M = torch.randn((1, M, 2 * M), device=dev)
RHS_P = torch.randn((1, M, P), device=dev)
t('trsm_packed N (M x M rhs)', lambda: R.trsm_packed_(L, pk, RHS_M))
t('trsm_packed N (M x B rhs)', lambda: R.trsm_packed_(L, pk, RHS_B))
t('trsm_packed N (M x P rhs)', lambda: R.trsm_packed_(L, pk, RHS_P))
t('trsm_packed T (M x 3M rhs)', lambda: R.trsm_packed_(L, pk, RHS_3M, transpose=True))
t('trsm_packed T (M x 2M rhs)', lambda: R.trsm_packed_(L, pk, RHS_2M, transpose=True))
t('trsm_packed T (M x P rhs)', lambda: R.trsm_packed_(L, pk, RHS_P, transpose=True))
A = torch.randn((1, M, B), device=dev) * 0.1
Cm = torch.randn((1, M, M), device=dev) * 0.1
t('gemm tri A A^T (M,M,B)', lambda: R.gemm(A, A, transB=True, tri=True))
t('gemm tri C C^T (M,M,M)', lambda: R.gemm(Cm, Cm, transB=True, tri=True))
t('copy_ltu (M)', lambda: R.copy_ltu(Cm))
mt = torch.randn((1, M, P), device=dev)
t('gemm A^T mt (B,P,M)', lambda: R.gemm(A, mt, transA=True))
G1 = R.gemm(A, mt, transA=True)
t('reduce sumsqdiff (B x P)', lambda: R.reduce(R.RED_SUMSQDIFF, Y, G1))
t('reduce sumsq A (M x B)', lambda: R.reduce(R.RED_SUMSQ, A))
t('reduce dot (M x M)', lambda: R.reduce(R.RED_DOT, Cm, Cm))
t('sumlogdiag', lambda: R.sumlogdiag(L))
t('gemm Phi T (M,M,M) NN', lambda: R.gemm(Cm, Cm))
t('gemm A Y (M,P,B)', lambda: R.gemm(A, Y))
t('gemm Phi mt (M,P,M)', lambda: R.gemm(Cm, mt))
coef = torch.ones((1, 6), device=dev)
t('svgp_bwd_assemble', lambda: R.svgp_bwd_assemble(Cm, Cm, Cm, mt, mt, coef))
E3 = R.svgp_bwd_assemble(Cm, Cm, Cm, mt, mt, coef)
t('gemm E_R A (M,B,M) NN', lambda: R.gemm(E3[:, :, 2 * M:], A))
dK = torch.randn((1, M, B), device=dev)
t('gemm w Y^T acc (M,B,P)', lambda: R.gemm(mt, Y, transB=True, beta=1.0, C=dK))
F2 = torch.empty((1, M, 2 * M), device=dev)
t('transpose x2 (M)', lambda: (R.transpose(E3[:, :, :M], out=F2[:, :, :M]), R.transpose(E3[:, :, M:2 * M], out=F2[:, :, M:])))
t('kbuild_bwd Kuf', lambda: R.kbuild_bwd(0, Z, X, ls, var, dK, need_dX2=False))
t('kbuild_bwd Kuu', lambda: R.kbuild_bwd(0, Z, None, ls, var, Cm))
eye = torch.eye(M, device=dev).unsqueeze(0).contiguous()
t('gemm Sbar W (M,M,M) NN', lambda: R.gemm(Cm, W, alpha=2.0))
t('axpby_dev (M x M)', lambda: R.axpby_dev(coef[:, 0], Cm, coef[:, 1], Cm))
t('get_diag', lambda: R.get_diag(Cm))
t('torch eye+expand (M)', lambda: torch.eye(M, device=dev).unsqueeze(0).expand(1, M, M).contiguous())
tot = sum(v for _, v in res)
print('sum of listed ops: %.1f us' % tot)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'step_breakdown.json'), 'w'), indent=1)
