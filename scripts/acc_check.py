import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import ops, _raw as R
from oracle import torch_ref
dev = torch.device('cuda:0')
for N, Din in [(512, 2), (1024, 2), (512, 8)]:
    rng = np.random.RandomState(2)
    d = dict(X=rng.uniform(-3, 3, (1, N, Din)), Y=rng.randn(1, N, 1), noise=rng.rand(1, 1) * 0.1 + 0.05,
             ls=rng.rand(1, Din) * 0.5 + 0.8, var=rng.rand(1, 1) + 0.5)
    r = {k: torch.tensor(v) for k, v in d.items()}
    want = float(torch_ref.gp_log_pdf(0, r['X'], r['Y'], r['noise'], r['ls'], r['var'], jitter=1e-6))
    r32 = {k: torch.tensor(v, dtype=torch.float32) for k, v in d.items()}
    cpu32 = float(torch_ref.gp_log_pdf(0, r32['X'], r32['Y'], r32['noise'], r32['ls'], r32['var'], jitter=1e-6))
    t = {k: torch.tensor(v, dtype=torch.float32, device=dev) for k, v in d.items()}
    got = float(ops.gp_log_pdf(0, t['X'], t['Y'], t['noise'], t['ls'], t['var'], jitter=1e-6)[0])
    # old substitution path for comparison
    K = R.kbuild_fwd(0, t['X'], None, t['ls'], t['var'], diag_add=t['noise'], diag_const=1e-6)
    K64 = K.double().cpu().numpy()[0]
    print('N', N, 'D', Din, 'cond(K) %.2e' % np.linalg.cond(K64))
    L, info = R.potrf_(K.clone())
    LY = R.trsm_(L, t['Y'].clone())
    old = float(-R.sumlogdiag(L) - 0.5 * R.reduce(R.RED_SUMSQ, LY) - 0.5 * N * np.log(2 * np.pi))
    print('  f64 ref %.4f | torch cpu f32 (LAPACK) err %.3e | packed+TC=%s err %.3e | old substitution f32 err %.3e' % (
        want, abs(cpu32 - want) / abs(want), os.environ.get('MXF_GEMM_TC', '1'), abs(got - want) / abs(want), abs(old - want) / abs(want)))
