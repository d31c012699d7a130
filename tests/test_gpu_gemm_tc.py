"""tcgen05 (3xTF32) GEMM against float64 matmul.  The kernel must be FP32-accurate: the bound used is
|err| <= 4e-6 * (|A| |B|)_ij (3 * 2^-22 per product + fp32 accumulation), i.e. ~50x tighter than a single
TF32 pass would achieve."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # m, n, k, transB, S, tri, alpha, beta
    (128, 128, 32, True, 1, False, 1.0, 0.0),
    (128, 128, 64, True, 1, False, 1.0, 0.0),
    (256, 384, 128, True, 1, False, 0.5, 0.0),
    (1024, 1024, 1024, True, 1, False, 1.0, 0.0),
    (100, 200, 64, True, 2, False, 1.0, 0.3),
    (896, 896, 128, True, 1, True, -1.0, 1.0),
    (1024, 4096, 1024, True, 1, False, 1.0, 0.0),
    (128, 128, 32, False, 1, False, 1.0, 0.0),
    (256, 384, 128, False, 1, False, 0.5, 0.0),
    (1024, 4096, 1024, False, 1, False, 1.0, 0.0),
    (100, 200, 72, False, 2, False, -1.0, 1.0),
    (1000, 64, 128, False, 1, False, 1.0, 1.0),
    (64, 64, 4096, True, 1, False, 1.0, 0.0),
    # persistent kernel (>= 2 waves of 128 x 128 tiles, 256 <= k <= 2048): NT / NN, beta, ragged edges with an odd number
    # of K steps (the two converter groups swap roles from tile to tile), batch, lower tiles only
    (2048, 2560, 512, True, 1, False, 1.0, 0.0),
    (2048, 2560, 512, False, 1, False, -1.0, 1.0),
    (2000, 2500, 288, True, 1, False, 0.5, 0.3),
    (2000, 2500, 288, False, 1, False, 0.5, 0.3),
    (1024, 2560, 256, False, 2, False, 1.0, 0.0),
    (3072, 3072, 512, True, 1, True, -1.0, 1.0),
    (512, 28392, 512, False, 1, False, -1.0, 1.0),
]


@pytest.mark.parametrize('case', CASES)
def test_gemm_tc_matches_f64(cuda, case):
    from mxfusion_b200 import _raw, _lib
    m, n, k, tb, S, tri, alpha, beta = case
    rng = np.random.RandomState(m + n + k)
    A = rng.randn(S, m, k).astype(np.float32)
    B = (rng.randn(S, n, k) if tb else rng.randn(S, k, n)).astype(np.float32)
    C0 = rng.randn(S, m, n).astype(np.float32)
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    Bop = np.swapaxes(B64, -1, -2) if tb else B64
    want = alpha * (A64 @ Bop) + beta * C0.astype(np.float64)
    bound = 4e-6 * (np.abs(alpha) * (np.abs(A64) @ np.abs(Bop)) + np.abs(beta * C0)) + 1e-7
    C = torch.as_tensor(C0.copy(), device=cuda)
    _raw.gemm(torch.as_tensor(A, device=cuda), torch.as_tensor(B, device=cuda), False, tb, alpha=alpha, beta=beta,
              C=C, tri=tri)
    got = C.cpu().numpy().astype(np.float64)
    if tri:
        # only tiles touching the lower triangle are defined; compare the lower triangle, rest must be untouched
        mask = np.tril(np.ones((m, n), dtype=bool))
        err = np.abs(got - want)[:, mask]
        assert np.all(err <= bound[:, mask]), float(np.max(err / bound[:, mask]))
    else:
        err = np.abs(got - want)
        assert np.all(err <= bound), (float(np.max(err / bound)), float(np.max(err)))
