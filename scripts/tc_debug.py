import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import _raw
dev = torch.device('cuda:0')
for (m, n, k, tb) in [(128, 128, 256, True), (128, 128, 256, False), (128, 64, 512, True), (256, 512, 1024, True), (256, 512, 1024, False), (2048, 4096, 512, True)]:
    rng = np.random.RandomState(0)
    A = rng.randn(1, m, k).astype(np.float32)
    B = (rng.randn(1, n, k) if tb else rng.randn(1, k, n)).astype(np.float32)
    want = A.astype(np.float64) @ (np.swapaxes(B, -1, -2) if tb else B).astype(np.float64)
    got = _raw.gemm(torch.as_tensor(A, device=dev), torch.as_tensor(B, device=dev), False, tb)
    torch.cuda.synchronize()
    g = got.cpu().numpy().astype(np.float64)
    err = np.abs(g - want)
    print((m, n, k, 'NT' if tb else 'NN'), 'max err', err.max(), 'mean err', err.mean(), 'max |want|', np.abs(want).max(),
          'frac bad', float((err > 1e-2).mean()), flush=True)
    if err.max() > 1e-2:
        bad = np.argwhere(err[0] > 1e-2)
        print('   first bad idx', bad[:5].tolist(), 'rows bad', np.unique(bad[:, 0])[:10], 'cols bad', np.unique(bad[:, 1])[:10])
        print('   got[0,:4,:4]', g[0, :4, :4].round(3).tolist(), 'want', want[0, :4, :4].round(3).tolist())
