import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import _raw as R, _lib
dev = torch.device('cuda:0')
lib = _lib.lib()
prof = torch.zeros((32,), dtype=torch.int64, device=dev)
lib.mxf_debug_set_prof.argtypes = [ctypes.c_void_p]
M = 128
W = torch.randn((M, M), device=dev)
A0 = (W @ W.t() / M + torch.eye(M, device=dev)).unsqueeze(0)
R.potrf_packed_(A0.clone())
lib.mxf_debug_set_prof(ctypes.c_void_p(prof.data_ptr()))
R.potrf_packed_(A0.clone())
torch.cuda.synchronize()
p = prof.cpu().tolist()
names = ['load'] + ['chol%d' % j if k == 0 else 'solve%d' % j if k == 1 else 'upd%d' % j for j in range(4) for k in range(3)] + ['facdone', 'invdiag', 'invoff', 'store']
print('stamps', p[:17])
for i in range(1, 17):
    if p[i]:
        prev = max(x for x in p[:i] if x)
        print(i, names[i] if i < len(names) else '', p[i] - prev)
lib.mxf_debug_set_prof(ctypes.c_void_p(0))
