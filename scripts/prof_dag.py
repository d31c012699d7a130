"""Phase stamps of the diagonal tickets of the tile-dataflow potrf (csrc/chol_dag.cu): where the critical path goes."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import _raw, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
Z = torch.rand((n, 8), generator=g) * 6 - 3
A = (torch.exp(-0.5 * torch.cdist(Z.double(), Z.double()) ** 2) + 1e-3 * torch.eye(n, dtype=torch.float64)).float().to(dev)[None]
work = A.clone()
pack = _raw.new_pack(work)
info = torch.zeros((1,), dtype=torch.int32, device=dev)
for _ in range(3):
    work.copy_(A)
    _raw.potrf_packed_(work, info, pack)
torch.cuda.synchronize()
prof = torch.zeros((16 * 16 + 16 * 8,), dtype=torch.int64, device=dev)
fn = _lib.lib().mxf_debug_set_dag_prof
fn.argtypes = [ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(prof.data_ptr()) == 0
work.copy_(A)
_raw.potrf_packed_(work, info, pack)
torch.cuda.synchronize()
fn(None)
pall = prof.cpu().numpy()
p = pall[:256].reshape(16, 16)
inner = pall[256:].reshape(16, 8)
T = (n + 63) // 64
names = ['load', 'upd', 'waitW', 'ldW', 'trsm', 'pubL', 'syrk', 'diag', 'store', 'pubW']
print('ticket  ' + ' '.join('%7s' % s for s in names[1:]) + '   start_us   end_us (globaltimer, rel. to ticket 0)')
t0 = p[0, 14]
for c in range(T):
    d = [p[c, i + 1] - p[c, i] for i in range(9)]
    if c == 0:
        d = [0] * 6 + [p[0, 7] - p[0, 0], p[0, 8] - p[0, 7], p[0, 9] - p[0, 8]]
    print('%6d  ' % c + ' '.join('%7d' % x for x in d) + '   %8.1f %8.1f' % ((p[c, 14] - t0) / 1e3, (p[c, 15] - t0) / 1e3))
print('diag block phases (cycles): identity+panel0  upd0+panel1  upd1+panel2  upd2+panel3  upd3')
for c in range(T):
    q = inner[c]
    print('%6d  %8d %8d %8d %8d %8d' % (c, q[0] - q[5], q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3]))
