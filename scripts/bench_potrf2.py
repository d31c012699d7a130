"""Two n x n factorisations on two streams inside one CUDA graph (what the SVGP step does with Kuu and S) vs one alone."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxfusion_b200 import _raw
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
Z = torch.rand((n, 8), generator=g) * 6 - 3
A = (torch.exp(-0.5 * torch.cdist(Z.double(), Z.double()) ** 2) + 1e-3 * torch.eye(n, dtype=torch.float64)).float().to(dev)[None]
B0, B1 = A.clone(), A.clone()
p0, p1 = _raw.new_pack(B0), _raw.new_pack(B1)
i0 = torch.zeros((1,), dtype=torch.int32, device=dev); i1 = torch.zeros((1,), dtype=torch.int32, device=dev)
side, main = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        _raw.potrf_packed_(B1, i1, p1)
    _raw.potrf_packed_(B0, i0, p0)
    cur.wait_stream(side)


def one():
    _raw.potrf_packed_(B0, i0, p0)


ctas_list = [int(a) for a in sys.argv[2:]] or [_raw.lib().mxf_potrf_dag_ctas(0)]
for ctas, (name, fn) in [(c, nf) for c in ctas_list for nf in (('one', one), ('two concurrent', both))]:
    _raw.lib().mxf_potrf_dag_ctas(ctas)          # CTAs per matrix of the persistent dataflow kernel (captured with the graph)
    main.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(main):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=main):
        fn()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    print('ctas/matrix %d  %s: %.1f us per replay (in place on an already factored matrix: timing only)' % (
        ctas, name, 1e3 * a.elapsed_time(b) / 50))
