"""Per-kernel GPU time of the graph-replayed SVGP step (torch profiler / CUPTI), headline shapes."""
import os, sys, json, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity

dev = torch.device('cuda:0')
n = 65536
X, Y, Z = bench.synthetic(n=n)
infr, loop = bench.build_inference(X, Y, Z, bench.N_ROWS, 1, data_resident=True, device=dev)
state = {}
STEPS, WARM = 20, 6

def on_step(k, loss):
    if k == WARM:
        torch.cuda.synchronize()
        state['prof'] = profile(activities=[ProfilerActivity.CUDA])
        state['prof'].__enter__()
    elif k == WARM + STEPS:
        torch.cuda.synchronize()
        state['prof'].__exit__(None, None, None)

infr.run(X=X, Y=Y, max_iter=100, learning_rate=1e-2, max_steps=WARM + STEPS, on_step=on_step)
ev = state['prof'].events()
tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    if e.device_type is not None and 'cuda' in str(e.device_type).lower():
        name = e.name[:80]
        tot[name][0] += 1
        tot[name][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
s = sum(v[1] for v in tot.values())
print('total kernel time per step: %.1f us (sum over kernels, streams overlap not removed)' % (s / STEPS))
rows = sorted(tot.items(), key=lambda kv: -kv[1][1])
for k, v in rows[:40]:
    print('%-82s %5.1f /step %9.1f us/step %5.1f%%' % (k, v[0] / STEPS, v[1] / STEPS, 100 * v[1] / s))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump([(k, v[0] / STEPS, v[1] / STEPS) for k, v in rows], open(os.path.join(ROOT, 'gpurun_out', 'profile_step.json'), 'w'), indent=1)

# ---- timeline of one replayed step: (start offset us, duration us, name) for every kernel >= 6 us ----------------
kern = [e for e in ev if e.device_type is not None and 'cuda' in str(e.device_type).lower()]
kern.sort(key=lambda e: e.time_range.start)
# one step = the kernels between two consecutive adam_kernel launches (take the middle of the trace)
adam = [i for i, e in enumerate(kern) if 'adam_kernel' in e.name]
if len(adam) > 12:
    a, b = adam[10], adam[11]
    t0 = kern[a].time_range.end
    print('--- timeline of one step (offsets from the end of the previous Adam), kernels >= 6 us; span %.1f us' %
          (kern[b].time_range.end - t0))
    last_end = t0
    for e in kern[a + 1:b + 1]:
        d = e.time_range.end - e.time_range.start
        if d >= 6:
            nm = e.name.replace('void mxf::', '').replace('void at::native::', 'at::')[:60]
            print('%8.1f %7.1f  %s' % (e.time_range.start - t0, d, nm))
