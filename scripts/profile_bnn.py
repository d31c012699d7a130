"""Per-kernel GPU time of the graph-replayed BNN step (config 4) with the torch profiler (CUPTI)."""
import collections, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0], '1', '--no-cpu']
from torch.profiler import profile, ProfilerActivity
import mxfusion_b200 as mf
from mxfusion_b200.components.distributions import Normal
from mxfusion_b200.components.functions import MXFusionGluonFunction
from mxfusion_b200.inference import (GradBasedInference, StochasticVariationalInference, create_Gaussian_meanfield,
                                     MinibatchInferenceLoop)
N, B, H, S = 100000, 4096, 50, 3
dev = torch.device('cuda:0')
mf.config.DEFAULT_DTYPE = 'float32'
g = torch.Generator().manual_seed(0)
x = torch.rand((N, 1), generator=g) * 2 - 1
y = torch.sin(3 * x) + 0.05 * torch.randn((N, 1), generator=g)
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(1, H), torch.nn.Tanh(), torch.nn.Linear(H, H), torch.nn.Tanh(), torch.nn.Linear(H, 1))
m = mf.Model()
m.N = mf.Variable()
m.f = MXFusionGluonFunction(net, num_outputs=1, broadcastable=False)
m.x = mf.Variable(shape=(m.N, 1))
m.v = mf.Variable(shape=(1,), transformation=mf.components.PositiveTransformation(), initial_value=0.01)
m.r = m.f(m.x)
for _, v in m.r.factor.parameters.items():
    v.set_prior(Normal(mean=torch.tensor([0.]), variance=torch.tensor([1.])))
m.y = Normal.define_variable(mean=m.r, variance=m.v, shape=(m.N, 1))
q = create_Gaussian_meanfield(model=m, observed=[m.y, m.x])
alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=[m.y, m.x])
loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.y: N / float(B)}, rng=np.random.RandomState(0))
infr = GradBasedInference(inference_algorithm=alg, grad_loop=loop, context=dev)
infr.initialize(y=(N, 1), x=(N, 1))
for _, v in m.r.factor.parameters.items():
    infr.params[q[v].factor.mean] = v.initial_value
    infr.params[q[v].factor.variance] = torch.full(v.shape, 1e-6)
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


infr.run(max_iter=3, learning_rate=1e-3, max_steps=WARM + STEPS, on_step=on_step, y=y, x=x)
tot = collections.defaultdict(lambda: [0, 0.0])
for e in state['prof'].events():
    if e.device_type is not None and 'cuda' in str(e.device_type).lower():
        tot[e.name[:90]][0] += 1
        tot[e.name[:90]][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
s = sum(v[1] for v in tot.values())
print('kernel time per step %.1f us over %.1f launches' % (s / STEPS, sum(v[0] for v in tot.values()) / STEPS))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
    print('%-92s %5.1f /step %7.1f us/step' % (k, v[0] / STEPS, v[1] / STEPS))
