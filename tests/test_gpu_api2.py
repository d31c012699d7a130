"""Round-2 GPU tests through the public API: the mean-field BNN MC-ELBO (BASELINE config 4) against an injected-noise
float64 restatement, a non-positive-definite factorisation surfacing as InferenceError, and the 2-rank NCCL step."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ShapeKeyedNoise(object):
    """Injected standard normals (the MockMXNetRandomGenerator pattern, testutils.py:58-93): one fixed tensor per requested
    shape, so that the restatement can use the same draws whatever the order of the graph walk."""
    in_kernel = False

    def __init__(self, seed, device='cuda:0'):
        self.rng, self.draws, self.device = np.random.RandomState(seed), {}, device

    def sample_normal(self, loc=0, scale=1, shape=None, dtype=None, out=None, ctx=None):
        key = tuple(shape)
        if key not in self.draws:
            self.draws[key] = self.rng.randn(*key)
        return torch.as_tensor(self.draws[key], dtype=torch.float64 if dtype in ('float64', None) else torch.float32,
                               device=ctx if ctx is not None else self.device)


def test_bnn_mc_elbo_through_the_api_matches_injected_noise_restatement(cuda, monkeypatch):
    """variational.py:91-108 + normal.py:52-92 + function_evaluation.py:72-96 on the kernels (multi-tensor Normal log-pdf,
    fused MLP) vs torch float64 with the same weight noise: loss and every gradient."""
    import mxfusion_b200 as mf
    from mxfusion_b200.components.distributions import Normal
    from mxfusion_b200.components.functions import MXFusionGluonFunction
    from mxfusion_b200.inference import (GradBasedInference, StochasticVariationalInference, create_Gaussian_meanfield,
                                         BatchInferenceLoop)
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float64')
    H1, H2, S, N = 8, 6, 3, 257
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(1, H1), torch.nn.Tanh(), torch.nn.Linear(H1, H2), torch.nn.Tanh(),
                              torch.nn.Linear(H2, 1)).double()
    rng = np.random.RandomState(0)
    x = rng.rand(N, 1) * 2 - 1
    y = np.sin(3 * x) + 0.05 * rng.randn(N, 1)
    m = mf.Model()
    m.N = mf.Variable()
    m.f = MXFusionGluonFunction(net, num_outputs=1, broadcastable=False)
    m.x = mf.Variable(shape=(m.N, 1))
    m.v = mf.Variable(shape=(1,), transformation=mf.components.PositiveTransformation(), initial_value=0.01)
    m.r = m.f(m.x)
    for _, v in m.r.factor.parameters.items():
        v.set_prior(Normal(mean=torch.tensor([0.]).double(), variance=torch.tensor([1.]).double()))
    m.y = Normal.define_variable(mean=m.r, variance=m.v, shape=(m.N, 1))
    observed = [m.y, m.x]
    q = create_Gaussian_meanfield(model=m, observed=observed)
    noise = ShapeKeyedNoise(11)
    weights = list(m.r.factor.parameters.items())
    for _, v in weights:
        q[v].factor._rand_gen = noise
    alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=observed)
    infr = GradBasedInference(inference_algorithm=alg, grad_loop=BatchInferenceLoop(), context=cuda)
    infr.initialize(y=y.shape, x=x.shape)
    qvar = 0.05
    for _, v in weights:
        infr.params[q[v].factor.mean] = v.initial_value
        infr.params[q[v].factor.variance] = torch.full(v.shape, qvar).double()
    scale = 3.0
    ex = alg.create_executor(data_def=infr.observed_variable_UUIDs, params=infr.params, var_ties=infr.params.var_ties,
                             rv_scaling={m.y.uuid: scale})
    loss, lg = ex(None, torch.tensor(y, device=cuda), torch.tensor(x, device=cuda))
    lg.backward()
    # ---- restatement (float64, CPU): per-sample loop over the network, Normal log-densities, mean over samples ----------
    sp = torch.nn.functional.softplus
    inv_sp = lambda v: np.log(np.expm1(v))
    mus = [v.initial_value.clone().double().requires_grad_() for _, v in weights]
    rhos = [torch.full(v.shape, inv_sp(qvar), dtype=torch.float64).requires_grad_() for _, v in weights]
    nv_u = torch.tensor([inv_sp(0.01)], dtype=torch.float64, requires_grad=True)
    ws = [mu.unsqueeze(0) + torch.as_tensor(noise.draws[(S,) + tuple(mu.shape)]) * torch.sqrt(sp(r)).unsqueeze(0)
          for mu, r in zip(mus, rhos)]

    def logn(v, mean, var):
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(var) - torch.square(v - mean) / (2 * var)
    xt, yt = torch.tensor(x), torch.tensor(y)
    outs = []
    for s in range(S):
        h = torch.tanh(xt @ ws[0][s].T + ws[1][s])
        h = torch.tanh(h @ ws[2][s].T + ws[3][s])
        outs.append((h @ ws[4][s].T + ws[5][s]).unsqueeze(0))
    f = torch.cat(outs, 0)
    logp = scale * logn(yt.unsqueeze(0), f, sp(nv_u)).mean(0).sum()
    for w in ws:
        logp = logp + logn(w, torch.zeros((), dtype=torch.float64), torch.ones((), dtype=torch.float64)).mean(0).sum()
    logq = sum(logn(w, mu.unsqueeze(0), sp(r).unsqueeze(0)).mean(0).sum() for w, mu, r in zip(ws, mus, rhos))
    want = -(logp - logq)
    want.backward()
    np.testing.assert_allclose(float(loss.detach()), float(want.detach()), rtol=2e-8)      # measured 1.3e-9
    def close(g, w, what):        # every gradient within 1e-7 of its max-norm (measured 5e-9; float64 kernels)
        assert np.max(np.abs(g - w)) <= 1e-7 * np.max(np.abs(w)), (what, np.max(np.abs(g - w)), np.max(np.abs(w)))
    for (name, v), mu, r in zip(weights, mus, rhos):
        g = infr.params.param_dict[q[v].factor.mean.uuid].tensor.grad.cpu().numpy()
        close(g, mu.grad.numpy(), name + ' mean')
        g = infr.params.param_dict[q[v].factor.variance.uuid].tensor.grad.cpu().numpy()
        close(g, r.grad.numpy(), name + ' variance')
    g = infr.params.param_dict[m.v.uuid].tensor.grad.cpu().numpy()
    np.testing.assert_allclose(g, nv_u.grad.numpy(), rtol=1e-6)


def test_non_positive_definite_factorisation_raises_inference_error(cuda, monkeypatch):
    """svgp_regression.py:70-72 / gp_regression.py:58-60: without enough jitter potrf fails; the reference surfaces an
    MXNetError at its next synchronisation, here the device-side `info` is raised as InferenceError by the loop."""
    import mxfusion_b200 as mf
    from mxfusion_b200.common.exceptions import InferenceError
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float32')
    rng = np.random.RandomState(0)
    N, M, B = 512, 64, 128
    X = rng.uniform(-1, 1, (N, 1)).astype(np.float32)
    Y = np.sin(X).astype(np.float32)
    Z = np.repeat(rng.uniform(-1, 1, (M // 2, 1)), 2, axis=0).astype(np.float32)      # duplicated inducing points: Kuu singular
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 1))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.Z = mf.Variable(shape=(M, 1), initial_value=Z)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=RBF(1, lengthscale=3.0), noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, 1))
    m.Y.factor.svgp_log_pdf.jitter = 0.
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / float(B)}, rng=np.random.RandomState(1))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, context=cuda)
    with pytest.raises(InferenceError, match='positive definite'):
        infr.run(X=X, Y=Y, max_iter=2, learning_rate=1e-2)
    # the record is cleared by the raise: later runs on this device are not poisoned
    from mxfusion_b200 import ops
    assert int(ops.info_accumulator(cuda).item()) == 0


_NCCL_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import mxfusion_b200 as mf
from tests.test_gpu_api import _svgp_model
from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
mf.config.DEFAULT_DTYPE = 'float64'
np.random.seed(0)
N, B, M = 800, 100, 16
X = np.random.uniform(-3., 3., (N, 1)); Y = np.sin(X) + np.random.randn(N, 1) * 0.05
shard = N // world
Xs, Ys = X[rank * shard:(rank + 1) * shard], Y[rank * shard:(rank + 1) * shard]
np.random.seed(5 + rank)                       # default inducing inputs differ per rank: the stepper must broadcast rank 0's
m = _svgp_model(mf, N, 1, M, B)
loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / B}, rng=np.random.RandomState(7))
infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, context=dev)
infr.initialize(X=(shard, 1), Y=(shard, 1))
infr.params[m.Y.factor._extra_graphs[0].qU_cov_W] = np.eye(M) * 0.1
infr.run(X=Xs, Y=Ys, max_iter=1, learning_rate=0.05, max_steps=3)
flat = infr.params.flat.detach().cpu().numpy()
np.save(os.path.join(%r, 'flat_rank%%d.npy' %% rank), flat)
open(os.path.join(%r, 'exchange_rank%%d.txt' %% rank), 'w').write(infr._grad_loop.last_stepper.exchange)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_step_equals_the_concatenated_batch_step(cuda, tmp_path, monkeypatch):
    """The GPU twin of tests/test_dist_gloo.py: 2 ranks (NCCL), each on its half of the rows with rv_scaling N/B and the
    all-reduced gradient averaged, take the same 3 Adam steps as ONE process on the concatenated 2B batches with
    rv_scaling N/(2B); the replicas end bit-identical to each other and the initial parameters come from rank 0."""
    import mxfusion_b200 as mf
    from tests.test_gpu_api import _svgp_model
    from mxfusion_b200.inference.minibatch_loop import RolloverBatchSampler
    from oracle import torch_ref
    script = tmp_path / 'worker.py'
    script.write_text(_NCCL_WORKER % (ROOT, str(tmp_path), str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29741')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
                        '127.0.0.1', '--master-port', '29741', str(script)], env=env, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stderr[-3000:]
    f0, f1 = np.load(tmp_path / 'flat_rank0.npy'), np.load(tmp_path / 'flat_rank1.npy')
    np.testing.assert_array_equal(f0, f1)
    exch = [(tmp_path / ('exchange_rank%d.txt' % r)).read_text() for r in range(2)]
    print('gradient exchange used by the ranks:', exch)
    assert exch[0] == exch[1] and exch[0] in ('p2p-kernel', 'nccl')
    # single process on the concatenated batches
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float64')
    np.random.seed(0)
    N, B, M, world = 800, 100, 16, 2
    X = np.random.uniform(-3., 3., (N, 1))
    Y = np.sin(X) + np.random.randn(N, 1) * 0.05
    shard = N // world
    np.random.seed(5)
    m = _svgp_model(mf, N, 1, M, B)
    from mxfusion_b200.inference import GradBasedInference, MAP
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), context=cuda)
    infr.initialize(X=(2 * B, 1), Y=(2 * B, 1))
    infr.params[m.Y.factor._extra_graphs[0].qU_cov_W] = np.eye(M) * 0.1
    ex = infr.inference_algorithm.create_executor(data_def=infr.observed_variable_UUIDs, params=infr.params,
                                                  var_ties=infr.params.var_ties, rv_scaling={m.Y.uuid: N / (2.0 * B)})
    samplers = [RolloverBatchSampler(shard, B, rng=np.random.RandomState(7)) for _ in range(world)]
    idx = [s.epoch_indices()[0] for s in samplers]
    from mxfusion_b200 import ops
    p = infr.params
    p.refresh_leaves()
    for step in range(3):
        rows = np.concatenate([r * shard + idx[r][step * B:(step + 1) * B] for r in range(world)])
        p.gflat.zero_()
        loss, lg = ex(None, torch.tensor(X[rows], device=cuda), torch.tensor(Y[rows], device=cuda))
        lg.backward()
        ops.R.adam_step_(p.flat, p.gflat, p.adam_m, p.adam_v, p.adam_t, lr=0.05, rescale=1.0 / B)     # g_concat is already the ranks' average
    want = p.flat.detach().cpu().numpy()
    diff = np.abs(f0 - want)
    k = int(np.argmax(diff))
    print('max |diff| %.3e at %d (values %.6e / %.6e); > 1e-9: %d of %d' % (diff[k], k, f0[k], want[k], int((diff > 1e-9).sum()), diff.size))
    # float64; the two formulations sum the batch in a different order (all-reduce of two B-row gradients vs one 2B-row
    # launch with atomics), and Adam's first steps turn a relative gradient difference d into a relative update difference ~d
    np.testing.assert_allclose(f0, want, rtol=1e-8, atol=1e-10)      # measured max |diff| 2.4e-11


_P2P_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from mxfusion_b200.inference._p2p import PeerBucket
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
out = {}
for dt, name in ((torch.float32, 'f32'), (torch.float64, 'f64')):
    for n in (1, 7, 4096, 1085447):
        b = PeerBucket(n, dt, dev)
        g = torch.Generator(device='cpu').manual_seed(100 * rank + n %% 97)
        for rep in range(3):                       # repeated calls: the flags reset themselves
            mine = torch.randn(n, generator=g, dtype=dt).to(dev)
            b.grad.copy_(mine)
            got = b.all_reduce_(scale=0.5).clone()
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            want = parts[0].clone()
            for q in parts[1:]:
                want += q                           # rank order, like the kernel
            want *= 0.5
            assert torch.equal(got, want), (name, n, rep, float((got - want).abs().max()))
        b.check()
        out['%%s_%%d' %% (name, n)] = got.cpu().numpy()
# the launch is capturable: 5 replays of [fill -> all-reduce] inside one CUDA graph each
b = PeerBucket(4099, torch.float32, dev)
src = torch.full((4099,), float(rank + 1), device=dev)
s = torch.cuda.Stream(device=dev)
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    b.grad.copy_(src); b.all_reduce_()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    b.grad.copy_(src)
    b.all_reduce_()
for rep in range(5):
    src.fill_(float(rank + 1 + rep))
    g.replay()
    torch.cuda.synchronize()
    want = float(sum(r + 1 + rep for r in range(world)))
    assert bool((b.grad == want).all()), (rep, float(b.grad[0]), want)
b.check()
np.savez(os.path.join(%r, 'p2p_rank%%d.npz' %% rank), **out)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_peer_memory_allreduce_kernel(tmp_path):
    """csrc/allreduce_p2p.cu on two B200s: in-place scale * sum over the ranks, bit-identical to the rank-ordered sum and
    identical on both ranks, f32 / f64, sizes that are not multiples of 16 bytes or of the rank count, repeated calls
    (self-resetting flags), and replayed from a CUDA graph."""
    script = tmp_path / 'p2p_worker.py'
    script.write_text(_P2P_WORKER % (ROOT, str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29743', MXF_P2P_TIMEOUT_S='10')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
                        '127.0.0.1', '--master-port', '29743', str(script)], env=env, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stderr[-3000:]
    a, b = np.load(tmp_path / 'p2p_rank0.npz'), np.load(tmp_path / 'p2p_rank1.npz')
    for k in a.files:
        np.testing.assert_array_equal(a[k], b[k])
