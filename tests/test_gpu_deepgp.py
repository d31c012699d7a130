"""Deep GP (BASELINE config 5; no reference implementation -- "unpinned by the reference") on the CUDA kernels: the
scenarios of tests/test_deepgp.py through the real binding, float64 against the independent dense oracle and float32
against float64."""
import numpy as np
import pytest
import torch

from oracle import deepgp as odgp
from tests.test_deepgp import build_dgp, oracle_value

pytestmark = pytest.mark.gpu


@pytest.fixture()
def mfd(monkeypatch, request):
    import mxfusion_b200 as mf
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', request.param)
    return mf


@pytest.mark.parametrize('mfd', ['float64', 'float32'], indirect=True)
@pytest.mark.parametrize('widths,P', [([3, 3], 1), ([4, 2], 2), ([2, 2, 2], 1)])
def test_deep_gp_on_the_kernels_matches_the_dense_oracle(cuda, mfd, widths, P):
    f64 = mfd.config.DEFAULT_DTYPE == 'float64'
    rng = np.random.RandomState(1)
    B, M, S = 300, 24, 3
    X = rng.uniform(-2, 2, (B, widths[0]))
    Y = rng.randn(B, P)
    eps = [rng.randn(S, B, widths[l + 1]) for l in range(len(widths) - 1)]
    kinds = ('rbf', 'matern52')
    jit = 1e-6 if f64 else 1e-4
    tdt = torch.float64 if f64 else torch.float32
    model, infr, vals = build_dgp(mfd, X, Y, widths, M, S, [e.astype(np.float64 if f64 else np.float32) for e in eps],
                                  kinds=kinds, jitter=jit, device=cuda)
    ex = infr.inference_algorithm.create_executor(data_def=infr.observed_variable_UUIDs, params=infr.params,
                                                  var_ties=infr.params.var_ties, rv_scaling={model.Y.uuid: 2.5})
    loss, lg = ex(None, torch.tensor(X, dtype=tdt, device=cuda), torch.tensor(Y, dtype=tdt, device=cuda))
    lg.backward()
    want = oracle_value(vals, X, Y, eps, kinds, jit, 2.5)
    print('MEASURED dgp value rel %.2e' % (abs(-float(loss) - want.mean()) / abs(want.mean())))
    np.testing.assert_allclose(-float(loss), want.mean(), rtol=1e-9 if f64 else 2.5e-4)     # f32 measured <= 2.5e-5
    from oracle import torch_ref
    kk = [torch_ref.RBF if k == 'rbf' else torch_ref.MATERN52 for k in kinds]
    nl = len(widths)
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    T = dict(Z=[t(z) for z in vals['Z']], ls=[t(v) for v in vals['ls']], var=[t(v) for v in vals['var']],
             m=[t(v) for v in vals['m']], W=[t(v) for v in vals['W']], d=[t(v) for v in vals['d']], noise=t(vals['noise']))
    ref = odgp.dgp_elbo_torch([kk[l % 2] for l in range(nl)], torch.tensor(X), torch.tensor(Y), T['Z'], T['ls'], T['var'], T['m'],
                              T['W'], T['d'], T['noise'], [torch.tensor(e) for e in eps], jitter=jit, scale=2.5)
    (-ref).backward()
    post = model.Y.factor._extra_graphs[0]
    gmax = max(float(T[k][l].grad.abs().max()) for k in ('m', 'W', 'Z') for l in range(nl))
    # f32 (jitter 1e-4) measured: two layers <= 1.5e-4 of the gradient's own max-norm; three layers 3e-2 / 1.3e-2 of the
    # largest gradient entry (three ill-conditioned solves in a row)
    rel, floor = (1e-6, 1e-9) if f64 else ((2e-3, 1e-3) if nl == 2 else (1e-1, 5e-2))
    for l in range(nl):
        for name, var, w in (('m', post.qU_mean[l], T['m'][l].grad), ('W', post.qU_cov_W[l], T['W'][l].grad),
                             ('Z', getattr(model, 'Z%d' % l), T['Z'][l].grad)):
            g = infr.params.param_dict[var.uuid].tensor.grad.double().cpu().numpy()
            err = float(np.max(np.abs(g - w.numpy())))
            print('MEASURED dgp grad %s%d own %.2e gmax %.2e' % (name, l, err / float(w.abs().max()), err / gmax))
            assert err <= rel * float(w.abs().max()) + floor * gmax, (l, name, err, float(w.abs().max()), gmax)


def test_deep_gp_trains_on_the_gpu_with_in_kernel_noise(cuda, monkeypatch):
    """Two layers, float32, CUDA-graph step with the in-kernel Philox draws (fresh noise on every replay)."""
    import mxfusion_b200 as mf
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import DeepGPRegression
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float32')
    rng = np.random.RandomState(2)
    np.random.seed(3)
    N, B = 4096, 512
    X = rng.uniform(-2, 2, (N, 2))
    Y = np.sign(X[:, :1]) * np.cos(X[:, 1:]) + 0.05 * rng.randn(N, 1)
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 2))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.05)
    m.Y = DeepGPRegression.define_variable(X=m.X, kernels=[RBF(2, name='k0'), RBF(2, name='k1')], noise_var=m.noise_var,
                                           shape=(m.N, 1), num_inducing=32)
    m.Y.factor.dgp_log_pdf.jitter = 1e-4
    m.Y.factor.dgp_log_pdf.num_samples = 4
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / float(B)}, rng=np.random.RandomState(4))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, context=cuda)
    losses = [float(l) for l in infr.run(X=X.astype(np.float32), Y=Y.astype(np.float32), max_iter=12, learning_rate=0.02)]
    assert np.isfinite(losses).all() and np.mean(losses[-3:]) < np.mean(losses[:3])
