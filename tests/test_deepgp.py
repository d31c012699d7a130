"""Deep GP (BASELINE config 5) -- NOT in the reference, so "unpinned by the reference"; pinned instead by
(a) the one-layer reduction to the reference's SVGP fixture (known answer -32.72563540745786) and (b) an independent dense
formulation (oracle/deepgp.py) with injected noise, value and gradients.  CPU tier: the host layer on the stand-in
binding; GPU tier (test_gpu_deepgp.py) runs the same scenarios on the kernels."""
import numpy as np
import pytest
import torch

from oracle import deepgp as odgp
from tests.test_host_api import mf  # noqa: F401  (fixture)


def build_dgp(mf, X, Y, widths, M, num_samples, eps, kinds=('rbf', 'matern52'), jitter=1e-6, scaling=1.0, seed=0, device=None):
    """widths = [D, H1, ..]; returns (model, infr, executor-ready dict of the parameter values)."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF, Matern52
    from mxfusion_b200.components.distributions.random_gen import MockMXNetRandomGenerator
    from mxfusion_b200.modules.gp_modules import DeepGPRegression
    from mxfusion_b200.inference import Inference, MAP
    rng = np.random.RandomState(seed)
    nl = len(widths)
    P = Y.shape[1]
    kerns, vals = [], dict(Z=[], ls=[], var=[], m=[], W=[], d=[])
    for l in range(nl):
        cls = RBF if kinds[l % len(kinds)] == 'rbf' else Matern52
        ls, var = rng.uniform(0.7, 1.5, (1,)), rng.uniform(0.6, 1.4, (1,))
        kerns.append(cls(input_dim=widths[l], variance=var, lengthscale=ls, name='k_l%d' % l))
        out = widths[l + 1] if l + 1 < nl else P
        vals['Z'].append(rng.uniform(-2, 2, (M, widths[l])))
        vals['ls'].append(ls)
        vals['var'].append(var)
        vals['m'].append(0.5 * rng.randn(M, out))
        vals['W'].append(0.3 * rng.randn(M, M) / np.sqrt(M))
        vals['d'].append(rng.uniform(0.3, 0.9, (M,)))
    vals['noise'] = 0.07
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, widths[0]))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=vals['noise'])
    zs = [mf.Variable(shape=z.shape, initial_value=z) for z in vals['Z']]
    for l, z in enumerate(zs):
        setattr(m, 'Z%d' % l, z)
    gen = None
    if eps is not None:
        flat = torch.cat([torch.as_tensor(e).reshape(-1) for e in eps]) if len(eps) > 1 else torch.as_tensor(eps[0]).reshape(-1)
        gen = _SeqGen([torch.as_tensor(e) for e in eps])
    m.Y = DeepGPRegression.define_variable(X=m.X, kernels=kerns, noise_var=m.noise_var, inducing_inputs=zs, shape=(m.N, P),
                                           rand_gen=gen)
    gp = m.Y.factor
    gp.dgp_log_pdf.jitter = jitter
    gp.dgp_log_pdf.num_samples = num_samples
    infr = Inference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    post = gp._extra_graphs[0]
    for l in range(nl):
        infr.params[post.qU_mean[l]] = vals['m'][l]
        infr.params[post.qU_cov_W[l]] = vals['W'][l]
        infr.params[post.qU_cov_diag[l]] = vals['d'][l]
    return m, infr, vals


class _SeqGen(object):
    """Injected noise, one tensor per hidden layer in order (the MockMXNetRandomGenerator pattern, testutils.py:58-93)."""
    in_kernel = False

    def __init__(self, tensors):
        self.tensors, self.i = tensors, 0

    def sample_normal(self, loc=0, scale=1, shape=None, dtype=None, out=None, ctx=None):
        t = self.tensors[self.i % len(self.tensors)]
        self.i += 1
        return t.reshape(shape).to(ctx) if ctx is not None else t.reshape(shape)


def oracle_value(vals, X, Y, eps, kinds, jitter, scaling):
    from oracle import kernels as ok
    kk = [ok.RBF if k == 'rbf' else ok.MATERN52 for k in kinds]
    nl = len(vals['Z'])
    return odgp.dgp_elbo_np([kk[l % len(kk)] for l in range(nl)], X, Y, vals['Z'], vals['ls'], vals['var'], vals['m'], vals['W'],
                            vals['d'], vals['noise'], eps, jitter=jitter, scale=scaling)


def test_one_layer_deep_gp_is_the_svgp_bound_on_the_reference_fixture(mf):  # noqa: F811
    """testing/modules/svgpregression_test.py:41-56 fixture; ELBO = -32.72563540745786 (BASELINE.md known answer)."""
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import DeepGPRegression
    from mxfusion_b200.inference import Inference, MAP
    np.random.seed(0)
    X, Y, Z = np.random.rand(10, 3), np.random.rand(10, 1), np.random.rand(3, 3)
    qU_mean, qU_cov_W, qU_cov_diag = np.random.rand(3, 1), np.random.rand(3, 3), np.random.rand(3,)
    noise_var, lengthscale, variance = np.random.rand(1), np.random.rand(3), np.random.rand(1)
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 3))
    m.Z = mf.Variable(shape=(3, 3), initial_value=Z)
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=noise_var)
    kernel = RBF(input_dim=3, ARD=True, variance=variance, lengthscale=lengthscale)
    m.Y = DeepGPRegression.define_variable(X=m.X, kernels=[kernel], noise_var=m.noise_var, inducing_inputs=[m.Z], shape=(m.N, 1))
    gp = m.Y.factor
    gp.dgp_log_pdf.jitter = 1e-8
    infr = Inference(MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    post = gp._extra_graphs[0]
    infr.params[post.qU_mean[0]] = qU_mean
    infr.params[post.qU_cov_W[0]] = qU_cov_W
    infr.params[post.qU_cov_diag[0]] = qU_cov_diag
    loss, _ = infr.run(X=X, Y=Y)
    assert abs(-float(loss) - (-32.72563540745786)) < 1e-9


@pytest.mark.parametrize('widths,P', [([3, 3], 1), ([4, 2], 2), ([2, 2, 2], 1)])
def test_deep_gp_bound_and_gradients_match_the_dense_oracle(mf, widths, P):  # noqa: F811
    rng = np.random.RandomState(1)
    B, M, S = 17, 6, 3
    X = rng.uniform(-2, 2, (B, widths[0]))
    Y = rng.randn(B, P)
    eps = [rng.randn(S, B, widths[l + 1]) for l in range(len(widths) - 1)]
    kinds = ('rbf', 'matern52')
    model, infr, vals = build_dgp(mf, X, Y, widths, M, S, eps, kinds=kinds, jitter=1e-6, scaling=2.5)
    ex = infr.inference_algorithm.create_executor(data_def=infr.observed_variable_UUIDs, params=infr.params,
                                                  var_ties=infr.params.var_ties, rv_scaling={model.Y.uuid: 2.5})
    loss, lg = ex(None, torch.tensor(X), torch.tensor(Y))
    want = oracle_value(vals, X, Y, eps, kinds, 1e-6, 2.5)
    np.testing.assert_allclose(-float(loss), want.mean(), rtol=1e-9)
    # gradients with respect to the (constrained-space) values, via the torch restatement of the dense formulation
    from oracle import torch_ref
    kk = [torch_ref.RBF if k == 'rbf' else torch_ref.MATERN52 for k in kinds]
    nl = len(widths)
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    T = dict(Z=[t(z) for z in vals['Z']], ls=[t(v) for v in vals['ls']], var=[t(v) for v in vals['var']],
             m=[t(v) for v in vals['m']], W=[t(v) for v in vals['W']], d=[t(v) for v in vals['d']], noise=t(vals['noise']))
    ref = odgp.dgp_elbo_torch([kk[l % 2] for l in range(nl)], torch.tensor(X), torch.tensor(Y), T['Z'], T['ls'], T['var'], T['m'],
                              T['W'], T['d'], T['noise'], [torch.tensor(e) for e in eps], jitter=1e-6, scale=2.5)
    (-ref).backward()
    lg.backward()
    gp = model.Y.factor
    post = gp._extra_graphs[0]
    sig = lambda u: 1.0 / (1.0 + np.exp(-u))
    for l in range(nl):
        # unconstrained parameters: chain rule through the softplus (var_trans.py:75) for d
        g = infr.params.param_dict[post.qU_mean[l].uuid].tensor.grad.numpy()
        np.testing.assert_allclose(g, T['m'][l].grad.numpy(), rtol=1e-6, atol=1e-9)
        g = infr.params.param_dict[post.qU_cov_W[l].uuid].tensor.grad.numpy()
        np.testing.assert_allclose(g, T['W'][l].grad.numpy(), rtol=1e-6, atol=1e-9)
        p = infr.params.param_dict[post.qU_cov_diag[l].uuid].tensor
        np.testing.assert_allclose(p.grad.numpy(), T['d'][l].grad.numpy() * sig(p.detach().numpy()), rtol=1e-6, atol=1e-9)
        zvar = getattr(model, 'Z%d' % l)
        g = infr.params.param_dict[zvar.uuid].tensor.grad.numpy()
        np.testing.assert_allclose(g, T['Z'][l].grad.numpy(), rtol=1e-6, atol=1e-8)


def test_deep_gp_trains_and_predicts(mf):  # noqa: F811
    """Two layers on a step function: the bound improves under Adam and the prediction algorithm returns finite moments."""
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop, TransferInference
    from mxfusion_b200.inference.prediction import ModulePredictionAlgorithm
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import DeepGPRegression
    rng = np.random.RandomState(2)
    np.random.seed(3)
    N = 120
    X = rng.uniform(-2, 2, (N, 1))
    Y = np.sign(X) + 0.05 * rng.randn(N, 1)
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 1))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.05)
    m.Y = DeepGPRegression.define_variable(X=m.X, kernels=[RBF(1, name='k0'), RBF(1, name='k1')], noise_var=m.noise_var,
                                           shape=(m.N, 1), num_inducing=8)
    m.Y.factor.dgp_log_pdf.jitter = 1e-6
    m.Y.factor.dgp_log_pdf.num_samples = 4
    loop = MinibatchInferenceLoop(batch_size=40, rv_scaling={m.Y: N / 40.}, rng=np.random.RandomState(4))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop)
    losses = [float(l) for l in infr.run(X=X, Y=Y, max_iter=30, learning_rate=0.02)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    xt = np.linspace(-2, 2, 9)[:, None]
    infr2 = TransferInference(ModulePredictionAlgorithm(model=m, observed=[m.X], target_variables=[m.Y]),
                              infr_params=infr.params)
    mu, var = infr2.run(X=xt)[0]
    assert tuple(mu.shape) == (1, 9, 1) and tuple(var.shape) == (1, 9, 1)
    assert bool(torch.isfinite(mu).all()) and bool((var > 0).all())
