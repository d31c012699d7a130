"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python source (/root/reference) on the stand-in
`mxnet` package of tests/golden/_mxnet_standin (torch CPU float64; see its __init__ for what that pins).

    python tests/golden/make_golden.py          # only in the build container: /root/reference must exist

Scenarios follow the reference's tests and notebooks:
  kernels   testing/components/distributions/gp/kernel_test.py (RBF / Matern, ARD, sample axis, active_dims)
  svgp      testing/modules/svgpregression_test.py:41-115 fixture (+ Matern52, P=2, rv_scaling), value and gradients
  gp        testing/modules/gpregression_test.py:40-96 fixture, value, gradients, cached L / LinvY
  gp_nb     examples/notebooks/gp_regression.ipynb: loss at init, 100 Adam steps (printed -16.903135093930537)
  svgp_mb   MinibatchInferenceLoop trajectory (shuffled rollover batches, rv_scaling, grads / B)
  normal    testing/components/distributions/normal_test.py:35-109 (log_pdf, reparameterised draw with injected eps)
  svi       StochasticVariationalInference on a conjugate toy model with injected posterior samples
  gpdist    GaussianProcess / ConditionalGaussianProcess distributions (testing/components/distributions/gp/gp_test.py,
            cond_gp_test.py): log-pdf and draws with injected standard normals, with a sample axis, P = 1 and 3
  combo     kernel algebra (kernel_test.py:163-300: Linear / Bias / White / Add / Multiply) and the three GP modules with
            combination kernels: values and gradients
  sparsegp  testing/modules/sparsegpregression_test.py:41-196 fixture (+ Matern, P=1, larger M): bound, gradients,
            cached wv / L / LA, mean / variance prediction in the four modes
Each .npz stores the inputs next to the outputs, so tests need nothing but the file.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_mxnet_standin'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import mxnet as mx  # noqa: E402  (the stand-in)
import mxfusion  # noqa: E402  (the reference)
from mxfusion import Model, Variable  # noqa: E402
from mxfusion.common import config  # noqa: E402
from mxfusion.components.variables import PositiveTransformation  # noqa: E402
from mxfusion.components.distributions import Normal  # noqa: E402
from mxfusion.components.distributions.gp.gp import GaussianProcess  # noqa: E402
from mxfusion.components.distributions.gp.cond_gp import ConditionalGaussianProcess  # noqa: E402
from mxfusion.components.distributions.gp.kernels import RBF, Matern12, Matern32, Matern52, Linear, Bias, White  # noqa: E402
from mxfusion.modules.gp_modules import GPRegression, SVGPRegression, SparseGPRegression  # noqa: E402
from mxfusion.inference import (Inference, GradBasedInference, MAP, BatchInferenceLoop, MinibatchInferenceLoop,  # noqa: E402
                                StochasticVariationalInference, create_Gaussian_meanfield)
from mxfusion.inference import TransferInference, ModulePredictionAlgorithm  # noqa: E402
from mxfusion.util.testutils import MockMXNetRandomGenerator  # noqa: E402

config.DEFAULT_DTYPE = 'float64'
DT = 'float64'
KERNELS = {'rbf': RBF, 'matern12': Matern12, 'matern32': Matern32, 'matern52': Matern52}


def nd(a):
    return mx.nd.array(a, dtype=DT)


def save(name, **arrays):
    np.savez(os.path.join(HERE, name + '.npz'), **{k: np.asarray(v) for k, v in arrays.items()})
    print('wrote', name, sorted(arrays.keys()))


def grads_of(infr, variables):
    out = {}
    for name, var in variables.items():
        p = infr.params.param_dict[var.uuid]
        out[name] = p.grad().asnumpy().copy()
    return out


# ------------------------------------------------------------------------------------------------ kernels
def golden_kernels():
    rng = np.random.RandomState(0)
    out = {}
    for kname, cls in KERNELS.items():
        for ard in (False, True):
            for S in (1, 3):
                D, N, N2 = 3, 7, 5
                X = rng.rand(S, N, D)
                X2 = rng.rand(S, N2, D)
                ls = rng.rand(S, D if ard else 1) + 0.5
                var = rng.rand(S, 1) + 0.5
                k = cls(input_dim=D, ARD=ard, dtype=DT)
                params = {k.name + '_lengthscale': nd(ls), k.name + '_variance': nd(var)}
                tag = '%s_ard%d_S%d' % (kname, int(ard), S)
                out[tag + '_X'], out[tag + '_X2'], out[tag + '_ls'], out[tag + '_var'] = X, X2, ls, var
                out[tag + '_K'] = k.K(mx.nd, nd(X), **params).asnumpy()
                out[tag + '_K2'] = k.K(mx.nd, nd(X), nd(X2), **params).asnumpy()
                out[tag + '_Kdiag'] = k.Kdiag(mx.nd, nd(X), **params).asnumpy()
    # active_dims (kernel_test.py active-dims cases): only dims [0, 2] of a 4-d input
    X = rng.rand(1, 6, 4)
    ls, var = rng.rand(1, 2) + 0.5, rng.rand(1, 1) + 0.5
    k = RBF(input_dim=2, ARD=True, active_dims=[0, 2], dtype=DT)
    out['active_X'], out['active_ls'], out['active_var'] = X, ls, var
    out['active_K'] = k.K(mx.nd, nd(X), **{'rbf_lengthscale': nd(ls), 'rbf_variance': nd(var)}).asnumpy()
    save('kernels', **out)


# ------------------------------------------------------------------------------------------------ SVGP fixture
def svgp_case(kname, P, seed, rv_scaling=None, N=10, M=3, Din=3, jitter=1e-8):
    np.random.seed(seed)
    X, Y, Z = np.random.rand(N, Din), np.random.rand(N, P), np.random.rand(M, Din)
    qU_mean, qU_cov_W, qU_cov_diag = np.random.rand(M, P), np.random.rand(M, M), np.random.rand(M,)
    noise_var, lengthscale, variance = np.random.rand(1), np.random.rand(Din), np.random.rand(1)
    m = Model()
    m.N = Variable()
    m.X = Variable(shape=(m.N, Din))
    m.Z = Variable(shape=(M, Din), initial_value=nd(Z))
    m.noise_var = Variable(transformation=PositiveTransformation(), initial_value=nd(noise_var))
    kernel = KERNELS[kname](input_dim=Din, ARD=True, variance=nd(variance), lengthscale=nd(lengthscale), dtype=DT)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, P), dtype=DT)
    gp = m.Y.factor
    gp.svgp_log_pdf.jitter = jitter
    loop = MinibatchInferenceLoop(batch_size=N, rv_scaling={m.Y: rv_scaling}) if rv_scaling else BatchInferenceLoop()
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, dtype=DT)
    infr.initialize(X=X.shape, Y=Y.shape)
    post = gp._extra_graphs[0]
    infr.params[post.qU_mean] = nd(qU_mean)
    infr.params[post.qU_cov_W] = nd(qU_cov_W)
    infr.params[post.qU_cov_diag] = nd(qU_cov_diag)
    executor = infr.create_executor()
    with mx.autograd.record():
        loss, loss_g = executor(mx.nd.zeros(1), nd(X), nd(Y))
        loss_g.backward()
    g = grads_of(infr, dict(Z=m.Z, noise_var=m.noise_var, qU_mean=post.qU_mean, qU_cov_W=post.qU_cov_W,
                            qU_cov_diag=post.qU_cov_diag, lengthscale=kernel.lengthscale, variance=kernel.variance))
    return dict(X=X, Y=Y, Z=Z, qU_mean=qU_mean, qU_cov_W=qU_cov_W, qU_cov_diag=qU_cov_diag, noise_var=noise_var,
                lengthscale=lengthscale, variance=variance, jitter=jitter, rv_scaling=rv_scaling or 1.0,
                loss=loss.asnumpy(), **{'grad_' + k: v for k, v in g.items()})


def golden_svgp():
    out = {}
    cases = [('rbf', 1, 0, None), ('matern52', 1, 1, None), ('rbf', 2, 2, None), ('rbf', 1, 3, 12.5),
             ('matern32', 2, 4, 3.0), ('matern12', 1, 5, None)]
    for i, (kname, P, seed, sc) in enumerate(cases):
        r = svgp_case(kname, P, seed, sc)
        out['case%d_kernel' % i] = kname
        for k, v in r.items():
            out['case%d_%s' % (i, k)] = v
    out['n_cases'] = len(cases)
    save('svgp_fixture', **out)


# ------------------------------------------------------------------------------------------------ exact GP fixture
def golden_gp():
    out = {}
    for i, (kname, P, seed) in enumerate([('rbf', 2, 0), ('matern52', 1, 1), ('matern32', 3, 2)]):
        np.random.seed(seed)
        N, Din = 10, 3
        X, Y = np.random.rand(N, Din), np.random.rand(N, P)
        noise_var, lengthscale, variance = np.random.rand(1), np.random.rand(Din), np.random.rand(1)
        m = Model()
        m.N = Variable()
        m.X = Variable(shape=(m.N, Din))
        m.noise_var = Variable(transformation=PositiveTransformation(), initial_value=nd(noise_var))
        kernel = KERNELS[kname](input_dim=Din, ARD=True, variance=nd(variance), lengthscale=nd(lengthscale), dtype=DT)
        m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P), dtype=DT)
        infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), dtype=DT)
        infr.initialize(X=X.shape, Y=Y.shape)
        executor = infr.create_executor()
        with mx.autograd.record():
            loss, loss_g = executor(mx.nd.zeros(1), nd(X), nd(Y))
            loss_g.backward()
        g = grads_of(infr, dict(noise_var=m.noise_var, lengthscale=kernel.lengthscale, variance=kernel.variance))
        post = m.Y.factor._extra_graphs[0]
        r = dict(X=X, Y=Y, noise_var=noise_var, lengthscale=lengthscale, variance=variance, loss=loss.asnumpy(),
                 L=infr.params[post.L].asnumpy(), LinvY=infr.params[post.LinvY].asnumpy(),
                 **{'grad_' + k: v for k, v in g.items()})
        out['case%d_kernel' % i] = kname
        for k, v in r.items():
            out['case%d_%s' % (i, k)] = v
    out['n_cases'] = 3
    save('gp_fixture', **out)


# ------------------------------------------------------------------------------------------------ GP notebook
def golden_gp_notebook():
    np.random.seed(0)
    X = np.random.uniform(-3., 3., (20, 1))
    Y = np.sin(X) + np.random.randn(20, 1) * 0.05
    m = Model()
    m.N = Variable()
    m.X = Variable(shape=(m.N, 1))
    m.noise_var = Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=1, variance=1, lengthscale=1)
    m.Y = GPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    loss0, _ = infr.create_executor()(mx.nd.zeros(1), nd(X), nd(Y))
    infr.run(X=nd(X), Y=nd(Y), max_iter=100, learning_rate=0.05, verbose=False)
    loss1, _ = infr.create_executor()(mx.nd.zeros(1), nd(X), nd(Y))
    save('gp_notebook', X=X, Y=Y, loss_init=loss0.asnumpy(), loss_final=loss1.asnumpy(),
         variance=infr.params[m.kernel.variance].asnumpy(), lengthscale=infr.params[m.kernel.lengthscale].asnumpy(),
         noise_var=infr.params[m.noise_var].asnumpy(), printed_loss=-16.903135093930537,
         printed_params=np.array([0.616992, 1.649073, 0.002251]))
    print('  stand-in vs the notebook\'s printed loss:', float(loss1.asnumpy()), 'vs -16.903135093930537')


# ------------------------------------------------------------------------------------------------ SVGP minibatch loop
def golden_svgp_minibatch():
    np.random.seed(0)
    N, B, M, epochs = 203, 20, 8, 3
    X = np.random.uniform(-3., 3., (N, 1))
    Y = np.sin(X) + np.random.randn(N, 1) * 0.05
    Z0 = np.linspace(-3, 3, M)[:, None]
    m = Model()
    m.N = Variable()
    m.X = Variable(shape=(m.N, 1))
    m.noise_var = Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=1, variance=1, lengthscale=1)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1), num_inducing=M)
    m.Y.factor.svgp_log_pdf.jitter = 1e-6
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / B})
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop)
    infr.initialize(X=(N, 1), Y=(N, 1))
    post = m.Y.factor._extra_graphs[0]
    infr.params[m.Y.factor.inducing_inputs] = nd(Z0)
    infr.params[post.qU_mean] = nd(np.zeros((M, 1)))
    infr.params[post.qU_cov_W] = nd(np.eye(M) * 0.1)
    infr.params[post.qU_cov_diag] = nd(np.ones(M) * 0.5)
    np.random.seed(123)                        # the sampler's shuffles come from the NumPy global generator
    infr.run(X=nd(X), Y=nd(Y), max_iter=epochs, learning_rate=0.05, verbose=False)
    loss_full, _ = infr.create_executor()(mx.nd.zeros(1), nd(X), nd(Y))
    save('svgp_minibatch', X=X, Y=Y, Z0=Z0, N=N, B=B, M=M, epochs=epochs, shuffle_seed=123, lr=0.05,
         final_Z=infr.params[m.Y.factor.inducing_inputs].asnumpy(), final_qU_mean=infr.params[post.qU_mean].asnumpy(),
         final_qU_cov_diag=infr.params[post.qU_cov_diag].asnumpy(),
         final_lengthscale=infr.params[m.kernel.lengthscale].asnumpy(),
         final_variance=infr.params[m.kernel.variance].asnumpy(), final_noise_var=infr.params[m.noise_var].asnumpy(),
         final_full_data_loss_at_batch_scaling=loss_full.asnumpy())


# ------------------------------------------------------------------------------------------------ prediction
def golden_predict():
    """svgpregression_test.py:171-242 / gpregression_test.py:169-226: mean / variance prediction in the four modes."""
    out = {}
    np.random.seed(0)
    N, M, Din, P = 10, 3, 3, 1
    X, Y, Z = np.random.rand(N, Din), np.random.rand(N, P), np.random.rand(M, Din)
    qU_mean, qU_cov_W, qU_cov_diag = np.random.rand(M, P), np.random.rand(M, M), np.random.rand(M,)
    noise_var, lengthscale, variance = np.random.rand(1), np.random.rand(Din), np.random.rand(1)
    Xt = np.random.rand(5, Din)
    out.update(X=X, Y=Y, Z=Z, qU_mean=qU_mean, qU_cov_W=qU_cov_W, qU_cov_diag=qU_cov_diag, noise_var=noise_var,
               lengthscale=lengthscale, variance=variance, Xt=Xt)
    for module in ('svgp', 'gp'):
        m = Model()
        m.N = Variable()
        m.X = Variable(shape=(m.N, Din))
        m.noise_var = Variable(transformation=PositiveTransformation(), initial_value=nd(noise_var))
        kernel = RBF(input_dim=Din, ARD=True, variance=nd(variance), lengthscale=nd(lengthscale), dtype=DT)
        if module == 'svgp':
            m.Z = Variable(shape=(M, Din), initial_value=nd(Z))
            m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                                 shape=(m.N, P), dtype=DT)
            m.Y.factor.svgp_log_pdf.jitter = 1e-8
        else:
            m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P), dtype=DT)
        gp = m.Y.factor
        infr = Inference(MAP(model=m, observed=[m.X, m.Y]), dtype=DT)
        infr.initialize(X=X.shape, Y=Y.shape)
        if module == 'svgp':
            post = gp._extra_graphs[0]
            infr.params[post.qU_mean] = nd(qU_mean)
            infr.params[post.qU_cov_W] = nd(qU_cov_W)
            infr.params[post.qU_cov_diag] = nd(qU_cov_diag)
        infr.run(X=nd(X), Y=nd(Y))
        alg = gp.svgp_predict if module == 'svgp' else gp.gp_predict
        for noise_free in (True, False):
            for diag in (True, False):
                alg.noise_free, alg.diagonal_variance = noise_free, diag
                infr2 = TransferInference(ModulePredictionAlgorithm(m, observed=[m.X], target_variables=[m.Y]),
                                          infr_params=infr.params, dtype=np.float64)
                res = infr2.run(X=nd(Xt))[0]
                tag = '%s_nf%d_diag%d' % (module, int(noise_free), int(diag))
                out[tag + '_mean'] = res[0].asnumpy()
                out[tag + '_var'] = res[1].asnumpy()
    save('predict', **out)


# ------------------------------------------------------------------------------------------------ Normal
def golden_normal():
    rng = np.random.RandomState(0)
    out = {}
    S, shape = 3, (5, 2)
    mean, var, rv = rng.randn(S, *shape), rng.rand(S, *shape) + 0.2, rng.randn(S, *shape)
    m = Normal.define_variable(shape=shape, dtype=DT).factor
    variables = {m.mean.uuid: nd(mean), m.variance.uuid: nd(var), m.random_variable.uuid: nd(rv)}
    out['mean'], out['var'], out['rv'] = mean, var, rv
    out['log_pdf'] = m.log_pdf(F=mx.nd, variables=variables).asnumpy()
    m.log_pdf_scaling = 2.5
    out['log_pdf_scaled_2p5'] = m.log_pdf(F=mx.nd, variables=variables).asnumpy()
    eps = rng.randn(S, *shape)
    m2 = Normal.define_variable(shape=shape, dtype=DT, rand_gen=MockMXNetRandomGenerator(nd(eps.flatten()))).factor
    variables = {m2.mean.uuid: nd(mean[:1]), m2.variance.uuid: nd(var[:1])}
    out['eps'] = eps
    out['draw'] = m2.draw_samples(F=mx.nd, variables=variables, num_samples=S).asnumpy()
    save('normal', **out)


# ------------------------------------------------------------------------------------------------ mean-field SVI
def golden_svi():
    rng = np.random.RandomState(1)
    N, S = 6, 4
    y = rng.randn(N, 1)
    eps_mu = rng.randn(S, 1)
    m = Model()
    m.mu = Normal.define_variable(mean=nd([0.]), variance=nd([4.]), shape=(1,), dtype=DT)
    m.s2 = Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=nd([0.7]))
    m.y = Normal.define_variable(mean=mxfusion.components.functions.operators.broadcast_to(m.mu, (N, 1)),
                                 variance=mxfusion.components.functions.operators.broadcast_to(m.s2, (N, 1)),
                                 shape=(N, 1), dtype=DT)
    q = create_Gaussian_meanfield(model=m, observed=[m.y], dtype=DT)
    q.mu.factor._rand_gen = MockMXNetRandomGenerator(nd(eps_mu.flatten()))
    alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=[m.y])
    infr = GradBasedInference(inference_algorithm=alg, dtype=DT)
    infr.initialize(y=y.shape)
    infr.params[q.mu.factor.mean] = nd([0.3])
    infr.params[q.mu.factor.variance] = nd([0.5])
    executor = infr.create_executor()
    with mx.autograd.record():
        loss, loss_g = executor(mx.nd.zeros(1), nd(y))
        loss_g.backward()
    g = grads_of(infr, dict(q_mean=q.mu.factor.mean, q_var=q.mu.factor.variance, s2=m.s2))
    save('svi_toy', y=y, eps=eps_mu, S=S, prior_mean=0., prior_var=4., s2=0.7, q_mean=0.3, q_var=0.5,
         loss=loss.asnumpy(), **{'grad_' + k: v for k, v in g.items()})


# ------------------------------------------------------------------------------------------------ sparse GP (Titsias)
def sparsegp_case(kname, P, seed, N=10, M=3, Din=3, jitter=1e-8, ard=True):
    np.random.seed(seed)
    X, Y, Z = np.random.rand(N, Din), np.random.rand(N, P), np.random.rand(M, Din)
    noise_var, lengthscale, variance = np.random.rand(1), np.random.rand(Din if ard else 1), np.random.rand(1)
    Xt = np.random.rand(5, Din)
    m = Model()
    m.N = Variable()
    m.X = Variable(shape=(m.N, Din))
    m.Z = Variable(shape=(M, Din), initial_value=nd(Z))
    m.noise_var = Variable(transformation=PositiveTransformation(), initial_value=nd(noise_var))
    kernel = KERNELS[kname](input_dim=Din, ARD=ard, variance=nd(variance), lengthscale=nd(lengthscale), dtype=DT)
    m.Y = SparseGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                             shape=(m.N, P), dtype=DT)
    gp = m.Y.factor
    gp.sgp_log_pdf.jitter = jitter
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), dtype=DT)
    infr.initialize(X=X.shape, Y=Y.shape)
    executor = infr.create_executor()
    with mx.autograd.record():
        loss, loss_g = executor(mx.nd.zeros(1), nd(X), nd(Y))
        loss_g.backward()
    g = grads_of(infr, dict(Z=m.Z, noise_var=m.noise_var, lengthscale=kernel.lengthscale, variance=kernel.variance))
    post = gp._extra_graphs[0]
    r = dict(X=X, Y=Y, Z=Z, noise_var=noise_var, lengthscale=lengthscale, variance=variance, jitter=jitter, Xt=Xt,
             loss=loss.asnumpy(), wv=infr.params[post.wv].asnumpy(), L=infr.params[post.L].asnumpy(),
             LA=infr.params[post.LA].asnumpy(), **{'grad_' + k: v for k, v in g.items()})
    for noise_free in (True, False):
        for diag in (True, False):
            gp.sgp_predict.noise_free, gp.sgp_predict.diagonal_variance = noise_free, diag
            infr2 = TransferInference(ModulePredictionAlgorithm(m, observed=[m.X], target_variables=[m.Y]),
                                      infr_params=infr.params, dtype=np.float64)
            res = infr2.run(X=nd(Xt))[0]
            tag = 'pred_nf%d_diag%d' % (int(noise_free), int(diag))
            r[tag + '_mean'] = res[0].asnumpy()
            r[tag + '_var'] = res[1].asnumpy()
    return r


def golden_sparsegp():
    out = {}
    cases = [('rbf', 2, 0, {}), ('matern52', 1, 1, {}), ('matern32', 3, 2, dict(N=17, M=5)),
             ('rbf', 1, 3, dict(N=40, M=8, Din=2, ard=False, jitter=1e-6)), ('matern12', 1, 4, {})]
    for i, (kname, P, seed, kw) in enumerate(cases):
        r = sparsegp_case(kname, P, seed, **kw)
        out['case%d_kernel' % i] = kname
        for k, v in r.items():
            out['case%d_%s' % (i, k)] = v
    out['n_cases'] = len(cases)
    save('sparsegp_fixture', **out)


# ------------------------------------------------------------------------------------------------ kernel algebra
def combo_kernel(spec, Din):
    """spec -> a reference kernel object.  Initial values are filled in by the caller through infr.params / K kwargs."""
    if spec == 'linear':
        return Linear(input_dim=Din, ARD=False, dtype=DT)
    if spec == 'linear_ard':
        return Linear(input_dim=Din, ARD=True, dtype=DT)
    if spec == 'bias':
        return Bias(input_dim=Din, dtype=DT)
    if spec == 'white':
        return White(input_dim=Din, dtype=DT)
    if spec == 'rbf+linear_ard':
        return RBF(input_dim=Din, ARD=True, dtype=DT) + Linear(input_dim=Din, ARD=True, dtype=DT)
    if spec == 'rbf*matern32':
        return RBF(input_dim=Din, ARD=False, dtype=DT) * Matern32(input_dim=Din, ARD=True, dtype=DT)
    if spec == 'rbf+rbf+bias':
        return RBF(input_dim=Din, dtype=DT) + RBF(input_dim=Din, ARD=True, dtype=DT) + Bias(input_dim=Din, dtype=DT)
    if spec == '(matern52+white)*linear':
        return (Matern52(input_dim=Din, dtype=DT) + White(input_dim=Din, dtype=DT)) * Linear(input_dim=Din, dtype=DT)
    raise ValueError(spec)


COMBO_SPECS = ['linear', 'linear_ard', 'bias', 'white', 'rbf+linear_ard', 'rbf*matern32', 'rbf+rbf+bias',
               '(matern52+white)*linear']


def golden_combo_kernels():
    rng = np.random.RandomState(7)
    out = {'specs': np.array(COMBO_SPECS)}
    Din, N, N2 = 3, 6, 4
    for i, spec in enumerate(COMBO_SPECS):
        for S in (1, 2):
            k = combo_kernel(spec, Din)
            X, X2 = rng.rand(S, N, Din), rng.rand(S, N2, Din)
            names = sorted(k.parameters.keys())
            vals = {n: rng.rand(S, *k.parameters[n].shape) + 0.3 for n in names}
            params = {n: nd(v) for n, v in vals.items()}
            tag = 'k%d_S%d' % (i, S)
            out[tag + '_X'], out[tag + '_X2'] = X, X2
            out[tag + '_names'] = np.array(names)
            for n in names:
                out[tag + '_p_' + n] = vals[n]
            out[tag + '_K'] = k.K(mx.nd, nd(X), **params).asnumpy()
            out[tag + '_K2'] = k.K(mx.nd, nd(X), nd(X2), **params).asnumpy()
            out[tag + '_Kdiag'] = k.Kdiag(mx.nd, nd(X), **params).asnumpy()
    save('combo_kernels', **out)


def combo_module_case(module, spec, seed, N=12, M=4, Din=3, P=2):
    np.random.seed(seed)
    X, Y, Z = np.random.rand(N, Din), np.random.rand(N, P), np.random.rand(M, Din)
    noise_var = np.random.rand(1) + 0.1
    m = Model()
    m.N = Variable()
    m.X = Variable(shape=(m.N, Din))
    m.noise_var = Variable(transformation=PositiveTransformation(), initial_value=nd(noise_var))
    kernel = combo_kernel(spec, Din)
    if module == 'gp':
        m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P), dtype=DT)
    else:
        m.Z = Variable(shape=(M, Din), initial_value=nd(Z))
        cls = SVGPRegression if module == 'svgp' else SparseGPRegression
        m.Y = cls.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z, shape=(m.N, P),
                                  dtype=DT)
        (m.Y.factor.svgp_log_pdf if module == 'svgp' else m.Y.factor.sgp_log_pdf).jitter = 1e-6
    gp = m.Y.factor
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), dtype=DT)
    infr.initialize(X=X.shape, Y=Y.shape)
    r = dict(X=X, Y=Y, Z=Z, noise_var=noise_var)
    names = sorted(kernel.parameters.keys())
    r['names'] = np.array(names)
    for n in names:
        v = np.random.rand(*kernel.parameters[n].shape) + 0.3
        infr.params[kernel.parameters[n]] = nd(v)
        r['p_' + n] = v
    if module == 'svgp':
        post = gp._extra_graphs[0]
        for nm in ('qU_mean', 'qU_cov_W', 'qU_cov_diag'):
            v = np.random.rand(*getattr(post, nm).shape)
            infr.params[getattr(post, nm)] = nd(v)
            r[nm] = v
    executor = infr.create_executor()
    with mx.autograd.record():
        loss, loss_g = executor(mx.nd.zeros(1), nd(X), nd(Y))
        loss_g.backward()
    gvars = {'noise_var': m.noise_var}
    if module != 'gp':
        gvars['Z'] = m.Z
    if module == 'svgp':
        gvars.update(qU_mean=post.qU_mean, qU_cov_W=post.qU_cov_W, qU_cov_diag=post.qU_cov_diag)
    for n in names:
        gvars['p_' + n] = kernel.parameters[n]
    g = grads_of(infr, gvars)
    r['loss'] = loss.asnumpy()
    r.update({'grad_' + k: v for k, v in g.items()})
    return r


def golden_combo_modules():
    out = {}
    cases = [('svgp', 'rbf+linear_ard', 0), ('svgp', 'rbf*matern32', 1), ('gp', '(matern52+white)*linear', 2),
             ('gp', 'rbf+rbf+bias', 3), ('sparsegp', 'rbf+linear_ard', 4), ('sparsegp', 'linear', 5)]
    for i, (module, spec, seed) in enumerate(cases):
        r = combo_module_case(module, spec, seed)
        out['case%d_module' % i], out['case%d_spec' % i] = module, spec
        for k, v in r.items():
            out['case%d_%s' % (i, k)] = v
    out['n_cases'] = len(cases)
    save('combo_modules', **out)


# ------------------------------------------------------------------------------------------------ GP distributions
def golden_gp_distributions():
    rng = np.random.RandomState(11)
    out = {}
    i = 0
    for kname in ('rbf', 'matern52'):
        for S in (1, 3):
            for P in (1, 3):
                N, Nc, Din, ns = 7, 5, 2, 4
                X, Xc = rng.rand(S, N, Din), rng.rand(S, Nc, Din)
                Y, Yc = rng.randn(S, N, P), rng.randn(S, Nc, P)
                ls, var = rng.rand(S, Din) + 0.5, rng.rand(S, 1) + 0.5
                die = rng.randn(ns, N, P)
                kern = KERNELS[kname](input_dim=Din, ARD=True, dtype=DT)
                kp = {kern.name + '_lengthscale': nd(ls), kern.name + '_variance': nd(var)}
                X_var, Xc_var, Yc_var = Variable(shape=(N, Din)), Variable(shape=(Nc, Din)), Variable(shape=(Nc, P))
                tag = 'c%d' % i
                out[tag + '_kernel'] = kname
                for k, v in dict(X=X, Xc=Xc, Y=Y, Yc=Yc, ls=ls, var=var, die=die).items():
                    out[tag + '_' + k] = v
                # prior GP
                gp = GaussianProcess.define_variable(X=X_var, kernel=kern, shape=(N, P), dtype=DT,
                                                     rand_gen=MockMXNetRandomGenerator(nd(die.flatten()))).factor
                variables = {gp.X.uuid: nd(X), gp.random_variable.uuid: nd(Y)}
                variables.update({getattr(gp, n).uuid: v for n, v in kp.items()})
                out[tag + '_gp_log_pdf'] = gp.log_pdf(F=mx.nd, variables=variables).asnumpy()
                variables1 = {gp.X.uuid: nd(X[:1])}
                variables1.update({getattr(gp, n).uuid: nd(v.asnumpy()[:1]) for n, v in kp.items()})
                out[tag + '_gp_draw'] = gp.draw_samples(F=mx.nd, variables=variables1, num_samples=ns).asnumpy()
                # conditional GP
                kern2 = KERNELS[kname](input_dim=Din, ARD=True, dtype=DT)
                cgp = ConditionalGaussianProcess.define_variable(
                    X=X_var, X_cond=Xc_var, Y_cond=Yc_var, kernel=kern2, shape=(N, P), dtype=DT,
                    rand_gen=MockMXNetRandomGenerator(nd(die.flatten()))).factor
                variables = {cgp.X.uuid: nd(X), cgp.X_cond.uuid: nd(Xc), cgp.Y_cond.uuid: nd(Yc),
                             cgp.random_variable.uuid: nd(Y)}
                variables.update({getattr(cgp, n).uuid: v for n, v in kp.items()})
                out[tag + '_cgp_log_pdf'] = cgp.log_pdf(F=mx.nd, variables=variables).asnumpy()
                variables1 = {cgp.X.uuid: nd(X[:1]), cgp.X_cond.uuid: nd(Xc[:1]), cgp.Y_cond.uuid: nd(Yc[:1])}
                variables1.update({getattr(cgp, n).uuid: nd(v.asnumpy()[:1]) for n, v in kp.items()})
                out[tag + '_cgp_draw'] = cgp.draw_samples(F=mx.nd, variables=variables1, num_samples=ns).asnumpy()
                i += 1
    out['n_cases'] = i
    save('gp_distributions', **out)


# ------------------------------------------------------------------------------------------------ heteroscedastic SVGP
def golden_svgp_hetero():
    """svgp_regression.py:61-67: noise_var with one value per data point, shape (N, 1) and (N, P)."""
    out = {}
    for i, (P, cols, seed) in enumerate([(1, 1, 0), (2, 1, 1), (2, 2, 2)]):
        np.random.seed(seed)
        N, M, Din = 11, 4, 3
        X, Y, Z = np.random.rand(N, Din), np.random.rand(N, P), np.random.rand(M, Din)
        qU_mean, qU_cov_W, qU_cov_diag = np.random.rand(M, P), np.random.rand(M, M), np.random.rand(M,)
        noise_var = np.random.rand(N, cols) + 0.1
        lengthscale, variance = np.random.rand(Din) + 0.3, np.random.rand(1) + 0.3
        m = Model()
        m.N = Variable()
        m.X = Variable(shape=(m.N, Din))
        m.Z = Variable(shape=(M, Din), initial_value=nd(Z))
        m.noise_var = Variable(shape=(N, cols), transformation=PositiveTransformation(), initial_value=nd(noise_var))
        kernel = RBF(input_dim=Din, ARD=True, variance=nd(variance), lengthscale=nd(lengthscale), dtype=DT)
        m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                             shape=(m.N, P), dtype=DT)
        gp = m.Y.factor
        gp.svgp_log_pdf.jitter = 1e-8
        infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), dtype=DT)
        infr.initialize(X=X.shape, Y=Y.shape)
        post = gp._extra_graphs[0]
        infr.params[post.qU_mean] = nd(qU_mean)
        infr.params[post.qU_cov_W] = nd(qU_cov_W)
        infr.params[post.qU_cov_diag] = nd(qU_cov_diag)
        executor = infr.create_executor()
        with mx.autograd.record():
            loss, loss_g = executor(mx.nd.zeros(1), nd(X), nd(Y))
            loss_g.backward()
        g = grads_of(infr, dict(Z=m.Z, noise_var=m.noise_var, qU_mean=post.qU_mean, qU_cov_W=post.qU_cov_W,
                                qU_cov_diag=post.qU_cov_diag, lengthscale=kernel.lengthscale, variance=kernel.variance))
        r = dict(X=X, Y=Y, Z=Z, qU_mean=qU_mean, qU_cov_W=qU_cov_W, qU_cov_diag=qU_cov_diag, noise_var=noise_var,
                 lengthscale=lengthscale, variance=variance, loss=loss.asnumpy(), **{'grad_' + k: v for k, v in g.items()})
        for k, v in r.items():
            out['case%d_%s' % (i, k)] = v
    out['n_cases'] = 3
    save('svgp_hetero', **out)


# ------------------------------------------------------------------------------------------------ sampling prediction
def golden_sampling_prediction():
    """*SamplingPrediction of the three modules (gpregression_test.py:255-307 pattern) with injected standard normals:
    diagonal and full-covariance draws, noisy and noise-free."""
    from mxfusion.modules.gp_modules.gp_regression import GPRegressionSamplingPrediction
    from mxfusion.modules.gp_modules.svgp_regression import SVGPRegressionSamplingPrediction
    from mxfusion.modules.gp_modules.sparsegp_regression import SparseGPRegressionSamplingPrediction
    out = {}
    np.random.seed(3)
    N, M, Din, P, Nt, ns = 10, 3, 3, 2, 6, 4
    X, Y, Z = np.random.rand(N, Din), np.random.rand(N, P), np.random.rand(M, Din)
    qU_mean, qU_cov_W, qU_cov_diag = np.random.rand(M, P), np.random.rand(M, M), np.random.rand(M,)
    noise_var, lengthscale, variance = np.random.rand(1) + 0.1, np.random.rand(Din) + 0.3, np.random.rand(1) + 0.3
    Xt = np.random.rand(Nt, Din)
    die = np.random.randn(ns, Nt, P)
    out.update(X=X, Y=Y, Z=Z, qU_mean=qU_mean, qU_cov_W=qU_cov_W, qU_cov_diag=qU_cov_diag, noise_var=noise_var,
               lengthscale=lengthscale, variance=variance, Xt=Xt, die=die)
    for module in ('gp', 'svgp', 'sparsegp'):
        m = Model()
        m.N = Variable()
        m.X = Variable(shape=(m.N, Din))
        m.noise_var = Variable(transformation=PositiveTransformation(), initial_value=nd(noise_var))
        kernel = RBF(input_dim=Din, ARD=True, variance=nd(variance), lengthscale=nd(lengthscale), dtype=DT)
        if module == 'gp':
            m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P), dtype=DT)
        else:
            m.Z = Variable(shape=(M, Din), initial_value=nd(Z))
            cls = SVGPRegression if module == 'svgp' else SparseGPRegression
            m.Y = cls.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                      shape=(m.N, P), dtype=DT)
            (m.Y.factor.svgp_log_pdf if module == 'svgp' else m.Y.factor.sgp_log_pdf).jitter = 1e-8
        gp = m.Y.factor
        infr = Inference(MAP(model=m, observed=[m.X, m.Y]), dtype=DT)
        infr.initialize(X=X.shape, Y=Y.shape)
        if module == 'svgp':
            post = gp._extra_graphs[0]
            infr.params[post.qU_mean] = nd(qU_mean)
            infr.params[post.qU_cov_W] = nd(qU_cov_W)
            infr.params[post.qU_cov_diag] = nd(qU_cov_diag)
        infr.run(X=nd(X), Y=nd(Y))
        cls = {'gp': GPRegressionSamplingPrediction, 'svgp': SVGPRegressionSamplingPrediction,
               'sparsegp': SparseGPRegressionSamplingPrediction}[module]
        name = {'gp': 'gp_predict', 'svgp': 'svgp_predict', 'sparsegp': 'sgp_predict'}[module]
        for noise_free in (True, False):
            for diag in (True, False):
                alg = cls(gp._module_graph, gp._extra_graphs[0], [gp._module_graph.X],
                          rand_gen=MockMXNetRandomGenerator(nd(die.flatten())))
                alg.noise_free, alg.diagonal_variance, alg.jitter = noise_free, diag, 1e-6
                gp.attach_prediction_algorithms(targets=gp.output_names, conditionals=gp.input_names, algorithm=alg,
                                                alg_name=name)
                infr2 = TransferInference(ModulePredictionAlgorithm(m, observed=[m.X], target_variables=[m.Y],
                                                                    num_samples=ns),
                                          infr_params=infr.params, dtype=np.float64)
                out['%s_nf%d_diag%d' % (module, int(noise_free), int(diag))] = infr2.run(X=nd(Xt))[0].asnumpy()
    save('sampling_prediction', **out)


# ------------------------------------------------------------------------------------------------ posterior forward sampling
def golden_vpfs():
    """VariationalPosteriorForwardSampling (forward_sampling.py:99-157) on the conjugate toy model of `svi`: the latent
    mean is drawn from q (injected noise), the observation from the model's likelihood given that draw (injected noise)."""
    from mxfusion.inference import VariationalPosteriorForwardSampling
    rng = np.random.RandomState(5)
    N, S = 6, 4
    y = rng.randn(N, 1)
    eps_mu, eps_y = rng.randn(S, 1), rng.randn(S, N, 1)
    m = Model()
    m.mu = Normal.define_variable(mean=nd([0.]), variance=nd([4.]), shape=(1,), dtype=DT)
    m.s2 = Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=nd([0.7]))
    m.y = Normal.define_variable(mean=mxfusion.components.functions.operators.broadcast_to(m.mu, (N, 1)),
                                 variance=mxfusion.components.functions.operators.broadcast_to(m.s2, (N, 1)),
                                 shape=(N, 1), dtype=DT, rand_gen=MockMXNetRandomGenerator(nd(eps_y.flatten())))
    q = create_Gaussian_meanfield(model=m, observed=[m.y], dtype=DT)
    q.mu.factor._rand_gen = MockMXNetRandomGenerator(nd(eps_mu.flatten()))
    alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=[m.y])
    infr = GradBasedInference(inference_algorithm=alg, dtype=DT)
    infr.initialize(y=y.shape)
    infr.params[q.mu.factor.mean] = nd([0.3])
    infr.params[q.mu.factor.variance] = nd([0.5])
    infr2 = VariationalPosteriorForwardSampling(S, [], infr, [m.y, m.mu], dtype=np.float64)
    res = infr2.run()
    save('vpfs_toy', y=y, eps_mu=eps_mu, eps_y=eps_y, S=S, q_mean=0.3, q_var=0.5, s2=0.7,
         sample_y=res[0].asnumpy(), sample_mu=res[1].asnumpy())


if __name__ == '__main__':
    golden_kernels()
    golden_svgp()
    golden_gp()
    golden_gp_notebook()
    golden_svgp_minibatch()
    golden_normal()
    golden_svi()
    golden_predict()
    golden_sparsegp()
    golden_svgp_hetero()
    golden_sampling_prediction()
    golden_vpfs()
    golden_gp_distributions()
    golden_combo_kernels()
    golden_combo_modules()
