"""Host-side mirror of the reference API (Model / Variable / Module / Inference.run) exercised end to
end on the CPU with the CUDA binding replaced by tests/raw_standin.py.  Anchors are the reference's own
printed numbers (examples/notebooks/gp_regression.ipynb) and the known answers on its test fixtures
(BASELINE.md section 2).  The same scenarios run on the real kernels in tests/test_gpu_api.py."""
import numpy as np
import pytest
import torch


@pytest.fixture()
def mf(monkeypatch):
    import mxfusion_b200 as mf
    from mxfusion_b200 import ops
    from tests import raw_standin
    monkeypatch.setattr(ops, 'R', raw_standin)
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float64')
    monkeypatch.setattr(mf.config, 'MXNET_DEFAULT_DEVICE', 'cpu')
    return mf


def gp_notebook_model(mf):
    """examples/notebooks/gp_regression.ipynb cells 4-10."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import GPRegression
    np.random.seed(0)
    X = np.random.uniform(-3., 3., (20, 1))
    Y = np.sin(X) + np.random.randn(20, 1) * 0.05
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 1))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=1, variance=1, lengthscale=1)
    m.Y = GPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1))
    return m, X, Y


def test_gp_notebook_initial_loss_and_training(mf):
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    loss, _ = infr.create_executor()(None, torch.tensor(X), torch.tensor(Y))
    assert abs(float(loss) - (-8.321443970764)) < 1e-8            # BASELINE.md known answer
    infr.run(X=X, Y=Y, max_iter=100, learning_rate=0.05)
    loss, _ = infr.create_executor()(None, torch.tensor(X), torch.tensor(Y))
    # notebook cell 12 prints -16.903135093930537 after the same 100 Adam steps
    assert abs(float(loss) - (-16.903135093930537)) < 2e-3
    got = [float(infr.params[v]) for v in (m.kernel.variance, m.kernel.lengthscale, m.noise_var)]
    np.testing.assert_allclose(got, [0.616992, 1.649073, 0.002251], rtol=2e-2)   # notebook cell 14


def test_svgp_module_matches_reference_fixture(mf):
    """testing/modules/svgpregression_test.py:41-115 (GPy assert replaced by the pinned known answer)."""
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import Inference, MAP
    np.random.seed(0)
    D, X, Y, Z = 1, np.random.rand(10, 3), np.random.rand(10, 1), np.random.rand(3, 3)
    qU_mean, qU_cov_W, qU_cov_diag = np.random.rand(3, 1), np.random.rand(3, 3), np.random.rand(3,)
    noise_var, lengthscale, variance = np.random.rand(1), np.random.rand(3), np.random.rand(1)
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 3))
    m.Z = mf.Variable(shape=(3, 3), initial_value=Z)
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=noise_var)
    kernel = RBF(input_dim=3, ARD=True, variance=variance, lengthscale=lengthscale)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, D))
    gp = m.Y.factor
    gp.svgp_log_pdf.jitter = 1e-8
    infr = Inference(MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    infr.params[gp._extra_graphs[0].qU_mean] = qU_mean
    infr.params[gp._extra_graphs[0].qU_cov_W] = qU_cov_W
    infr.params[gp._extra_graphs[0].qU_cov_diag] = qU_cov_diag
    loss, _ = infr.run(X=X, Y=Y)
    assert abs(-float(loss) - (-32.72563540745786)) < 1e-9
    np.testing.assert_allclose(infr.params[m.noise_var].numpy(), noise_var, rtol=1e-12)
    np.testing.assert_allclose(infr.params[kernel.lengthscale].numpy(), lengthscale, rtol=1e-12)


def test_svgp_minibatch_training_decreases_loss_and_uses_rollover(mf):
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    np.random.seed(0)
    N, B, M = 203, 20, 8
    X = np.random.uniform(-3., 3., (N, 1))
    Y = np.sin(X) + np.random.randn(N, 1) * 0.05
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 1))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=1, variance=1, lengthscale=1)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1),
                                         num_inducing=M)
    m.Y.factor.svgp_log_pdf.jitter = 1e-6
    for resident in (True, False):
        np.random.seed(1)
        loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / B}, data_resident=resident)
        infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop)
        infr.initialize(X=(N, 1), Y=(N, 1))
        infr.params[m.Y.factor.inducing_inputs] = np.linspace(-3, 3, M)[:, None]
        infr.params[m.Y.factor._extra_graphs[0].qU_cov_W] = np.eye(M) * 0.1
        losses = infr.run(X=X, Y=Y, max_iter=4, learning_rate=0.05)
        losses = [float(l) for l in losses]
        assert losses[-1] < losses[0]
        assert int(infr.params.adam_t.item()) == (4 * N) // B       # rollover keeps the remainders: 40 steps
        if resident:
            ref = losses
        else:
            np.testing.assert_allclose(losses, ref, rtol=1e-9)       # both data paths see the same batches


def test_rollover_sampler_is_bit_exact_with_oracle():
    from mxfusion_b200.inference import RolloverBatchSampler
    from oracle.loop import RolloverBatchSampler as Oracle
    a = RolloverBatchSampler(103, 10, rng=np.random.RandomState(5))
    b = Oracle(103, 10, np.random.RandomState(5))
    for _ in range(5):
        idx, nfull = a.epoch_indices()
        want = list(b.epoch())
        assert nfull == len(want)
        np.testing.assert_array_equal(idx.reshape(nfull, 10), np.stack(want))


def test_meanfield_svi_bnn_like_model_trains(mf):
    """Mean-field MC-ELBO (BASELINE config 4 shape, testing/inference/meanfield_test.py:62-105 pattern)."""
    from mxfusion_b200.components.distributions import Normal
    from mxfusion_b200.components.functions import MXFusionGluonFunction
    from mxfusion_b200.inference import (GradBasedInference, StochasticVariationalInference,
                                         create_Gaussian_meanfield, BatchInferenceLoop)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(1, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1)).double()
    np.random.seed(0)
    x = np.random.rand(50, 1) * 2 - 1
    y = np.sin(3 * x) + 0.05 * np.random.randn(50, 1)
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
    alg = StochasticVariationalInference(num_samples=3, model=m, posterior=q, observed=observed)
    infr = GradBasedInference(inference_algorithm=alg, grad_loop=BatchInferenceLoop())
    infr.initialize(y=y.shape, x=x.shape)
    for v_name, v in m.r.factor.parameters.items():
        infr.params[q[v].factor.mean] = v.initial_value
        infr.params[q[v].factor.variance] = torch.full(v.shape, 1e-6).double()
    l0, _ = infr.create_executor()(None, torch.tensor(y), torch.tensor(x))
    infr.run(max_iter=150, learning_rate=1e-2, y=y, x=x)
    l1, _ = infr.create_executor()(None, torch.tensor(y), torch.tensor(x))
    assert float(l1) < float(l0)


def test_forward_sampling_draws_from_a_gp_prior(mf):
    """ForwardSamplingAlgorithm over a model whose output has a GaussianProcess prior (forward_sampling.py:24-56,
    gp.py:123-153): with injected standard normals the draw is chol(K) @ die."""
    from mxfusion_b200.components.distributions import GaussianProcess, MockMXNetRandomGenerator
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.inference import Inference, ForwardSamplingAlgorithm
    from oracle import kernels as ok
    rng = np.random.RandomState(3)
    N, ns = 9, 5
    X = rng.rand(N, 2)
    die = rng.randn(ns, N, 1)
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 2))
    kern = RBF(input_dim=2, variance=1.3, lengthscale=0.7)
    m.Y = GaussianProcess.define_variable(X=m.X, kernel=kern, shape=(m.N, 1),
                                          rand_gen=MockMXNetRandomGenerator(torch.tensor(die.flatten())))
    infr = Inference(ForwardSamplingAlgorithm(m, observed=[m.X], num_samples=ns, target_variables=[m.Y]))
    infr.initialize(X=X.shape)
    with torch.no_grad():
        got = infr.run(X=X)[0].numpy()
    K = ok.K(0, X[None], np.array([[0.7]]), np.array([[1.3]]))[0]
    want = np.linalg.cholesky(K) @ die
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)


def test_inference_save_load_round_trip(mf, tmp_path):
    """Inference.save / load (inference.py:179-310): one zip with the reference's six members; a freshly built inference
    of the same topology (different UUIDs) gets the trained values back and reproduces the loss."""
    import zipfile
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    infr.run(X=X, Y=Y, max_iter=15, learning_rate=0.05)
    loss, _ = infr.create_executor()(None, torch.tensor(X), torch.tensor(Y))
    path = str(tmp_path / 'inference.zip')
    infr.save(path)
    with zipfile.ZipFile(path) as zf:
        names = set(zf.namelist())
    assert len(names) == 6 and 'version.json' in names          # serialization.py:26-135 archive layout
    m2, _, _ = gp_notebook_model(mf)
    infr2 = GradBasedInference(inference_algorithm=MAP(model=m2, observed=[m2.X, m2.Y]))
    infr2.initialize(X=X.shape, Y=Y.shape)
    before, _ = infr2.create_executor()(None, torch.tensor(X), torch.tensor(Y))
    assert abs(float(before) - float(loss)) > 1e-3
    infr2.load(path)
    after, _ = infr2.create_executor()(None, torch.tensor(X), torch.tensor(Y))
    np.testing.assert_allclose(float(after), float(loss), rtol=1e-12)
    for a, b in ((m.kernel.variance, m2.kernel.variance), (m.kernel.lengthscale, m2.kernel.lengthscale),
                 (m.noise_var, m2.noise_var)):
        np.testing.assert_allclose(infr2.params[b].numpy(), infr.params[a].numpy(), rtol=1e-12)


def test_load_is_applied_once_and_training_continues(mf, tmp_path):
    """A checkpoint is applied once (at load, or at the end of initialize when loaded before it): later run() calls
    continue from the trained state instead of being reset to the archive."""
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.run(X=X, Y=Y, max_iter=10, learning_rate=0.05)
    path = str(tmp_path / 'ckpt.zip')
    infr.save(path)
    saved_noise = float(infr.params[m.noise_var])
    # load BEFORE initialize: picked up by initialize()
    m2, _, _ = gp_notebook_model(mf)
    infr2 = GradBasedInference(inference_algorithm=MAP(model=m2, observed=[m2.X, m2.Y]))
    infr2.load(path)
    infr2.initialize(X=X.shape, Y=Y.shape)
    np.testing.assert_allclose(float(infr2.params[m2.noise_var]), saved_noise, rtol=1e-12)
    infr2.run(X=X, Y=Y, max_iter=30, learning_rate=0.05)
    trained = float(infr2.params[m2.noise_var])
    assert abs(trained - saved_noise) > 1e-6
    infr2.run(X=X, Y=Y, max_iter=1, learning_rate=1e-9)
    np.testing.assert_allclose(float(infr2.params[m2.noise_var]), trained, rtol=1e-6)   # not reset to the checkpoint
    # the plain Inference.run path sees a checkpoint loaded before initialize as well
    from mxfusion_b200.inference import Inference
    m3, _, _ = gp_notebook_model(mf)
    infr3 = Inference(MAP(model=m3, observed=[m3.X, m3.Y]))
    infr3.load(path)
    infr3.run(X=X, Y=Y)
    np.testing.assert_allclose(float(infr3.params[m3.noise_var]), saved_noise, rtol=1e-12)


def test_load_matches_by_name_and_rejects_mismatches(mf, tmp_path):
    from mxfusion_b200.common.exceptions import SerializationError
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import GPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.run(X=X, Y=Y, max_iter=5, learning_rate=0.05)
    path = str(tmp_path / 'ckpt.zip')
    infr.save(path)

    def rebuilt(noise_first, in_dim=1):
        mm = mf.Model()
        mm.N = mf.Variable()
        if noise_first:          # same model written down in another order
            mm.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
            mm.X = mf.Variable(shape=(mm.N, in_dim))
        else:
            mm.X = mf.Variable(shape=(mm.N, in_dim))
            mm.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
        mm.kernel = RBF(input_dim=in_dim, variance=1, lengthscale=1)
        mm.Y = GPRegression.define_variable(X=mm.X, kernel=mm.kernel, noise_var=mm.noise_var, shape=(mm.N, 1))
        return mm
    m2 = rebuilt(True)
    infr2 = GradBasedInference(inference_algorithm=MAP(model=m2, observed=[m2.X, m2.Y]))
    infr2.initialize(X=X.shape, Y=Y.shape)
    infr2.load(path)
    for a, b in ((m.kernel.variance, m2.kernel.variance), (m.kernel.lengthscale, m2.kernel.lengthscale),
                 (m.noise_var, m2.noise_var)):
        np.testing.assert_allclose(infr2.params[b].numpy(), infr.params[a].numpy(), rtol=1e-12)
    # a parameter whose shape changed is an error, not a skip
    m3 = rebuilt(False, in_dim=1)
    m3.extra = mf.Variable(shape=(2,), initial_value=np.zeros(2))
    infr3 = GradBasedInference(inference_algorithm=MAP(model=m3, observed=[m3.X, m3.Y]))
    infr3.initialize(X=X.shape, Y=Y.shape)
    with pytest.raises(SerializationError):
        infr3.load(path)
    X5 = np.random.rand(5, 1)
    m4 = rebuilt(False)
    infr4 = GradBasedInference(inference_algorithm=MAP(model=m4, observed=[m4.X, m4.Y]))
    infr4.initialize(X=X5.shape, Y=(5, 1))           # cached posterior quantities now have N = 5, the archive N = 20
    with pytest.raises(SerializationError):
        infr4.load(path)


def test_gp_modules_draw_samples_by_default(mf):
    """The three GP modules attach a draw-samples algorithm by default (gp_regression.py:373-378,
    svgp_regression.py:398-403, sparsegp_regression.py:372-377), so ForwardSampling works on models that contain them.
    With injected noise the draws are the reference's: chol(K + noise I) eps for the exact GP; U -> F | U -> Y for the
    inducing-point modules."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.components.distributions.random_gen import MockMXNetRandomGenerator
    from mxfusion_b200.modules.gp_modules import GPRegression, SVGPRegression, SparseGPRegression
    from mxfusion_b200.inference import Inference
    from mxfusion_b200.inference.forward_sampling import ForwardSamplingAlgorithm
    from oracle import kernels as ok
    rng = np.random.RandomState(3)
    N, M, ns = 7, 4, 3
    X, Z = rng.rand(N, 2), rng.rand(M, 2)
    die = rng.randn(ns * N)

    def model(kind):
        m = mf.Model()
        m.N = mf.Variable()
        m.X = mf.Variable(shape=(m.N, 2))
        m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.3)
        kern = RBF(input_dim=2, variance=1.3, lengthscale=0.7)
        gen = MockMXNetRandomGenerator(torch.tensor(die))
        if kind == 'gp':
            m.Y = GPRegression.define_variable(X=m.X, kernel=kern, noise_var=m.noise_var, shape=(m.N, 1), rand_gen=gen)
        else:
            cls = SVGPRegression if kind == 'svgp' else SparseGPRegression
            m.Z = mf.Variable(shape=(M, 2), initial_value=Z)
            m.Y = cls.define_variable(X=m.X, kernel=kern, noise_var=m.noise_var, inducing_inputs=m.Z, shape=(m.N, 1))
        return m
    m = model('gp')
    infr = Inference(ForwardSamplingAlgorithm(m, observed=[m.X], num_samples=ns, target_variables=[m.Y]))
    infr.initialize(X=X.shape)
    with torch.no_grad():
        got = infr.run(X=X)[0].numpy()
    K = ok.K(0, X[None], np.array([[0.7]]), np.array([[1.3]]))[0] + 0.3 * np.eye(N)
    want = np.linalg.cholesky(K) @ die.reshape(ns, N, 1)
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)
    for kind in ('svgp', 'sgp'):
        m = model(kind)
        infr = Inference(ForwardSamplingAlgorithm(m, observed=[m.X], num_samples=ns, target_variables=[m.Y]))
        infr.initialize(X=X.shape)
        with torch.no_grad():
            y = infr.run(X=X)[0]
        assert tuple(y.shape) == (ns, N, 1) and bool(torch.isfinite(y).all())


def test_executor_sees_shape_constants_updated_after_its_creation(mf):
    """The minibatch loop calls update_shape_constants(batch) AFTER the executor exists (grad_based_inference.py:80-86 of
    the reference): the executor must read the live constants (m.N = batch size), not a snapshot taken at creation."""
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    ex = infr.inference_algorithm.create_executor(data_def=infr.observed_variable_UUIDs, params=infr.params,
                                                  var_ties=infr.params.var_ties)
    assert ex._constants[m.N.uuid] == 20
    infr.params.update_constants({m.N: 7})
    assert ex._constants[m.N.uuid] == 7
    t = torch.arange(3.)
    infr.params.update_constants({'some_array': t})
    assert ex._constants['some_array'].dtype == torch.float64 and ex._constants['some_array'] is ex._constants['some_array']


def test_sgd_optimizer_takes_plain_gradient_steps(mf):
    """run(optimizer='sgd') (grad_based_inference.py:67 hands any optimiser name to gluon.Trainer): one step is
    w -= lr * grad on the unconstrained parameters; other names raise InferenceError."""
    from mxfusion_b200.inference import GradBasedInference, MAP
    from mxfusion_b200.common.exceptions import InferenceError
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    def hyper(i, mm):           # unconstrained values of the three trained hyper-parameters
        return np.concatenate([i.params._params[v.uuid].tensor.detach().numpy().ravel()
                               for v in (mm.noise_var, mm.kernel.lengthscale, mm.kernel.variance)])
    w0 = hyper(infr, m)
    infr.run(X=X, Y=Y, optimizer='sgd', learning_rate=1e-4, max_iter=1)
    step = (w0 - hyper(infr, m)) / 1e-4                       # = the gradient of the first step
    assert np.abs(step).min() > 0
    m2, _, _ = gp_notebook_model(mf)
    infr2 = GradBasedInference(inference_algorithm=MAP(model=m2, observed=[m2.X, m2.Y]))
    infr2.initialize(X=X.shape, Y=Y.shape)
    np.testing.assert_array_equal(hyper(infr2, m2), w0)
    infr2.run(X=X, Y=Y, optimizer='sgd', learning_rate=2e-4, max_iter=1)
    step2 = (w0 - hyper(infr2, m2)) / 2e-4
    np.testing.assert_allclose(step, step2, rtol=1e-6)        # same gradient, twice the step
    with pytest.raises(InferenceError, match='sgd'):
        infr2.run(X=X, Y=Y, optimizer='rmsprop', learning_rate=1e-3, max_iter=1)


def test_pack_grads_can_leave_one_segment_alone(mf):
    """The data-parallel step exchanges the gradient of the last bucket segment early (inference/_stepper.py): the final
    gradient gather must then skip that segment and still write every other one."""
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    ex = infr.inference_algorithm.create_executor(data_def=infr.observed_variable_UUIDs, params=infr.params,
                                                  var_ties=infr.params.var_ties)
    p = infr.params
    p.refresh_leaves()
    p.setup_fused_transforms(ex._var_trans)
    ex.pretransformed = p._fused_uuids
    p.transform_all_()
    p.clear_leaf_grads()
    loss, lg = ex(None, torch.tensor(X), torch.tensor(Y))
    lg.backward()
    full = torch.full_like(p.gflat, 7.0)
    p.pack_grads_(out=full)
    segs = [s for s in p._segments if (s[1].tleaf if s[4] == 1 else s[1].tensor).grad is not None]
    assert len(segs) >= 2
    skip = segs[-1]
    part = torch.full_like(p.gflat, 7.0)
    p.pack_grads_(out=part, skip_offset=skip[2])
    assert bool((part[skip[2]:skip[2] + skip[3]] == 7.0).all())
    keep = torch.ones_like(part, dtype=torch.bool)
    keep[skip[2]:skip[2] + skip[3]] = False
    np.testing.assert_array_equal(part[keep].numpy(), full[keep].numpy())
