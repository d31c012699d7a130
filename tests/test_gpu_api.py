"""The public API on the real kernels: the reference's notebook numbers and fixture known answers, CUDA-graph
replay vs eager launches, resident vs host-streamed minibatches, and a full-size step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def mf64(monkeypatch):
    import mxfusion_b200 as mf
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float64')
    return mf


def test_gp_notebook_on_gpu(cuda, mf64):
    """examples/notebooks/gp_regression.ipynb: loss at init -8.321443970764 (known answer), loss after 100 Adam
    steps -16.903135093930537 (cell 12), learned parameters 0.616992 / 1.649073 / 0.002251 (cell 14)."""
    from tests.test_host_api import gp_notebook_model
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf64)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    Xt, Yt = torch.tensor(X, device=cuda), torch.tensor(Y, device=cuda)
    loss, _ = infr.create_executor()(None, Xt, Yt)
    assert abs(float(loss) - (-8.321443970764)) < 1e-8
    infr.run(X=X, Y=Y, max_iter=100, learning_rate=0.05)
    loss, _ = infr.create_executor()(None, Xt, Yt)
    assert abs(float(loss) - (-16.903135093930537)) < 2e-3
    got = [float(infr.params[v]) for v in (m.kernel.variance, m.kernel.lengthscale, m.noise_var)]
    np.testing.assert_allclose(got, [0.616992, 1.649073, 0.002251], rtol=2e-2)


def _svgp_model(mf, N, Din, M, B):
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=Din, variance=1, lengthscale=1)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1),
                                         num_inducing=M)
    m.Y.factor.svgp_log_pdf.jitter = 1e-6
    return m


@pytest.mark.parametrize('use_graph,resident', [(True, True), (False, True), (True, False)])
def test_svgp_minibatch_paths_agree(cuda, mf64, use_graph, resident):
    """CUDA-graph replay, eager launches and host-streamed batches must give the same trajectory."""
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    np.random.seed(0)
    N, B, M = 1003, 100, 16
    X = np.random.uniform(-3., 3., (N, 1))
    Y = np.sin(X) + np.random.randn(N, 1) * 0.05
    out = []
    for ug, res in ((False, True), (use_graph, resident)):
        np.random.seed(5)
        m = _svgp_model(mf64, N, 1, M, B)
        loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / B}, data_resident=res, use_cuda_graph=ug,
                                      rng=np.random.RandomState(7))
        infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop)
        infr.initialize(X=(N, 1), Y=(N, 1))
        infr.params[m.Y.factor.inducing_inputs] = np.linspace(-3, 3, M)[:, None]
        infr.params[m.Y.factor._extra_graphs[0].qU_cov_W] = np.eye(M) * 0.1
        losses = [float(l) for l in infr.run(X=X, Y=Y, max_iter=3, learning_rate=0.05)]
        out.append((losses, infr.params[m.kernel.lengthscale].cpu().numpy()))
    np.testing.assert_allclose(out[1][0], out[0][0], rtol=1e-8)
    np.testing.assert_allclose(out[1][1], out[0][1], rtol=1e-8)
    assert out[0][0][-1] < out[0][0][0]


def test_svgp_fixture_through_api(cuda, mf64):
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import Inference, MAP
    mf = mf64
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
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, 1))
    gp = m.Y.factor
    gp.svgp_log_pdf.jitter = 1e-8
    infr = Inference(MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    infr.params[gp._extra_graphs[0].qU_mean] = qU_mean
    infr.params[gp._extra_graphs[0].qU_cov_W] = qU_cov_W
    infr.params[gp._extra_graphs[0].qU_cov_diag] = qU_cov_diag
    loss, _ = infr.run(X=X, Y=Y)
    assert abs(-float(loss) - (-32.72563540745786)) < 1e-9


def test_full_size_step_is_finite_and_decreases(cuda):
    """BASELINE headline shapes (M=1024, D=8, B=4096, f32) on a 64k-row slice: size-independent properties."""
    import bench
    X, Y, Z = bench.synthetic(n=65536)
    infr, loop = bench.build_inference(X, Y, Z, 65536, 1, data_resident=True, device=cuda)
    losses = []
    infr.run(X=X, Y=Y, max_iter=3, learning_rate=1e-2, max_steps=40, on_step=lambda k, l: losses.append(l))
    ls = [float(l) for l in losses]
    assert all(np.isfinite(ls))
    assert np.mean(ls[-5:]) < np.mean(ls[:5])
    import mxfusion_b200 as mf
    mf.config.DEFAULT_DTYPE = 'float32'


def test_inference_save_load_round_trip_on_gpu(cuda, mf64, tmp_path):
    """Inference.save / load (inference.py:179-310) with the parameters living on the device: a freshly built inference of
    the same topology gets the trained values back (on the device) and reproduces the loss through the real kernels."""
    from tests.test_host_api import gp_notebook_model
    from mxfusion_b200.inference import GradBasedInference, MAP
    m, X, Y = gp_notebook_model(mf64)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]))
    infr.initialize(X=X.shape, Y=Y.shape)
    infr.run(X=X, Y=Y, max_iter=15, learning_rate=0.05)
    Xt, Yt = torch.tensor(X, device=cuda), torch.tensor(Y, device=cuda)
    loss, _ = infr.create_executor()(None, Xt, Yt)
    path = str(tmp_path / 'inference.zip')
    infr.save(path)
    m2, _, _ = gp_notebook_model(mf64)
    infr2 = GradBasedInference(inference_algorithm=MAP(model=m2, observed=[m2.X, m2.Y]))
    infr2.initialize(X=X.shape, Y=Y.shape)
    before, _ = infr2.create_executor()(None, Xt, Yt)
    assert abs(float(before) - float(loss)) > 1e-3
    infr2.load(path)
    after, _ = infr2.create_executor()(None, Xt, Yt)
    np.testing.assert_allclose(float(after), float(loss), rtol=1e-10)
    for a, b in ((m.kernel.variance, m2.kernel.variance), (m.kernel.lengthscale, m2.kernel.lengthscale),
                 (m.noise_var, m2.noise_var)):
        got, want = infr2.params[b], infr.params[a]
        assert got.is_cuda
        np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=1e-12)
    # training continues from the loaded state on the device
    infr2.run(X=X, Y=Y, max_iter=5, learning_rate=0.05)
    later, _ = infr2.create_executor()(None, Xt, Yt)
    assert float(later) < float(after)
