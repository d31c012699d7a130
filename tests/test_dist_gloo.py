"""Data-parallel minibatch VI on two CPU processes (gloo): rows sharded across ranks, one all-reduce of the flat
gradient bucket per step.  Checks that both ranks hold identical parameters and that the result equals a
single-process emulation that averages the two ranks' gradients (the CUDA binding is replaced by the CPU stand-in;
the NCCL path runs the same Stepper code, see bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(mf, N, M, B, seed):
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 1))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=1, variance=1, lengthscale=1)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1), num_inducing=M)
    m.Y.factor.svgp_log_pdf.jitter = 1e-6
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / B}, rng=np.random.RandomState(seed))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, context='cpu')
    infr.initialize(X=(N // 2, 1), Y=(N // 2, 1))
    post = m.Y.factor._extra_graphs[0]
    infr.params[m.Y.factor.inducing_inputs] = np.linspace(-3, 3, M)[:, None]
    infr.params[post.qU_mean] = np.zeros((M, 1))
    infr.params[post.qU_cov_W] = np.eye(M) * 0.1
    infr.params[post.qU_cov_diag] = np.ones(M) * 0.5
    return m, infr, loop


def _data(N):
    rng = np.random.RandomState(0)
    X = rng.uniform(-3., 3., (N, 1))
    Y = np.sin(X) + rng.randn(N, 1) * 0.05
    return X, Y


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import mxfusion_b200 as mf
    from mxfusion_b200 import ops
    from tests import raw_standin
    ops.R = raw_standin
    mf.config.DEFAULT_DTYPE = 'float64'
    mf.config.MXNET_DEFAULT_DEVICE = 'cpu'
    dist.init_process_group('gloo', rank=rank, world_size=world)
    N, M, B = 200, 6, 20
    X, Y = _data(N)
    sh = N // world
    m, infr, loop = _build(mf, N, M, B, seed=100 + rank)
    infr.run(X=X[rank * sh:(rank + 1) * sh], Y=Y[rank * sh:(rank + 1) * sh], max_iter=2, learning_rate=0.05)
    np.save(os.path.join(out_dir, 'flat_%d.npy' % rank), infr.params.flat.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_gradient_averaging(tmp_path, monkeypatch):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    f0, f1 = np.load(tmp_path / 'flat_0.npy'), np.load(tmp_path / 'flat_1.npy')
    np.testing.assert_array_equal(f0, f1)                      # replicas stay bit-identical

    # single-process emulation: per step, average the two ranks' gradients, one Adam update with rescale 1/B
    import mxfusion_b200 as mf
    from mxfusion_b200 import ops
    from mxfusion_b200.inference import RolloverBatchSampler
    from tests import raw_standin
    monkeypatch.setattr(ops, 'R', raw_standin)
    monkeypatch.setattr(mf.config, 'DEFAULT_DTYPE', 'float64')
    monkeypatch.setattr(mf.config, 'MXNET_DEFAULT_DEVICE', 'cpu')
    N, M, B = 200, 6, 20
    X, Y = _data(N)
    sh = N // 2
    m, infr, loop = _build(mf, N, M, B, seed=0)
    ex = infr.create_executor()
    p = infr.params
    p.refresh_leaves()
    samplers = [RolloverBatchSampler(sh, B, rng=np.random.RandomState(100 + r)) for r in range(2)]
    for epoch in range(2):
        idx = [s.epoch_indices() for s in samplers]
        for i in range(idx[0][1]):
            p.gflat.zero_()
            for r in range(2):
                sel = idx[r][0][i * B:(i + 1) * B] + r * sh
                loss, lg = ex(None, torch.tensor(X[sel]), torch.tensor(Y[sel]))
                lg.backward()
            raw_standin.adam_step_(p.flat, p.gflat, p.adam_m, p.adam_v, p.adam_t, lr=0.05, rescale=1.0 / B / 2)
    np.testing.assert_allclose(f0, p.flat.numpy(), rtol=1e-9, atol=1e-12)


def test_c_abi_library_exports_every_declared_symbol():
    """include/mxf_b200.h vs libmxf_b200.so: the library loads without a GPU and exports each declared entry point
    (no compute call is made)."""
    import re
    import ctypes
    from mxfusion_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, 'include', 'mxf_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(mxf_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    bound = set(_lib._declare(lib).keys())
    assert set(declared) <= bound | {'mxf_debug_set_prof'}, sorted(set(declared) - bound)
    assert lib.mxf_version() >= 100


def test_ops_raise_without_cuda_tensors():
    from mxfusion_b200 import ops, _lib
    with pytest.raises(_lib.MXFusionB200Error):
        ops.potrf(torch.eye(4).unsqueeze(0))
    with pytest.raises(_lib.MXFusionB200Error):
        ops.svgp_log_pdf(0, torch.zeros(1, 4, 2), torch.zeros(1, 4, 1), torch.zeros(1, 2, 2), torch.ones(1, 1),
                         torch.zeros(1, 2, 1), torch.zeros(1, 2, 2), torch.ones(1, 2), torch.ones(1, 1), torch.ones(1, 1))
