"""Host-side algebra of mxfusion_b200/ops.py checked on the CPU: the hand-derived adjoints of the fused
SVGP / exact-GP bounds and of the primitives must equal autograd through the op-for-op restatement of the
reference (oracle/torch_ref.py, float64).  `ops.R` (the CUDA binding) is replaced by tests/raw_standin.py;
the same checks run against the real kernels in tests/test_gpu_ops.py."""
import numpy as np
import pytest
import torch

from oracle import torch_ref
from oracle import svgp as osvgp
from oracle import gp as ogp


@pytest.fixture()
def ops(monkeypatch):
    from mxfusion_b200 import ops as _ops
    from tests import raw_standin
    monkeypatch.setattr(_ops, 'R', raw_standin)
    return _ops


def _svgp_inputs(rng, S, B, M, Din, P, ard=True):
    d = dict(X=rng.rand(S, B, Din), Y=rng.rand(S, B, P), Z=rng.rand(S, M, Din), noise=rng.rand(S, 1) + 0.1,
             mu=rng.rand(S, M, P), W=rng.rand(S, M, M), dv=rng.rand(S, M) + 0.1,
             ls=rng.rand(S, Din if ard else 1) + 0.5, var=rng.rand(S, 1) + 0.5)
    return {k: torch.tensor(v, requires_grad=True) for k, v in d.items()}


@pytest.mark.parametrize('kind', [0, 1, 2, 3])
@pytest.mark.parametrize('dims', [(1, 10, 3, 3, 1), (2, 17, 5, 2, 3)])
def test_fused_svgp_value_and_gradients(ops, kind, dims):
    S, B, M, Din, P = dims
    rng = np.random.RandomState(0)
    a = _svgp_inputs(rng, S, B, M, Din, P)
    b = {k: v.detach().clone().requires_grad_() for k, v in a.items()}
    gout = torch.tensor(rng.randn(S))
    want = torch_ref.svgp_log_pdf(kind, a['X'], a['Y'], a['Z'], a['noise'], a['mu'], a['W'], a['dv'], a['ls'],
                                  a['var'], jitter=1e-6, log_pdf_scaling=3.5)
    (want * gout).sum().backward()
    got = ops.svgp_log_pdf(kind, b['X'], b['Y'], b['Z'], b['noise'], b['mu'], b['W'], b['dv'], b['ls'], b['var'],
                           jitter=1e-6, log_pdf_scaling=3.5)
    (got * gout).sum().backward()
    np.testing.assert_allclose(got.detach().numpy(), want.detach().numpy(), rtol=1e-10)
    for k in a:
        np.testing.assert_allclose(b[k].grad.numpy(), a[k].grad.numpy(), rtol=2e-7, atol=1e-9, err_msg=k)


def test_fused_svgp_matches_numpy_oracle_on_reference_fixture(ops):
    """testing/modules/svgpregression_test.py:41-56 fixture; known answer from BASELINE.md."""
    np.random.seed(0)
    X = np.random.rand(10, 3); Y = np.random.rand(10, 1); Z = np.random.rand(3, 3)
    qU_mean = np.random.rand(3, 1); qU_cov_W = np.random.rand(3, 3); qU_cov_diag = np.random.rand(3,)
    noise_var = np.random.rand(1); lengthscale = np.random.rand(3); variance = np.random.rand(1)
    t = lambda a: torch.tensor(a[None])
    got = ops.svgp_log_pdf(0, t(X), t(Y), t(Z), t(noise_var), t(qU_mean), t(qU_cov_W), t(qU_cov_diag),
                           t(lengthscale), t(variance), jitter=1e-8)
    assert abs(float(got[0]) - (-32.72563540745786)) < 1e-9
    want = osvgp.svgp_log_pdf(0, X[None], Y[None], Z[None], noise_var[None], qU_mean[None], qU_cov_W[None],
                              qU_cov_diag[None], lengthscale[None], variance[None], jitter=1e-8)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)


def test_fused_svgp_sample_axis_broadcast(ops):
    """X sampled (S=3) while parameters carry S=1: gradients of broadcast operands are summed over S."""
    rng = np.random.RandomState(1)
    a = _svgp_inputs(rng, 1, 9, 4, 2, 1)
    a['X'] = torch.tensor(rng.rand(3, 9, 2), requires_grad=True)
    b = {k: v.detach().clone().requires_grad_() for k, v in a.items()}
    ex = lambda t: t.expand((3,) + tuple(t.shape[1:]))
    want = torch_ref.svgp_log_pdf(0, a['X'], ex(a['Y']), ex(a['Z']), ex(a['noise']), ex(a['mu']), ex(a['W']),
                                  ex(a['dv']), ex(a['ls']), ex(a['var']), jitter=1e-6)
    want.mean().backward()
    got = ops.svgp_log_pdf(0, b['X'], b['Y'], b['Z'], b['noise'], b['mu'], b['W'], b['dv'], b['ls'], b['var'],
                           jitter=1e-6)
    got.mean().backward()
    for k in a:
        np.testing.assert_allclose(b[k].grad.numpy(), a[k].grad.numpy(), rtol=2e-7, atol=1e-10, err_msg=k)


@pytest.mark.parametrize('kind', [0, 3])
def test_fused_gp_value_and_gradients(ops, kind):
    rng = np.random.RandomState(2)
    S, N, Din, P = 2, 12, 3, 2
    d = dict(X=rng.rand(S, N, Din), Y=rng.rand(S, N, P), noise=rng.rand(S, 1) + 0.1, ls=rng.rand(S, Din) + 0.5,
             var=rng.rand(S, 1) + 0.5)
    a = {k: torch.tensor(v, requires_grad=True) for k, v in d.items()}
    b = {k: v.detach().clone().requires_grad_() for k, v in a.items()}
    gout = torch.tensor(rng.randn(S))
    want = torch_ref.gp_log_pdf(kind, a['X'], a['Y'], a['noise'], a['ls'], a['var'], jitter=1e-6)
    (want * gout).sum().backward()
    got, L, LinvY = ops.gp_log_pdf(kind, b['X'], b['Y'], b['noise'], b['ls'], b['var'], jitter=1e-6)
    (got * gout).sum().backward()
    np.testing.assert_allclose(got.detach().numpy(), want.detach().numpy(), rtol=1e-10)
    for k in a:
        np.testing.assert_allclose(b[k].grad.numpy(), a[k].grad.numpy(), rtol=2e-7, atol=1e-9, err_msg=k)
    wantL = ogp.gp_log_pdf(kind, d['X'], d['Y'], d['noise'], d['ls'], d['var'], jitter=1e-6)
    np.testing.assert_allclose(L.numpy(), wantL[1], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(LinvY.numpy(), wantL[2], rtol=1e-9, atol=1e-12)


def test_primitive_adjoints(ops):
    rng = np.random.RandomState(3)
    S, n, k = 2, 7, 4
    Wm = rng.randn(S, n, n)
    A0 = Wm @ np.swapaxes(Wm, -1, -2) + n * np.eye(n)
    B0 = rng.randn(S, n, k)
    for transpose in (False, True):
        A = torch.tensor(A0, requires_grad=True)
        B = torch.tensor(B0, requires_grad=True)
        A2 = A.detach().clone().requires_grad_()
        B2 = B.detach().clone().requires_grad_()
        Lr = torch.linalg.cholesky(A)
        Xr = torch_ref.trsm(Lr, B, transpose) * 0.7
        (Xr.sin().sum() + torch_ref.sumlogdiag(Lr).sum()).backward()
        L = ops.potrf(A2)
        X = ops.trsm(L, B2, transpose=transpose, alpha=0.7)
        (X.sin().sum() + ops.sumlogdiag(L).sum()).backward()
        # the gradient wrt a symmetric input is defined up to its symmetric part
        sym = lambda g: 0.5 * (g + g.transpose(-1, -2))
        np.testing.assert_allclose(sym(A2.grad).numpy(), sym(A.grad).numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(B2.grad.numpy(), B.grad.numpy(), rtol=1e-8, atol=1e-10)
    # gemm2 / syrk / make_diagonal / softplus
    P = torch.tensor(rng.randn(S, 5, 3), requires_grad=True)
    Q = torch.tensor(rng.randn(S, 5, 4), requires_grad=True)
    v = torch.tensor(rng.randn(S, 3), requires_grad=True)
    P2, Q2, v2 = [t.detach().clone().requires_grad_() for t in (P, Q, v)]
    ref = 1.3 * torch.matmul(P.transpose(-1, -2), Q)
    ref2 = torch.matmul(P.transpose(-1, -2), P) + torch.diag_embed(torch.nn.functional.softplus(v))
    (ref.cos().sum() + ref2.sin().sum()).backward()
    got = ops.gemm2(P2, Q2, True, False, alpha=1.3)
    got2 = ops.syrk(P2, transpose=True) + ops.make_diagonal(ops.softplus(v2))
    (got.cos().sum() + got2.sin().sum()).backward()
    for x, y in ((P2, P), (Q2, Q), (v2, v)):
        np.testing.assert_allclose(x.grad.numpy(), y.grad.numpy(), rtol=1e-9, atol=1e-12)
