"""The fused bounds and differentiable primitives on the real kernels (C ABI through ctypes) against the
oracle: values vs the NumPy restatement, gradients vs autograd through the torch restatement, both
float64 (tight) and float32 (the reference's only stated fp32 tolerance is rtol 1e-4 / atol 1e-5 for
an elementwise density; for the ill-conditioned solves here fp32 is held to 2e-3 relative on values
and 2e-2 of the gradient's max-norm)."""
import numpy as np
import pytest
import torch

from oracle import torch_ref, svgp as osvgp, gp as ogp

pytestmark = pytest.mark.gpu


def _svgp_inputs(rng, S, B, M, Din, P):
    return dict(X=rng.uniform(-2, 2, (S, B, Din)), Y=rng.randn(S, B, P), Z=rng.uniform(-2, 2, (S, M, Din)),
                noise=rng.rand(S, 1) * 0.2 + 0.05, mu=rng.randn(S, M, P) * 0.3,
                W=0.3 * np.eye(M)[None] + 0.05 * rng.randn(S, M, M), dv=rng.rand(S, M) * 0.3 + 0.2,
                ls=rng.rand(S, Din) * 0.5 + 0.8, var=rng.rand(S, 1) + 0.5)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('kind', [0, 1, 2, 3])
@pytest.mark.parametrize('dims', [(1, 10, 3, 3, 1), (2, 300, 70, 4, 2), (1, 1000, 257, 8, 1)])
def test_fused_svgp(cuda, prec, kind, dims):
    from mxfusion_b200 import ops
    S, B, M, Din, P = dims
    rng = np.random.RandomState(0)
    a = _svgp_inputs(rng, S, B, M, Din, P)
    tdt = torch.float64 if prec == 'f64' else torch.float32
    jit = 1e-6 if prec == 'f64' else 1e-4
    r = {k: torch.tensor(v, requires_grad=True) for k, v in a.items()}
    t = {k: torch.tensor(v, dtype=tdt, device=cuda, requires_grad=True) for k, v in a.items()}
    gout = rng.randn(S)
    want = torch_ref.svgp_log_pdf(kind, r['X'], r['Y'], r['Z'], r['noise'], r['mu'], r['W'], r['dv'], r['ls'],
                                  r['var'], jitter=jit, log_pdf_scaling=3.5)
    (want * torch.tensor(gout)).sum().backward()
    got = ops.svgp_log_pdf(kind, t['X'], t['Y'], t['Z'], t['noise'], t['mu'], t['W'], t['dv'], t['ls'], t['var'],
                           jitter=jit, log_pdf_scaling=3.5)
    (got * torch.tensor(gout, dtype=tdt, device=cuda)).sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-9 if prec == 'f64' else 2e-3)
    for k in a:
        g, w = t[k].grad.double().cpu().numpy(), r[k].grad.numpy()
        tol = 1e-7 if prec == 'f64' else 2e-2
        assert np.max(np.abs(g - w)) <= tol * (1e-3 + np.max(np.abs(w))), (k, np.max(np.abs(g - w)), np.max(np.abs(w)))


def test_fused_svgp_reference_fixture_known_answer(cuda):
    """testing/modules/svgpregression_test.py:41-56; ELBO = -32.72563540745786."""
    from mxfusion_b200 import ops
    np.random.seed(0)
    X = np.random.rand(10, 3); Y = np.random.rand(10, 1); Z = np.random.rand(3, 3)
    qU_mean = np.random.rand(3, 1); qU_cov_W = np.random.rand(3, 3); qU_cov_diag = np.random.rand(3,)
    noise_var = np.random.rand(1); lengthscale = np.random.rand(3); variance = np.random.rand(1)
    t = lambda a: torch.tensor(a[None], device=cuda)
    got = ops.svgp_log_pdf(0, t(X), t(Y), t(Z), t(noise_var), t(qU_mean), t(qU_cov_W), t(qU_cov_diag),
                           t(lengthscale), t(variance), jitter=1e-8)
    assert abs(float(got[0]) - (-32.72563540745786)) < 1e-9


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('kind', [0, 3])
@pytest.mark.parametrize('dims', [(2, 12, 3, 2), (1, 512, 2, 1)])
def test_fused_gp(cuda, prec, kind, dims):
    from mxfusion_b200 import ops
    S, N, Din, P = dims
    rng = np.random.RandomState(2)
    d = dict(X=rng.uniform(-3, 3, (S, N, Din)), Y=rng.randn(S, N, P), noise=rng.rand(S, 1) * 0.1 + 0.05,
             ls=rng.rand(S, Din) * 0.5 + 0.8, var=rng.rand(S, 1) + 0.5)
    tdt = torch.float64 if prec == 'f64' else torch.float32
    r = {k: torch.tensor(v, requires_grad=True) for k, v in d.items()}
    t = {k: torch.tensor(v, dtype=tdt, device=cuda, requires_grad=True) for k, v in d.items()}
    want = torch_ref.gp_log_pdf(kind, r['X'], r['Y'], r['noise'], r['ls'], r['var'], jitter=1e-6)
    want.sum().backward()
    got, L, LinvY = ops.gp_log_pdf(kind, t['X'], t['Y'], t['noise'], t['ls'], t['var'], jitter=1e-6)
    got.sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-9 if prec == 'f64' else 2e-3)
    for k in d:
        g, w = t[k].grad.double().cpu().numpy(), r[k].grad.numpy()
        tol = 1e-7 if prec == 'f64' else 2e-2
        assert np.max(np.abs(g - w)) <= tol * (1e-3 + np.max(np.abs(w))), (k, np.max(np.abs(g - w)), np.max(np.abs(w)))
    if prec == 'f64':
        wl = ogp.gp_log_pdf(kind, d['X'], d['Y'], d['noise'], d['ls'], d['var'], jitter=1e-6)
        np.testing.assert_allclose(L.cpu().numpy(), wl[1], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(LinvY.cpu().numpy(), wl[2], rtol=1e-8, atol=1e-10)


def test_primitive_adjoints_on_gpu(cuda):
    from mxfusion_b200 import ops
    rng = np.random.RandomState(3)
    S, n, k = 2, 97, 33
    Wm = rng.randn(S, n, n)
    A0 = Wm @ np.swapaxes(Wm, -1, -2) + n * np.eye(n)
    B0 = rng.randn(S, n, k)
    for transpose in (False, True):
        A = torch.tensor(A0, requires_grad=True)
        B = torch.tensor(B0, requires_grad=True)
        A2 = torch.tensor(A0, device=cuda, requires_grad=True)
        B2 = torch.tensor(B0, device=cuda, requires_grad=True)
        Lr = torch.linalg.cholesky(A)
        (torch_ref.trsm(Lr, B, transpose).sin().sum() + torch_ref.sumlogdiag(Lr).sum()).backward()
        L = ops.potrf(A2)
        (ops.trsm(L, B2, transpose=transpose).sin().sum() + ops.sumlogdiag(L).sum()).backward()
        sym = lambda g: 0.5 * (g + g.transpose(-1, -2))
        np.testing.assert_allclose(sym(A2.grad).cpu().numpy(), sym(A.grad).numpy(), rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(B2.grad.cpu().numpy(), B.grad.numpy(), rtol=1e-7, atol=1e-10)


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path must fail loudly off the GPU."""
    from mxfusion_b200 import ops, _lib
    x = torch.zeros((1, 4, 2))
    with pytest.raises(_lib.MXFusionB200Error):
        ops.kernel_matrix(0, x, None, torch.ones((1, 1)), torch.ones((1, 1)))


@pytest.mark.parametrize('name,kind,B,M,Din', [('C2 SVGP M=512 D=8 RBF B=2048', 0, 2048, 512, 8),
                                               ('C3 SVGP M=1024 D=16 Matern52 B=4096', 3, 4096, 1024, 16),
                                               ('H  SVGP M=1024 D=8 RBF B=4096', 0, 4096, 1024, 8)])
def test_fused_svgp_at_baseline_shapes_f32(cuda, name, kind, B, M, Din):
    """BASELINE.json configs 2, 3 and the headline shape in float32 (the reference's default dtype) against the float64
    restatement: value within 2e-3 relative, every gradient within 2e-2 of its max-norm (fp32 with a jittered Kuu)."""
    from mxfusion_b200 import ops
    rng = np.random.RandomState(7)
    X = rng.uniform(-3, 3, (1, B, Din))
    Y = (np.sin(X).sum(-1, keepdims=True) / np.sqrt(Din) + 0.05 * rng.randn(1, B, 1))
    a = dict(X=X, Y=Y, Z=X[:, rng.permutation(B)[:M]].copy(), noise=np.full((1, 1), 0.05), mu=0.1 * rng.randn(1, M, 1),
             W=0.05 * rng.randn(1, M, M) / np.sqrt(M), dv=np.full((1, M), 0.5), ls=np.full((1, 1), 1.3), var=np.full((1, 1), 1.1))
    r = {k: torch.tensor(v, requires_grad=True) for k, v in a.items()}
    t = {k: torch.tensor(v, dtype=torch.float32, device=cuda, requires_grad=True) for k, v in a.items()}
    want = torch_ref.svgp_log_pdf(kind, r['X'], r['Y'], r['Z'], r['noise'], r['mu'], r['W'], r['dv'], r['ls'], r['var'],
                                  jitter=1e-4, log_pdf_scaling=25.0)
    want.sum().backward()
    got = ops.svgp_log_pdf(kind, t['X'], t['Y'], t['Z'], t['noise'], t['mu'], t['W'], t['dv'], t['ls'], t['var'],
                           jitter=1e-4, log_pdf_scaling=25.0)
    got.sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=2e-3)
    for k in a:
        if k in ('X', 'Y'):
            continue
        g, w = t[k].grad.double().cpu().numpy(), r[k].grad.numpy()
        assert np.max(np.abs(g - w)) <= 2e-2 * (1e-3 + np.max(np.abs(w))), (name, k, np.max(np.abs(g - w)), np.max(np.abs(w)))
