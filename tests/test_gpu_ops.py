"""The fused bounds and differentiable primitives on the real kernels (C ABI through ctypes) against the
oracle: values vs the NumPy restatement, gradients vs autograd through the torch restatement, both
float64 (tight) and float32 (the reference's only stated fp32 tolerance is rtol 1e-4 / atol 1e-5 for
an elementwise density; the fused bounds in fp32 are held to ~10x what a B200 measures against the float64 restatement:
values 2e-5 relative (measured <= 2.1e-6), every gradient within 2e-4 of its own max-norm (measured <= 2.2e-4 / typically
1e-5) plus 2e-5 of the call's largest gradient entry for gradients that are cancellation residuals (measured <= 4.1e-6)."""
import numpy as np
import pytest
import torch

from oracle import torch_ref, svgp as osvgp, gp as ogp

pytestmark = pytest.mark.gpu


# float32 gates (~10x the errors measured on a B200, printed with -s as MEASURED lines): see module docstring
F32_VALUE, F32_GRAD, F32_FLOOR = 2e-5, 2e-4, 2e-5


def _f32_close(tag, got, want, vtol):
    rel = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-30)))
    print('MEASURED value %s %.2e' % (tag, rel))
    np.testing.assert_allclose(got, want, rtol=vtol)


def _grad_close(tag, grads, tol, floor):
    """grads: {name: (got, want)}.  Every gradient within tol of its own max-norm + floor x the largest gradient entry of
    the call (a gradient that is a cancellation residual is held to the call's resolution, not to its own size)."""
    gmax = max(float(np.max(np.abs(w))) for _, w in grads.values())
    for k, (g, w) in grads.items():
        err, sc = float(np.max(np.abs(g - w))), float(np.max(np.abs(w)))
        print('MEASURED grad %s %s err/own %.2e err/gmax %.2e' % (tag, k, err / max(sc, 1e-300), err / max(gmax, 1e-300)))
        assert err <= tol * sc + floor * gmax, (tag, k, err, sc, gmax)


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
    _f32_close('svgp %s %s' % (dims, kind), got.detach().cpu().numpy(), want.detach().numpy(), 1e-9 if prec == 'f64' else F32_VALUE)
    _grad_close('svgp %s %s %s' % (dims, kind, prec), {k: (t[k].grad.double().cpu().numpy(), r[k].grad.numpy()) for k in a},
                1e-7 if prec == 'f64' else F32_GRAD, 1e-10 if prec == 'f64' else F32_FLOOR)


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
    _f32_close('gp %s %s' % (dims, kind), got.detach().cpu().numpy(), want.detach().numpy(), 1e-9 if prec == 'f64' else F32_VALUE)
    _grad_close('gp %s %s %s' % (dims, kind, prec), {k: (t[k].grad.double().cpu().numpy(), r[k].grad.numpy()) for k in d},
                1e-7 if prec == 'f64' else F32_GRAD, 1e-10 if prec == 'f64' else F32_FLOOR)
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
    restatement, at the module's float32 gates (F32_VALUE / F32_GRAD / F32_FLOOR)."""
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
    _f32_close(name, got.detach().cpu().numpy(), want.detach().numpy(), F32_VALUE)
    _grad_close(name, {k: (t[k].grad.double().cpu().numpy(), r[k].grad.numpy()) for k in a if k not in ('X', 'Y')},
                F32_GRAD, F32_FLOOR)
