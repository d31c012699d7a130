"""Parity of every C-ABI kernel (through mxfusion_b200._raw -> ctypes -> libmxf_b200.so) with the
NumPy oracle on the same seeded inputs.  Tolerances: float64 follows the reference's own
`np.allclose` defaults (rtol 1e-5, atol 1e-8; testing/modules/svgpregression_test.py:115) tightened
to 1e-9 where the arithmetic is benign; float32 uses the only fp32 tolerance the reference states
(rtol 1e-4, atol 1e-5; testing/components/distributions/normal_test.py:63-67) unless noted."""
import numpy as np
import pytest
import torch

from oracle import kernels as ok, linalg as ol, normal as on, loop as oloop

pytestmark = pytest.mark.gpu

KINDS = [ok.RBF, ok.MATERN12, ok.MATERN32, ok.MATERN52]
DT = {'f64': (torch.float64, np.float64, 1e-9, 1e-11), 'f32': (torch.float32, np.float32, 2e-4, 2e-5)}


def T(a, cuda, dt):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=cuda)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('shape', [(1, 10, 7, 3, True), (2, 33, 130, 8, False), (1, 70, 300, 16, True),
                                   (3, 5, 1, 1, False), (1, 257, 513, 2, False)])
def test_kbuild_cross(cuda, prec, kind, shape):
    from mxfusion_b200 import _raw
    S, N, N2, D, ard = shape
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(1)
    X = rng.uniform(-2, 2, (S, N, D)).astype(ndt)
    X2 = rng.uniform(-2, 2, (S, N2, D)).astype(ndt)
    ls = rng.uniform(0.5, 2.0, (S, D if ard else 1)).astype(ndt)
    var = rng.uniform(0.5, 2.0, (S, 1)).astype(ndt)
    want = ok.K(kind, X.astype(np.float64), ls.astype(np.float64), var.astype(np.float64), X2.astype(np.float64))
    got = _raw.kbuild_fwd(kind, T(X, cuda, tdt), T(X2, cuda, tdt), T(ls, cuda, tdt), T(var, cuda, tdt))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=rtol, atol=atol * 10 if prec == 'f32' else atol)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('shape', [(1, 300, 512, 16, True), (2, 129, 600, 16, False), (1, 1000, 1100, 8, True),
                                   (1, 128, 128, 5, False), (3, 77, 260, 12, True), (1, 4096, 1024, 16, False),
                                   (1, 50, 130, 16, True)])
def test_kbuild_cross_tensor_core_path(cuda, kind, shape):
    """The tcgen05 + TMA-store K-build (csrc/kbuild_tc.cuh; stationary.py:102's gemm2 on the tensor pipe, 3xTF32) against the
    f64 oracle at the f32 tolerance of the streaming kernel, and against the streaming FMA kernel itself; ragged row tiles,
    ragged / multiple 512-column groups, D <= 8 (one K slice) and 8 < D <= 16 (two), sample axis, and a width TMA cannot
    store without touching the neighbours (N2 = 130, not a multiple of 4: must fall back to the streaming kernel, not fail)."""
    from mxfusion_b200 import _raw
    S, N, N2, D, ard = shape
    tdt, ndt, rtol, atol = DT['f32']
    rng = np.random.RandomState(11)
    X = rng.uniform(-2, 2, (S, N, D)).astype(ndt)
    X2 = rng.uniform(-2, 2, (S, N2, D)).astype(ndt)
    ls = rng.uniform(0.5, 2.0, (S, D if ard else 1)).astype(ndt)
    var = rng.uniform(0.5, 2.0, (S, 1)).astype(ndt)
    want = ok.K(kind, X.astype(np.float64), ls.astype(np.float64), var.astype(np.float64), X2.astype(np.float64))
    args = (kind, T(X, cuda, tdt), T(X2, cuda, tdt), T(ls, cuda, tdt), T(var, cuda, tdt))
    old = _raw.kbuild_tc_threshold(1 << 62)
    try:
        fma = _raw.kbuild_fwd(*args).cpu().numpy()
        _raw.kbuild_tc_threshold(0)
        before = _raw.launch_count()
        got = _raw.kbuild_fwd(*args)
        # sentinel-filled strided output: the TMA stores must clip at the edges of the (N, N2) view
        big = torch.full((S, N + 3, N2 + 4 - N2 % 4 + 8), -7.0, dtype=tdt, device=cuda)
        view = big[:, :N, :N2]
        _raw.kbuild_fwd(*args, out=view)
        assert _raw.launch_count() - before == 2
    finally:
        _raw.kbuild_tc_threshold(old)
    got = got.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol * 10)
    np.testing.assert_allclose(got, fma, rtol=2e-5, atol=2e-6)
    big = big.cpu().numpy()
    np.testing.assert_array_equal(big[:, :N, :N2], got)
    assert (big[:, N:, :] == -7.0).all() and (big[:, :, N2:] == -7.0).all()


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('shape', [(2, 300, 16), (1, 1024, 8), (1, 132, 5), (1, 640, 12)])
def test_kbuild_symmetric_tensor_core_path(cuda, kind, shape):
    """K(X, X) + eye * (noise + jitter) on the tcgen05 kernel: exact kernel value on the diagonal, the diagonal term folded
    into the store, 256- and 512-column groups; against the f64 oracle and the streaming kernel."""
    from mxfusion_b200 import _raw
    S, N, D = shape
    tdt, ndt, rtol, atol = DT['f32']
    rng = np.random.RandomState(12)
    X = rng.uniform(-2, 2, (S, N, D)).astype(ndt)
    ls = rng.uniform(0.5, 2.0, (S, D)).astype(ndt)
    var = rng.uniform(0.5, 2.0, (S, 1)).astype(ndt)
    noise = rng.uniform(0.1, 0.2, (S, 1)).astype(ndt)
    want = ok.K(kind, X.astype(np.float64), ls.astype(np.float64), var.astype(np.float64)) + \
        np.eye(N)[None] * (noise.astype(np.float64)[..., None] + 1e-3)
    args = (kind, T(X, cuda, tdt), None, T(ls, cuda, tdt), T(var, cuda, tdt))
    old = _raw.kbuild_tc_threshold(1 << 62)
    try:
        fma = _raw.kbuild_fwd(*args, diag_add=T(noise, cuda, tdt), diag_const=1e-3).cpu().numpy()
        _raw.kbuild_tc_threshold(0)
        got = _raw.kbuild_fwd(*args, diag_add=T(noise, cuda, tdt), diag_const=1e-3).cpu().numpy()
    finally:
        _raw.kbuild_tc_threshold(old)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol * 10)
    np.testing.assert_allclose(got, fma, rtol=2e-5, atol=2e-6)
    d = np.arange(N)
    np.testing.assert_array_equal(got[:, d, d], fma[:, d, d])


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('kind', KINDS)
def test_kbuild_symmetric_with_diag(cuda, prec, kind):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(2)
    S, N, D = 2, 67, 3
    X = rng.uniform(-2, 2, (S, N, D)).astype(ndt)
    ls = rng.uniform(0.5, 2.0, (S, D)).astype(ndt)
    var = rng.uniform(0.5, 2.0, (S, 1)).astype(ndt)
    noise = rng.uniform(0.1, 0.2, (S, 1)).astype(ndt)
    want = ok.K(kind, X.astype(np.float64), ls.astype(np.float64), var.astype(np.float64)) + \
        np.eye(N)[None] * (noise.astype(np.float64)[..., None] + 1e-3)
    got = _raw.kbuild_fwd(kind, T(X, cuda, tdt), None, T(ls, cuda, tdt), T(var, cuda, tdt),
                          diag_add=T(noise, cuda, tdt), diag_const=1e-3)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=rtol, atol=atol * 10 if prec == 'f32' else atol)


def test_kbuild_sample_broadcast(cuda):
    """X shared over samples (S=1) while the kernel parameters are sampled (S=3)."""
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(3)
    X = rng.rand(1, 20, 4)
    X2 = rng.rand(1, 9, 4)
    ls = rng.uniform(0.5, 2.0, (3, 4))
    var = rng.uniform(0.5, 2.0, (3, 1))
    want = ok.K(ok.RBF, np.repeat(X, 3, 0), ls, var, np.repeat(X2, 3, 0))
    got = _raw.kbuild_fwd(ok.RBF, T(X, cuda, torch.float64), T(X2, cuda, torch.float64),
                          T(ls, cuda, torch.float64), T(var, cuda, torch.float64))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('sym', [False, True])
@pytest.mark.parametrize('ard', [False, True])
def test_kbuild_bwd_matches_autograd_of_restatement(cuda, kind, sym, ard):
    """The reference never tests gradients; the oracle is autograd through the
    op-for-op torch restatement (oracle/torch_ref.py) in float64."""
    from mxfusion_b200 import _raw
    from oracle import torch_ref
    rng = np.random.RandomState(4)
    S, N, N2, D = 2, 37, 37 if sym else 53, 5
    X = torch.tensor(rng.uniform(-2, 2, (S, N, D)), requires_grad=True)
    X2 = None if sym else torch.tensor(rng.uniform(-2, 2, (S, N2, D)), requires_grad=True)
    ls = torch.tensor(rng.uniform(0.5, 2.0, (S, D if ard else 1)), requires_grad=True)
    var = torch.tensor(rng.uniform(0.5, 2.0, (S, 1)), requires_grad=True)
    G = torch.tensor(rng.randn(S, N, N2))
    Kt = torch_ref.K(kind, X, ls, var, X2)
    (Kt * G).sum().backward()
    dX, dX2, dls, dvar = _raw.kbuild_bwd(kind, X.detach().to(cuda), None if sym else X2.detach().to(cuda),
                                         ls.detach().to(cuda), var.detach().to(cuda), G.to(cuda))
    np.testing.assert_allclose(dX.cpu().numpy(), X.grad.numpy(), rtol=1e-7, atol=1e-9)
    if not sym:
        np.testing.assert_allclose(dX2.cpu().numpy(), X2.grad.numpy(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dls.cpu().numpy(), ls.grad.numpy(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dvar.cpu().numpy(), var.grad.numpy(), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('ta,tb', [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize('mnk', [(5, 7, 3), (130, 65, 33), (257, 300, 129), (64, 64, 64), (1, 1, 1)])
def test_gemm(cuda, prec, ta, tb, mnk):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    m, n, k = mnk
    rng = np.random.RandomState(5)
    S = 2
    A = rng.randn(S, k, m) if ta else rng.randn(S, m, k)
    B = rng.randn(S, n, k) if tb else rng.randn(S, k, n)
    C0 = rng.randn(S, m, n)
    want = 0.7 * ol.gemm2(A, B, ta, tb) + 0.3 * C0
    C = T(C0, cuda, tdt)
    _raw.gemm(T(A, cuda, tdt), T(B, cuda, tdt), ta, tb, alpha=0.7, beta=0.3, C=C)
    np.testing.assert_allclose(C.cpu().numpy(), want, rtol=rtol, atol=atol * 50 if prec == 'f32' else atol)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('n', [1, 3, 64, 65, 130, 300, 513])
def test_potrf(cuda, prec, n):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(6)
    S = 2
    W = rng.randn(S, n, n)
    A = W @ np.swapaxes(W, -1, -2) + n * np.eye(n)[None]
    want = ol.potrf(A)
    L, info = _raw.potrf_(T(A, cuda, tdt))
    assert info.cpu().tolist() == [0, 0]
    got = L.cpu().numpy()
    assert np.all(np.triu(got, 1) == 0), "strict upper triangle must be zeroed (MXNet potrf convention)"
    np.testing.assert_allclose(got, want, rtol=rtol * 5, atol=atol * 100 if prec == 'f32' else atol * 100)


def test_potrf_reports_first_bad_pivot(cuda):
    from mxfusion_b200 import _raw
    A = np.eye(100)[None].repeat(2, 0)
    A[1, 70, 70] = -1.0
    L, info = _raw.potrf_(T(A, cuda, torch.float64))
    assert info.cpu().tolist() == [0, 71]


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('transpose', [False, True])
@pytest.mark.parametrize('n,nrhs', [(1, 1), (3, 1), (64, 5), (65, 130), (200, 257), (513, 33)])
def test_trsm(cuda, prec, transpose, n, nrhs):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(7)
    S = 2
    Lm = np.tril(rng.randn(S, n, n)) * 0.1 + 2.0 * np.eye(n)[None]
    B = rng.randn(S, n, nrhs)
    want = ol.trsm(Lm, B, transpose=transpose, alpha=0.5)
    Bt = T(B, cuda, tdt)
    _raw.trsm_(T(Lm, cuda, tdt), Bt, transpose=transpose, alpha=0.5)
    np.testing.assert_allclose(Bt.cpu().numpy(), want, rtol=rtol * 5, atol=atol * 100)


def test_trsm_broadcast_factor(cuda):
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(8)
    Lm = np.tril(rng.randn(1, 50, 50)) * 0.1 + 2.0 * np.eye(50)[None]
    B = rng.randn(3, 50, 7)
    want = ol.trsm(np.repeat(Lm, 3, 0), B)
    Bt = T(B, cuda, torch.float64)
    _raw.trsm_(T(Lm, cuda, torch.float64), Bt)
    np.testing.assert_allclose(Bt.cpu().numpy(), want, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_square_matrix_utilities(cuda, prec):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(9)
    A = rng.randn(2, 77, 77).astype(ndt)
    At = T(A, cuda, tdt)
    At_T = np.swapaxes(A, -1, -2)
    np.testing.assert_array_equal(_raw.copy_ltu(At).cpu().numpy(), np.tril(A) + np.swapaxes(np.tril(A, -1), -1, -2))
    np.testing.assert_allclose(_raw.symmetrize(At, 0.5).cpu().numpy(), 0.5 * (A + At_T), rtol=1e-6)
    np.testing.assert_array_equal(_raw.tril(At).cpu().numpy(), np.tril(A))
    np.testing.assert_array_equal(_raw.tril(At, strict=True).cpu().numpy(), np.tril(A, -1))
    R = rng.randn(2, 45, 131).astype(ndt)
    np.testing.assert_array_equal(_raw.transpose(T(R, cuda, tdt)).cpu().numpy(), np.swapaxes(R, -1, -2))
    d = rng.randn(2, 77).astype(ndt)
    want = A + ol.make_diagonal(d) + np.eye(77, dtype=ndt) * ndt(0.25)
    got = _raw.add_diag_(T(A, cuda, tdt), T(d, cuda, tdt), 0.25).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(_raw.get_diag(At).cpu().numpy(), ol.make_diagonal_backward(A))


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('shape', [(2, 1, 1), (2, 33, 7), (1, 1000, 1024), (3, 517, 129)])
def test_reductions(cuda, prec, shape):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(10)
    a = rng.uniform(0.5, 1.5, shape)
    b = rng.randn(*shape)
    at, bt = T(a, cuda, tdt), T(b, cuda, tdt)
    for op, want in [(_raw.RED_SUM, a.sum((1, 2))), (_raw.RED_SUMSQ, (a * a).sum((1, 2))),
                     (_raw.RED_DOT, (a * b).sum((1, 2))), (_raw.RED_SUMLOG, np.log(a).sum((1, 2)))]:
        got = _raw.reduce(op, at, bt if op == _raw.RED_DOT else None, scale=0.5)
        np.testing.assert_allclose(got.cpu().numpy(), 0.5 * want, rtol=rtol, atol=atol * 100)
    A = rng.uniform(0.5, 2, (2, 40, 40))
    np.testing.assert_allclose(_raw.sumlogdiag(T(A, cuda, tdt)).cpu().numpy(), ol.sumlogdiag(A), rtol=rtol)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_normal_logpdf_sum_and_adjoint(cuda, prec):
    from mxfusion_b200 import _raw
    from oracle import torch_ref
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(11)
    S, shape = 3, (41, 5)
    x = rng.randn(S, *shape)
    m = rng.randn(1, *shape)
    v = rng.uniform(0.5, 2.0, (1,) + shape)
    want = 1.7 * on.factor_reduce(on.log_pdf(m, v, x))
    got = _raw.normal_logpdf_sum(T(x, cuda, tdt), T(m, cuda, tdt), T(v, cuda, tdt), scale=1.7)
    np.testing.assert_allclose(got.cpu().numpy()[0], want, rtol=rtol * 5)
    xt = torch.tensor(x, requires_grad=True)
    mt = torch.tensor(m, requires_grad=True)
    vt = torch.tensor(v, requires_grad=True)
    (1.7 * torch.sum(torch.mean(torch_ref.normal_log_pdf(mt, vt, xt), dim=0)) * 0.3).backward()
    gx, gm, gv = _raw.normal_logpdf_sum_bwd(T(x, cuda, tdt), T(m, cuda, tdt), T(v, cuda, tdt),
                                            torch.tensor([0.3], dtype=tdt, device=cuda), scale=1.7)
    np.testing.assert_allclose(gx.cpu().numpy(), xt.grad.numpy(), rtol=rtol * 5, atol=atol)
    np.testing.assert_allclose(gm.cpu().numpy(), mt.grad.numpy(), rtol=rtol * 5, atol=atol * 10)
    np.testing.assert_allclose(gv.cpu().numpy(), vt.grad.numpy(), rtol=rtol * 5, atol=atol * 10)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_normal_reparam_injected_eps(cuda, prec):
    """normal_test.py:78-109 pattern: the standard-normal draw is injected (MockMXNetRandomGenerator)."""
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(12)
    S, shape = 4, (13, 3)
    eps = rng.randn(S, *shape)
    m = rng.randn(1, *shape)
    v = rng.uniform(0.5, 2.0, (1,) + shape)
    got = _raw.normal_reparam(T(m, cuda, tdt), T(v, cuda, tdt), S, eps=T(eps, cuda, tdt))
    np.testing.assert_allclose(got.cpu().numpy(), on.draw_samples(m, v, eps), rtol=rtol, atol=atol)


def test_normal_reparam_philox_moments(cuda):
    from mxfusion_b200 import _raw
    m = torch.zeros((1, 1 << 20), device=cuda)
    v = torch.ones((1, 1 << 20), device=cuda)
    w, eps = _raw.normal_reparam(m, v, 4, seed=123, offset=0, return_eps=True)
    assert torch.equal(w, eps)
    z = w.double().cpu().numpy().ravel()
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3
    assert abs(((z - z.mean()) ** 4).mean() - 3) < 3e-2
    w2 = _raw.normal_reparam(m, v, 4, seed=123, offset=0)
    assert torch.equal(w, w2), "counter-based stream must be reproducible"
    w3 = _raw.normal_reparam(m, v, 4, seed=123, offset=1)
    assert not torch.equal(w, w3)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_adam_matches_mxnet_update_rule(cuda, prec):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(13)
    n = 1000
    w = rng.randn(n)
    m = np.zeros(n)
    v = np.zeros(n)
    wt, mt, vt = T(w, cuda, tdt), T(m, cuda, tdt), T(v, cuda, tdt)
    cnt = torch.zeros((1,), dtype=torch.int32, device=cuda)
    for t in range(1, 6):
        g = rng.randn(n)
        w, m, v = oloop.adam_step(w, g, m, v, t, lr=0.01, rescale_grad=1. / 16)
        _raw.adam_step_(wt, T(g, cuda, tdt), mt, vt, cnt, lr=0.01, rescale=1. / 16)
    assert int(cnt.item()) == 5
    np.testing.assert_allclose(wt.cpu().numpy(), w, rtol=rtol, atol=atol)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('momentum', [0.0, 0.9])
def test_sgd_matches_mxnet_update_rule(cuda, prec, momentum):
    """mx.optimizer.SGD.update (python/mxnet/optimizer/optimizer.py, the Trainer('sgd') of grad_based_inference.py:67):
    mom = momentum * mom - lr * rescale * g; w += mom   (plain w -= lr * rescale * g without momentum)."""
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(15)
    n = 1000
    w = rng.randn(n)
    mom = np.zeros(n)
    wt, mt = T(w, cuda, tdt), T(mom, cuda, tdt)
    cnt = torch.zeros((1,), dtype=torch.int32, device=cuda)
    for t in range(1, 6):
        g = rng.randn(n)
        mom = momentum * mom - 0.01 * (g / 16)
        w = w + mom
        _raw.sgd_step_(wt, T(g, cuda, tdt), mt if momentum else None, cnt, lr=0.01, momentum=momentum, rescale=1. / 16)
    assert int(cnt.item()) == 5
    np.testing.assert_allclose(wt.cpu().numpy(), w, rtol=rtol, atol=atol)


def test_gather_rows_bit_exact(cuda):
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(14)
    src = rng.randn(1000, 9).astype(np.float32)
    idx = rng.permutation(1000).astype(np.int64)
    off = torch.tensor([128], dtype=torch.int64, device=cuda)
    got = _raw.gather_rows(torch.as_tensor(src, device=cuda), torch.as_tensor(idx, device=cuda), off, 64)
    np.testing.assert_array_equal(got.cpu().numpy(), src[idx[128:192]])


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('n', [1, 3, 64, 65, 128, 130, 300, 513, 1024, 1536, 2048])
def test_potrf_packed_and_pack_contents(cuda, prec, n):
    """GEMM-based potrf (tcgen05 updates in f32): same factor as numpy.linalg.cholesky, MXNet zero-upper convention,
    and a pack whose blocks really are the inverses of the diagonal blocks / the transpose of L."""
    from mxfusion_b200 import _raw, _lib
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(6)
    S = 2
    W = rng.randn(S, n, n)
    A = W @ np.swapaxes(W, -1, -2) + n * np.eye(n)[None]
    want = ol.potrf(A)
    L, info, pack = _raw.potrf_packed_(T(A, cuda, tdt))
    assert info.cpu().tolist() == [0, 0]
    got = L.cpu().numpy()
    assert np.all(np.triu(got, 1) == 0)
    if prec == 'f32':
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * np.sqrt(n))
    else:
        np.testing.assert_allclose(got, want, rtol=rtol * 5, atol=atol * 100)
    NB = _lib.lib().mxf_tri_block(_lib.dtype_code(L))
    nblk = (n + NB - 1) // NB
    pk = pack.cpu().numpy().astype(np.float64)
    dinv = pk[:, :nblk * NB * NB].reshape(S, nblk, NB, NB)
    dinvT = pk[:, nblk * NB * NB:2 * nblk * NB * NB].reshape(S, nblk, NB, NB)
    ldt = (n + 3) & ~3
    LT = pk[:, 2 * nblk * NB * NB:2 * nblk * NB * NB + n * ldt].reshape(S, n, ldt)[:, :, :n]
    np.testing.assert_allclose(LT, np.swapaxes(got, -1, -2), rtol=0, atol=0)
    for b in range(nblk):
        k0, k1 = b * NB, min(n, (b + 1) * NB)
        blk = want[:, k0:k1, k0:k1]
        inv = np.linalg.inv(blk)
        np.testing.assert_allclose(dinv[:, b, :k1 - k0, :k1 - k0], inv, rtol=rtol * 20, atol=atol * 100)
        np.testing.assert_allclose(dinvT[:, b, :k1 - k0, :k1 - k0], np.swapaxes(inv, -1, -2), rtol=rtol * 20,
                                   atol=atol * 100)
    pk2 = _raw.tri_pack(L).cpu().numpy().astype(np.float64)       # pack of an existing factor: same contents
    np.testing.assert_allclose(pk2[:, :2 * nblk * NB * NB], pk[:, :2 * nblk * NB * NB], rtol=rtol * 20, atol=atol * 100)
    LT2 = pk2[:, 2 * nblk * NB * NB:2 * nblk * NB * NB + n * ldt].reshape(S, n, ldt)[:, :, :n]
    np.testing.assert_array_equal(LT2, LT)


def test_potrf_packed_reports_first_bad_pivot(cuda):
    from mxfusion_b200 import _raw
    A = np.eye(300)[None].repeat(2, 0)
    A[1, 170, 170] = -1.0
    L, info, _ = _raw.potrf_packed_(T(A, cuda, torch.float32))
    assert info.cpu().tolist() == [0, 171]


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('transpose', [False, True])
@pytest.mark.parametrize('n,nrhs', [(1, 1), (3, 1), (64, 5), (65, 130), (200, 257), (513, 33), (1024, 1024),
                                    (1024, 4096), (300, 3072)])
def test_trsm_packed(cuda, prec, transpose, n, nrhs):
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(7)
    S = 2
    Lm = np.tril(rng.randn(S, n, n)) * (0.5 / np.sqrt(n)) + 2.0 * np.eye(n)[None]
    B = rng.randn(S, n, nrhs)
    want = ol.trsm(Lm, B, transpose=transpose, alpha=0.5)
    Lt, Bt = T(Lm, cuda, tdt), T(B, cuda, tdt)
    pack = _raw.tri_pack(Lt)
    _raw.trsm_packed_(Lt, pack, Bt, transpose=transpose, alpha=0.5)
    # well-conditioned factor: fp32 must stay at fp32 accuracy (3xTF32 updates, inverted diagonal blocks)
    if prec == 'f32':
        np.testing.assert_allclose(Bt.cpu().numpy(), want, rtol=5e-5, atol=5e-5)
    else:
        np.testing.assert_allclose(Bt.cpu().numpy(), want, rtol=rtol * 5, atol=atol * 100)


def test_trsm_packed_shared_factor(cuda):
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(8)
    Lm = np.tril(rng.randn(1, 150, 150)) * 0.05 + 2.0 * np.eye(150)[None]
    B = rng.randn(3, 150, 40)
    want = ol.trsm(np.repeat(Lm, 3, 0), B)
    Lt, Bt = T(Lm, cuda, torch.float32), T(B, cuda, torch.float32)
    _raw.trsm_packed_(Lt, _raw.tri_pack(Lt), Bt)
    np.testing.assert_allclose(Bt.cpu().numpy(), want, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('transpose', [False, True])
@pytest.mark.parametrize('n,nrhs', [(256, 100), (512, 2048), (768, 260), (1024, 1024), (1024, 5124), (2048, 512), (300, 64)])
def test_trsm_solve_large_blocks(cuda, prec, transpose, n, nrhs):
    """Out-of-place solve through the hierarchically built inverse blocks (256 / 512): a few large GEMMs."""
    from mxfusion_b200 import _raw
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(17)
    S = 2 if n <= 1024 else 1
    Lm = np.tril(rng.randn(S, n, n)) * (0.5 / np.sqrt(n)) + 2.0 * np.eye(n)[None]
    B = rng.randn(S, n, nrhs)
    want = ol.trsm(Lm, B, transpose=transpose)
    Lt, Bt = T(Lm, cuda, tdt), T(B, cuda, tdt)
    X = _raw.trsm_solve(Lt, _raw.tri_pack(Lt), Bt, transpose=transpose)
    if prec == 'f32':
        np.testing.assert_allclose(X.cpu().numpy(), want, rtol=5e-5, atol=5e-5)
    else:
        np.testing.assert_allclose(X.cpu().numpy(), want, rtol=1e-9, atol=1e-10)


def test_trsm_solve_after_potrf_matches_lapack(cuda):
    """potrf_packed's own pack (built during the factorisation) drives the large-block solve."""
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(18)
    n = 1024
    W = rng.randn(1, n, n)
    A = W @ np.swapaxes(W, -1, -2) / n + np.eye(n)[None]
    B = rng.randn(1, n, 300)
    Lw = ol.potrf(A)
    L, info, pack = _raw.potrf_packed_(T(A, cuda, torch.float32))
    X = _raw.trsm_solve(L, pack, T(B, cuda, torch.float32))
    np.testing.assert_allclose(X.cpu().numpy(), ol.trsm(Lw, B), rtol=2e-4, atol=2e-4)
    Xt = _raw.trsm_solve(L, pack, T(B, cuda, torch.float32), transpose=True)
    np.testing.assert_allclose(Xt.cpu().numpy(), ol.trsm(Lw, B, transpose=True), rtol=2e-4, atol=2e-4)


def test_potrf_packed_two_level_large(cuda):
    """n > 1024: two-level blocking (512-wide panels and trailing updates through the inverted 512 blocks); the pack must
    still drive correct solves."""
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(21)
    n = 4096
    W = rng.randn(n, n).astype(np.float32)
    A = (W @ W.T / n + np.eye(n, dtype=np.float32)).astype(np.float64)
    want = np.linalg.cholesky(A)
    L, info, pack = _raw.potrf_packed_(T(A[None], cuda, torch.float32))
    assert info.cpu().tolist() == [0]
    got = L.cpu().numpy()[0]
    assert np.all(np.triu(got, 1) == 0)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5)
    B = rng.randn(1, n, 64)
    X = _raw.trsm_solve(L, pack, T(B, cuda, torch.float32))
    np.testing.assert_allclose(X.cpu().numpy()[0], np.linalg.solve(want, B[0]), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('ard', [False, True])
def test_kbuild_bwd_without_column_gradient(cuda, kind, ard):
    """X2 needs no gradient (the minibatch rows of K(Z, X)): the lengthscale gradient is accumulated directly."""
    from mxfusion_b200 import _raw
    from oracle import torch_ref
    rng = np.random.RandomState(5)
    S, N, N2, D = 2, 70, 300, 8
    X = torch.tensor(rng.uniform(-2, 2, (S, N, D)), requires_grad=True)
    X2 = torch.tensor(rng.uniform(-2, 2, (S, N2, D)))
    ls = torch.tensor(rng.uniform(0.5, 2.0, (S, D if ard else 1)), requires_grad=True)
    var = torch.tensor(rng.uniform(0.5, 2.0, (S, 1)), requires_grad=True)
    G = torch.tensor(rng.randn(S, N, N2))
    (torch_ref.K(kind, X, ls, var, X2) * G).sum().backward()
    dX, dX2, dls, dvar = _raw.kbuild_bwd(kind, X.detach().to(cuda), X2.to(cuda), ls.detach().to(cuda),
                                         var.detach().to(cuda), G.to(cuda), need_dX2=False)
    assert dX2 is None
    np.testing.assert_allclose(dX.cpu().numpy(), X.grad.numpy(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dls.cpu().numpy(), ls.grad.numpy(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dvar.cpu().numpy(), var.grad.numpy(), rtol=1e-7, atol=1e-9)


# ------------------------------------------------------------------------------------------- fused dense-tanh network
MLP_CASES = [
    # S, Sx, B, widths, bias, shared-first-layer
    (3, 1, 4096, (1, 50, 50, 1), True, False),       # BASELINE config 4 (bnn_regression.ipynb: H=50, S=3)
    (1, 1, 100, (3, 7, 2), True, False),
    (2, 2, 130, (5, 64, 64, 64, 3), True, False),    # 4 dense layers at the width limit, x sampled too
    (4, 1, 65, (2, 33, 1), False, False),            # no bias, ragged last CTA
    (3, 1, 200, (4, 16, 2), True, True),             # first layer's weight shared by all samples (stride 0)
]


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('case', MLP_CASES)
def test_mlp_tanh_forward_and_adjoint(cuda, prec, case):
    """csrc/mlp.cu against the per-sample loop of the reference (oracle/mlp.py) and autograd of a float64 torch
    restatement for the weight / bias gradients."""
    from mxfusion_b200 import ops
    from oracle import mlp as omlp
    S, Sx, B, widths, bias, shared0 = case
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(len(widths) * 100 + B)
    x = rng.uniform(-1, 1, (Sx, B, widths[0]))
    Ws = [rng.randn(1 if (shared0 and l == 0) else S, widths[l + 1], widths[l]) / np.sqrt(widths[l]) for l in range(len(widths) - 1)]
    bs = [rng.randn(1 if (shared0 and l == 0) else S, widths[l + 1]) * 0.3 if bias else None for l in range(len(widths) - 1)]
    gout = rng.randn(S, B, widths[-1])
    tW = [T(w, cuda, tdt).requires_grad_() for w in Ws]
    tb = [None if b is None else T(b, cuda, tdt).requires_grad_() for b in bs]
    out = ops.mlp_tanh(T(x, cuda, tdt), tW, tb)
    want = omlp.mlp_tanh(x, Ws, bs)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=rtol, atol=atol * 10)
    (out * T(gout, cuda, tdt)).sum().backward()
    rW = [torch.tensor(w, requires_grad=True) for w in Ws]
    rb = [None if b is None else torch.tensor(b, requires_grad=True) for b in bs]
    h = torch.tensor(x)
    for l in range(len(Ws)):
        h = torch.matmul(h, rW[l].transpose(-1, -2))
        if rb[l] is not None:
            h = h + rb[l].unsqueeze(-2)
        if l + 1 < len(Ws):
            h = torch.tanh(h)
    (h * torch.tensor(gout)).sum().backward()
    gt = 1e-9 if prec == 'f64' else 2e-4
    for l in range(len(Ws)):
        w = rW[l].grad.numpy()
        g = tW[l].grad.double().cpu().numpy()
        assert np.max(np.abs(g - w)) <= gt * (1.0 + np.max(np.abs(w))), ('W', l, np.max(np.abs(g - w)))
        if bias:
            w = rb[l].grad.numpy()
            g = tb[l].grad.double().cpu().numpy()
            assert np.max(np.abs(g - w)) <= gt * (1.0 + np.max(np.abs(w))), ('b', l, np.max(np.abs(g - w)))


def test_mlp_tanh_rejects_wide_layers(cuda):
    from mxfusion_b200 import _raw, _lib
    x = torch.zeros((1, 8, 65), device=cuda)
    W = [torch.zeros((1, 4, 65), device=cuda)]
    with pytest.raises(_lib.MXFusionB200Error):
        _raw.mlp_tanh_fwd(x, W, [None])


def test_reparam_draw_is_fresh_on_every_graph_replay(cuda):
    """The in-kernel Philox stream is keyed by host scalars, which a CUDA graph freezes at capture time; the device step
    counter (bumped by the fused Adam kernel) is mixed into the counter at run time so that every replayed optimiser
    step draws new noise, while a replay at the same step reproduces the draw."""
    from mxfusion_b200 import _raw
    m = torch.zeros((1, 4096), device=cuda)
    v = torch.ones((1, 4096), device=cuda)
    ctr = torch.zeros((1,), dtype=torch.int32, device=cuda)
    _raw.normal_reparam(m, v, 3, seed=7, offset=5, step_counter=ctr)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        w = _raw.normal_reparam(m, v, 3, seed=7, offset=5, step_counter=ctr)
    outs = []
    for step in (0, 1, 1, 2):
        ctr.fill_(step)
        g.replay()
        outs.append(w.clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[3])
    assert torch.equal(outs[1], outs[2])
    eager = _raw.normal_reparam(m, v, 3, seed=7, offset=5, step_counter=ctr)
    assert torch.equal(eager, outs[3])
    z = torch.cat([o.flatten() for o in outs])
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    # without a counter the stream is the plain (seed, offset) one
    a = _raw.normal_reparam(m, v, 3, seed=7, offset=5)
    ctr.zero_()
    b = _raw.normal_reparam(m, v, 3, seed=7, offset=5, step_counter=ctr)
    assert torch.equal(a, b)


@pytest.mark.parametrize('shape', [(1000, 1, 5000, 5000, False), (1024, 1, 28392, 28392, False), (130, 3, 4099, 4100, False),
                                   (64, 8, 8192, 8192, True), (5, 4, 4096, 4096, False)])
def test_gemm_skinny_long_k(cuda, shape):
    """The long-k matrix-vector kernel (b += L^-1 K(Z, X_c) Y_c of the streamed sparse-GP statistics) against float64,
    including a k % 4 tail on a padded leading dimension and beta accumulation."""
    from mxfusion_b200 import _raw
    m, n, k, lda, tb = shape
    rng = np.random.RandomState(k + n)
    Abuf = rng.randn(1, m, lda).astype(np.float32)
    Bm = (rng.randn(1, n, k) if tb else rng.randn(1, k, n)).astype(np.float32)
    C0 = rng.randn(1, m, n).astype(np.float32)
    A_t = torch.as_tensor(Abuf, device=cuda)[:, :, :k]
    C = torch.as_tensor(C0.copy(), device=cuda)
    _raw.gemm(A_t, torch.as_tensor(Bm, device=cuda), False, tb, alpha=0.5, beta=1.0, C=C)
    A64, B64 = Abuf[:, :, :k].astype(np.float64), Bm.astype(np.float64)
    Bop = np.swapaxes(B64, -1, -2) if tb else B64
    want = 0.5 * (A64 @ Bop) + C0
    bound = 2e-6 * (np.abs(A64) @ np.abs(Bop) + np.abs(C0)) + 1e-6
    assert np.all(np.abs(C.cpu().numpy() - want) <= bound)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('shared', [(True, True), (False, True), (False, False)])
def test_reparam_adjoint(cuda, prec, shared):
    """mxf_normal_reparam_bwd against autograd of eps * sqrt(v) + m (normal.py:89-92), with operands shared by the samples
    (gradient summed over S) or sampled."""
    from mxfusion_b200 import ops
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(4)
    S, shape = 3, (37, 5)
    m = rng.randn(1 if shared[0] else S, *shape)
    v = rng.rand(1 if shared[1] else S, *shape) + 0.2
    eps, gw = rng.randn(S, *shape), rng.randn(S, *shape)
    tm, tv = T(m, cuda, tdt).requires_grad_(), T(v, cuda, tdt).requires_grad_()
    w = ops.normal_draw(tm, tv, S, eps=T(eps, cuda, tdt))
    (w * T(gw, cuda, tdt)).sum().backward()
    rm, rv = torch.tensor(m, requires_grad=True), torch.tensor(v, requires_grad=True)
    ((torch.tensor(eps) * torch.sqrt(rv) + rm) * torch.tensor(gw)).sum().backward()
    np.testing.assert_allclose(tm.grad.cpu().numpy(), rm.grad.numpy(), rtol=rtol * 10, atol=atol * 10)
    np.testing.assert_allclose(tv.grad.cpu().numpy(), rv.grad.numpy(), rtol=rtol * 10, atol=atol * 10)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_normal_logpdf_multi(cuda, prec):
    """The multi-tensor Normal log-density (all Normal factors of a graph walk in one launch each way) against the oracle,
    entry by entry: sampled / shared operands, one-element (constant prior) operands, different sample counts and scales."""
    from mxfusion_b200 import ops
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(9)
    spec = [  # (S_x, S_m, S_v, shape, scalar_m, scalar_v, scale)
        (3, 1, 1, (50, 1), False, False, 1.0), (3, 1, 1, (50, 50), True, True, 1.0), (1, 3, 1, (64, 1), False, True, 24.4),
        (3, 3, 3, (7,), False, False, 0.5), (1, 1, 1, (5, 2), False, False, 2.0)]
    spec = spec * 4          # 20 entries: more than one table (16 per launch)
    entries, ref, leaves = [], 0.0, []
    for Sx, Sm, Sv, shape, scm, scv, scale in spec:
        x = rng.randn(Sx, *shape)
        m = rng.randn(Sm, *((1,) * len(shape) if scm else shape))
        v = rng.rand(Sv, *((1,) * len(shape) if scv else shape)) + 0.3
        S = max(Sx, Sm, Sv)
        lp = on.log_pdf(np.broadcast_to(m, (S,) + shape) if not scm else m, v, np.broadcast_to(x, (S,) + shape))
        ref += scale * np.sum(np.mean(np.broadcast_to(lp, (S,) + shape), axis=0))
        tx = T(x, cuda, tdt).requires_grad_()
        tm = T(m, cuda, tdt).requires_grad_(not scm)
        tv = T(v, cuda, tdt).requires_grad_(not scv)
        entries.append((tx, tm, tv, scale))
        leaves.append((x, m, v, scale, S, shape, scm, scv))
    got = ops.normal_log_pdf_sum_multi(entries)
    np.testing.assert_allclose(float(got), ref, rtol=rtol * 5)
    (got * 1.7).sum().backward()
    for (tx, tm, tv, _), (x, m, v, scale, S, shape, scm, scv) in zip(entries, leaves):
        rx, rm, rv = torch.tensor(x, requires_grad=True), torch.tensor(m, requires_grad=True), torch.tensor(v, requires_grad=True)
        lp = -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(rv) - (rx - rm) ** 2 / (2 * rv)
        (1.7 * scale * lp.expand((S,) + shape).sum() / S).backward()
        pairs = [(tx, rx)] + ([] if scm else [(tm, rm)]) + ([] if scv else [(tv, rv)])
        for t, r in pairs:
            np.testing.assert_allclose(t.grad.cpu().numpy(), r.grad.numpy(), rtol=rtol * 20, atol=atol * 20)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_normal_draw_multi(cuda, prec):
    """Batched reparameterised draws (one launch for all entries, one for their adjoints): moments of the noise, exact
    adjoints given the realised noise, independent streams per entry, and a different draw per optimiser step."""
    from mxfusion_b200 import ops
    tdt, ndt, rtol, atol = DT[prec]
    rng = np.random.RandomState(2)
    shapes = [(50, 1), (50,), (50, 50), (1, 50), (1,), (333, 7)] * 3        # 18 entries: two tables
    S = 4
    ms = [T(rng.randn(1, *sh), cuda, tdt).requires_grad_() for sh in shapes]
    vs = [T(rng.rand(1, *sh) + 0.3, cuda, tdt).requires_grad_() for sh in shapes]
    ctr = torch.zeros((1,), dtype=torch.int32, device=cuda)
    ws = ops.normal_draw_multi([(m, v, S) for m, v in zip(ms, vs)], seed=3, offsets=list(range(1, len(shapes) + 1)),
                               step_counter=ctr)
    gws = [T(rng.randn(S, *sh), cuda, tdt) for sh in shapes]
    sum((w * g).sum() for w, g in zip(ws, gws)).backward()
    zs = []
    for m, v, w, g, sh in zip(ms, vs, ws, gws, shapes):
        assert tuple(w.shape) == (S,) + sh
        eps = ((w - m) / torch.sqrt(v)).detach()
        zs.append(eps.flatten())
        np.testing.assert_allclose(m.grad.cpu().numpy(), g.sum(0, keepdim=True).cpu().numpy(), rtol=rtol * 10, atol=atol * 10)
        want = (g * eps * 0.5 / torch.sqrt(v.detach())).sum(0, keepdim=True)
        np.testing.assert_allclose(v.grad.cpu().numpy(), want.cpu().numpy(), rtol=max(rtol * 100, 1e-6), atol=atol * 100)
    z = torch.cat(zs).double()
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    # entries 0 and 6 have the same shape but different offsets: different noise
    assert not torch.allclose(zs[0], zs[6])
    ctr.fill_(1)
    ws2 = ops.normal_draw_multi([(m.detach(), v.detach(), S) for m, v in zip(ms, vs)], seed=3,
                                offsets=list(range(1, len(shapes) + 1)), step_counter=ctr)
    assert not torch.allclose(ws2[2], ws[2].detach())
    ctr.fill_(0)
    ws3 = ops.normal_draw_multi([(m.detach(), v.detach(), S) for m, v in zip(ms, vs)], seed=3,
                                offsets=list(range(1, len(shapes) + 1)), step_counter=ctr)
    assert torch.equal(ws3[2], ws[2].detach())


@pytest.mark.parametrize('n', [1, 5, 64, 100, 128, 192, 257, 512, 1000, 1024])
def test_potrf_dataflow_factor_and_inverse(cuda, n):
    """f32, n <= 1024: the single-launch tile-dataflow kernel (csrc/chol_dag.cu) returns the factor AND its explicit
    inverse W = L^-1 / W^T in the pack (the operands that make every later trsm one GEMM); an ill-conditioned RBF Gram
    matrix (the bench's Kuu with jitter 1e-6 has cond ~1e6) keeps the LAPACK-class backward error."""
    from mxfusion_b200 import _raw, _lib
    rng = np.random.RandomState(23)
    S = 3
    Z = rng.uniform(-3, 3, (S, n, 4))
    r2 = ((Z[:, :, None, :] - Z[:, None, :, :]) ** 2).sum(-1)
    A = np.exp(-0.5 * r2) + 1e-4 * np.eye(n)[None]
    At = T(A, cuda, torch.float32)
    A32 = At.cpu().numpy().astype(np.float64)
    L, info, pack = _raw.potrf_packed_(At.clone())
    assert info.cpu().tolist() == [0] * S
    got = L.cpu().numpy().astype(np.float64)
    assert np.all(np.triu(got, 1) == 0)
    # backward error of the factorisation (what LAPACK guarantees): |L L^T - A| <= c n eps |L||L^T|
    resid = np.abs(got @ np.swapaxes(got, -1, -2) - A32)
    bound = np.abs(got) @ np.abs(np.swapaxes(got, -1, -2))
    assert np.max(resid) <= 4 * 6e-8 * max(8, np.sqrt(n)) * np.max(bound), np.max(resid) / np.max(bound)
    NB = 128
    nblk, ldt = (n + NB - 1) // NB, (n + 3) & ~3
    top = _lib.lib().mxf_tri_top_block(_lib.dtype_code(L), n)
    assert top == max(4, ldt)
    off = (2 * nblk * NB * NB + n * ldt + 3) & ~3
    pk = pack.cpu().numpy().astype(np.float64)
    W = pk[:, off:off + top * top].reshape(S, top, top)[:, :n, :n]
    WT = pk[:, off + top * top:off + 2 * top * top].reshape(S, top, top)[:, :n, :n]
    np.testing.assert_array_equal(WT, np.swapaxes(W, -1, -2))
    assert np.all(np.triu(W, 1) == 0)
    # W is the inverse of the computed factor: |W L - I| small relative to |W||L|
    eye = np.eye(n)[None]
    res = np.abs(W @ got - eye)
    bnd = np.abs(W) @ np.abs(got)
    assert np.max(res) <= 4 * 6e-8 * max(8, n ** 0.5) * np.max(bnd), np.max(res) / np.max(bnd)
    # the inverse-only mode (pack of an existing factor) gives the same blocks
    pk2 = _raw.tri_pack(L).cpu().numpy().astype(np.float64)
    W2 = pk2[:, off:off + top * top].reshape(S, top, top)[:, :n, :n]
    np.testing.assert_allclose(W2, W, rtol=1e-4, atol=1e-5 * np.max(np.abs(W)))


def test_potrf_dataflow_two_concurrent_streams_and_graph(cuda):
    """Two factorisations on two streams inside one captured CUDA graph (the SVGP step runs chol(Kuu) and chol(S) this way):
    the ticket order makes the schedule independent of how many CTAs of each launch are resident."""
    from mxfusion_b200 import _raw
    rng = np.random.RandomState(24)
    n = 1024
    Ws = rng.randn(2, n, n) / np.sqrt(n)
    A = Ws @ np.swapaxes(Ws, -1, -2) + 0.1 * np.eye(n)[None]
    want = np.linalg.cholesky(A)
    A0, A1 = T(A[:1], cuda, torch.float32), T(A[1:], cuda, torch.float32)
    B0, B1 = A0.clone(), A1.clone()
    p0, p1 = _raw.new_pack(B0), _raw.new_pack(B1)
    i0 = torch.zeros((1,), dtype=torch.int32, device=cuda)
    i1 = torch.zeros((1,), dtype=torch.int32, device=cuda)
    side, main = torch.cuda.Stream(device=cuda), torch.cuda.Stream(device=cuda)

    def body():
        B0.copy_(A0)
        B1.copy_(A1)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            _raw.potrf_packed_(B1, i1, p1)
        _raw.potrf_packed_(B0, i0, p0)
        cur.wait_stream(side)
    main.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(main):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=main):
        body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert i0.item() == 0 and i1.item() == 0
    np.testing.assert_allclose(B0.cpu().numpy()[0], want[0], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(B1.cpu().numpy()[0], want[1], rtol=2e-4, atol=2e-5)
