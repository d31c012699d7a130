"""Reference-restated CPU path in PyTorch (test infrastructure + CPU baseline).

The same operations, in the same order, as ``oracle/svgp.py`` / ``oracle/gp.py``
/ ``oracle/kernels.py`` (which follow ``svgp_regression.py:43-109``,
``gp_regression.py:55-70``, ``stationary.py:90-107``, ``rbf.py:71-72``,
``matern.py:84-88``), issued as LAPACK/BLAS calls through torch on the CPU:
``torch.linalg.cholesky`` -> potrf, ``solve_triangular`` -> trsm, ``matmul``
-> gemm2/syrk.  Because it is differentiable it is (a) the gradient oracle for
the hand-written CUDA backward kernels (the reference never tests gradients,
SURVEY.md section 4) and (b) the "reference-restated CPU" baseline that
``bench.py`` times: forward + autograd backward + Adam with
gradients divided by the batch size (``minibatch_loop.py:90-91``).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import math
import torch

RBF, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3


def r2(X, lengthscale, X2=None):
    ls = lengthscale.unsqueeze(-2)
    if X2 is None:
        xsc = X / ls
        amat = torch.matmul(xsc, xsc.transpose(-1, -2)) * -2
        dg = torch.sum(torch.square(xsc), dim=-1)
        amat = amat + dg.unsqueeze(-1)
        amat = amat + dg.unsqueeze(-2)
    else:
        x1 = X / ls
        x2 = X2 / ls
        amat = torch.matmul(x1, x2.transpose(-1, -2)) * -2
        amat = amat + torch.sum(torch.square(x1), dim=-1, keepdim=True)
        amat = amat + torch.sum(torch.square(x2), dim=-1).unsqueeze(-2)
    return amat


def K(kind, X, lengthscale, variance, X2=None):
    R2 = r2(X, lengthscale, X2)
    var = variance.unsqueeze(-1)
    if kind == RBF:
        return torch.exp(R2 / -2) * var
    R = torch.sqrt(torch.clamp(R2, min=1e-14))
    if kind == MATERN52:
        return (1 + math.sqrt(5) * R + 5 / 3. * R2) * torch.exp(-math.sqrt(5) * R) * var
    if kind == MATERN32:
        return (1 + math.sqrt(3) * R) * torch.exp(-math.sqrt(3) * R) * var
    return torch.exp(-R) * var


def softplus(x):
    return torch.nn.functional.softplus(x)


def sumlogdiag(A):
    return torch.sum(torch.log(torch.diagonal(A, dim1=-2, dim2=-1)), dim=-1)


def trsm(L, B, transpose=False):
    if transpose:
        return torch.linalg.solve_triangular(L.transpose(-1, -2), B, upper=True)
    return torch.linalg.solve_triangular(L, B, upper=False)


def svgp_log_pdf(kind, X, Y, Z, noise_var, mu, S_W, S_diag, lengthscale, variance,
                 jitter=0.0, log_pdf_scaling=1.0):
    """svgp_regression.py:61-109, homoscedastic noise (S,1), no mean."""
    D = Y.shape[-1]
    M = Z.shape[-2]
    noise_var = noise_var.unsqueeze(-2)
    beta_sum = D * torch.sum(1 / noise_var, dim=-1)
    eye = torch.eye(M, dtype=Z.dtype).unsqueeze(0)
    Kuu = K(kind, Z, lengthscale, variance)
    if jitter > 0.:
        Kuu = Kuu + eye * jitter
    Kuf = K(kind, Z, lengthscale, variance, X)
    Kff_diag = torch.zeros(X.shape[:-1], dtype=X.dtype) + variance
    S = torch.matmul(S_W, S_W.transpose(-1, -2)) + torch.diag_embed(S_diag)
    psi1Y = torch.matmul(Kuf, Y / noise_var)
    L = torch.linalg.cholesky(Kuu)
    Ls = torch.linalg.cholesky(S)
    LinvLs = trsm(L, Ls)
    Linvmu = trsm(L, mu)
    LinvKuf = trsm(L, Kuf)
    KfuKuuInvmu = torch.matmul(LinvKuf.transpose(-1, -2), Linvmu)
    KfuKuuInvLs = torch.matmul(LinvKuf.transpose(-1, -2), LinvLs)
    LinvKufY = trsm(L, psi1Y)
    KL_u = (M / 2. + sumlogdiag(Ls)) * D - sumlogdiag(L) * D \
        - torch.sum(torch.square(LinvLs), dim=(-1, -2)) / 2. * D \
        - torch.sum(torch.square(Linvmu), dim=(-1, -2)) / 2.
    logL = -torch.sum(torch.square(Y) / noise_var + math.log(2. * math.pi) +
                      torch.log(noise_var), dim=(-1, -2)) / 2.
    logL = logL - torch.sum(Kff_diag * beta_sum, dim=-1) / 2.
    logL = logL - torch.sum(torch.square(KfuKuuInvmu) / noise_var, dim=(-1, -2)) / 2.
    logL = logL - torch.sum(torch.square(KfuKuuInvLs) * beta_sum.unsqueeze(-1), dim=(-1, -2)) / 2.
    logL = logL + torch.sum(torch.square(LinvKuf) * beta_sum.unsqueeze(-2), dim=(-1, -2)) / 2.
    logL = logL + torch.sum(Linvmu * LinvKufY, dim=(-1, -2))
    return log_pdf_scaling * logL + KL_u


def gp_log_pdf(kind, X, Y, noise_var, lengthscale, variance, jitter=0.0):
    """gp_regression.py:55-70."""
    N = X.shape[-2]
    D = Y.shape[-1]
    eye = torch.eye(N, dtype=X.dtype).unsqueeze(0)
    Kxx = K(kind, X, lengthscale, variance) + eye * noise_var.unsqueeze(-2)
    if jitter > 0.:
        Kxx = Kxx + eye * jitter
    L = torch.linalg.cholesky(Kxx)
    LinvY = trsm(L, Y)
    logdet_l = sumlogdiag(torch.abs(L))
    tmp = torch.sum((torch.square(LinvY) + math.log(2. * math.pi)).reshape(Y.shape[0], -1), dim=-1)
    return -logdet_l * D - tmp / 2


def sparsegp_log_pdf(kind, X, Y, Z, noise_var, lengthscale, variance, jitter=0.0):
    """sparsegp_regression.py:62-100 (collapsed bound; log_pdf_scaling is not applied by the reference)."""
    D = Y.shape[-1]
    M = Z.shape[-2]
    noise_var_m = noise_var.unsqueeze(-2)
    eye = torch.eye(M, dtype=Z.dtype).unsqueeze(0)
    Kuu = K(kind, Z, lengthscale, variance)
    if jitter > 0.:
        Kuu = Kuu + eye * jitter
    Kuf = K(kind, Z, lengthscale, variance, X)
    Kff_diag = torch.zeros(X.shape[:-1], dtype=X.dtype) + variance
    L = torch.linalg.cholesky(Kuu)
    LinvKuf = trsm(L, Kuf)
    A = eye + torch.matmul(LinvKuf, LinvKuf.transpose(-1, -2)) / noise_var_m
    LA = torch.linalg.cholesky(A)
    LAInvLinvKufY = trsm(LA, torch.matmul(LinvKuf, Y))
    logL = -D * sumlogdiag(LA)
    logL = logL - torch.sum(torch.square(Y) / noise_var_m + math.log(2. * math.pi) + torch.log(noise_var_m),
                            dim=(-1, -2)) / 2
    logL = logL + torch.sum(torch.square(LAInvLinvKufY) / (2 * torch.square(noise_var_m)), dim=(-1, -2))
    logL = logL - D * torch.sum(Kff_diag / (2 * noise_var), dim=-1)
    logL = logL + D * torch.sum(torch.square(LinvKuf) / (2. * noise_var_m), dim=(-1, -2))
    return logL


def normal_log_pdf(mean, variance, rv, scaling=1.0):
    """normal.py:67-69."""
    logvar = math.log(2 * math.pi) / -2 + torch.log(variance) / -2
    return (logvar + torch.square(rv - mean) / (-2 * variance)) * scaling


class AdamMX(object):
    """mx.optimizer.Adam as driven by gluon.Trainer.step(batch_size)
    (minibatch_loop.py:71-74, 90-91); see oracle/loop.py:adam_step."""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    @torch.no_grad()
    def step(self, batch_size=1):
        self.t += 1
        lr_t = self.lr * math.sqrt(1. - self.b2 ** self.t) / (1. - self.b1 ** self.t)
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad / batch_size
            m.mul_(self.b1).add_(g, alpha=1. - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1. - self.b2)
            p.sub_(lr_t * m / (v.sqrt() + self.eps))
            p.grad = None


class SVGPStepCPU(object):
    """One reference iteration (minibatch_loop.py:81-92) of the SVGP MAP
    objective on the CPU: softplus transforms (inference_alg.py:79-80), ELBO,
    backward, Adam.  Parameters are stored unconstrained as the reference does
    (inference_parameters.py:163-170)."""

    def __init__(self, kind, Z, noise_var, lengthscale, variance, qU_mean, qU_cov_W,
                 qU_cov_diag, jitter, scaling, lr, dtype=torch.float32):
        inv = lambda y: torch.log(torch.expm1(torch.as_tensor(y, dtype=dtype)))
        t = lambda a: torch.as_tensor(a, dtype=dtype).clone()
        self.kind, self.jitter, self.scaling = kind, jitter, scaling
        self.Z = t(Z).requires_grad_()
        self.noise_u = inv(noise_var).requires_grad_()
        self.ls_u = inv(lengthscale).requires_grad_()
        self.var_u = inv(variance).requires_grad_()
        self.mu = t(qU_mean).requires_grad_()
        self.W = t(qU_cov_W).requires_grad_()
        self.d_u = inv(qU_cov_diag).requires_grad_()
        self.params = [self.Z, self.noise_u, self.ls_u, self.var_u, self.mu, self.W, self.d_u]
        self.opt = AdamMX(self.params, lr)

    def loss(self, Xb, Yb):
        un = lambda a: a.unsqueeze(0)
        logL = svgp_log_pdf(self.kind, un(Xb), un(Yb), un(self.Z), un(softplus(self.noise_u)),
                            un(self.mu), un(self.W), un(softplus(self.d_u)),
                            un(softplus(self.ls_u)), un(softplus(self.var_u)),
                            jitter=self.jitter, log_pdf_scaling=self.scaling)
        return -torch.sum(torch.mean(logL, dim=0))   # factor_graph.py:223, map.py:83-84

    def step(self, Xb, Yb, batch_size):
        loss = self.loss(Xb, Yb)
        loss.backward()
        self.opt.step(batch_size)
        return float(loss)     # loss.asscalar(), minibatch_loop.py:92
