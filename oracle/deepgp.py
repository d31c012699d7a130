"""TEST INFRASTRUCTURE (oracle): an INDEPENDENT dense formulation of the two-/L-layer deep-GP bound with doubly-stochastic
variational inference (Salimbeni & Deisenroth 2017), in NumPy and in differentiable torch float64.

The reference (amzn/MXFusion) has NO deep GP -- parity of this path is "unpinned by the reference".  What pins it:
  * with ONE layer the bound must equal SVGPRegressionLogPdf's (svgp_regression.py:43-109) on the reference's own test
    fixture (ELBO = -32.72563540745786), and
  * this file computes every layer through explicit inverses and log-determinants (no Cholesky, no whitening):
        mean = Kfu Kuu^-1 m,   var = kff - diag(Kfu Kuu^-1 Kuf) + diag(Kfu Kuu^-1 S Kuu^-1 Kuf),
        KL   = 1/2 [ P (tr(Kuu^-1 S) - M + log|Kuu| - log|S|) + tr(m^T Kuu^-1 m) ]
    which shares no code path with mxfusion_b200/modules/gp_modules/deep_gp.py.
Only tests/ and bench.py's cpu_baseline leg may import this module."""
import math

import numpy as np
import torch

from . import kernels as ok


def _layer_np(kind, H, Z, ls, var, m, S, jitter):
    M, P = Z.shape[0], m.shape[1]
    Kuu = ok.K(kind, Z[None], ls[None], var[None])[0] + jitter * np.eye(M)
    Kuf = ok.K(kind, Z[None], ls[None], var[None], H[None])[0]
    kdiag = np.full((H.shape[0],), var[0])
    Kinv = np.linalg.inv(Kuu)
    A = Kinv @ Kuf
    mean = A.T @ m
    v = kdiag - np.sum(Kuf * A, axis=0) + np.sum(A * (S @ A), axis=0)
    kl = 0.5 * (P * (np.trace(Kinv @ S) - M + np.linalg.slogdet(Kuu)[1] - np.linalg.slogdet(S)[1]) + np.sum(m * (Kinv @ m)))
    return mean, v, kl


def dgp_elbo_np(kinds, X, Y, Zs, lss, variances, ms, Ws, ds, noise_var, eps, jitter=0.0, scale=1.0):
    """kinds[l], Zs[l] (M, D_l), lss[l] (1|D_l,), variances[l] (1,), ms[l] (M, D_{l+1}), Ws[l] (M, M), ds[l] (M,) (already
    positive), eps[l] (S, B, D_{l+1}) for the hidden layers.  Returns the per-sample bound (S,)."""
    nl = len(kinds)
    S_mc = eps[0].shape[0] if nl > 1 else 1
    out = []
    for s in range(S_mc):
        h, kls = X, 0.0
        for l in range(nl):
            Sl = Ws[l] @ Ws[l].T + np.diag(ds[l])
            mean, v, kl = _layer_np(kinds[l], h, Zs[l], lss[l], variances[l], ms[l], Sl, jitter)
            kls += kl
            if l < nl - 1:
                if mean.shape[1] == h.shape[1]:
                    mean = mean + h
                h = mean + np.sqrt(np.maximum(v, 0.0))[:, None] * eps[l][s]
        B, P = Y.shape
        data = -0.5 * B * P * (math.log(2 * math.pi) + math.log(noise_var)) - np.sum((Y - mean) ** 2) / (2 * noise_var) \
            - P * np.sum(v) / (2 * noise_var)
        out.append(scale * data - kls)
    return np.array(out)


def _K_t(kind, A, B, ls, var):
    from . import torch_ref
    return torch_ref.K(kind, A[None], ls[None], var[None], None if B is None else B[None])[0]


def dgp_elbo_torch(kinds, X, Y, Zs, lss, variances, ms, Ws, ds, noise_var, eps, jitter=0.0, scale=1.0):
    """Same dense formulation, differentiable (float64 tensors): the gradient oracle.  Returns the MC average (scalar)."""
    nl = len(kinds)
    S_mc = eps[0].shape[0] if nl > 1 else 1
    tot = 0.0
    for s in range(S_mc):
        h, kls = X, 0.0
        for l in range(nl):
            M, P = Zs[l].shape[0], ms[l].shape[1]
            Sl = Ws[l] @ Ws[l].T + torch.diag(ds[l])
            Kuu = _K_t(kinds[l], Zs[l], None, lss[l], variances[l]) + jitter * torch.eye(M, dtype=X.dtype)
            Kuf = _K_t(kinds[l], Zs[l], h, lss[l], variances[l])
            Kinv = torch.linalg.inv(Kuu)
            A = Kinv @ Kuf
            mean = A.T @ ms[l]
            v = variances[l][0] - torch.sum(Kuf * A, dim=0) + torch.sum(A * (Sl @ A), dim=0)
            kls = kls + 0.5 * (P * (torch.trace(Kinv @ Sl) - M + torch.logdet(Kuu) - torch.logdet(Sl)) +
                               torch.sum(ms[l] * (Kinv @ ms[l])))
            if l < nl - 1:
                if mean.shape[1] == h.shape[1]:
                    mean = mean + h
                h = mean + torch.sqrt(torch.clamp(v, min=0.0))[:, None] * eps[l][s]
        B, P = Y.shape
        data = -0.5 * B * P * (math.log(2 * math.pi) + torch.log(noise_var)) - torch.sum((Y - mean) ** 2) / (2 * noise_var) \
            - P * torch.sum(v) / (2 * noise_var)
        tot = tot + scale * data - kls
    return tot / S_mc
