"""Normal log-density, reparameterised draw and the factor-graph reduction,
restated in NumPy (test infrastructure).

``mxfusion/components/distributions/normal.py:52-92``,
``mxfusion/models/factor_graph.py:223-224`` (sum(mean(logpdf, axis=0))),
``mxfusion/inference/variational.py:103-108``.
"""
import numpy as np


def log_pdf(mean, variance, random_variable, log_pdf_scaling=1.0):
    """normal.py:67-70 -- elementwise, broadcast over the sample axis."""
    logvar = np.log(2 * np.pi) / -2 + np.log(variance) / -2
    return (logvar + np.square(random_variable - mean) / (-2 * variance)) * log_pdf_scaling


def draw_samples(mean, variance, eps):
    """normal.py:89-92 with the standard-normal draw ``eps`` injected, as the
    reference tests do through MockMXNetRandomGenerator
    (``mxfusion/util/testutils.py:58-93``)."""
    return eps * np.sqrt(variance) + mean


def factor_reduce(log_pdf_values):
    """factor_graph.py:223-224: F.sum(expectation(.)) = sum(mean(., axis=0))."""
    return np.sum(np.mean(log_pdf_values, axis=0))


def meanfield_elbo_terms(w_mean, w_var, eps, prior_mean, prior_var):
    """One mean-field weight tensor of the MC-ELBO (variational.py:103-108):
    draw w = eps*sqrt(v)+mu, return (w, sum-mean log p(w), sum-mean log q(w))."""
    w = draw_samples(w_mean[None], w_var[None], eps)
    lp = factor_reduce(log_pdf(prior_mean[None], prior_var[None], w))
    lq = factor_reduce(log_pdf(w_mean[None], w_var[None], w))
    return w, lp, lq
