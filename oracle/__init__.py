"""CPU oracle for the MXFusion VI/GP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy/SciPy (float64 unless a dtype is passed), the
arithmetic that amzn/MXFusion performs through MXNet operators on the
per-iteration ELBO path.  Every function cites the reference file:line it
follows.  Nothing under ``mxfusion_b200/`` may import it: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do, and only
as the checker / reported baseline -- never as the product path.

Parity pinning
--------------
* ``import mxnet`` is impossible in this image, and the
  arithmetic of the reference lives in MXNet (``mxnet>=1.3``, unpinned,
  ``requirements/requirements.txt:1``), which is absent from /root/reference.
  MXNet operator semantics are restated from the MXNet 1.x operator docs:
  ``potrf`` -> lower L, ``trsm(A,B)`` -> A^-1 B, ``syrk(A)`` -> A A^T,
  ``gemm2(A,B,ta,tb)``, ``sumlogdiag``, Adam with ``rescale_grad``.
* The oracle is pinned three ways (tests/test_oracle.py):
  1. known answers on the reference's own seed-0 fixtures
     (``testing/modules/svgpregression_test.py:41-56``,
     ``testing/modules/gpregression_test.py:40-48``, the GP notebook);
  2. an independent formulation of every quantity (dense Hensman bound,
     ``scipy.stats.multivariate_normal``, ``scipy.stats.norm``), replacing the
     GPy asserts of the reference tests (GPy is absent too);
  3. fixtures under ``tests/golden/*.npz`` produced by running the reference's
     *own Python source* from /root/reference on top of a stand-in for the
     ``mxnet`` module (``tests/golden/make_golden.py`` +
     ``tests/golden/_mxnet_standin``): kernels, SVGP / GP values AND gradients,
     the minibatch loop's trajectory, Normal, mean-field SVI.  The stand-in is
     itself validated by reproducing the loss the reference's GP notebook
     prints after 100 Adam steps (-16.903135 vs -16.903127 here).
"""
from . import kernels, linalg, transforms, normal, gp, svgp, loop  # noqa: F401
