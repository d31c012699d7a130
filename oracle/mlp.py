"""Function evaluation of a Dense/tanh Gluon block with sampled weights, restated in NumPy (test infrastructure).

``mxfusion/components/functions/function_evaluation.py:72-96``: for a non-broadcastable function the reference loops
over the S samples, slices every input / weight at sample i (or 0 when it is shared), runs the block
(``mxfusion_gluon_function.py:97-111`` with the sampled weights injected, ``:166-194``) and concatenates the outputs
on the sample axis.  The block of the BNN notebooks (``bnn_regression.ipynb`` cell 6) is
``Dense(H, tanh) -> Dense(H, tanh) -> Dense(1)``; MXNet ``Dense`` computes ``x W^T + b`` with W of shape (out, in).

Only tests/ may import this.
"""
import numpy as np


def mlp_tanh(x, weights, biases):
    """x (S|1,B,in); weights[l] (S|1,out,in); biases[l] (S|1,out) or None -> (S,B,out_L), the per-sample loop."""
    S = max([x.shape[0]] + [w.shape[0] for w in weights])
    outs = []
    for i in range(S):                                              # function_evaluation.py:80-93
        h = x[i if x.shape[0] > 1 else 0]
        for l, (W, b) in enumerate(zip(weights, biases)):
            Wi = W[i if W.shape[0] > 1 else 0]
            h = h @ Wi.T
            if b is not None:
                h = h + b[i if b.shape[0] > 1 else 0]
            if l + 1 < len(weights):
                h = np.tanh(h)
        outs.append(h[None])
    return np.concatenate(outs, axis=0)                             # :94-96
