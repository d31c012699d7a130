"""ctypes binding of libmxf_b200.so (the C ABI declared in include/mxf_b200.h).

There is no fallback: if the shared library is missing or a tensor is not on a
CUDA device the call raises.  PyTorch is used only for device memory and the
current stream.
"""
import ctypes
import os
from ctypes import c_int, c_int64, c_longlong, c_uint64, c_double, c_void_p, c_size_t

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmxf_b200.so')

F32, F64 = 0, 1
_DTYPES = {torch.float32: F32, torch.float64: F64}

_lib = None


class MXFusionB200Error(RuntimeError):
    pass


def _declare(lib):
    i, l, d, p, u, z = c_int, c_int64, c_double, c_void_p, c_uint64, c_size_t
    sig = {
        'mxf_version': (c_int, []),
        'mxf_launch_count': (u, []),
        'mxf_kbuild_fwd': (i, [i, i, p, p, p, i, p, p, d, p, l, i, i, i, i, l, l, l, l, l, l, p]),
        'mxf_kbuild_tc_threshold': (c_longlong, [c_longlong]),
        'mxf_kbuild_bwd_workspace_bytes': (z, [i, i, i, i, i]),
        'mxf_potrf_dag_ctas': (i, [i]),
        'mxf_debug_set_prof': (i, [p]),
        'mxf_debug_set_dag_prof': (i, [p]),
        'mxf_allreduce_p2p_flag_bytes': (z, []),
        'mxf_allreduce_p2p': (i, [i, p, i, i, l, d, z, d, p, p]),
        'mxf_kbuild_bwd': (i, [i, i, p, p, p, i, p, p, l, p, p, p, p, i, i, i, i, l, l, l, l, l, p, z, p]),
        'mxf_gemm': (i, [i, i, i, i, i, i, d, p, l, l, p, l, l, d, p, l, l, i, i, p]),
        'mxf_potrf': (i, [i, p, l, l, i, i, p, p]),
        'mxf_trsm': (i, [i, i, i, i, d, p, l, l, p, l, l, i, p]),
        'mxf_tri_block': (i, [i]),
        'mxf_tri_pack_elems': (z, [i, i]),
        'mxf_tri_pack': (i, [i, p, l, l, i, i, p, p]),
        'mxf_potrf_packed': (i, [i, p, l, l, i, i, p, p, p]),
        'mxf_trsm_packed': (i, [i, i, i, i, d, p, l, l, p, l, p, l, l, i, p]),
        'mxf_tri_top_block': (i, [i, i]),
        'mxf_trsm_packed_oop': (i, [i, i, i, i, p, l, l, p, l, p, l, l, p, l, l, i, p]),
        'mxf_copy_ltu': (i, [i, p, l, l, p, l, l, i, i, p]),
        'mxf_copy_ltu_sum': (i, [i, p, l, l, i, p, l, i, p]),
        'mxf_copy2d': (i, [i, p, l, l, p, l, l, i, i, i, p]),
        'mxf_axpby2d': (i, [i, p, p, l, l, p, p, l, l, p, l, l, i, i, i, p]),
        'mxf_tri_pack_layout': (i, [i, i, p]),
        'mxf_info_max': (i, [p, p, i, p]),
        'mxf_symmetrize': (i, [i, d, p, l, l, p, l, l, i, i, p]),
        'mxf_tril': (i, [i, i, p, l, l, p, l, l, i, i, p]),
        'mxf_transpose': (i, [i, p, l, l, p, l, l, i, i, i, p]),
        'mxf_reduce': (i, [i, i, p, l, l, p, l, l, i, l, l, d, p, p]),
        'mxf_sumlogdiag': (i, [i, p, l, l, i, i, p, p]),
        'mxf_add_diag': (i, [i, p, l, l, p, l, d, i, i, p]),
        'mxf_get_diag': (i, [i, p, l, l, p, l, i, i, p]),
        'mxf_axpby_dev': (i, [i, p, p, l, p, p, l, p, l, i, l, p]),
        'mxf_softplus_fwd': (i, [i, p, d, p, l, p]),
        'mxf_softplus_bwd': (i, [i, p, p, p, l, p]),
        'mxf_svgp_bwd_assemble': (i, [i, p, p, p, p, p, p, p, l, l, i, i, i, p]),
        'mxf_svgp_bound_fwd': (i, [i, i, i, i, i, d, p, p, p, p, p, p, p, p, l, p, l, p, p, p, p]),
        'mxf_svgp_coef_bwd': (i, [i, i, i, i, d, p, p, p, p, p, p, p, p, p, p, p]),
        'mxf_normal_logpdf_sum': (i, [i, p, l, p, l, p, l, i, l, d, p, p]),
        'mxf_normal_logpdf_sum_bwd': (i, [i, p, l, p, l, p, l, i, l, d, p, p, p, p, p]),
        'mxf_normal_reparam': (i, [i, p, p, l, p, l, i, l, u, u, p, p, p, p]),
        'mxf_normal_logpdf_multi': (i, [i, i, p, p, p, p, p, p, p, p, p, p, p, p]),
        'mxf_normal_logpdf_multi_bwd': (i, [i, i, p, p, p, p, p, p, p, p, p, p, p, p, p, p, p]),
        'mxf_normal_reparam_multi': (i, [i, i, p, p, p, p, p, p, u, p, p, p, p, p]),
        'mxf_normal_reparam_multi_bwd': (i, [i, i, p, p, p, p, p, p, p, p, p, p]),
        'mxf_normal_reparam_bwd': (i, [i, p, p, p, l, l, i, l, p, p, p]),
        'mxf_adam_step': (i, [i, p, p, p, p, l, d, d, d, d, d, p, p]),
        'mxf_sgd_step': (i, [i, p, p, p, l, d, d, d, p, p]),
        'mxf_gather_rows': (i, [i, p, l, p, p, l, p, p]),
        'mxf_params_transform': (i, [i, i, p, p, p, p, p, p, p]),
        'mxf_params_pack_grads': (i, [i, i, p, p, p, p, p, p, p]),
        'mxf_mlp_tanh_fwd': (i, [i, i, p, p, l, p, p, p, p, p, i, i, p]),
        'mxf_mlp_tanh_bwd': (i, [i, i, p, p, l, p, p, p, p, p, p, p, i, i, p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTS = None


def lib():
    """Load (once) and return the shared library; raise loudly if it is absent."""
    global _lib, EXPORTS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MXFusionB200Error(
                "libmxf_b200.so is not built (expected at %s). Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C mxfusion_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        loaded = ctypes.CDLL(LIB_PATH)
        EXPORTS = _declare(loaded)
        _lib = loaded
    return _lib


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise MXFusionB200Error("unsupported dtype %s (float32/float64 only)" % t.dtype)


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MXFusionB200Error(
                "mxfusion_b200 operators run only on CUDA tensors (got a %s tensor); there is no CPU path."
                % t.device)


def check(rc, what):
    if rc != 0:
        if rc > 0:
            raise MXFusionB200Error("%s: CUDA error %d" % (what, rc))
        names = {-1: 'invalid argument', -2: 'unsupported dtype', -3: 'not implemented for this shape',
                 -4: 'workspace too small'}
        raise MXFusionB200Error("%s: %s" % (what, names.get(rc, rc)))


def launch_count():
    return int(lib().mxf_launch_count())
