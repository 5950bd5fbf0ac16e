"""ctypes binding of libfemol_b200.so (the C-ABI declared in include/femol_b200.h).

There is no CPU fallback: using any compute entry point without the built library or without a
CUDA device raises.  torch is used for device memory, streams and torch.distributed only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libfemol_b200.so')

T3, Q4, T6, Q8 = 0, 1, 2, 3
ETYPE_NODES = {T3: 3, Q4: 4, T6: 6, Q8: 8}
PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_JACOBI = 0, 1, 2
S_NOT_CONVERGED, S_BREAKDOWN, S_STAGNATED = 1, 2, 3

c_int, c_i64, c_dbl, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
P = c_vp  # every device pointer / stream / plan handle crosses as void*

# name -> (restype, argtypes); mirrors include/femol_b200.h one to one
SIGNATURES = {
    'femol_last_error': (ctypes.c_char_p, []),
    'femol_version': (ctypes.c_char_p, []),
    'femol_device_count': (c_int, []),
    'femol_launch_count': (c_i64, [c_int]),
    'femol_ke_batch': (c_int, [c_int, c_int, c_i64, P, P, P, P, P]),
    'femol_ke_batch_var': (c_int, [c_int, c_int, c_i64, P, P, P, c_int, P, P]),
    'femol_me_batch': (c_int, [c_int, c_int, c_i64, P, P, P, P, P]),
    'femol_plan_create': (c_int, [c_i64, c_int, c_int, c_i64, P, P, ctypes.POINTER(c_vp)]),
    'femol_plan_destroy': (c_int, [P]),
    'femol_plan_info': (c_int, [P, ctypes.POINTER(c_i64)]),
    'femol_plan_export_csr': (c_int, [P, P, P, P]),
    'femol_plan_export_adjacency': (c_int, [P, P, P, P]),
    'femol_plan_export_scatter_map': (c_int, [P, P, P]),
    'femol_assemble': (c_int, [P, P, c_int, P, c_int, P, c_dbl, P, P, P, P]),
    'femol_spmv': (c_int, [P, P, P, P, P]),
    'femol_spmv_csr': (c_int, [c_i64, P, P, P, P, P, P]),
    'femol_to_f32': (c_int, [P, P, P, P]),
    'femol_apply_f32': (c_int, [P, P, P, P, c_int, P, P, c_dbl, P]),
    'femol_plan_set_grid': (c_int, [P, c_i64, c_i64]),
    'femol_plan_set_matrix_free': (c_int, [P, c_i64, c_i64, P, P, P, P, P]),
    'femol_stress_q4': (c_int, [P, P, P, P, P, P, P, P]),
    'femol_apply': (c_int, [P, P, P, P, P]),
    'femol_apply_epilogue': (c_int, [P, P, P, P, c_int, P, P, c_dbl, P]),
    'femol_pcg_workspace_doubles': (c_i64, [P]),
    'femol_pcg': (c_int, [P, P, P, P, c_dbl, c_int, c_int, P, ctypes.POINTER(c_int), ctypes.POINTER(c_dbl), P]),
    'femol_gmg_create': (c_int, [c_i64, c_i64, c_int, P, P, c_int, c_dbl, P, ctypes.POINTER(c_vp)]),
    'femol_gmg_destroy': (c_int, [P]),
    'femol_gmg_setup': (c_int, [P, P, P, P, P, P]),
    'femol_gmg_info': (c_int, [P, ctypes.POINTER(c_i64)]),
    'femol_gmg_create_lattice': (c_int, [c_i64, c_i64, c_int, P, P, P, P, P, P, c_int, c_dbl, P, ctypes.POINTER(c_vp)]),
    'femol_gmg_setup_lattice': (c_int, [P, P, c_int, P, c_int, P, P, P]),
    'femol_gmg_apply': (c_int, [P, P, P, P]),
    'femol_gmg_clone_workspace': (c_int, [P, ctypes.POINTER(c_vp)]),
    'femol_gmg_attach_modulated': (c_int, [P, P]),
    'femol_plan_create_view': (c_int, [P, ctypes.POINTER(c_vp)]),
    'femol_pcg_gmg_workspace_doubles': (c_i64, [P]),
    'femol_pcg_gmg': (c_int, [P, P, P, P, P, c_dbl, c_int, P, ctypes.POINTER(c_int), ctypes.POINTER(c_dbl), P]),
    'femol_graph_stats': (c_int, [ctypes.POINTER(c_i64), c_int]),
    'femol_dist_unique_id': (c_int, [P]),
    'femol_dist_init': (c_int, [P, c_int, c_int, ctypes.POINTER(c_vp)]),
    'femol_dist_destroy': (c_int, [P]),
    'femol_pcg_dist_workspace_doubles': (c_i64, [P, c_i64]),
    'femol_pcg_dist': (c_int, [P, P, c_i64, c_int, P, P, P, P, P, P, P, P, c_dbl, c_int, c_int, P,
                       ctypes.POINTER(c_int), ctypes.POINTER(c_dbl), P]),
    'femol_gmg_truncate': (c_int, [P, c_int]),
    'femol_gmg_set_dist': (c_int, [P, P, c_int, P, P, P, P, P, P, P, P]),
    'femol_gmg_dist_refresh': (c_int, [P, P]),
    'femol_gmg_dist_exchange': (c_int, [P, c_int, P, c_int, P]),
    'femol_dist_allreduce_sum': (c_int, [P, P, c_i64, P]),
    'femol_gmg_dist_p2p_sizes': (c_int, [P, ctypes.POINTER(c_i64)]),
    'femol_dist_p2p_open': (c_int, [P, c_i64, c_i64, P]),
    'femol_dist_p2p_attach': (c_int, [P, P]),
    'femol_dist_p2p_status': (c_int, [P, ctypes.POINTER(c_i64)]),
    'femol_pcg_gmg_dist_workspace_doubles': (c_i64, [P]),
    'femol_pcg_gmg_dist': (c_int, [P, P, P, P, P, c_dbl, c_int, P, ctypes.POINTER(c_int), ctypes.POINTER(c_dbl), P]),
    'femol_spmm': (c_int, [P, P, P, c_int, P, c_int, c_int, P]),
    'femol_gram_workspace_doubles': (c_i64, [c_int, c_int]),
    'femol_gram': (c_int, [c_i64, c_int, c_int, P, c_int, P, c_int, P, P, P]),
    'femol_block_gemm': (c_int, [c_i64, c_int, c_int, P, c_int, P, P, c_int, c_dbl, c_dbl, P]),
    'femol_block_residual': (c_int, [c_i64, c_int, P, c_int, P, c_int, P, P, c_int, P]),
    'femol_precond_setup': (c_int, [P, P, c_int, P, P]),
    'femol_precond_apply': (c_int, [P, P, P, P, c_int, P, c_int, c_int, P]),
    'femol_block_pcg_workspace_doubles': (c_i64, [P, c_int]),
    'femol_block_pcg': (c_int, [P, P, P, P, P, c_int, P, c_int, c_int, c_int, c_dbl, P, P]),
    'femol_reduce_workspace_doubles': (c_i64, []),
    'femol_sens_compliance': (c_int, [P, P, c_int, P, c_int, P, P, c_dbl, c_dbl, c_dbl, P, P, P, P]),
    'femol_sens_eig': (c_int, [P, P, c_int, P, c_int, P, P, c_dbl, c_dbl, c_dbl, c_dbl, P, P, P]),
    'femol_filter_grid': (c_int, [c_i64, c_i64, c_int, c_int, P, P, P, P, P]),
    'femol_filter_list': (c_int, [c_i64, P, P, P, P, P, P, P]),
    'femol_oc_update': (c_int, [c_i64, P, P, c_int, c_dbl, P, P, P, P, P, P]),
    'femol_mac': (c_int, [c_i64, c_int, P, P, P, P, P]),
    'femol_dot': (c_int, [c_i64, P, P, P, P, P]),
}

_lib = None


class FemolError(RuntimeError):
    pass


class FemolConvergenceWarning(RuntimeWarning):
    """An iterative solve (PCG / LOBPCG) stopped without reaching its tolerance.  The reference uses direct
    solvers (spsolve / eigh), so its callers never check; here the condition is reported instead of silent."""


def load():
    """Load the shared library (symbols only; needs no GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FemolError('libfemol_b200.so is not built (%s). Run `python femol_b200/csrc/build.py` '
                         '(or __graft_entry__.build()); femol_b200 has no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_gpu():
    lib = load()
    if not torch.cuda.is_available() or lib.femol_device_count() < 1:
        raise FemolError('femol_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return lib


def check(rc, what=''):
    if rc < 0:
        msg = load().femol_last_error().decode(errors='replace')
        raise FemolError('%s failed (%d): %s' % (what, rc, msg))
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def graph_stats(reset=False):
    """CUDA-graph use of femol_pcg_gmg: captures attempted / instantiated, graph launches, side-stream solves."""
    out = (c_i64 * 4)()
    check(load().femol_graph_stats(out, 1 if reset else 0), 'femol_graph_stats')
    return dict(zip(('captures', 'instantiated', 'graph_launches', 'side_stream_solves'), [int(v) for v in out]))


def call(name, *args):
    lib = require_gpu()
    return check(getattr(lib, name)(*args), name)
