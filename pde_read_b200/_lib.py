"""ctypes binding of libpderead_b200.so (the C ABI declared in include/pderead_b200.h).

There is no CPU fallback: importing this module without the built library, or calling into it
without a CUDA device, raises.  Build with ``python -m pde_read_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

MAX_HIDDEN_LAYERS = 16
MAX_INPUT_DIM = 8
MAX_SPATIAL_ORDER = 4
MAX_TIME_ORDER = 2
MAX_LIBRARY_COLUMNS = 256

ACT_CODES = {"Tanh": 0, "Sin": 1, "Rational": 2}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDEREAD_LIB") or os.path.join(_HERE, "libpderead_b200.so")


class PdereadError(RuntimeError):
    pass


class pderead_net(Structure):
    _fields_ = [
        ("num_hidden_layers", c_int32),
        ("width", c_int32),
        ("input_dim", c_int32),
        ("activation", c_int32),
        ("weight", c_void_p * (MAX_HIDDEN_LAYERS + 1)),
        ("bias", c_void_p * (MAX_HIDDEN_LAYERS + 1)),
        ("rat_a", c_void_p * MAX_HIDDEN_LAYERS),
        ("rat_b", c_void_p * MAX_HIDDEN_LAYERS),
    ]


class pderead_net_grad(Structure):
    _fields_ = [
        ("weight", c_void_p * (MAX_HIDDEN_LAYERS + 1)),
        ("bias", c_void_p * (MAX_HIDDEN_LAYERS + 1)),
        ("rat_a", c_void_p * MAX_HIDDEN_LAYERS),
        ("rat_b", c_void_p * MAX_HIDDEN_LAYERS),
    ]


_NET = POINTER(pderead_net)
_GRAD = POINTER(pderead_net_grad)

# name -> (restype, argtypes); kept in one table so the CPU test can check that the library
# exports every symbol the header declares.
PROTOTYPES = {
    "pderead_version": (c_char_p, []),
    "pderead_status_string": (c_char_p, [c_int]),
    "pderead_last_cuda_error": (c_int, []),
    "pderead_last_cuda_error_where": (ctypes.c_char_p, []),
    "pderead_sol_derivatives_workspace_bytes": (c_size_t, [_NET, c_int, c_int, c_int64, c_int]),
    "pderead_sol_derivatives_forward": (c_int, [_NET, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                                c_size_t, c_void_p]),
    "pderead_sol_derivatives_backward": (c_int, [_NET, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, _GRAD,
                                                 c_float, c_void_p, c_size_t, c_void_p]),
    "pderead_mlp_workspace_bytes": (c_size_t, [_NET, c_int64, c_int]),
    "pderead_mlp_forward": (c_int, [_NET, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pderead_mlp_backward": (c_int, [_NET, c_void_p, c_int64, c_void_p, c_void_p, _GRAD, c_float, c_void_p, c_size_t,
                                     c_void_p]),
    "pderead_residual_workspace_bytes": (c_size_t, [_NET, _NET, c_int, c_int, c_int64, c_int]),
    "pderead_residual_forward": (c_int, [_NET, _NET, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                         c_size_t, c_void_p]),
    "pderead_collocation_loss_grad": (c_int, [_NET, _NET, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                              c_void_p, _GRAD, _GRAD, c_float, c_void_p, c_size_t, c_void_p]),
    "pderead_library_num_columns": (c_int, [c_int, c_int]),
    "pderead_library_multi_indices": (c_int, [c_int, c_int, POINTER(c_int32)]),
    "pderead_library_dense": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "pderead_library_gram_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pderead_library_gram": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "pderead_dense_gram": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "pderead_extraction_gram_workspace_bytes": (c_size_t, [_NET, _NET, c_int, c_int, c_int64]),
    "pderead_extraction_gram": (c_int, [_NET, _NET, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "pderead_rfe_workspace_bytes": (c_size_t, [c_int]),
    "pderead_rfe_gram": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "pderead_input_moments": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "pderead_bn_fold": (c_int, [_NET, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "pderead_bn_grad_sums": (c_int, [_NET, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pderead_bn_backward": (c_int, [_NET, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pderead_bn_input_grad": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "pderead_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                  c_float, c_float, c_void_p, c_void_p]),
    "pderead_sgd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_int, c_float, c_float,
                                 c_void_p, c_void_p]),
    "pderead_loss_tail_pack": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pderead_loss_tail_unpack": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pderead_peer_area_bytes": (c_size_t, [c_int, c_int]),
    "pderead_peer_alloc": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "pderead_peer_open": (c_int, [c_void_p, c_void_p]),
    "pderead_peer_close": (c_int, [c_void_p]),
    "pderead_peer_free": (c_int, [c_void_p]),
    "pderead_peer_allreduce": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pderead_generate_points": (c_int, [POINTER(c_double), c_int, c_int64, c_uint64, c_uint64, c_void_p, c_void_p]),
    "pderead_fma_peak_probe": (c_int, [c_int, c_int, c_void_p, POINTER(c_double), c_void_p]),
    "pderead_debug_stage_events": (c_int, [POINTER(c_void_p), c_int]),
    "pderead_debug_stage_events_recorded": (c_int, []),
    "pderead_debug_stage_tags": (c_int, [POINTER(c_int32), c_int]),
    "pderead_debug_set_tensor_cores": (c_int, [c_int]),
    "pderead_debug_set_plain_mlp": (c_int, [c_int]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises PdereadError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PdereadError(
            "%s not found: the CUDA library is not built. Run `python -m pde_read_b200.build` "
            "(needs nvcc; there is no CPU fallback)." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.pderead_status_string(rc).decode()
        if rc == -5:
            msg += " (cudaError %d at %s)" % (lib.pderead_last_cuda_error(),
                                               (lib.pderead_last_cuda_error_where() or b"?").decode())
        raise PdereadError("%s failed: %s [%d]" % (what or "pderead call", msg, rc))
