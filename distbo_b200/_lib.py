"""ctypes binding of the C-ABI library ``libdistbo_b200.so`` (see include/distbo_b200.h).

The library is the product: there is NO CPU fallback.  Importing this module without the built
library, or calling into it without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdistbo_b200.so")

_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_d = C.c_double

# name -> (restype, argtypes); mirrors include/distbo_b200.h one to one
SIGNATURES = {
    "distbo_version": (_i, []),
    "distbo_launch_count": (_i64, [_i]),
    "distbo_debug_timeline": (_i, [_vp]),
    "distbo_embed_fwd_f64": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                  _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp]),
    "distbo_embed_fwd_f32": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                  _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp]),
    "distbo_embed_bwd_f64": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                  _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _i64, _vp]),
    "distbo_embed_f32_on_tensor_cores": (_i, [_i, _i, _i, _i, _i, _i]),
    "distbo_embed_workspace": (_i64, [_i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "distbo_py_shuffle_i64": (_i, [_vp, _vp, _vp, _i64]),
    "distbo_gather_rows": (_i, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "distbo_task_gram_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "distbo_scale_theta_f64": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "distbo_gram_f64": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _d, _vp, _vp]),
    "distbo_plan_create": (_i, [_i, C.POINTER(_vp)]),
    "distbo_plan_destroy": (_i, [_vp]),
    "distbo_potrf_f64": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "distbo_trtri_f64": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "distbo_lauum_f64": (_i, [_vp, _vp, _vp, _vp]),
    "distbo_nll_workspace": (_i64, [_i]),
    "distbo_nll_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "distbo_gram_grad_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp]),
    "distbo_task_gram_bwd_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "distbo_task_tri_f64": (_i, [_vp, _i, _vp, _vp, _vp]),
    "distbo_task_tri_bwd_f64": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "distbo_blr_workspace": (_i64, [_i64, _i, _i, _i, _i, _i]),
    "distbo_blr_fit_f64": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "distbo_blr_score_blocks": (_i, []),
    "distbo_blr_score_f64": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp,
                                  _vp, _vp, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "distbo_adam_f64": (_i, [_vp, _vp, _vp, _vp, _i64, _d, _d, _d, _d, _i64, _d, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "distbo_score_chunk": (_i, []),
    "distbo_score_workspace": (_i64, [_i]),
    "distbo_score_timing": (_i, [_i]),
    "distbo_score_timing_read": (_i, [C.POINTER(_d), C.POINTER(_i64)]),
    "distbo_score_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _d, _d, _d,
                              _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "distbo_trtri_levels": (_i, [_vp]),
    "distbo_trtri_range_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "distbo_trtri_i8_first_level": (_i, [_i, _i]),
    "distbo_trtri_i8_sched_words": (_i64, [_i, _i]),
    "distbo_trtri_i8_plan": (_i, [_i, _i, _vp]),
    "distbo_trtri_i8_workspace": (_i64, [_i]),
    "distbo_trtri_i8_f64": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "distbo_lauum_i8_sched_words": (_i64, [_i]),
    "distbo_lauum_i8_plan": (_i, [_i, _vp]),
    "distbo_lauum_i8_workspace": (_i64, [_i]),
    "distbo_lauum_i8_f64": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "distbo_score_i8_workspace": (_i64, [_i]),
    "distbo_score_i8_prepare": (_i, [_vp, _i, _vp, _vp, _vp]),
    "distbo_score_i8_f64": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i,
                                 _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None


class DistboError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DistboError(
            "distbo_b200: %s is missing - build it with `make` or `python -c 'import __graft_entry__ as g; "
            "g.build()'`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        kind = {-1: "bad argument", -2: "CUDA error"}.get(rc, "error %d" % rc)
        raise DistboError("%s failed: %s" % (what, kind))


def ptr(t):
    """Device pointer of a torch tensor (or NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
