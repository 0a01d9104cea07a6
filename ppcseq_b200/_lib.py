"""ctypes loader for libppcseq_b200.so (the C ABI declared in include/ppcseq_b200.h)."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PpcseqError(RuntimeError):
    """Carries the integer PPCSEQ_E* code of include/ppcseq_b200.h in `.rc`."""

    def __init__(self, msg, rc=None):
        super().__init__(msg)
        self.rc = rc


EDIVERGED = 6      # PPCSEQ_EDIVERGED
ECOMM = 7          # PPCSEQ_ECOMM


def library_path() -> str:
    # PPCSEQ_B200_LIB: another build of the SAME library (A/B timing of kernel variants, profiles/tools/ab.sh)
    return os.environ.get("PPCSEQ_B200_LIB") or os.path.join(_HERE, "libppcseq_b200.so")


c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)
I32, I64, DBL, VP, INT = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int

class NutsOpts(ctypes.Structure):
    """ppcseq_nuts_opts (include/ppcseq_b200.h)."""
    _fields_ = [("chains", I32), ("iter", I32), ("warmup", I32), ("max_treedepth", I32), ("adapt_init_buffer", I32),
                ("adapt_term_buffer", I32), ("adapt_window", I32), ("threads", I32), ("adapt_delta", DBL),
                ("adapt_gamma", DBL), ("adapt_kappa", DBL), ("adapt_t0", DBL), ("stepsize", DBL), ("init_radius", DBL),
                ("seed", ctypes.c_uint64), ("init", c_double_p)]


class AdviOpts(ctypes.Structure):
    """ppcseq_advi_opts (include/ppcseq_b200.h)."""
    _fields_ = [("iter", I32), ("grad_samples", I32), ("elbo_samples", I32), ("eval_elbo", I32), ("output_samples", I32),
                ("adapt_engaged", I32), ("adapt_iter", I32), ("reserved", I32), ("eta", DBL), ("tol_rel_obj", DBL),
                ("init_radius", DBL), ("seed", ctypes.c_uint64), ("init", c_double_p)]


# name -> (restype, argtypes); mirrors include/ppcseq_b200.h one to one
SIGNATURES = {
    "ppcseq_last_error": (ctypes.c_char_p, []),
    "ppcseq_abi_version": (INT, []),
    "ppcseq_launch_count": (I64, []),
    "ppcseq_model_create": (INT, [I32, I32, I32, I32, c_int32_p, c_double_p, c_double_p, DBL, INT, c_void_pp]),
    "ppcseq_model_create_shard": (INT, [I32, I32, I32, I32, I32, I32, c_int32_p, c_double_p, c_double_p, DBL, INT,
                                        c_void_pp]),
    "ppcseq_model_create_multi": (INT, [I32, I32, I32, I32, c_int32_p, c_double_p, c_double_p, DBL, I32, c_int32_p, c_void_pp]),
    "ppcseq_model_free": (None, [VP]),
    "ppcseq_model_set_exclusion": (INT, [VP, c_int32_p, I64]),
    "ppcseq_model_set_design_path": (INT, [VP, INT]),
    "ppcseq_model_dims": (INT, [VP, c_int32_p, c_int32_p, c_int32_p, c_int32_p, ctypes.POINTER(I64)]),
    "ppcseq_comm_create": (INT, [VP, I32, I32, I32, I32, c_uint8_p]),
    "ppcseq_comm_connect": (INT, [VP, c_uint8_p]),
    "ppcseq_comm_status": (INT, [VP, c_int32_p]),
    "ppcseq_model_status": (INT, [VP, c_int32_p]),
    "ppcseq_log_prob_grad": (INT, [VP, I32, c_double_p, INT, INT, c_double_p, c_double_p]),
    "ppcseq_log_prob_grad_device": (INT, [VP, I32, VP, INT, INT, VP, VP, VP]),
    "ppcseq_log_prob_grad_partial_device": (INT, [VP, I32, VP, INT, VP, VP, VP]),
    "ppcseq_finalize_hyper_device": (INT, [VP, I32, VP, VP, INT, INT, VP, VP, VP]),
    "ppcseq_exposure_grad": (INT, [VP, c_double_p, c_double_p]),
    "ppcseq_exposure_grad_device": (INT, [VP, VP, VP, VP]),
    "ppcseq_summarise_draws": (INT, [INT, c_double_p, I32, I64, DBL, c_double_p, c_double_p, c_double_p, c_double_p]),
    "ppcseq_flags": (INT, [VP, c_double_p, c_double_p, c_double_p, c_double_p, c_uint8_p, c_uint8_p, c_int32_p, c_int32_p]),
    "ppcseq_fit_from_draws": (INT, [VP, c_double_p, I32, c_void_pp]),
    "ppcseq_fit_free": (None, [VP]),
    "ppcseq_fit_num_draws": (INT, [VP, c_int32_p]),
    "ppcseq_fit_get_draws": (INT, [VP, I64, I64, c_double_p]),
    "ppcseq_fit_param_mean": (INT, [VP, I64, I64, c_double_p]),
    "ppcseq_fit_info": (INT, [VP, c_double_p, I32]),
    "ppcseq_nuts_default_opts": (INT, [ctypes.POINTER(NutsOpts)]),
    "ppcseq_advi_default_opts": (INT, [ctypes.POINTER(AdviOpts)]),
    "ppcseq_sample_nuts": (INT, [VP, ctypes.POINTER(NutsOpts), c_void_pp]),
    "ppcseq_advi_meanfield": (INT, [VP, ctypes.POINTER(AdviOpts), c_void_pp]),
    "ppcseq_ppc_summary": (INT, [VP, INT, I64, DBL, DBL, ctypes.c_uint64, c_double_p, c_double_p, c_double_p, c_double_p]),
    "ppcseq_ppc_draws": (INT, [VP, DBL, ctypes.c_uint64, c_double_p]),
    "ppcseq_prep_table": (INT, [I64, ctypes.POINTER(I64), ctypes.POINTER(I64), VP, I32, c_double_p, c_uint8_p, I64, I32, c_void_pp]),
    "ppcseq_prep_dims": (INT, [VP, c_int32_p, c_int32_p, c_int32_p]),
    "ppcseq_prep_fetch": (INT, [VP, ctypes.POINTER(I64), ctypes.POINTER(I64), ctypes.POINTER(I64), c_int32_p]),
    "ppcseq_prep_free": (None, [VP]),
    "ppcseq_tmm_factors": (INT, [I32, I32, c_int32_p, c_int32_p, I32, I32, c_double_p, c_double_p, c_int32_p]),
    "ppcseq_device_alloc": (INT, [INT, I64, c_void_pp]),
    "ppcseq_device_free": (INT, [INT, VP]),
    "ppcseq_memcpy_h2d": (INT, [VP, VP, I64, VP]),
    "ppcseq_memcpy_d2h": (INT, [VP, VP, I64, VP]),
    "ppcseq_stream_sync": (INT, [VP, VP]),
    "ppcseq_time_log_prob_grad_device": (INT, [VP, I32, VP, INT, INT, VP, VP, VP, I32, INT, c_float_p]),
    "ppcseq_measure_fp64_peak": (INT, [INT, c_double_p]),
}


def lib():
    """The loaded library.  Raises PpcseqError if it has not been built -- never falls back."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise PpcseqError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C ppcseq_b200/csrc`).  ppcseq_b200 has no CPU fallback.")
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        msg = lib().ppcseq_last_error()
        raise PpcseqError(f"ppcseq_b200 error {rc}: {msg.decode() if msg else '?'}", rc)
