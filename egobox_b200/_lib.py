"""ctypes binding of libegobox_gpu.so (the C ABI in include/egobox_gpu.h).

There is no Python/CPU fallback: if the shared library is missing the import
of any compute entry point raises, and if no CUDA device is visible every
compute call returns EGX_CUDA_ERROR (surfaced as ``GpuError``)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libegobox_gpu.so")

EGX_OK, EGX_NOT_POSITIVE_DEFINITE, EGX_ILL_CONDITIONED_FT, EGX_ILL_CONDITIONED_F, \
    EGX_INVALID_VALUE, EGX_CUDA_ERROR = range(6)
STATUS_NAMES = {0: "OK", 1: "NOT_POSITIVE_DEFINITE", 2: "ILL_CONDITIONED_FT", 3: "ILL_CONDITIONED_F",
                4: "INVALID_VALUE", 5: "CUDA_ERROR"}
NUM_STAGES = 13
STAGE_NAMES = ["corr_build", "potrf_diag", "trsm_panel", "syrk_gemm", "gls", "backsolve",
               "cross_corr", "var_finish", "small_batch", "gemm_lookahead", "ozaki_slice", "ozaki_syrk", "theta_grad"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/egobox_gpu.h declares
SIGNATURES = {
    "egx_device_count": (C.c_int, []),
    "egx_last_error": (C.c_char_p, []),
    "egx_version": (C.c_char_p, []),
    "egx_gp_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, _dp,
                                C.c_double, C.c_double, _dp, C.c_int, C.c_double]),
    "egx_gp_destroy": (None, [_vp]),
    "egx_gp_dims": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "egx_gp_reduced_likelihood": (C.c_int, [_vp, _dp, _dp]),
    "egx_gp_reduced_likelihood_batch": (C.c_int, [_vp, _dp, C.c_int, _dp, _ip]),
    "egx_gp_finalize": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "egx_gp_async_slots": (C.c_int, [_vp, C.c_int]),
    "egx_gp_release_workspaces": (C.c_int, [_vp]),
    "egx_gp_eval_begin": (C.c_int, [_vp, C.c_int, _dp]),
    "egx_gp_eval_end": (C.c_int, [_vp, C.c_int, _dp]),
    "egx_gp_reduced_likelihood_grad": (C.c_int, [_vp, _dp, C.c_double, _dp, _dp]),
    "egx_gp_reduced_likelihood_grad_analytic": (C.c_int, [_vp, _dp, _dp, _dp]),
    "egx_gp_download_chol": (C.c_int, [_vp, _dp]),
    "egx_gp_predict": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_predict_var": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_predict_valvar": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp]),
    "egx_gp_predict_valvar_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "egx_gp_predict_gradients": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_predict_var_gradients": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_predict_gradients_dev": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "egx_gp_predict_var_gradients_dev": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "egx_moe_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double]),
    "egx_moe_destroy": (None, [_vp]),
    "egx_moe_set_heaviside_factor": (C.c_int, [_vp, C.c_double]),
    "egx_moe_set_expert": (C.c_int, [_vp, C.c_int, _vp]),
    "egx_moe_parameters": (C.c_int, [_vp, _dp, _dp, _dp]),
    "egx_moe_predict_probas": (C.c_int, [_vp, _dp, C.c_int, _dp, _ip]),
    "egx_moe_predict_probas_derivatives": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_moe_predict": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp]),
    "egx_symmetric_eig": (C.c_int, [C.c_int, _dp, _dp]),
    "egx_lhs_sample": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, C.c_ulonglong, _dp]),
    "egx_shuffled_indices": (C.c_int, [C.c_int, C.c_ulonglong, _ip]),
    "egx_pls_rotations": (C.c_int, [_dp, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "egx_release_cached_memory": (None, []),
    "egx_gp_covariance": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_sample": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp]),
    "egx_gp_correlation_matrix": (C.c_int, [_vp, _dp, _dp]),
    "egx_gp_cross_correlation": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_set_profiling": (C.c_int, [_vp, C.c_int]),
    "egx_gp_reset_profile": (C.c_int, [_vp]),
    "egx_gp_get_profile": (C.c_int, [_vp, _dp, C.POINTER(C.c_longlong)]),
    "egx_gp_set_force_blocked": (C.c_int, [_vp, C.c_int]),
    "egx_gp_set_lookahead": (C.c_int, [_vp, C.c_int]),
    "egx_gp_timer_start": (C.c_int, [_vp]),
    "egx_gp_timer_stop": (C.c_int, [_vp, _dp]),
}



EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_void_p)


class GpParamsStruct(C.Structure):
    """egx_gp_params (include/egobox_gpu.h)."""
    _fields_ = [("corr", C.c_int), ("mean", C.c_int), ("theta_tuning", C.c_int),
                ("theta_init", _dp), ("n_theta_init", C.c_int),
                ("theta_bounds", _dp), ("n_theta_bounds", C.c_int),
                ("active", _ip), ("n_active", C.c_int),
                ("n_start", C.c_int), ("max_eval", C.c_int), ("nugget", C.c_double),
                ("w_star", _dp), ("kpls_dim", C.c_int), ("device", C.c_int),
                ("seed", C.c_ulonglong), ("cobyla_rhobeg", C.c_double), ("cobyla_ftol_rel", C.c_double),
                ("optimizer", C.c_int), ("chain_rank", C.c_int), ("chain_world", C.c_int),
                ("exchange", EXCHANGE_FN), ("exchange_user", C.c_void_p)]


EGX_OPT_COBYLA, EGX_OPT_LBFGSB = 0, 1
OBJECTIVE_FN = C.CFUNCTYPE(C.c_double, _dp, C.c_int, C.c_void_p)
OBJECTIVE_GRAD_FN = C.CFUNCTYPE(C.c_double, _dp, C.c_int, _dp, C.c_void_p)
_pp = C.POINTER(GpParamsStruct)
SIGNATURES.update({
    "egx_gp_params_default": (None, [_pp]),
    "egx_gp_fit": (C.c_int, [_pp, _dp, C.c_int, C.c_int, _dp, C.POINTER(_vp)]),
    "egx_gp_model_destroy": (None, [_vp]),
    "egx_gp_model_dims": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "egx_gp_model_theta": (C.c_int, [_vp, _dp]),
    "egx_gp_model_variance": (C.c_double, [_vp]),
    "egx_gp_model_likelihood": (C.c_double, [_vp]),
    "egx_gp_model_n_evals": (C.c_longlong, [_vp]),
    "egx_gp_model_inner_params": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "egx_gp_model_normalization": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "egx_gp_model_context": (_vp, [_vp]),
    "egx_gp_model_predict": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_model_predict_var": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_model_predict_valvar": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp]),
    "egx_gp_model_predict_gradients": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_model_predict_var_gradients": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_model_covariance": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_gp_model_sample": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp]),
    "egx_bound_cobyla_minimize": (C.c_int, [OBJECTIVE_FN, C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_double,
                                            C.c_double, C.c_int, _dp, _dp, _ip]),
    "egx_bound_lbfgs_minimize": (C.c_int, [OBJECTIVE_GRAD_FN, C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_double,
                                           C.c_double, C.c_int, _dp, _dp, _ip]),
    "egx_prepare_multistart": (C.c_int, [C.c_int, _dp, _dp, C.c_int, C.c_ulonglong, _dp]),
    "egx_comm_init": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]),
    "egx_comm_destroy": (None, [_vp]),
    "egx_comm_rank": (C.c_int, [_vp]),
    "egx_comm_size": (C.c_int, [_vp]),
    "egx_comm_allgather": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_argmin_allreduce": (C.c_int, [_vp, _dp, _dp, C.c_int, _ip]),
})



class SgpParamsStruct(C.Structure):
    """egx_sgp_params (include/egobox_gpu.h)."""
    _fields_ = [("corr", C.c_int), ("method", C.c_int), ("theta_fixed", C.c_int),
                ("theta_init", _dp), ("n_theta_init", C.c_int),
                ("theta_bounds", _dp), ("n_theta_bounds", C.c_int),
                ("noise_fixed", C.c_int), ("noise_init", C.c_double), ("noise_lo", C.c_double),
                ("noise_hi", C.c_double), ("z", _dp), ("n_inducings", C.c_int),
                ("n_start", C.c_int), ("max_eval", C.c_int), ("nugget", C.c_double),
                ("w_star", _dp), ("kpls_dim", C.c_int), ("device", C.c_int), ("seed", C.c_ulonglong),
                ("cobyla_rhobeg", C.c_double), ("cobyla_ftol_rel", C.c_double)]


_sp = C.POINTER(SgpParamsStruct)
_llp = C.POINTER(C.c_longlong)
SIGNATURES.update({
    "egx_sgp_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                 _dp, C.c_int, C.c_double]),
    "egx_sgp_destroy": (None, [_vp]),
    "egx_sgp_reduced_likelihood": (C.c_int, [_vp, _dp, C.c_double, C.c_double, _dp]),
    "egx_sgp_finalize": (C.c_int, [_vp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]),
    "egx_sgp_predict": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_sgp_predict_var": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_sgp_set_profiling": (C.c_int, [_vp, C.c_int]),
    "egx_sgp_get_profile": (C.c_int, [_vp, _dp, _llp]),
    "egx_sgp_params_default": (None, [_sp]),
    "egx_sgp_fit": (C.c_int, [_sp, _dp, C.c_int, C.c_int, _dp, C.POINTER(_vp)]),
    "egx_sgp_model_destroy": (None, [_vp]),
    "egx_sgp_model_dims": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "egx_sgp_model_theta": (C.c_int, [_vp, _dp]),
    "egx_sgp_model_variance": (C.c_double, [_vp]),
    "egx_sgp_model_noise_variance": (C.c_double, [_vp]),
    "egx_sgp_model_likelihood": (C.c_double, [_vp]),
    "egx_sgp_model_n_evals": (C.c_longlong, [_vp]),
    "egx_sgp_model_inducings": (C.c_int, [_vp, _dp]),
    "egx_sgp_model_woodbury": (C.c_int, [_vp, _dp, _dp]),
    "egx_sgp_model_context": (_vp, [_vp]),
    "egx_sgp_model_predict": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_sgp_model_predict_var": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "egx_sgp_model_sample": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp]),
    "egx_sgp_sample": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp]),
})

_lib = None


class GpuError(RuntimeError):
    """A C-ABI call returned EGX_CUDA_ERROR / EGX_INVALID_VALUE."""

    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), msg))
        self.status = status


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "egobox_b200: %s is missing -- build it with `python -m egobox_b200._build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().egx_last_error().decode("utf-8", "replace")


def device_count():
    return int(load().egx_device_count())
