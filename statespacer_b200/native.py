"""ctypes binding of the in-tree libssgpu.so (include/ssgpu.h).  There is no CPU fallback: if the
library is missing or a call fails the error is raised, never papered over."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSGPU_LIB: another build of the same library (A/B runs of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("SSGPU_LIB") or os.path.join(_HERE, "libssgpu.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class SsgpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libssgpu status {status}: {msg}")
        self.status = status


class Matrix(C.Structure):
    """ssgpu_matrix"""
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("diag", C.c_int),
                ("n_slices", C.c_int), ("stride_series", C.c_longlong), ("stride_variant", C.c_longlong)]


class Model(C.Structure):
    """ssgpu_model"""
    _fields_ = [("N", C.c_int), ("p", C.c_int), ("m", C.c_int), ("r", C.c_int),
                ("a", Matrix), ("P_inf", Matrix), ("P_star", Matrix), ("Z", Matrix), ("T", Matrix),
                ("R", Matrix), ("Q", Matrix)]


KALMAN_FIELDS = ["loglik", "a_pred", "a_fil", "a_smooth", "P_pred", "P_fil", "V", "P_inf_pred", "P_star_pred",
                 "P_inf_fil", "P_star_fil", "yfit", "v", "v_norm", "eta", "e", "Fmat", "eta_var", "D", "a_fc",
                 "P_inf_fc", "P_star_fc", "P_fc", "Tstat_observation", "Tstat_state", "r_vec", "Nmat"]


class KalmanOut(C.Structure):
    """ssgpu_kalman_out"""
    _fields_ = [("initialisation_steps", ip)] + [(n, dp) for n in KALMAN_FIELDS]


class CovBlock(C.Structure):
    """ssgpu_cov_block"""
    _fields_ = [("target", C.c_int), ("offset", C.c_int), ("n", C.c_int), ("theta_offset", C.c_int)]


class SimComponent(C.Structure):
    """ssgpu_sim_component"""
    _fields_ = [("m", C.c_int), ("r", C.c_int), ("repeat_Q", C.c_int), ("draw_initial", C.c_int),
                ("a", C.c_void_p), ("Z", C.c_void_p), ("T", C.c_void_p), ("R", C.c_void_p), ("Q", C.c_void_p),
                ("P_star", C.c_void_p), ("nZ", C.c_int), ("nT", C.c_int), ("nR", C.c_int), ("nQ", C.c_int),
                ("Q_root", C.c_void_p), ("P_root", C.c_void_p)]


class SimModel(C.Structure):
    """ssgpu_sim_model"""
    _fields_ = [("m", C.c_int), ("r", C.c_int), ("a", C.c_void_p), ("P_inf", C.c_void_p), ("P_star", C.c_void_p),
                ("Z", C.c_void_p), ("T", C.c_void_p), ("R", C.c_void_p), ("Q", C.c_void_p),
                ("nZ", C.c_int), ("nT", C.c_int), ("nR", C.c_int), ("nQ", C.c_int)]


# every symbol include/ssgpu.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = ["ssgpu_version", "ssgpu_last_error", "ssgpu_set_device", "ssgpu_device_count", "ssgpu_release",
           "ssgpu_launch_count", "ssgpu_LogLikC", "ssgpu_KalmanC", "ssgpu_FastSmootherC", "ssgpu_SimulateC",
           "ssgpu_sym_root", "ssgpu_ExtractComponentC", "ssgpu_loglik_batch", "ssgpu_loglik_batch_dev",
           "ssgpu_last_kernel", "ssgpu_fp64_peak", "ssgpu_plan_state_order", "ssgpu_SimulateC_dev",
           "ssgpu_FastSmootherC_dev", "ssgpu_ExtractComponentC_dev", "ssgpu_draw_layout_dev", "ssgpu_last_sim_kernel_ms",
           "ssgpu_loglik_grad_batch", "ssgpu_loglik_grad_batch_dev", "ssgpu_loglik_theta_batch_dev",
           "ssgpu_SimSmoother", "ssgpu_kalman_batch", "ssgpu_kalman_batch_dev", "ssgpu_collapse", "ssgpu_forecast", "ssgpu_SimulateModel", "ssgpu_plan_blocks", "ssgpu_core_flops_per_step", "ssgpu_plan_thread_kernel", "ssgpu_plan_core",
           "ssgpu_use_devices", "ssgpu_devices_in_use", "ssgpu_loglik_cov_batch_dev"]

KERNEL_AUTO, KERNEL_GENERIC, KERNEL_COLS, KERNEL_MMA, KERNEL_MMA_DENSE, KERNEL_WSM, KERNEL_BLK, KERNEL_BLK_LANES, KERNEL_CORE = 0, 1, 2, 3, 4, 5, 6, 7, 8


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                              "statespacer_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.ssgpu_last_error.restype = C.c_char_p
        L.ssgpu_last_kernel.restype = C.c_char_p
        L.ssgpu_launch_count.restype = C.c_longlong
        i, ll, vp = C.c_int, C.c_longlong, C.c_void_p
        L.ssgpu_LogLikC.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, i, vp, i, vp, i, vp, i, dp]
        L.ssgpu_KalmanC.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, i, vp, i, vp, i, vp, i, i,
                                    C.POINTER(KalmanOut)]
        L.ssgpu_FastSmootherC.argtypes = [vp, i, i, i, i, i, vp, vp, vp, vp, i, vp, i, vp, i, vp, i, i, i, vp, vp, vp]
        L.ssgpu_SimulateC.argtypes = [i] * 8 + [vp, vp, i, vp, i, vp, i, vp, i, vp, i, i, i, vp, ll, vp, vp,
                                                vp, vp, vp, vp]
        L.ssgpu_SimulateC_dev.argtypes = [i] * 8 + [vp, vp, i, vp, i, vp, i, vp, i, vp, i, i, vp, ll, vp, vp,
                                                    vp, vp, vp, vp]
        L.ssgpu_FastSmootherC_dev.argtypes = [vp, i, i, i, i, i, vp, vp, vp, vp, i, vp, i, vp, i, vp, i, i, vp, vp, vp]
        L.ssgpu_ExtractComponentC_dev.argtypes = [vp, i, i, i, vp, i, i, vp, vp]
        L.ssgpu_draw_layout_dev.argtypes = [vp, vp, i, i, ll, i, vp]
        L.ssgpu_last_sim_kernel_ms.argtypes = [dp, dp, dp]
        L.ssgpu_sym_root.argtypes = [vp, i, vp]
        L.ssgpu_ExtractComponentC.argtypes = [vp, i, i, i, vp, i, i, vp]
        L.ssgpu_loglik_batch.argtypes = [vp, ll, i, C.POINTER(Model), i, vp]
        L.ssgpu_loglik_batch_dev.argtypes = [vp, ll, i, C.POINTER(Model), i, vp, vp]
        L.ssgpu_fp64_peak.argtypes = [i, i, i, dp, dp]
        d = C.c_double
        L.ssgpu_loglik_grad_batch.argtypes = [vp, ll, C.POINTER(Model), vp, i, vp, vp, d, i, vp, vp]
        L.ssgpu_loglik_grad_batch_dev.argtypes = [vp, ll, C.POINTER(Model), vp, i, vp, vp, d, i, vp, vp, vp]
        L.ssgpu_loglik_theta_batch_dev.argtypes = [vp, ll, i, C.POINTER(Model), vp, i, vp, vp, i, vp, vp]
        L.ssgpu_plan_state_order.argtypes = [vp, i, vp, vp, vp]
        L.ssgpu_SimulateModel.argtypes = [i, i, i, i, C.POINTER(SimComponent), vp, i, vp, ll, vp, vp, vp, vp, i, vp, vp, vp]
        L.ssgpu_plan_blocks.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp]
        L.ssgpu_core_flops_per_step.argtypes = [i, i]
        L.ssgpu_plan_thread_kernel.argtypes = [vp] * 7 + [i, i, i] + [vp] * 6
        L.ssgpu_plan_core.argtypes = [vp] * 7 + [i, i, i] + [vp] * 4
        L.ssgpu_collapse.argtypes = [vp, i, i, i, vp, i, vp, i, vp, vp, vp]
        L.ssgpu_forecast.argtypes = [i, i, i, i, vp, vp, vp, i, vp, i, vp, i, vp, i, vp, i, vp, vp, vp, vp]
        L.ssgpu_kalman_batch.argtypes = [vp, ll, C.POINTER(Model), C.POINTER(KalmanOut)]
        L.ssgpu_kalman_batch_dev.argtypes = [vp, ll, C.POINTER(Model), C.POINTER(KalmanOut), vp]
        L.ssgpu_SimSmoother.argtypes = [i, i, i, i, C.POINTER(SimComponent), vp, i, C.POINTER(SimModel), i, vp, vp, vp, vp,
                                        ll, vp, vp, vp, i, vp, vp, vp]
        L.ssgpu_device_count.argtypes = [ip]
        L.ssgpu_set_device.argtypes = [i]
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise SsgpuError(status, lib().ssgpu_last_error().decode())
