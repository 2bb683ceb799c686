"""ctypes binding of libfwi_b200.so (C ABI declared in include/fwi_b200.h).

The shared library is built IN-TREE by ``fwiflow/jl_b200/csrc/Makefile`` (nvcc, sm_100a).
There is no Python / CPU fallback: if the library is missing or a CUDA device is not
usable, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# FWI_B200_LIB: developer override used to time kernel variants built under variants/ (always a CUDA build of csrc/)
LIB_PATH = os.environ.get("FWI_B200_LIB") or os.path.join(HERE, "libfwi_b200.so")

c_dp = ctypes.POINTER(ctypes.c_double)
c_fp = ctypes.POINTER(ctypes.c_float)
c_ip = ctypes.POINTER(ctypes.c_int)

# every symbol include/fwi_b200.h declares
SYMBOLS = [
    "fwi_b200_cufd", "fwi_b200_forward", "fwi_b200_backward", "fwi_b200_obscalc",
    "fwi_b200_misfit_and_gradient", "fwi_b200_gradient_multi", "fwi_b200_last_error", "fwi_b200_release",
    "fwi_b200_plan_create", "fwi_b200_plan_destroy", "fwi_b200_plan_set_model", "fwi_b200_plan_set_stf",
    "fwi_b200_plan_set_obs", "fwi_b200_plan_load_obs_files", "fwi_b200_plan_run",
    "fwi_b200_plan_result_device", "fwi_b200_plan_result_count", "fwi_b200_plan_get_result",
    "fwi_b200_plan_get_traces", "fwi_b200_plan_write_obs_files", "fwi_b200_plan_info",
    "fwi_b200_plan_shot_geometry", "fwi_b200_plan_launch_count", "fwi_b200_plan_get_field",
    "fwi_b200_plan_time_kernel", "fwi_b200_version", "fwi_b200_grid_info", "fwi_b200_para_info",
    "fwi_b200_timelapse", "fwi_b200_set_option", "fwi_b200_cufd_ex", "fwi_b200_plan_set_layout",
    "fwi_b200_plan_set_velocities", "fwi_b200_plan_get_velocity_gradients",
]

ERRORS = {-1: "ERR_ARG", -2: "ERR_IO", -3: "ERR_JSON", -4: "ERR_CFL", -5: "ERR_CUDA", -6: "ERR_UNSUPPORTED",
          -7: "ERR_GEOM"}


class FwiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fwi_b200 {ERRORS.get(code, code)}: {msg}")
        self.code = code


def build(force=False, extra=""):
    """Compile libfwi_b200.so for sm_100a (cross-compiles without a GPU)."""
    csrc = os.path.join(HERE, "csrc")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    cmd = ["make", "-C", csrc]
    if extra:
        cmd.append(f"EXTRA={extra}")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FwiError(-5, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the FWI path)")
    L = ctypes.CDLL(LIB_PATH)
    host_sig = [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                c_ip, ctypes.c_char_p]
    L.fwi_b200_cufd.argtypes = host_sig
    L.fwi_b200_cufd_ex.argtypes = host_sig + [ctypes.c_int, ctypes.c_int]
    L.fwi_b200_plan_set_layout.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.fwi_b200_plan_set_velocities.argtypes = [ctypes.c_void_p] + [c_dp] * 6 + [ctypes.c_int, ctypes.c_int]
    L.fwi_b200_plan_get_velocity_gradients.argtypes = [ctypes.c_void_p] + [c_dp] * 4
    L.fwi_b200_misfit_and_gradient.argtypes = [c_dp] * 9 + [ctypes.c_int, ctypes.c_int, c_ip, ctypes.c_char_p]
    L.fwi_b200_gradient_multi.argtypes = [c_dp] * 9 + [ctypes.c_int, c_ip, ctypes.c_int, c_ip, ctypes.c_char_p]
    L.fwi_b200_grid_info.argtypes = [ctypes.c_char_p, c_ip]
    L.fwi_b200_para_info.argtypes = [ctypes.c_char_p, c_ip]
    L.fwi_b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    pp = ctypes.POINTER(c_dp)
    L.fwi_b200_timelapse.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), pp, pp, pp, c_dp, ctypes.c_int, c_ip,
                                     ctypes.c_int, c_ip, c_dp, pp, pp, pp]
    L.fwi_b200_forward.argtypes = [c_dp] * 5 + [ctypes.c_int, ctypes.c_int, c_ip, ctypes.c_char_p]
    L.fwi_b200_obscalc.argtypes = [c_dp] * 5 + [ctypes.c_int, ctypes.c_int, c_ip, ctypes.c_char_p]
    L.fwi_b200_backward.argtypes = [c_dp] * 8 + [ctypes.c_int, ctypes.c_int, c_ip, ctypes.c_char_p]
    L.fwi_b200_last_error.restype = ctypes.c_char_p
    L.fwi_b200_version.restype = ctypes.c_char_p
    L.fwi_b200_release.restype = None
    vp = ctypes.c_void_p
    L.fwi_b200_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int, ctypes.c_int, c_ip,
                                       ctypes.c_int]
    L.fwi_b200_plan_destroy.argtypes = [vp]
    L.fwi_b200_plan_destroy.restype = None
    L.fwi_b200_plan_set_model.argtypes = [vp, c_dp, c_dp, c_dp]
    L.fwi_b200_plan_set_stf.argtypes = [vp, c_dp]
    L.fwi_b200_plan_set_obs.argtypes = [vp, ctypes.c_int, c_fp]
    L.fwi_b200_plan_load_obs_files.argtypes = [vp]
    L.fwi_b200_plan_run.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int]
    L.fwi_b200_plan_result_device.argtypes = [vp]
    L.fwi_b200_plan_result_device.restype = vp
    L.fwi_b200_plan_result_count.argtypes = [vp]
    L.fwi_b200_plan_result_count.restype = ctypes.c_size_t
    L.fwi_b200_plan_get_result.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.fwi_b200_plan_get_traces.argtypes = [vp, ctypes.c_int, ctypes.c_int, c_fp]
    L.fwi_b200_plan_write_obs_files.argtypes = [vp]
    L.fwi_b200_plan_info.argtypes = [vp] + [c_ip] * 8
    L.fwi_b200_plan_shot_geometry.argtypes = [vp, ctypes.c_int, c_ip, c_ip, c_ip, c_ip, c_ip]
    L.fwi_b200_plan_launch_count.argtypes = [vp]
    L.fwi_b200_plan_launch_count.restype = ctypes.c_longlong
    L.fwi_b200_plan_get_field.argtypes = [vp, ctypes.c_int, ctypes.c_int, c_fp]
    L.fwi_b200_plan_time_kernel.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, c_fp, c_dp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise FwiError(rc, lib().fwi_b200_last_error().decode(errors="replace"))
