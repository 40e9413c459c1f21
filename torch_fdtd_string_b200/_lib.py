"""ctypes binding of libsfdtd.so (include/sfdtd.h).  There is NO fallback: if the CUDA
library is missing or fails to load, importing the stepper raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFDTD_LIB") or os.path.join(HERE, "libsfdtd.so")     # SFDTD_LIB: A/B builds of the same ABI

SFDTD_ABI_VERSION = 2
SFDTD_F64, SFDTD_F32 = 0, 1
SURFACE_INTEGRAL, MANUFACTURED, SAVE_STATE, SKIP_AUX = 1, 2, 4, 8
ST_SOLVER_CAP, ST_OUTER_CAP, ST_HAMMER_CAP, ST_BOW_WINDOW, ST_RANGE = 1, 2, 4, 8, 16

EXPORTS = ["sfdtd_forward", "sfdtd_last_error", "sfdtd_abi_version", "sfdtd_launch_count",
           "sfdtd_measure_fma_peak", "sfdtd_plan_create", "sfdtd_forward_plan", "sfdtd_plan_destroy",
           "sfdtd_synth_controls", "sfdtd_postprocess", "sfdtd_postprocess_f32"]


class Array(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("bs", ctypes.c_int64), ("ts", ctypes.c_int64)]


class Args(ctypes.Structure):
    _fields_ = (
        [("abi_version", ctypes.c_int32), ("dtype", ctypes.c_int32), ("flags", ctypes.c_uint32),
         ("B", ctypes.c_int32), ("group_size", ctypes.c_int32), ("Nt", ctypes.c_int32),
         ("Nx_t1", ctypes.c_int32), ("Nx_l1", ctypes.c_int32), ("n_0", ctypes.c_int32),
         ("max_iter", ctypes.c_int32),
         ("k", ctypes.c_float), ("theta_t", ctypes.c_float), ("lambda_c", ctypes.c_float),
         ("relative_order", ctypes.c_float)]
        + [(n, Array) for n in ("state_u", "state_z", "kappa", "alpha", "p_a", "f0", "pos", "T60",
                                "x_b", "v_b", "F_b", "wid", "phi_0", "phi_1",
                                "x_H", "w_H", "M_r", "alpha_H", "u_H")]
        + [("bow_mask", ctypes.c_void_p), ("hammer_mask", ctypes.c_void_p), ("xax", ctypes.c_void_p)]
        + [(n, Array) for n in ("uout", "zout", "v_r", "F_H", "u_H_out")]
        + [("sig0", ctypes.c_void_p), ("sig1", ctypes.c_void_p), ("status", ctypes.c_void_p),
           ("counters", ctypes.c_void_p), ("synth", ctypes.c_void_p)]
    )


SYNTH_KEYS = ["f0_a", "f0_b", "mod_frq", "mod_amp", "vib_t0", "x_b1", "x_b2", "v_b1", "v_b2", "F_b1", "F_b2",
              "pulloff", "wid", "v_H"]


class Synth(ctypes.Structure):
    """struct sfdtd_synth: per-string scalars the stepper synthesises the control curves from"""
    _fields_ = ([("Nt_full", ctypes.c_int32), ("t_0", ctypes.c_int32), ("sr", ctypes.c_double)]
                + [(n, ctypes.c_void_p) for n in SYNTH_KEYS])


_lib = None


def load():
    """Loads libsfdtd.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m torch_fdtd_string_b200.build` "
                "(there is no CPU or PyTorch fallback for the stepper)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.sfdtd_forward.argtypes = [ctypes.POINTER(Args), ctypes.c_void_p]
        lib.sfdtd_forward.restype = ctypes.c_int
        lib.sfdtd_last_error.restype = ctypes.c_char_p
        lib.sfdtd_abi_version.restype = ctypes.c_int
        lib.sfdtd_launch_count.restype = ctypes.c_int64
        lib.sfdtd_measure_fma_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        lib.sfdtd_measure_fma_peak.restype = ctypes.c_int
        lib.sfdtd_plan_create.argtypes = [ctypes.POINTER(Args), ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.sfdtd_plan_create.restype = ctypes.c_int
        lib.sfdtd_forward_plan.argtypes = [ctypes.c_void_p, ctypes.POINTER(Args), ctypes.c_void_p]
        lib.sfdtd_forward_plan.restype = ctypes.c_int
        lib.sfdtd_plan_destroy.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.sfdtd_plan_destroy.restype = ctypes.c_int
        lib.sfdtd_synth_controls.argtypes = [ctypes.POINTER(Args)] + [ctypes.POINTER(Array)] * 5 + [ctypes.c_void_p]
        lib.sfdtd_synth_controls.restype = ctypes.c_int
        lib.sfdtd_postprocess.argtypes = [ctypes.POINTER(Array), ctypes.POINTER(Array), ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int32, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.sfdtd_postprocess.restype = ctypes.c_int
        lib.sfdtd_postprocess_f32.argtypes = lib.sfdtd_postprocess.argtypes
        lib.sfdtd_postprocess_f32.restype = ctypes.c_int
        if lib.sfdtd_abi_version() != SFDTD_ABI_VERSION:
            raise RuntimeError("libsfdtd.so ABI version mismatch")
        _lib = lib
    return _lib
