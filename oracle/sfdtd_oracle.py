"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/sfdtd_oracle.c.

Exposes ``forward_fn`` with the signature, return list and in-place side effects
of the reference extension (reference src/model/cpp/simulator.cpp:14-59), fp64 on
CPU.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sfdtd_oracle.c")
LIB = os.path.join(HERE, "_build", "libsfdtd_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        # -ffp-contract=off: no FMA contraction, keep the reference's rounding points
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", LIB, SRC, "-lm"])
    return LIB


class _Args(ctypes.Structure):
    _P = ctypes.c_void_p
    _fields_ = [
        ("B", ctypes.c_int32), ("Nt", ctypes.c_int32), ("Nx_t1", ctypes.c_int32), ("Nx_l1", ctypes.c_int32),
        ("state_u", _P), ("state_z", _P),
        ("kappa", _P), ("alpha", _P), ("p_a", _P), ("f0", _P), ("pos", _P), ("T60", _P),
        ("x_b", _P), ("v_b", _P), ("F_b", _P), ("wid", _P), ("phi_0", _P), ("phi_1", _P),
        ("x_H", _P), ("w_H", _P), ("M_r", _P), ("alpha_H", _P), ("u_H", _P),
        ("bow_mask", _P), ("hammer_mask", _P),
        ("k", ctypes.c_float), ("theta_t", ctypes.c_float), ("lambda_c", ctypes.c_float),
        ("relative_order", ctypes.c_float),
        ("surface_integral", ctypes.c_int32), ("manufactured", ctypes.c_int32), ("n_0", ctypes.c_int32),
        ("uout", _P), ("zout", _P), ("v_r", _P), ("F_H", _P), ("u_H_out", _P), ("sig0", _P), ("sig1", _P),
        ("stats", _P), ("max_iter", ctypes.c_int32), ("solver", ctypes.c_int32),
    ]


_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.sfdtd_oracle_forward.argtypes = [ctypes.POINTER(_Args)]
        _lib.sfdtd_oracle_forward.restype = ctypes.c_int
    return _lib


def _c(t, shape=None):
    """contiguous fp64 CPU copy (or the tensor itself when already so)."""
    t = t.detach().to("cpu", torch.float64)
    if shape is not None:
        t = t.expand(shape)
    return t.contiguous()


last_stats = None


def forward_fn(state_u, state_z, string_params, bow_params, hammer_params,
               bow_mask, hammer_mask, constant, relative_error,
               surface_integral, manufactured, n_0, Nt, max_iter=1000, solver=0):
    """Drop-in for the reference ``forward_fn`` (simulator.cpp:14-27). fp64 only."""
    global last_stats
    lib = _get_lib()
    B, Nt_c, Nx_t1 = state_u.shape
    Nx_l1 = state_z.shape[2]
    assert Nt == Nt_c, (Nt, Nt_c)
    kappa, alpha, u0, v0, p_a, f0, pos, T60 = string_params
    x_b, v_b, F_b, phi_0, phi_1, wid = bow_params
    x_H, v_H, u_H, w_H, M_r, alpha_H = hammer_params

    su = _c(state_u); sz = _c(state_z)
    uH = _c(u_H, (B, Nt))
    keep = [su, sz, uH]

    def P(t, shape=None):
        c = _c(t, shape)
        keep.append(c)
        return c.data_ptr()

    outs = {n: torch.zeros(B, Nt, dtype=torch.float64) for n in ["uout", "zout", "v_r", "F_H", "u_H_out"]}
    sig0 = torch.zeros(B, dtype=torch.float64); sig1 = torch.zeros(B, dtype=torch.float64)
    stats = np.zeros(8, dtype=np.int64)
    a = _Args()
    a.B, a.Nt, a.Nx_t1, a.Nx_l1 = B, Nt, Nx_t1, Nx_l1
    a.state_u, a.state_z = su.data_ptr(), sz.data_ptr()
    a.kappa = P(kappa.reshape(-1)); a.alpha = P(alpha.reshape(-1)); a.p_a = P(p_a.reshape(-1))
    a.f0 = P(f0, (B, Nt)); a.pos = P(pos.reshape(-1)); a.T60 = P(T60.reshape(B, 4))
    a.x_b = P(x_b, (B, Nt)); a.v_b = P(v_b, (B, Nt)); a.F_b = P(F_b, (B, Nt)); a.wid = P(wid, (B, Nt))
    a.phi_0 = P(phi_0.reshape(-1)); a.phi_1 = P(phi_1.reshape(-1))
    a.x_H = P(x_H.reshape(-1)); a.w_H = P(w_H.reshape(-1)); a.M_r = P(M_r.reshape(-1))
    a.alpha_H = P(alpha_H.reshape(-1)); a.u_H = uH.data_ptr()
    bm = bow_mask.reshape(-1).to(torch.uint8).contiguous(); hm = hammer_mask.reshape(-1).to(torch.uint8).contiguous()
    a.bow_mask, a.hammer_mask = bm.data_ptr(), hm.data_ptr()
    a.k, a.theta_t, a.lambda_c = float(constant[0]), float(constant[1]), float(constant[2])
    a.relative_order = float(relative_error)
    a.surface_integral, a.manufactured, a.n_0 = int(bool(surface_integral)), int(bool(manufactured)), int(n_0)
    a.uout, a.zout, a.v_r = outs["uout"].data_ptr(), outs["zout"].data_ptr(), outs["v_r"].data_ptr()
    a.F_H, a.u_H_out = outs["F_H"].data_ptr(), outs["u_H_out"].data_ptr()
    a.sig0, a.sig1 = sig0.data_ptr(), sig1.data_ptr()
    a.stats = stats.ctypes.data
    a.max_iter = max_iter
    a.solver = solver
    rc = lib.sfdtd_oracle_forward(ctypes.byref(a))
    if rc < 0:
        raise RuntimeError(f"sfdtd_oracle_forward failed with status {rc}")
    last_stats = dict(outer_total=int(stats[0]), outer_max=int(stats[1]), hammer_total=int(stats[2]),
                      hammer_max=int(stats[3]), steps=int(stats[4]), capped=(rc == 1),
                      gs_sweeps=int(stats[5]), gs_sweeps_max=int(stats[6]), gs_solves=int(stats[7]))
    # in-place side effects of the reference (string.cpp:264-265,303)
    if su.data_ptr() != state_u.data_ptr():
        state_u.copy_(su.to(state_u.dtype))
    if sz.data_ptr() != state_z.data_ptr():
        state_z.copy_(sz.to(state_z.dtype))
    if uH.data_ptr() != u_H.data_ptr():
        u_H.copy_(uH.to(u_H.dtype))
    return [outs["uout"], outs["zout"], state_u, state_z, outs["v_r"], outs["F_H"], outs["u_H_out"],
            sig0.view(-1, 1, 1), sig1.view(-1, 1, 1)]
