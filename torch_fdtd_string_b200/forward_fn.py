"""Host-side mirror of the reference's native entry point, on top of the C ABI (include/sfdtd.h).

``forward_fn`` has the signature, return list and in-place side effects of the reference's
pybind function (reference src/model/cpp/simulator.cpp:14-27, 57-58): a module object exposing
it can be handed to the reference's ``process()`` in place of the JIT-built extension
(src/task/simulate.py:28-36).  PyTorch is used for device memory and streams only; all
arithmetic happens in libsfdtd.so.  There is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import Args, Array

_xax_cache = {}


def make_xax(Nx_t1, device):
    """float32 ``torch.linspace(1/N, 1, N)`` exactly as the reference builds it on the host
    (reference src/model/cpp/misc.cpp:26-27: ``float h = 1. / N; linspace(h, 1, N)``)."""
    key = (int(Nx_t1), str(device))
    if key not in _xax_cache:
        h = float(np.float32(1.0 / Nx_t1))
        _xax_cache[key] = torch.linspace(h, 1, int(Nx_t1), dtype=torch.float32).to(device)
    return _xax_cache[key]


def launch_count():
    return int(_lib.load().sfdtd_launch_count())


def _arr(t, batch_dim=0, time_dim=None):
    a = Array()
    a.ptr = t.data_ptr()
    a.bs = t.stride(batch_dim) if t.dim() > 0 else 0
    a.ts = t.stride(time_dim) if time_dim is not None else 0
    return a


def _time_arr(t, B, Nt):
    """(B,Nt) control curve; broadcastable shapes (B,), (B,1), (1,Nt), () give zero strides."""
    if t.dim() == 0:
        t = t.view(1, 1)
    if t.dim() == 1:
        t = t.view(-1, 1)
    t = t.expand(B, Nt) if (t.size(0) != B or t.size(1) != Nt) else t
    a = Array()
    a.ptr = t.data_ptr(); a.bs = t.stride(0); a.ts = t.stride(1)
    return a


def _vec_arr(t, B):
    t = t.reshape(-1)
    if t.numel() == 1 and B > 1:
        t = t.expand(B)
    assert t.numel() == B, (t.shape, B)
    a = Array()
    a.ptr = t.data_ptr(); a.bs = t.stride(0); a.ts = 0
    return a


def _call(args, stream=None):
    lib = _lib.load()
    s = torch.cuda.current_stream() if stream is None else stream
    rc = lib.sfdtd_forward(ctypes.byref(args), ctypes.c_void_p(s.cuda_stream))
    if rc != 0:
        raise RuntimeError(f"sfdtd_forward failed ({rc}): {lib.sfdtd_last_error().decode()}")


def step_strings(state_u, state_z, *, kappa, alpha, f0, pos, T60, x_b, v_b, F_b, wid, phi_0, phi_1,
                 x_H, w_H, M_r, alpha_H, u_H, bow_mask, hammer_mask, k, theta_t, lambda_c,
                 relative_order, Nt, group_size, surface_integral=True, save_state=False, skip_aux=False,
                 manufactured=False, n_0=0, p_a=None, max_iter=100, out=None, counters=False, stream=None, check=True):
    """Native API.  All tensors are float64 CUDA tensors.

    state_u/state_z: (B, Nt, Nx) when ``save_state`` (reference layout, updated in place), else
    (B, 2, Nx) holding rows [n-2, n-1] (overwritten with the last two rows).  Control curves
    (f0, x_b, v_b, F_b, wid) are (B, Nt) or anything broadcastable to it (time stride 0 allowed);
    u_H is (B, Nt) and updated in place.  Returns dict(uout, zout, v_r, F_H, u_H_out (B,Nt),
    sig0, sig1 (B), status (B) int32[, counters (B,4) int64]).
    """
    dev = state_u.device
    assert dev.type == "cuda", "the stepper runs on CUDA only"
    for t in (state_u, state_z, u_H):
        assert t.dtype == torch.float64 and t.stride(-1) == 1
    B, Nx_t1, Nx_l1 = state_u.size(0), state_u.size(2), state_z.size(2)
    f64 = dict(dtype=torch.float64, device=dev)
    if out is None:
        out = {n: torch.zeros(B, Nt, **f64) for n in ("uout", "zout", "v_r", "F_H", "u_H_out")}
    sig0 = torch.zeros(B, **f64); sig1 = torch.zeros(B, **f64)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cnt = torch.zeros(B, 4, dtype=torch.int64, device=dev) if counters else None
    bm = bow_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    hm = hammer_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    xax = make_xax(Nx_t1, dev)
    T60c = T60.reshape(B, 4).contiguous()
    if p_a is None:
        p_a = torch.zeros(B, **f64)
    keep = [bm, hm, xax, T60c, p_a]

    a = Args()
    a.abi_version = _lib.SFDTD_ABI_VERSION
    a.dtype = _lib.SFDTD_F64
    a.flags = ((_lib.SURFACE_INTEGRAL if surface_integral else 0) | (_lib.SAVE_STATE if save_state else 0)
               | (_lib.SKIP_AUX if skip_aux else 0) | (_lib.MANUFACTURED if manufactured else 0))
    a.B, a.group_size, a.Nt, a.Nx_t1, a.Nx_l1 = B, int(group_size), int(Nt), Nx_t1, Nx_l1
    a.n_0, a.max_iter = int(n_0), int(max_iter)
    a.k, a.theta_t, a.lambda_c, a.relative_order = float(k), float(theta_t), float(lambda_c), float(relative_order)
    a.state_u = _arr(state_u, 0, 1); a.state_z = _arr(state_z, 0, 1)
    a.kappa = _vec_arr(kappa, B); a.alpha = _vec_arr(alpha, B); a.p_a = _vec_arr(p_a, B); a.pos = _vec_arr(pos, B)
    a.f0 = _time_arr(f0, B, Nt)
    a.T60 = Array(); a.T60.ptr = T60c.data_ptr(); a.T60.bs = 4; a.T60.ts = 0
    a.x_b = _time_arr(x_b, B, Nt); a.v_b = _time_arr(v_b, B, Nt); a.F_b = _time_arr(F_b, B, Nt); a.wid = _time_arr(wid, B, Nt)
    a.phi_0 = _vec_arr(phi_0, B); a.phi_1 = _vec_arr(phi_1, B)
    a.x_H = _vec_arr(x_H, B); a.w_H = _vec_arr(w_H, B); a.M_r = _vec_arr(M_r, B); a.alpha_H = _vec_arr(alpha_H, B)
    assert u_H.shape == (B, Nt)
    a.u_H = _arr(u_H, 0, 1)
    a.bow_mask, a.hammer_mask, a.xax = bm.data_ptr(), hm.data_ptr(), xax.data_ptr()
    for n in ("uout", "zout", "v_r", "F_H", "u_H_out"):
        setattr(a, n, _arr(out[n], 0, 1))
    a.sig0, a.sig1, a.status = sig0.data_ptr(), sig1.data_ptr(), status.data_ptr()
    a.counters = cnt.data_ptr() if counters else None
    _call(a, stream)
    del keep
    if check:
        bits = 0
        for v in status.unique().tolist():
            bits |= int(v)
        if bits & (_lib.ST_RANGE | _lib.ST_BOW_WINDOW):
            raise RuntimeError(f"sfdtd: configuration outside the supported range (status bits {bits:#x}: "
                               "0x8 = bow window wider than the kernel supports, 0x10 = grid size out of range)")
    res = dict(out)
    res.update(sig0=sig0, sig1=sig1, status=status)
    if counters:
        res["counters"] = cnt
    return res


def forward_fn(state_u, state_z, string_params, bow_params, hammer_params,
               bow_mask, hammer_mask, constant, relative_error,
               surface_integral, manufactured, n_0, Nt):
    """Drop-in for the reference extension's ``forward_fn`` (simulator.cpp:14-27).

    Same positional arguments; returns ``[uout, zout, state_u, state_z, v_r, F_H, u_H/k, sig0, sig1]``
    and, like the reference, updates ``state_u``, ``state_z`` and ``hammer_params[2]`` in place.
    All strings of the call form one group (the reference's batch).  Tensors may live on the CPU
    or on CUDA and be float32 or float64: the arithmetic is float64 on the GPU either way (float32
    tensors are widened on entry and rounded on exit), results are returned on CUDA like the
    reference's ``device()`` (misc.cpp:13-15) does when a GPU is visible.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("torch_fdtd_string_b200.forward_fn needs a CUDA device (no CPU fallback)")
    dev = state_u.device if state_u.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    kappa, alpha, u0, v0, p_a, f0, pos, T60 = string_params
    x_b, v_b, F_b, phi_0, phi_1, wid = bow_params
    x_H, v_H, u_H, w_H, M_r, alpha_H = hammer_params
    B, Nt_c, _ = state_u.shape
    assert Nt == Nt_c, (Nt, Nt_c)

    def D(t):   # float64 on the device; aliases the input when it already is
        return t.detach().to(device=dev, dtype=torch.float64)

    def Dio(t):  # in/out tensor: (device copy with unit space stride, needs_copy_back)
        d = D(t)
        if d.stride(-1) != 1:
            d = d.contiguous()
        return d, d.data_ptr() != t.data_ptr()

    su, su_cp = Dio(state_u); sz, sz_cp = Dio(state_z)
    uH = D(u_H)
    if uH.dim() != 2 or uH.shape != (B, Nt):
        uH = uH.expand(B, Nt).contiguous()
    uH_cp = uH.data_ptr() != u_H.data_ptr()
    res = step_strings(
        su, sz, kappa=D(kappa), alpha=D(alpha), f0=D(f0), pos=D(pos), T60=D(T60),
        x_b=D(x_b), v_b=D(v_b), F_b=D(F_b), wid=D(wid), phi_0=D(phi_0), phi_1=D(phi_1),
        x_H=D(x_H), w_H=D(w_H), M_r=D(M_r), alpha_H=D(alpha_H), u_H=uH,
        bow_mask=bow_mask, hammer_mask=hammer_mask,
        k=constant[0], theta_t=constant[1], lambda_c=constant[2], relative_order=relative_error,
        Nt=Nt, group_size=B, surface_integral=bool(surface_integral), save_state=True, manufactured=bool(manufactured), n_0=n_0,
        p_a=D(p_a))
    # in-place side effects of the reference (string.cpp:264-265, 303)
    if su_cp: state_u.copy_(su)
    if sz_cp: state_z.copy_(sz)
    if uH_cp: u_H.copy_(uH)
    odt = state_u.dtype
    o = [res[n].to(odt) for n in ("uout", "zout")]
    return [o[0], o[1], state_u, state_z, res["v_r"].to(odt), res["F_H"].to(odt), res["u_H_out"].to(odt),
            res["sig0"].to(odt).view(-1, 1, 1), res["sig1"].to(odt).view(-1, 1, 1)]
