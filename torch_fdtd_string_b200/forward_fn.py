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


def _synth_struct(synth, B, dev, keep):
    """dict of the compact control description (sampler keys) -> ctypes sfdtd_synth with device pointers"""
    y = _lib.Synth()
    y.Nt_full = int(synth["Nt_full"]); y.t_0 = int(synth.get("t_0", 0)); y.sr = float(synth["sr"])
    for n in _lib.SYNTH_KEYS:
        t = synth[n].reshape(-1).to(device=dev, dtype=torch.float64).contiguous()
        assert t.numel() == B, (n, t.shape, B)
        keep.append(t)
        setattr(y, n, t.data_ptr())
    keep.append(y)
    return y


def build_args(state_u, state_z, *, kappa, alpha, pos, T60, phi_0, phi_1, x_H, w_H, M_r, alpha_H, bow_mask, hammer_mask,
               k, theta_t, lambda_c, relative_order, Nt, group_size, f0=None, x_b=None, v_b=None, F_b=None, wid=None,
               u_H=None, synth=None, surface_integral=True, save_state=False, skip_aux=False, manufactured=False, n_0=0,
               p_a=None, max_iter=100, out=None, counters=False, aux_outputs=True):
    """Fills a ``sfdtd_args`` for the C ABI.  Returns (args, results dict, keep-alive list).

    Control curves come either as (B,Nt)-broadcastable tensors (f0, x_b, v_b, F_b, wid, u_H) or as ``synth``: a dict of the
    per-string scalars of ``sfdtd_synth`` (keys ``_lib.SYNTH_KEYS`` + Nt_full, sr[, t_0]) that the stepper evaluates
    itself.  ``aux_outputs=False`` leaves v_r / F_H / u_H_out unallocated (audio only)."""
    dev = state_u.device
    assert dev.type == "cuda", "the stepper runs on CUDA only"
    # the arithmetic type of the call is the dtype of the state: float64 (SFDTD_F64, parity mode) or float32 (SFDTD_F32,
    # the reference's `precision: single`); states, u_H and the (B,Nt) outputs are of that type, parameters are float64
    sdt = state_u.dtype
    assert sdt in (torch.float64, torch.float32), sdt
    for t in (state_u, state_z):
        assert t.dtype == sdt and t.stride(-1) == 1
    B, Nx_t1, Nx_l1 = state_u.size(0), state_u.size(2), state_z.size(2)
    f64 = dict(dtype=torch.float64, device=dev)
    names = ("uout", "zout", "v_r", "F_H", "u_H_out") if aux_outputs else ("uout", "zout")
    if out is None:
        out = {n: torch.zeros(B, Nt, dtype=sdt, device=dev) for n in names}
    for n in names:
        assert out[n].dtype == sdt, (n, out[n].dtype, sdt)
    sig0 = torch.zeros(B, **f64); sig1 = torch.zeros(B, **f64)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cnt = torch.zeros(B, 4, dtype=torch.int64, device=dev) if counters else None
    bm = bow_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    hm = hammer_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    xax = make_xax(Nx_t1, dev)
    T60c = T60.reshape(B, 4).contiguous()
    if p_a is None:
        p_a = torch.zeros(B, **f64)
    keep = [bm, hm, xax, T60c, p_a, sig0, sig1, status, cnt, out]

    a = Args()
    a.abi_version = _lib.SFDTD_ABI_VERSION
    a.dtype = _lib.SFDTD_F64 if sdt == torch.float64 else _lib.SFDTD_F32
    a.flags = ((_lib.SURFACE_INTEGRAL if surface_integral else 0) | (_lib.SAVE_STATE if save_state else 0)
               | (_lib.SKIP_AUX if skip_aux else 0) | (_lib.MANUFACTURED if manufactured else 0))
    a.B, a.group_size, a.Nt, a.Nx_t1, a.Nx_l1 = B, int(group_size), int(Nt), Nx_t1, Nx_l1
    a.n_0, a.max_iter = int(n_0), int(max_iter)
    a.k, a.theta_t, a.lambda_c, a.relative_order = float(k), float(theta_t), float(lambda_c), float(relative_order)
    a.state_u = _arr(state_u, 0, 1); a.state_z = _arr(state_z, 0, 1)
    a.kappa = _vec_arr(kappa, B); a.alpha = _vec_arr(alpha, B); a.p_a = _vec_arr(p_a, B); a.pos = _vec_arr(pos, B)
    a.T60 = Array(); a.T60.ptr = T60c.data_ptr(); a.T60.bs = 4; a.T60.ts = 0
    a.phi_0 = _vec_arr(phi_0, B); a.phi_1 = _vec_arr(phi_1, B)
    a.x_H = _vec_arr(x_H, B); a.w_H = _vec_arr(w_H, B); a.M_r = _vec_arr(M_r, B); a.alpha_H = _vec_arr(alpha_H, B)
    if synth is None:
        a.f0 = _time_arr(f0, B, Nt)
        a.x_b = _time_arr(x_b, B, Nt); a.v_b = _time_arr(v_b, B, Nt); a.F_b = _time_arr(F_b, B, Nt); a.wid = _time_arr(wid, B, Nt)
        a.synth = None
    else:
        y = _synth_struct(synth, B, dev, keep)
        a.synth = ctypes.addressof(y)
    if u_H is not None:
        assert u_H.shape == (B, Nt) and u_H.dtype == sdt
        a.u_H = _arr(u_H, 0, 1)
    else:
        assert synth is not None, "u_H may only be omitted with synthesised controls"
    a.bow_mask, a.hammer_mask, a.xax = bm.data_ptr(), hm.data_ptr(), xax.data_ptr()
    for n in names:
        setattr(a, n, _arr(out[n], 0, 1))
    a.sig0, a.sig1, a.status = sig0.data_ptr(), sig1.data_ptr(), status.data_ptr()
    a.counters = cnt.data_ptr() if counters else None
    res = dict(out)
    res.update(sig0=sig0, sig1=sig1, status=status)
    if counters:
        res["counters"] = cnt
    return a, res, keep


def _check_status(status):
    bits = 0
    for v in status.unique().tolist():
        bits |= int(v)
    if bits & (_lib.ST_RANGE | _lib.ST_BOW_WINDOW):
        raise RuntimeError(f"sfdtd: configuration outside the supported range (status bits {bits:#x}: "
                           "0x8 = bow window wider than the kernel supports, 0x10 = grid size out of range)")


def step_strings(state_u, state_z, *, stream=None, check=True, **kw):
    """Native API.  CUDA tensors: states, u_H and outputs float64 or float32 (the call's arithmetic type), parameters and
    control curves float64 (see ``build_args`` for the arguments).

    state_u/state_z: (B, Nt, Nx) when ``save_state`` (reference layout, updated in place), else
    (B, 2, Nx) holding rows [n-2, n-1] (overwritten with the last two rows).  Control curves
    (f0, x_b, v_b, F_b, wid) are (B, Nt) or anything broadcastable to it (time stride 0 allowed), or
    ``synth`` scalars; u_H is (B, Nt) and updated in place.  Returns dict(uout, zout[, v_r, F_H, u_H_out]
    (B,Nt), sig0, sig1 (B), status (B) int32[, counters (B,4) int64]).  Asynchronous: the kernels are queued
    on the stream when this returns (``check=True`` reads the status words and therefore waits).
    """
    a, res, keep = build_args(state_u, state_z, **kw)
    with torch.cuda.device(state_u.device):
        _call(a, stream)
    if check:
        _check_status(res["status"])
    res["_keep"] = keep
    return res


class Plan:
    """``sfdtd_plan``: everything the stepper derives from the parameters before it can launch (one small device->host
    read).  ``run`` queues a call without any host synchronisation, so transfers and the next call can overlap."""

    def __init__(self, args, stream=None, keep=None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        s = torch.cuda.current_stream() if stream is None else stream
        rc = self._lib.sfdtd_plan_create(ctypes.byref(args), ctypes.c_void_p(s.cuda_stream), ctypes.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"sfdtd_plan_create failed ({rc}): {self._lib.sfdtd_last_error().decode()}")

    def run(self, args, stream=None):
        s = torch.cuda.current_stream() if stream is None else stream
        rc = self._lib.sfdtd_forward_plan(self._h, ctypes.byref(args), ctypes.c_void_p(s.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"sfdtd_forward_plan failed ({rc}): {self._lib.sfdtd_last_error().decode()}")

    def close(self, stream=None):
        if self._h:
            s = torch.cuda.current_stream() if stream is None else stream
            self._lib.sfdtd_plan_destroy(self._h, ctypes.c_void_p(s.cuda_stream))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_controls(synth, B, Nt, device, k=None):
    """The (B,Nt) control curves exactly as the stepper evaluates ``synth`` (sfdtd_synth_controls)."""
    lib = _lib.load()
    keep = []
    y = _synth_struct(synth, B, device, keep)
    a = Args()
    a.abi_version = _lib.SFDTD_ABI_VERSION; a.B = B; a.Nt = Nt; a.group_size = 1; a.Nx_t1 = 1; a.Nx_l1 = 1
    a.k = float(k if k is not None else 1.0 / synth["sr"])
    a.synth = ctypes.addressof(y)
    out = {n: torch.empty(B, Nt, dtype=torch.float64, device=device) for n in ("f0", "x_b", "v_b", "F_b", "u_H")}
    arrs = [_arr(out[n], 0, 1) for n in ("f0", "x_b", "v_b", "F_b", "u_H")]
    with torch.cuda.device(device):
        rc = lib.sfdtd_synth_controls(ctypes.byref(a), *[ctypes.byref(x) for x in arrs],
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(f"sfdtd_synth_controls failed ({rc}): {lib.sfdtd_last_error().decode()}")
    out["wid"] = synth["wid"].reshape(-1, 1).to(device=device, dtype=torch.float64).expand(B, Nt)
    return out


def postprocess(uout, zout, n0=2, n_samples=None, silence_db=-23.0, normalize=True, bits=24, pcm=("u", "z", "w"), stream=None,
                out=None):
    """Device-side NaN / silence flags, l-infinity gain and PCM quantisation of the audio (sfdtd_postprocess; reference
    src/task/simulate.py:333-337,416-425, src/utils/audio.py:42-48,72-76).  uout, zout: (B,Nt) float64 or float32 CUDA tensors.
    Returns dict(is_nan, is_silent (B) uint8, gain (B), pcm_u / pcm_z / pcm_w (B, pitch) uint8 rows of packed little-endian
    PCM_16 / PCM_24 samples, pitch = row bytes rounded up to 16)."""
    lib = _lib.load()
    B, Nt = uout.shape
    n_samples = Nt - n0 if n_samples is None else int(n_samples)
    dev = uout.device
    row = n_samples * (bits // 8)
    pitch = (row + 15) // 16 * 16
    res = dict(is_nan=torch.empty(B, dtype=torch.uint8, device=dev), is_silent=torch.empty(B, dtype=torch.uint8, device=dev),
               gain=torch.empty(B, dtype=torch.float64, device=dev), pitch=pitch, row_bytes=row)
    ptr = {}
    for kx in ("u", "z", "w"):
        if kx in pcm:
            res["pcm_" + kx] = out[kx] if out is not None else torch.empty(B, pitch, dtype=torch.uint8, device=dev)
            assert res["pcm_" + kx].shape == (B, pitch) and res["pcm_" + kx].is_contiguous()
            ptr[kx] = res["pcm_" + kx].data_ptr()
        else:
            ptr[kx] = None
    ua, za = _arr(uout, 0, 1), _arr(zout, 0, 1)
    s = torch.cuda.current_stream() if stream is None else stream
    with torch.cuda.device(dev):
        assert uout.dtype == zout.dtype and uout.dtype in (torch.float64, torch.float32)
        fn = lib.sfdtd_postprocess if uout.dtype == torch.float64 else lib.sfdtd_postprocess_f32
        rc = fn(ctypes.byref(ua), ctypes.byref(za), B, int(n0), n_samples, float(silence_db),
                                   1 if normalize else 0, int(bits), pitch, res["is_nan"].data_ptr(), res["is_silent"].data_ptr(),
                                   res["gain"].data_ptr(), ptr["u"], ptr["z"], ptr["w"], ctypes.c_void_p(s.cuda_stream))
    if rc != 0:
        raise RuntimeError(f"sfdtd_postprocess failed ({rc}): {lib.sfdtd_last_error().decode()}")
    return res


class deferred_checks:
    """Context manager: inside it ``forward_fn`` does not read the status words back (its only host synchronisation after
    the launch), so several reference batches can be in flight on several streams; the statuses are checked on exit.

        with deferred_checks():
            for s, args in zip(streams, batches):
                with torch.cuda.stream(s):
                    outs.append(forward_fn(*args))
    """
    active = False
    pending = []

    def __enter__(self):
        deferred_checks.active = True
        deferred_checks.pending = []
        return self

    def __exit__(self, *exc):
        deferred_checks.active = False
        pend, deferred_checks.pending = deferred_checks.pending, []
        if exc[0] is None:
            torch.cuda.synchronize()
            for st in pend:
                _check_status(st)
        return False


def forward_fn(state_u, state_z, string_params, bow_params, hammer_params,
               bow_mask, hammer_mask, constant, relative_error,
               surface_integral, manufactured, n_0, Nt):
    """Drop-in for the reference extension's ``forward_fn`` (simulator.cpp:14-27).

    Same positional arguments; returns ``[uout, zout, state_u, state_z, v_r, F_H, u_H/k, sig0, sig1]``
    and, like the reference, updates ``state_u``, ``state_z`` and ``hammer_params[2]`` in place.
    All strings of the call form one group (the reference's batch).  Tensors may live on the CPU
    or on CUDA.  float64 tensors run the fp64 kernels (parity mode); float32 tensors (the reference's
    ``precision: single``) run the fp32 kernels -- grid sizes from the reference's float32 evaluation of
    ``get_derived_vars``, states / solves / outputs in float32 -- unless ``SFDTD_WIDEN_F32=1``, which widens
    them to float64 on entry and rounds on exit.  Results are returned on CUDA like the reference's
    ``device()`` (misc.cpp:13-15) does when a GPU is visible.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("torch_fdtd_string_b200.forward_fn needs a CUDA device (no CPU fallback)")
    dev = state_u.device if state_u.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    kappa, alpha, u0, v0, p_a, f0, pos, T60 = string_params
    x_b, v_b, F_b, phi_0, phi_1, wid = bow_params
    x_H, v_H, u_H, w_H, M_r, alpha_H = hammer_params
    B, Nt_c, _ = state_u.shape
    assert Nt == Nt_c, (Nt, Nt_c)

    import os
    sdt = torch.float32 if (state_u.dtype == torch.float32 and os.environ.get("SFDTD_WIDEN_F32", "0") != "1"
                            and not manufactured) else torch.float64

    def D(t):   # float64 on the device; aliases the input when it already is
        return t.detach().to(device=dev, dtype=torch.float64)

    def Dio(t):  # in/out tensor: (device copy with unit space stride, needs_copy_back)
        d = t.detach().to(device=dev, dtype=sdt)
        if d.stride(-1) != 1:
            d = d.contiguous()
        return d, d.data_ptr() != t.data_ptr()

    su, su_cp = Dio(state_u); sz, sz_cp = Dio(state_z)
    uH = u_H.detach().to(device=dev, dtype=sdt)
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
        p_a=D(p_a), check=not deferred_checks.active)
    if deferred_checks.active:
        deferred_checks.pending.append(res["status"])
    # in-place side effects of the reference (string.cpp:264-265, 303)
    if su_cp: state_u.copy_(su)
    if sz_cp: state_z.copy_(sz)
    if uH_cp: u_H.copy_(uH)
    odt = state_u.dtype
    o = [res[n].to(odt) for n in ("uout", "zout")]
    return [o[0], o[1], state_u, state_z, res["v_r"].to(odt), res["F_H"].to(odt), res["u_H_out"].to(odt),
            res["sig0"].to(odt).view(-1, 1, 1), res["sig1"].to(odt).view(-1, 1, 1)]
