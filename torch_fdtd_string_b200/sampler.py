"""Compact nsynth-like parameter / excitation sampler (host side, PyTorch CPU RNG).

Draws the same distributions as the reference's ``String`` / ``Bow`` / ``Hammer`` modules under
``experiment=nsynth-like`` (reference src/model/simulator.py:11-597, src/utils/control.py,
src/configs/experiment/nsynth-like.yaml:31-56) but never materialises the reference's
(B, Nt, Nx) tensors: a string is ~40 scalars plus its two initial rows.  ``expand_controls``
turns the compact description into the (B, Nt) control curves on the device.

It is NOT RNG-stream compatible with the reference (the parity fixtures come from the
reference's own samplers); it exists so that the throughput workloads have the reference's
parameter statistics.
"""
import math

import torch

NSYNTH = dict(
    f0_min=98.0, f0_max=440.0, f0_diff_max=30.0, f0_mod_max=0.08, kappa_min=0.01, kappa_max=0.03,
    alpha_min=1.0, alpha_max=25.0, t60_min_1=10.0, t60_max_1=25.0, t60_min_2=10.0, t60_max_2=30.0,
    t60_diff_max=5.0, pos_min=0.3, pos_max=0.7, p_a_min=0.001, p_a_max=0.02, p_x_min=0.1, p_x_max=0.5,
    f0_inf=98.0, alpha_inf=1.0, lambda_c=1.0, relative_order=4,
    # Hammer / Bow defaults (simulator.py:419-425,531-537) with the nsynth-like overrides
    x_H_min=0.1, x_H_max=0.9, v_H_min=0.5, v_H_max=5.0, M_r_min=1.0, M_r_max=10.0, w_H_min=1000.0, w_H_max=3000.0,
    alpha_H=3.0, x_b_min=0.2, x_b_max=0.5, x_b_maxdiff=0.2, v_b_min=0.3, v_b_max=0.4, F_b_min=80.0, F_b_max=100.0,
    F_b_maxdiff=10.0, phi_0_min=2.0, phi_0_max=6.0, phi_1_min=0.0, phi_1_max=0.5, wid_min=3.0, wid_max=6.0,
    # sampling_T60: 'random' | 'fix' (t60_fixed at 1000 / 100 Hz, or lossless = all-zero decay times; simulator.py:376-385)
    sampling_T60="random", lossless=False, t60_fixed=20.0,
)


def get_theta(kappa_max, f0_inf, sr, lambda_c=1):
    """reference src/utils/fdm.py:125-141"""
    gamma = 2 * f0_inf
    kappa = gamma * kappa_max
    k = 1 / sr
    R = ((gamma ** 4 * k ** 2 + 4 * kappa ** 2 * math.pi ** 2) / (gamma ** 4 * k ** 2)) ** .5
    S = gamma ** 4 * k ** 2 * lambda_c ** 2 / (4 * kappa ** 2 * math.pi ** 4)
    theta = 0.5 + 2 * S * lambda_c ** 2 * (R - 1) ** 2 + math.pi ** 2 * S * (R - 1)
    assert theta < 1, theta
    return theta


def derived_grid(f0, kappa_rel, k, theta_t, lambda_c, alpha):
    """reference src/utils/fdm.py:101-123 (python twin of get_derived_vars); tensors or floats"""
    sq = (lambda x: x.sqrt()) if isinstance(f0, torch.Tensor) else math.sqrt
    fl = (lambda x: x.floor()) if isinstance(f0, torch.Tensor) else (lambda x: float(int(x)))
    gamma = 2 * f0
    kappa = gamma * kappa_rel
    K = sq((math.pi * kappa / gamma) ** 2) * (gamma / math.pi)
    lam = max(1, lambda_c)
    h = lam * sq((gamma ** 2 * k ** 2 + sq(gamma ** 4 * k ** 4 + 16 * K ** 2 * k ** 2 * (2 * theta_t - 1))) / (2 * (2 * theta_t - 1)))
    N_t = fl(1 / h)
    N_l = fl(1 / (lam * gamma * alpha * k))
    return N_t, N_l


def _u(lo, hi, n, gen):
    return (hi - lo) * torch.rand(n, generator=gen, dtype=torch.float64, device=gen.device) + lo


def sample_nsynth_like(B, sr=48000, length=1.0, excitation="pluck", seed=1234, cfg=None, device="cpu"):
    """-> dict of compact float64 tensors (+ python scalars).  ``device``: where the draws are made and the tensors live
    ("cpu": torch's CPU generator; a CUDA device: the GPU-side sampler -- same distributions, that device's generator, and
    the min-f0 scan over the whole curve, which dominates on the host, runs on the GPU)."""
    c = dict(NSYNTH)
    if cfg:
        c.update(cfg)
    device = torch.device(device)
    g = torch.Generator(device=device).manual_seed(seed)
    _rand = lambda n, **kw: torch.rand(n, generator=g, device=device, **kw)
    _randn = lambda n, **kw: torch.randn(n, generator=g, device=device, **kw)
    _ones = lambda: torch.ones(B, dtype=torch.bool, device=device)
    k = 1.0 / sr
    Nt = int(sr * length)
    theta_t = c["theta_t"] if c.get("theta_t") is not None else get_theta(c["kappa_max"], c["f0_min"], sr, c["lambda_c"])
    # masks (src/utils/misc.py:95-121)
    if excitation.endswith("bow"):
        bow = _ones(); ham = ~_ones()
    elif excitation.endswith("hammer"):
        bow = ~_ones(); ham = _ones()
    elif excitation.endswith("pluck"):
        bow = ~_ones(); ham = ~_ones()
    else:
        bow = _rand(B) > 0.5
        ham = (_rand(B) > 0.5) & ~bow
    pluck = ~(bow | ham)
    kappa = _u(c["kappa_min"], c["kappa_max"], B, g)
    # f0 (simulator.py:210-234): constant or glissando, optional vibrato, then pre-correction
    f0_con = _u(c["f0_min"], c["f0_max"], B, g)
    f0_1 = _u(c["f0_min"], c["f0_max"], B, g)
    f0_2 = torch.minimum(torch.maximum(_u(c["f0_min"], c["f0_max"], B, g), f0_1 - c["f0_diff_max"]), f0_1 + c["f0_diff_max"])
    tv = _randn(B) >= 0.5
    f0_a = torch.where(tv, f0_1, f0_con); f0_b = torch.where(tv, f0_2, f0_con)
    vib_off = _randn(B) >= 0.5                      # True -> no vibrato
    mod_frq = 5.0 * _rand(B, dtype=torch.float64) + 3.0
    mod_amp = c["f0_mod_max"] * _rand(B, dtype=torch.float64)
    vib_t0 = torch.floor((Nt // 2) * _rand(B, dtype=torch.float64))
    vib_sign = torch.sign(_randn(B, dtype=torch.float64))
    mod_amp = torch.where(vib_off, torch.zeros_like(mod_amp), mod_amp * vib_sign)
    Bc = (math.pi * kappa) ** 2                                        # Fletcher detune (fdm.py:143-158)
    w0 = (1 + (2 / math.pi) * Bc.sqrt() + 4 / math.pi ** 2 * Bc) * (1 + Bc).sqrt()
    f0_inf = c["f0_inf"] / float(w0.max())
    Nx_t, Nx_l = derived_grid(f0_inf, 0.0, k, theta_t, c["lambda_c"], c["alpha_inf"])
    Nx_t1, Nx_l1 = int(Nx_t) + 1, int(Nx_l) + 1
    alpha = _u(c["alpha_min"], c["alpha_max"], B, g)
    pos = _u(c["pos_min"], c["pos_max"], B, g)
    fmin, fmax = (1 / 240) * sr / 2, (1 / 4) * sr / 2
    T_f1 = _u(fmin + 1000, fmax, B, g)
    T_f2 = fmin + (T_f1 - 1000 - fmin) * _rand(B, dtype=torch.float64)
    T_t1 = _u(c["t60_min_1"], c["t60_max_1"], B, g)
    T_t2 = (T_t1 + _u(0, c["t60_diff_max"], B, g)).clamp(c["t60_min_2"], c["t60_max_2"])
    if c["sampling_T60"] == "fix":
        T_f1 = torch.full((B,), 1000.0, dtype=torch.float64, device=device); T_f2 = torch.full((B,), 100.0, dtype=torch.float64, device=device)
        T_t1 = torch.full((B,), 0.0 if c["lossless"] else float(c["t60_fixed"]), dtype=torch.float64, device=device); T_t2 = T_t1.clone()
    T60 = torch.stack([torch.stack([T_f1, T_t1], -1), torch.stack([T_f2, T_t2], -1)], 1)   # (B,2,2)
    # pluck shape (simulator.py:169-200; misc.py:60-72 triangular), rows n=0 and n=1 of state_u
    p_a = _u(c["p_a_min"], c["p_a_max"], B, g) * pluck
    p_x = _u(c["p_x_min"], c["p_x_max"], B, g)
    prm = dict(f0_a=f0_a / w0, f0_b=f0_b / w0, mod_frq=mod_frq, mod_amp=mod_amp, vib_t0=vib_t0)
    f0_lo = f0_min_over_time(prm, Nt, k)
    n_t, _ = derived_grid(f0_lo, kappa, k, theta_t, c["lambda_c"], alpha)
    N = Nx_t1
    n = (n_t + 1).view(-1, 1)
    i = torch.arange(N, dtype=torch.float64, device=device).view(1, -1)
    vl = (p_a / p_x).view(-1, 1) / n
    vr = (p_a / (1 - p_x)).view(-1, 1) / n
    left = (vl * i).clamp(min=0)
    right = (vr * (i + 1) - vr * (N - n + 1)).clamp(min=0).flip(1)
    u0 = torch.minimum(left, right) * pluck.view(-1, 1)
    state_u = torch.stack([u0, u0], 1).contiguous()                    # v0 = 0 -> rows 0 and 1 equal
    state_z = torch.zeros(B, 2, Nx_l1, dtype=torch.float64, device=device)
    # hammer (simulator.py:531-597)
    x_H = _u(c["x_H_min"], c["x_H_max"], B, g)
    v_H = _u(c["v_H_min"], c["v_H_max"], B, g)
    wgt = 1.0 - (v_H - c["v_H_min"]) / (c["v_H_max"] - c["v_H_min"])
    M_r = (c["M_r_max"] - c["M_r_min"]) * _rand(B, dtype=torch.float64) * wgt + c["M_r_min"]
    w_H = _u(c["w_H_min"], c["w_H_max"], B, g)
    alpha_H = torch.full((B,), c["alpha_H"], dtype=torch.float64, device=device)
    # bow (simulator.py:419-484)
    x_b1 = _u(c["x_b_min"], c["x_b_max"], B, g)
    x_b2 = (x_b1 + _u(-c["x_b_maxdiff"], c["x_b_maxdiff"], B, g)).clamp(c["x_b_min"], c["x_b_max"])
    v_b1 = _u(c["v_b_min"], c["v_b_max"], B, g); v_b2 = _u(c["v_b_min"], c["v_b_max"], B, g)
    F_b1 = _u(c["F_b_min"], c["F_b_max"], B, g)
    F_b2 = F_b1 + _u(-c["F_b_maxdiff"], c["F_b_maxdiff"], B, g).clamp(c["F_b_min"], c["F_b_max"])
    pulloff = torch.where(_rand(B) > 0.5,
                          (3 * length / 4) * _rand(B, dtype=torch.float64) + length / 4,
                          torch.full((B,), -1.0, dtype=torch.float64, device=device))
    phi_0 = _u(c["phi_0_min"], c["phi_0_max"], B, g); phi_1 = _u(c["phi_1_min"], c["phi_1_max"], B, g)
    wid = _u(c["wid_min"], c["wid_max"], B, g)
    return dict(
        B=B, sr=sr, Nt=Nt, k=k, theta_t=theta_t, lambda_c=c["lambda_c"], relative_order=c["relative_order"],
        Nx_t1=Nx_t1, Nx_l1=Nx_l1, bow_mask=bow, hammer_mask=ham, pluck_mask=pluck,
        kappa=kappa, alpha=alpha, pos=pos, T60=T60, p_a=p_a, p_x=p_x, state_u=state_u, state_z=state_z,
        f0_a=prm["f0_a"], f0_b=prm["f0_b"], mod_frq=mod_frq, mod_amp=mod_amp, vib_t0=vib_t0,
        x_H=x_H, v_H=v_H, M_r=M_r, w_H=w_H, alpha_H=alpha_H,
        x_b1=x_b1, x_b2=x_b2, v_b1=v_b1, v_b2=v_b2, F_b1=F_b1, F_b2=F_b2, pulloff=pulloff,
        phi_0=phi_0, phi_1=phi_1, wid=wid)


def _f0_curve(p, Nt, k, t):
    """f0(t) = glissando * (1 + vibrato) (control.py:12-45); t = 1..Nt (cumsum of ones)"""
    ramp = (t - 1) / max(Nt - 1, 1)
    f0 = p["f0_a"].view(-1, 1) + (p["f0_b"] - p["f0_a"]).view(-1, 1) * ramp
    dt = t - p["vib_t0"].view(-1, 1)
    vib = (dt > 0) * p["mod_amp"].view(-1, 1) * (1 - torch.cos(2 * math.pi * p["mod_frq"].view(-1, 1) * dt * k)) / 2
    return f0 + vib * f0


def f0_min_over_time(p, Nt, k, block=4096):
    lo = None
    for s in range(0, Nt, block):
        t = torch.arange(s + 1, min(s + block, Nt) + 1, dtype=torch.float64, device=p["f0_a"].device).view(1, -1)
        m = _f0_curve(p, Nt, k, t).min(dim=1).values
        lo = m if lo is None else torch.minimum(lo, m)
    return lo


def expand_controls(p, device, n_run=None, chunk=4096):
    """compact description (already on `device`) -> dict of (B,n_run) float64 control curves on the device:
    the first `n_run` samples (default: all `Nt`) of the curves defined over the full length `Nt`.  Evaluated `chunk`
    strings at a time into preallocated outputs, so that the temporaries stay small next to the curves themselves."""
    Nt, B = p["Nt"], p["B"]
    n_run = Nt if n_run is None else int(n_run)
    out = {k: torch.empty(B, n_run, dtype=torch.float64, device=device) for k in ("f0", "x_b", "v_b", "F_b", "u_H")}
    for s0 in range(0, B, chunk):
        sl = slice(s0, min(B, s0 + chunk))
        q = {k: (v[sl] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.size(0) == B else v) for k, v in p.items()}
        q["B"] = sl.stop - sl.start
        c = _expand_chunk(q, device, n_run)
        for k in out:
            out[k][sl] = c[k]
    out["wid"] = p["wid"].view(-1, 1).expand(B, n_run)                                 # time stride 0
    return out


def _expand_chunk(p, device, n_run):
    Nt, k, sr = p["Nt"], p["k"], p["sr"]
    B = p["B"]
    t = torch.arange(1, n_run + 1, dtype=torch.float64, device=device).view(1, -1)
    ramp = (t - 1) / max(Nt - 1, 1)
    lin = lambda a, b: a.view(-1, 1) + (b - a).view(-1, 1) * ramp
    f0 = _f0_curve(p, Nt, k, t)
    x_b = lin(p["x_b1"], p["x_b2"])
    v_b = lin(p["v_b1"], p["v_b2"]) * torch.tanh(t / sr * 10)                       # pre_shaper (misc.py:74-76)
    F_b = lin(p["F_b1"], p["F_b2"])
    # post_shaper (misc.py:78-82): tanh ramp-down ending at the pull-off time
    off = (Nt - (sr * p["pulloff"]).floor()).view(-1, 1)
    w = torch.tanh((Nt - (t - 1) - off).clamp(min=0) / sr * 100)
    F_b = torch.where(p["pulloff"].view(-1, 1) > 0, F_b * w, F_b)
    u_H = torch.zeros(B, n_run, dtype=torch.float64, device=device)                      # simulator.py:573-578
    u_H[:, :2] = -1e-3
    u_H[:, 1] += k * p["v_H"]
    return dict(f0=f0, x_b=x_b, v_b=v_b, F_b=F_b, u_H=u_H)


def fletcher_w0(kappa):
    """stiff-string detune factor f_1 / f0 (reference src/utils/fdm.py:143-158); f0 is pre-corrected by it"""
    Bc = (math.pi * kappa) ** 2
    return (1 + (2 / math.pi) * Bc.sqrt() + 4 / math.pi ** 2 * Bc) * (1 + Bc).sqrt()


def concat(parts):
    """concatenates sampler outputs of equal (Nx_t1, Nx_l1, Nt) along the string axis"""
    p0 = parts[0]
    for q in parts[1:]:
        assert (q["Nx_t1"], q["Nx_l1"], q["Nt"]) == (p0["Nx_t1"], p0["Nx_l1"], p0["Nt"])
    out = dict(p0)
    extra = [kx for kx in ("target_f0_a", "target_f0_b", "u0", "v_H_redrawn") if all(kx in q for q in parts)]
    for kx in TENSOR_KEYS + ["p_x", "pluck_mask"] + extra:
        out[kx] = torch.cat([q[kx] for q in parts], 0)
    out["B"] = sum(q["B"] for q in parts)
    if all(q.get("f0_inf_corrected") is not None for q in parts):
        out["f0_inf_corrected"] = min(q["f0_inf_corrected"] for q in parts)
    return out


TENSOR_KEYS = ["kappa", "alpha", "pos", "T60", "p_a", "state_u", "state_z", "f0_a", "f0_b", "mod_frq", "mod_amp",
               "vib_t0", "x_H", "v_H", "M_r", "w_H", "alpha_H", "x_b1", "x_b2", "v_b1", "v_b2", "F_b1", "F_b2",
               "pulloff", "phi_0", "phi_1", "wid", "bow_mask", "hammer_mask"]


def to_device(p, device, pin=False, non_blocking=False):
    q = dict(p)
    for kx in TENSOR_KEYS:
        t = p[kx]
        if pin and t.device.type == "cpu":
            t = t.pin_memory()
        q[kx] = t.to(device, non_blocking=non_blocking)
    return q


def compact_nbytes(p):
    return int(sum(p[kx].numel() * p[kx].element_size() for kx in TENSOR_KEYS))


SYNTH_KEYS = ["f0_a", "f0_b", "mod_frq", "mod_amp", "vib_t0", "x_b1", "x_b2", "v_b1", "v_b2", "F_b1", "F_b2", "pulloff", "wid", "v_H"]


def synth_dict(p):
    """compact description -> the ``synth`` argument of the stepper (struct sfdtd_synth): the control curves are evaluated
    inside the kernels from these per-string scalars, no (B,Nt) array exists"""
    d = {kx: p[kx] for kx in SYNTH_KEYS}
    d.update(Nt_full=p["Nt"], sr=float(p["sr"]), t_0=0)
    return d


def compact_args(p_dev, group_size, surface_integral=True, skip_aux=False, counters=False, out=None, n_run=None,
                 aux_outputs=True, su=None, sz=None, precision="double"):
    """(args, results, keep) of a synthesised-control call for compact device parameters (see forward_fn.build_args);
    precision "single" runs the fp32 kernels (states and outputs float32)"""
    from .forward_fn import build_args
    import torch as _t
    n_run = p_dev["Nt"] if n_run is None else int(n_run)
    sdt = _t.float32 if precision == "single" else _t.float64
    su = p_dev["state_u"].to(sdt, copy=True) if su is None else su
    sz = p_dev["state_z"].to(sdt, copy=True) if sz is None else sz
    return build_args(
        su, sz, kappa=p_dev["kappa"], alpha=p_dev["alpha"], pos=p_dev["pos"], T60=p_dev["T60"],
        phi_0=p_dev["phi_0"], phi_1=p_dev["phi_1"], x_H=p_dev["x_H"], w_H=p_dev["w_H"], M_r=p_dev["M_r"],
        alpha_H=p_dev["alpha_H"], bow_mask=p_dev["bow_mask"], hammer_mask=p_dev["hammer_mask"], k=p_dev["k"],
        theta_t=p_dev["theta_t"], lambda_c=p_dev["lambda_c"], relative_order=p_dev["relative_order"], Nt=n_run,
        group_size=group_size, synth=synth_dict(p_dev), surface_integral=surface_integral, save_state=False,
        skip_aux=skip_aux, p_a=p_dev["p_a"], counters=counters, out=out, aux_outputs=aux_outputs)


def run_compact(p_dev, group_size, surface_integral=True, skip_aux=False, counters=False, controls=None, out=None,
                n_run=None, synth=True, aux_outputs=True, precision="double"):
    """compact device params -> audio for the first `n_run` samples (default: the full length).  ``synth`` (default): the
    stepper synthesises the control curves itself; otherwise they are expanded to (B,n_run) arrays first (``controls`` may
    pass them in)."""
    from .forward_fn import step_strings, _call
    import torch as _t
    dev = p_dev["kappa"].device
    n_run = p_dev["Nt"] if n_run is None else int(n_run)
    if synth and controls is None:
        a, res, keep = compact_args(p_dev, group_size, surface_integral, skip_aux, counters, out, n_run, aux_outputs,
                                    precision=precision)
        with _t.cuda.device(dev):
            _call(a)
        res["_keep"] = keep
        return res
    c = controls if controls is not None else expand_controls(p_dev, dev, n_run)
    sdt = _t.float32 if precision == "single" else _t.float64
    su = p_dev["state_u"].to(sdt, copy=True); sz = p_dev["state_z"].to(sdt, copy=True)
    res = step_strings(
        su, sz, kappa=p_dev["kappa"], alpha=p_dev["alpha"], f0=c["f0"], pos=p_dev["pos"], T60=p_dev["T60"],
        x_b=c["x_b"], v_b=c["v_b"], F_b=c["F_b"], wid=c["wid"], phi_0=p_dev["phi_0"], phi_1=p_dev["phi_1"],
        x_H=p_dev["x_H"], w_H=p_dev["w_H"], M_r=p_dev["M_r"], alpha_H=p_dev["alpha_H"], u_H=c["u_H"].to(sdt, copy=True),
        bow_mask=p_dev["bow_mask"], hammer_mask=p_dev["hammer_mask"], k=p_dev["k"], theta_t=p_dev["theta_t"],
        lambda_c=p_dev["lambda_c"], relative_order=p_dev["relative_order"], Nt=n_run, group_size=group_size,
        surface_integral=surface_integral, save_state=False, skip_aux=skip_aux, p_a=p_dev["p_a"], counters=counters, out=out,
        check=False, aux_outputs=aux_outputs)
    return res
