"""RNG-stream-compatible compact sampler: the reference's ``String`` / ``Bow`` / ``Hammer`` parameter draws
(reference src/model/simulator.py:11-597, driven by src/task/simulate.py:121-162) restated so that

* the global torch RNG is consumed in the reference's order, with the reference's shapes and dtypes (uniforms are
  float32 draws cast to the working precision, src/utils/misc.py:84-90; the vibrato's are drawn in the working
  precision, src/utils/control.py:35-45; the pluck amplitude / position consume batch_size x Nt draws each,
  simulator.py:341-355; the pull-off draws are one or two scalars per string, simulator.py:461-465), so that
  ``torch.manual_seed(proc.seed)`` + a loop over batches yields the reference's dataset parameters, and
* nothing of size (B, Nt, Nx) is ever built: a batch is ~40 scalars per string plus its two initial state rows
  (the same compact description ``sampler.sample_nsynth_like`` produces; curves are synthesised in the stepper).

Covers the sampling modes 'random' / 'equidist' / 'fix' of every parameter, the pluck profiles triangular / smooth /
raised_cosine (simulator.py:169-190), the manufactured initial condition (:175-180) and ``randomize_each='batch'``.
Control curves agree with the reference's to rounding (~1e-16 relative; the reference divides the whole f0 curve by
the Fletcher factor, here the end points are divided).
"""
import math

import numpy as np
import torch

from . import sampler

STRING_DEFAULTS = dict(
    sampling_f0='random', sampling_kappa='random', sampling_alpha='random', sampling_pickup='random', sampling_T60='random',
    precorrect=True,
    f0_min=27.50, f0_max=440, f0_diff_max=50, f0_mod_max=0.02, f0_fixed=20,
    kappa_min=0., kappa_max=0.08, kappa_fixed=0.08, kappa_hammer=0.,
    alpha_min=1, alpha_max=25, alpha_fixed=3.,
    pos_min=0.3, pos_max=0.7, pos_fixed=0.5,
    lossless=False, t60_min_1=20., t60_max_1=30., t60_min_2=30., t60_max_2=30., t60_fixed=20., t60_diff_max=5.,
    sampling_p_a='random', sampling_p_x='random',
    p_a_min=0.001, p_a_max=0.01, p_a_fixed=0.01, p_x_min=0.100, p_x_max=0.90, p_x_fixed=0.50, pluck_profile=None)
BOW_DEFAULTS = dict(x_b_min=0.2, x_b_max=0.5, x_b_maxdiff=0.2, v_b_min=0.3, v_b_max=0.4, F_b_min=80, F_b_max=100,
                    F_b_maxdiff=10, do_pulloff=True, phi_0_max=6, phi_0_min=2, phi_1_max=0.5, phi_1_min=0., wid_min=3, wid_max=6)
HAMMER_DEFAULTS = dict(x_H_min=0.1, x_H_max=0.9, v_H_min=0.5, v_H_max=5, M_r_min=10.0, M_r_max=50.0, w_H_min=1000, w_H_max=3000,
                       alpha_fixed=None)


def _ru(lo, hi, size, dtype, weight=None):
    """src/utils/misc.py:84-90 random_uniform: a float32 draw, cast, weighted"""
    if not isinstance(size, tuple):
        size = (size,)
    if weight is None:
        weight = torch.ones(size, dtype=dtype)
    return (hi - lo) * torch.rand(size=size).to(dtype) * weight + lo


def _equi(lo, hi, n, dtype):
    return torch.linspace(lo, hi, n).to(dtype)


def _grid(f0, kappa_rel, k, theta_t, lambda_c, alpha):
    """src/utils/fdm.py:101-123 with the reference's own arithmetic (python floats or tensors)"""
    sq = (lambda x: x.pow(.5)) if isinstance(f0, torch.Tensor) else (lambda x: x ** .5)
    gamma = 2 * f0
    kappa = gamma * kappa_rel
    IHP = (np.pi * kappa / gamma) ** 2
    K = sq(IHP) * (gamma / np.pi)
    lam = int(1) if lambda_c <= 1 else lambda_c
    h = lam * sq((gamma ** 2 * k ** 2 + sq(gamma ** 4 * k ** 4 + 16 * K ** 2 * k ** 2 * (2 * theta_t - 1))) / (2 * (2 * theta_t - 1)))
    N_t = torch.floor(1 / h) if isinstance(h, torch.Tensor) else int(1 / h)
    h2 = lam * gamma * alpha * k
    N_l = torch.floor(1 / h2) if isinstance(h2, torch.Tensor) else int(1 / h2)
    return N_t, N_l


def _triangular(N, n, p_x, p_a):
    """src/utils/misc.py:60-72 on the single time slice that is plucked: p_x, p_a (B,1,1), n (B,1,1)"""
    vel_l = torch.where(p_x.le(0), torch.zeros_like(p_x), p_a / p_x / n)
    vel_r = torch.where(p_x.le(0), torch.zeros_like(p_x), p_a / (1 - p_x) / n)
    vel_l = ((vel_l * torch.ones_like(vel_l).repeat(1, 1, N)).cumsum(2) - vel_l).clamp(min=0)
    vel_r = ((vel_r * torch.ones_like(vel_r).repeat(1, 1, N)).cumsum(2) - vel_r * (N - n + 1)).clamp(min=0).flip(2)
    return torch.minimum(vel_l, vel_r)


def _raised_cosine(N, h, ctr, wid, n):
    """src/utils/misc.py:36-48"""
    xax = torch.linspace(h, 1, N).view(1, -1, 1)
    ctr = (ctr * n / N)
    wid = wid / N
    ind = torch.sign(torch.relu(-(xax - ctr - wid / 2) * (xax - ctr + wid / 2)))
    out = 0.5 * ind * (1 + torch.cos(2 * np.pi * (xax - ctr) / wid))
    return out / out.abs().sum(1, keepdim=True)


def get_masks(model_name, bs):
    """src/utils/misc.py:95-121 (disjoint=True)"""
    if model_name.endswith('bow'):
        return torch.ones(bs, dtype=torch.bool), torch.zeros(bs, dtype=torch.bool)
    if model_name.endswith('hammer'):
        return torch.zeros(bs, dtype=torch.bool), torch.ones(bs, dtype=torch.bool)
    if model_name.endswith('pluck'):
        return torch.zeros(bs, dtype=torch.bool), torch.zeros(bs, dtype=torch.bool)
    bow = torch.rand(size=(bs,)).gt(0.5)
    ham = torch.rand(size=(bs,)).gt(0.5)
    return bow, ham & ~bow


def sample_reference(batch_size, model_name, sr, length, theta_t, f0_inf, alpha_inf, lambda_c, precision='double',
                     string_kwargs=None, bow_kwargs=None, hammer_kwargs=None, manufactured=False, relative_order=4,
                     redraw_v_H=False):
    """One reference batch, drawn from the GLOBAL torch RNG exactly like reference ``simulate()`` does
    (src/task/simulate.py:148-162).  Returns the compact dict of ``sampler.sample_nsynth_like`` (CPU float64 tensors),
    plus ``target_f0_a/_b`` (the un-corrected end points) and ``u0`` (the initial displacement row)."""
    dt = torch.float64 if precision == 'double' else torch.float32
    sk = dict(STRING_DEFAULTS); sk.update(string_kwargs or {})
    bk = dict(BOW_DEFAULTS); bk.update(bow_kwargs or {})
    hk = dict(HAMMER_DEFAULTS); hk.update(hammer_kwargs or {})
    Bs = batch_size
    k = 1 / sr
    Nt = int(sr * length)
    # ---- masks (simulate.py:141-150) ----
    if model_name.endswith('pluck'):
        pluck_batch = True
    elif model_name == 'random':
        pluck_batch = None
    else:
        pluck_batch = False
    bow_mask, hammer_mask = get_masks(model_name, Bs)
    pluck_mask = ~(bow_mask | hammer_mask)
    ones = lambda: torch.ones(size=(Bs,), dtype=dt)

    # ---- String.initialize_config (simulator.py:150-154): kappa, f0, alpha, pickup, T60 ----
    if sk['sampling_kappa'] == 'random':
        kr = _ru(sk['kappa_min'], sk['kappa_max'], (Bs,), dt)
        kappa = kr * hammer_mask.logical_not() + (sk['kappa_hammer'] + kr) * hammer_mask
    elif sk['sampling_kappa'] == 'equidist':
        kappa = _equi(sk['kappa_min'], sk['kappa_max'], Bs, dt)
    else:
        kappa = sk['kappa_fixed'] * ones()

    mod_frq = torch.zeros(Bs, dtype=dt); mod_amp = torch.zeros(Bs, dtype=dt); vib_t0 = torch.zeros(Bs, dtype=dt)
    if sk['sampling_f0'] == 'random':
        f0_con = _ru(sk['f0_min'], sk['f0_max'], (Bs,), dt)
        f0_1 = _ru(sk['f0_min'], sk['f0_max'], (Bs,), dt)
        f0_2 = _ru(sk['f0_min'], sk['f0_max'], (Bs,), dt).clamp(f0_1 - sk['f0_diff_max'], f0_1 + sk['f0_diff_max'])
        tv = torch.randn((Bs,)).ge(0.5)
        f0_a = torch.where(tv, f0_1, f0_con); f0_b = torch.where(tv, f0_2, f0_con)
        vb_off = torch.randn((Bs,)).ge(0.5)                           # True: no vibrato (simulator.py:231-234)
        mod_frq = (5. * torch.rand(Bs, 1, dtype=dt) + 3.).view(-1)     # control.py:36 (mf = [3, 5])
        mod_amp = (sk['f0_mod_max'] * torch.rand(Bs, 1, dtype=dt)).view(-1)
        vib_t0 = torch.floor((Nt // 2) * torch.rand(Bs).view(-1, 1)).view(-1).to(dt)
        sign = torch.randn(Bs, 1, dtype=dt).sign().view(-1)
        mod_amp = torch.where(vb_off, torch.zeros_like(mod_amp), mod_amp * sign)
    elif sk['sampling_f0'] == 'equidist':
        f0_a = torch.linspace(sk['f0_min'], sk['f0_max'], Bs).to(dt); f0_b = f0_a.clone()
    else:
        ff = sk['f0_fixed']
        try:
            n_f = len(ff)
        except TypeError:
            n_f = 0
        if n_f > 1:
            f0_a = torch.tensor(list(ff), dtype=dt).view(-1) * ones()
            fmin = min(ff)
        else:
            f0_a = (ff if n_f == 0 else list(ff)[0]) * ones()
            fmin = ff if n_f == 0 else list(ff)[0]
        assert fmin >= f0_inf, f"f0_fixed (== {fmin}) should be >= than f0_inf (== {f0_inf})"
        f0_b = f0_a.clone()
    target_a, target_b = f0_a.clone(), f0_b.clone()
    if sk['precorrect']:
        Bc = (np.pi * kappa.view(-1, 1)) ** 2                          # fdm.stiff_string_modes, p = 1
        w0 = (1 * (1 + (2 / np.pi) * Bc ** .5 + 4 / np.pi ** 2 * Bc) * (1 + Bc * 1 ** 2) ** .5).view(-1)
        f0_inf = f0_inf / w0.flatten().max().item()
        f0_a = f0_a / w0; f0_b = f0_b / w0
    Nx_t, Nx_l = _grid(f0_inf, 0, k, theta_t, lambda_c, alpha_inf)

    if sk['sampling_alpha'] == 'random':
        alpha = _ru(sk['alpha_min'], sk['alpha_max'], (Bs,), dt)
    elif sk['sampling_alpha'] == 'equidist':
        alpha = _equi(sk['alpha_min'], sk['alpha_max'], Bs, dt)
    else:
        alpha = (alpha_inf if sk['alpha_fixed'] < alpha_inf else sk['alpha_fixed']) * ones()
    assert alpha.ge(alpha_inf).all()

    if sk['sampling_pickup'] == 'random':
        pos = _ru(sk['pos_min'], sk['pos_max'], (Bs,), dt)
    elif sk['sampling_pickup'] == 'equidist':
        pos = _equi(sk['pos_min'], sk['pos_max'], Bs, dt)
    else:
        pos = sk['pos_fixed'] * ones()

    if sk['sampling_T60'] == 'random':
        fmin_, fmax_ = (1 / 240) * sr / 2, (1 / 4) * sr / 2
        T_f1 = _ru(fmin_ + 1000, fmax_, (Bs,), dt)
        T_f2 = _ru(fmin_, T_f1 - 1000, (Bs,), dt)
        T_t1 = _ru(sk['t60_min_1'], sk['t60_max_1'], (Bs,), dt)
        T_t2 = (T_t1 + _ru(0, sk['t60_diff_max'], (Bs,), dt)).clamp(sk['t60_min_2'], sk['t60_max_2'])
    elif sk['sampling_T60'] == 'equidist':
        T_f1 = 1000. * ones(); T_f2 = 100. * ones()
        t1 = _equi(sk['t60_min_1'], sk['t60_max_1'], Bs - 1, dt); t2 = _equi(sk['t60_min_2'], sk['t60_max_2'], Bs - 1, dt)
        T_t1 = torch.cat([t1, torch.zeros(1, dtype=dt)]); T_t2 = torch.cat([t2, torch.zeros(1, dtype=dt)])
    else:
        T_f1 = 1000. * ones(); T_f2 = 100. * ones()
        T_t1 = (0. if sk['lossless'] else sk['t60_fixed']) * ones(); T_t2 = T_t1.clone()
    T60 = torch.stack([torch.stack([T_f1, T_t1], -1), torch.stack([T_f2, T_t2], -1)], 1)

    # ---- String.initialize_state (simulator.py:169-202) ----
    if pluck_batch:
        plucked = torch.ones(Bs, dtype=dt)
    elif isinstance(pluck_batch, bool):
        plucked = torch.zeros(Bs, dtype=dt)
    else:
        plucked = pluck_mask.to(dt)
    def draw_pp(mode, lo, hi, fixed):
        if mode == 'random':
            return _ru(lo, hi, (Bs, Nt), dt)[:, 0].clone()             # batch_size x Nt draws, only sample 0 is plucked
        if mode == 'equidist':
            return _equi(lo, hi, Bs, dt)
        return fixed * ones()
    p_a = draw_pp(sk['sampling_p_a'], sk['p_a_min'], sk['p_a_max'], sk['p_a_fixed']) * plucked
    p_x = draw_pp(sk['sampling_p_x'], sk['p_x_min'], sk['p_x_max'], sk['p_x_fixed']) * plucked
    prm = dict(f0_a=f0_a.double(), f0_b=f0_b.double(), mod_frq=mod_frq.double(), mod_amp=mod_amp.double(), vib_t0=vib_t0.double())
    f0_lo = sampler.f0_min_over_time(prm, Nt, k).to(dt)
    nx_t = _grid(f0_lo, kappa, k, theta_t, lambda_c, alpha)[0].view(-1, 1, 1)
    pa3, px3 = p_a.view(-1, 1, 1), p_x.view(-1, 1, 1)
    prof = sk['pluck_profile'] or 'triangular'
    if manufactured:
        px3 = torch.sign(px3) * 0.5
        tr = _triangular(Nx_t + 1, nx_t + 1, px3, torch.ones_like(px3)) - 1
        u0 = pa3 * torch.cos(np.pi * tr / 2).pow(2)
    elif prof == 'triangular':
        u0 = _triangular(Nx_t + 1, nx_t + 1, px3, pa3)
    elif prof == 'smooth':
        tr = _triangular(Nx_t + 1, nx_t + 1, px3, torch.ones_like(px3))
        u0 = pa3 * torch.sin(tr * math.pi / 2).pow(2)
    else:
        u0 = _raised_cosine(Nx_t + 1, 1 / Nx_t, px3, nx_t.div(10, rounding_mode='trunc'), nx_t.flatten() + 1).transpose(1, 2) * torch.sign(px3)
    u0 = u0.view(Bs, Nx_t + 1)
    state_u = torch.stack([u0, u0], 1).contiguous()                   # v0 = 0: rows 0 and 1 (fdm.py:92-98)
    state_z = torch.zeros(Bs, 2, Nx_l + 1, dtype=dt)
    p_a_out = p_a.abs()

    # ---- Bow (simulator.py:419-484) ----
    x_b1 = _ru(bk['x_b_min'], bk['x_b_max'], (Bs,), dt)
    x_b2 = (x_b1 + _ru(-bk['x_b_maxdiff'], bk['x_b_maxdiff'], (Bs,), dt)).clamp(bk['x_b_min'], bk['x_b_max'])
    v_b1 = _ru(bk['v_b_min'], bk['v_b_max'], (Bs,), dt); v_b2 = _ru(bk['v_b_min'], bk['v_b_max'], (Bs,), dt)
    F_b1 = _ru(bk['F_b_min'], bk['F_b_max'], (Bs,), dt)
    F_b2 = F_b1 + _ru(-bk['F_b_maxdiff'], bk['F_b_maxdiff'], (Bs,), dt).clamp(bk['F_b_min'], bk['F_b_max'])
    pulloff = torch.full((Bs,), -1.0, dtype=torch.float64)
    if bk['do_pulloff']:
        for b in range(Bs):
            if torch.rand([1])[0] > 0.5:
                po = (3 * length / 4) * torch.rand([1])[0] + (length / 4)
                n_on = int(sr * po)                                    # post_shaper: offset = Nt - int(sr * pulloff) (misc.py:79)
                pulloff[b] = (n_on + 0.5) / sr                         # any value whose floor(sr * .) is n_on
    phi_0 = (bk['phi_0_max'] - bk['phi_0_min']) * torch.rand(size=(Bs,)).to(dt) + bk['phi_0_min']
    phi_1 = (bk['phi_1_max'] - bk['phi_1_min']) * torch.rand(size=(Bs,)).to(dt) + bk['phi_1_min']
    wid = _ru(bk['wid_min'], bk['wid_max'], (Bs,), dt)

    # ---- Hammer (simulator.py:531-597) ----
    x_H = _ru(hk['x_H_min'], hk['x_H_max'], (Bs,), dt)
    v_H = _ru(hk['v_H_min'], hk['v_H_max'], (Bs,), dt)
    w = None if hk['v_H_max'] == hk['v_H_min'] else 1. - (v_H - hk['v_H_min']) / (hk['v_H_max'] - hk['v_H_min'])
    M_r = _ru(hk['M_r_min'], hk['M_r_max'], (Bs,), dt, weight=w)
    w_H = _ru(hk['w_H_min'], hk['w_H_max'], (Bs,), dt)
    if hk['alpha_fixed'] is None:
        alpha_H = (2 * _ru(0, 1, (Bs,), dt).ge(0.5) + 1).to(dt)
    else:
        alpha_H = hk['alpha_fixed'] * ones()

    # task.load_config with a hammer-v_H profile: Hammer.dump_parameter('v_H', profile) runs initialize_velocity again
    # (simulator.py:555-581), i.e. draws the strike velocities a second time after every other draw of the batch
    v_H_redrawn = _ru(hk['v_H_min'], hk['v_H_max'], (Bs,), dt) if redraw_v_H else v_H

    D = lambda t: t.to(torch.float64)
    return dict(
        v_H_redrawn=D(v_H_redrawn), f0_inf_corrected=float(f0_inf),
        B=Bs, sr=sr, Nt=Nt, k=k, theta_t=theta_t, lambda_c=lambda_c, relative_order=relative_order,
        Nx_t1=int(Nx_t) + 1, Nx_l1=int(Nx_l) + 1, bow_mask=bow_mask, hammer_mask=hammer_mask, pluck_mask=pluck_mask,
        kappa=D(kappa), alpha=D(alpha), pos=D(pos), T60=D(T60), p_a=D(p_a_out), p_x=D(p_x), state_u=D(state_u), state_z=D(state_z),
        u0=D(u0), f0_a=D(f0_a), f0_b=D(f0_b), target_f0_a=D(target_a), target_f0_b=D(target_b),
        mod_frq=D(mod_frq), mod_amp=D(mod_amp), vib_t0=D(vib_t0),
        x_H=D(x_H), v_H=D(v_H), M_r=D(M_r), w_H=D(w_H), alpha_H=D(alpha_H),
        x_b1=D(x_b1), x_b2=D(x_b2), v_b1=D(v_b1), v_b2=D(v_b2), F_b1=D(F_b1), F_b2=D(F_b2), pulloff=pulloff,
        phi_0=D(phi_0), phi_1=D(phi_1), wid=D(wid))
