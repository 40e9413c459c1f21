"""Generates the committed golden fixtures tests/golden/*.npz by running the
UNMODIFIED reference (its Python samplers + its compiled C++ extension,
oracle/_ref/forward_fn.so) in THIS container.  Not runnable on the GPU box
(/root/reference is absent there); the fixtures are what travels.

    python tests/golden/make_golden.py [case ...]

Each fixture stores the exact inputs handed to the reference's ``process()``
(reference src/task/simulate.py:16) in compact form (only the non-zero time rows
of state_u/state_z) and the outputs it returned.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_driver  # noqa: E402

sim, ext = ref_driver.import_reference()
import torch  # noqa: E402
import src.utils.fdm as fdm  # noqa: E402

SR = 48000

NSYNTH = dict(
    sr=SR, f0_inf=98.0, alpha_inf=1, lambda_c=1, relative_order=4, theta=("auto", 0.03, 98.0),
    string_kwargs=dict(
        sampling_f0='random', sampling_kappa='random', sampling_alpha='random',
        sampling_pickup='random', sampling_T60='random', precorrect=True,
        f0_min=98.0, f0_max=440.0, f0_diff_max=30, f0_mod_max=0.08,
        kappa_min=0.01, kappa_max=0.03, alpha_min=1., alpha_max=25.,
        t60_min_1=10., t60_max_1=25., t60_min_2=10., t60_max_2=30.,
        sampling_p_a='random', p_a_max=0.02, sampling_p_x='random', p_x_max=0.5),
    hammer_kwargs=dict(M_r_min=1.0, M_r_max=10., alpha_fixed=3),
    bow_kwargs=dict(),
)

ALLFIXED = dict(
    sr=SR, f0_inf=55.0, alpha_inf=20, lambda_c=1, relative_order=8, theta=("auto", 0.08, 55.0),
    string_kwargs=dict(
        sampling_f0='fix', sampling_kappa='fix', sampling_alpha='fix',
        sampling_pickup='fix', sampling_T60='fix', precorrect=True,
        f0_fixed=55.0, kappa_fixed=0.08, alpha_fixed=20., lossless=False,
        sampling_p_a='fix', p_a_fixed=0.02, sampling_p_x='fix', p_x_fixed=0.2),
    hammer_kwargs=dict(x_H_min=0.1, x_H_max=0.1, v_H_min=4.0, v_H_max=4.0, M_r_min=1.5, M_r_max=1.5,
                       w_H_min=2000, w_H_max=2000),
    bow_kwargs=dict(x_b_min=0.2, x_b_max=0.2, v_b_min=0.35, v_b_max=0.35, F_b_min=90, F_b_max=90.,
                    phi_0_max=9., phi_0_min=9., phi_1_max=0.01, phi_1_min=0.01, wid_min=4, wid_max=4),
)

LINEAR = dict(
    sr=SR, f0_inf=55.0, alpha_inf=1, lambda_c=1, relative_order=8, theta=("auto", 0.03, 55.0),
    string_kwargs=dict(
        sampling_f0='fix', sampling_kappa='fix', sampling_alpha='fix',
        sampling_pickup='random', sampling_T60='fix', precorrect=False,
        f0_fixed=55.0, f0_mod_max=0, lossless=False, t60_fixed=20., kappa_min=0.03, kappa_max=0.03,
        kappa_fixed=0.03, alpha_fixed=1., alpha_min=1., alpha_max=1.,
        sampling_p_a='fix', p_a_fixed=0.01, sampling_p_x='fix', p_x_fixed=0.3, pluck_profile='smooth'),
    hammer_kwargs=dict(x_H_min=0.5, x_H_max=0.5, v_H_min=2.5, v_H_max=2.5, M_r_min=10., M_r_max=10.,
                       w_H_min=3000, w_H_max=3000, alpha_fixed=3),
    bow_kwargs=dict(),
)

# hammered string with tension modulation on a finer grid (BASELINE config 4, scaled down)
FINEHAMMER = dict(
    sr=96000, f0_inf=55.0, alpha_inf=3, lambda_c=1, relative_order=8, theta=("auto", 0.01, 55.0),
    string_kwargs=dict(
        sampling_f0='fix', sampling_kappa='fix', sampling_alpha='fix',
        sampling_pickup='random', sampling_T60='fix', precorrect=False,
        f0_fixed=55.0, f0_mod_max=0, lossless=False, t60_fixed=20., kappa_fixed=0.01, alpha_fixed=3.,
        sampling_p_a='fix', p_a_fixed=0.01, sampling_p_x='fix', p_x_fixed=0.25, pluck_profile='smooth'),
    hammer_kwargs=dict(x_H_min=0.3, x_H_max=0.3, v_H_min=2.5, v_H_max=2.5, M_r_min=10., M_r_max=10.,
                       w_H_min=3000, w_H_max=3000, alpha_fixed=3),
    bow_kwargs=dict(),
)

PRESETS = dict(nsynth=NSYNTH, allfixed=ALLFIXED, linear=LINEAR, finehammer=FINEHAMMER)

CASES = dict(
    pluck_b1=dict(preset='nsynth', model='pluck', B=1, length=0.01),
    pluck_b3=dict(preset='nsynth', model='pluck', B=3, length=0.01),
    pluck_b3_pickup=dict(preset='nsynth', model='pluck', B=3, length=0.01, surface_integral=False),
    hammer_b3=dict(preset='nsynth', model='hammer', B=3, length=0.01),
    bow_b3=dict(preset='nsynth', model='bow', B=3, length=0.01),
    random_b6=dict(preset='nsynth', model='random', B=6, length=0.01, seed=7),
    hammer_b2_chunked=dict(preset='nsynth', model='hammer', B=2, length=0.01, chunk_length=0.002, seed=11),
    allfixed_bow_b1=dict(preset='allfixed', model='bow', B=1, length=0.01),
    allfixed_hammer_b1=dict(preset='allfixed', model='hammer', B=1, length=0.01),
    allfixed_pluck_b1=dict(preset='allfixed', model='pluck', B=1, length=0.01),
    manufactured_b1=dict(preset='linear', model='pluck', B=1, length=0.005, manufactured=True, chunk_length=0.001),
    finehammer_b1=dict(preset='finehammer', model='hammer', B=1, length=0.002),
    pluck_b24=dict(preset='nsynth', model='pluck', B=24, length=0.004),
    random_b24=dict(preset='nsynth', model='random', B=24, length=0.004, seed=3),
    pluck_b2_long=dict(preset='nsynth', model='pluck', B=2, length=0.1, seed=5),
    random_b4_long=dict(preset='nsynth', model='random', B=4, length=0.05, seed=9),
)


def compact_state(x):
    """(B,Nt,Nx) -> (row indices with any non-zero, rows)"""
    nz = (x != 0).any(dim=2).any(dim=0).nonzero().view(-1)
    return nz.numpy().astype(np.int64), x[:, nz, :].numpy().copy()


def run_case(name, spec):
    p = PRESETS[spec['preset']]
    sr = p['sr']
    mode, kap, f0m = p['theta']
    theta_t = fdm.get_theta(kap, f0m, sr)
    torch.manual_seed(spec.get('seed', 1234))
    captured = {}
    orig_process = sim.process

    def spy(root_dir, state_u, state_z, string_params, bow_params, hammer_params, bow_mask, hammer_mask,
            consts, Nt, chunk_size, *rest):
        captured.update(
            state_u=state_u.clone(), state_z=state_z.clone(),
            string_params=[t.clone() for t in string_params],
            bow_params=[t.clone() for t in bow_params],
            hammer_params=[t.clone() for t in hammer_params],
            bow_mask=bow_mask.clone(), hammer_mask=hammer_mask.clone(),
            consts=list(consts), Nt=Nt, chunk_size=chunk_size, rest=rest)
        return orig_process(root_dir, state_u, state_z, string_params, bow_params, hammer_params,
                            bow_mask, hammer_mask, consts, Nt, chunk_size, *rest)

    sim.process = spy
    t0 = time.time()
    try:
        with torch.no_grad():
            res, params, masks = sim.simulate(
                ref_driver.scratch_root(), spec['model'], sr, theta_t, spec['length'], spec['B'],
                p['f0_inf'], p['alpha_inf'], p['lambda_c'], cpu=True,
                chunk_length=spec.get('chunk_length', -1),
                string_kwargs=dict(p['string_kwargs']), hammer_kwargs=dict(p['hammer_kwargs']),
                bow_kwargs=dict(p['bow_kwargs']), precision='double',
                relative_order=p['relative_order'],
                surface_integral=spec.get('surface_integral', True),
                manufactured=spec.get('manufactured', False))
    finally:
        sim.process = orig_process
    dt = time.time() - t0
    uout, zout, state_u, state_z, v_r, F_H, u_H_o, sig0, sig1 = res
    c = captured
    su_idx, su_rows = compact_state(c['state_u'])
    sz_idx, sz_rows = compact_state(c['state_z'])
    sp = c['string_params']; bp = c['bow_params']; hp = c['hammer_params']
    B, Nt, Nx_t1 = c['state_u'].shape
    out = dict(
        # ---- inputs
        B=B, Nt=Nt, Nx_t1=Nx_t1, Nx_l1=c['state_z'].shape[2], sr=sr,
        chunk_size=c['chunk_size'], consts=np.array(c['consts'], dtype=np.float64),
        relative_order=p['relative_order'], surface_integral=spec.get('surface_integral', True),
        manufactured=spec.get('manufactured', False),
        state_u_idx=su_idx, state_u_rows=su_rows, state_z_idx=sz_idx, state_z_rows=sz_rows,
        kappa=sp[0].numpy(), alpha=sp[1].numpy(), p_a=sp[4].numpy(), f0=sp[5].numpy(), pos=sp[6].numpy(),
        T60=sp[7].numpy(),
        x_b=bp[0].numpy(), v_b=bp[1].numpy(), F_b=bp[2].numpy(), phi_0=bp[3].numpy(), phi_1=bp[4].numpy(),
        wid=bp[5].numpy(),
        x_H=hp[0].numpy(), u_H=hp[2].numpy(), w_H=hp[3].numpy(), M_r=hp[4].numpy(), alpha_H=hp[5].numpy(),
        bow_mask=c['bow_mask'].numpy(), hammer_mask=c['hammer_mask'].numpy(),
        # ---- outputs of the reference
        uout=uout.numpy(), zout=zout.numpy(), v_r_out=v_r.numpy(), F_H_out=F_H.numpy(), u_H_out=u_H_o.numpy(),
        sig0=sig0.numpy(), sig1=sig1.numpy(),
        state_u_last=state_u[:, -2:, :].numpy(), state_z_last=state_z[:, -2:, :].numpy(),
        state_u_abs_sum=state_u.abs().sum(dim=(1, 2)).numpy(), state_z_abs_sum=state_z.abs().sum(dim=(1, 2)).numpy(),
        u_H_inplace=params[2][2].numpy(),   # hammer_params[2] after the in-place update
        ref_seconds=dt,
    )
    if Nt <= 256 or spec.get('full_state'):
        out['state_u_full'] = state_u.numpy(); out['state_z_full'] = state_z.numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: B={B} Nt={Nt} Nx_t1={Nx_t1} Nx_l1={out['Nx_l1']} ref {dt:.1f}s -> "
          f"{os.path.getsize(path) / 1024:.0f} KiB; |uout|max={np.abs(out['uout']).max():.3e}", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        run_case(nm, CASES[nm])
