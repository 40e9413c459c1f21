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
sys.path.insert(0, HERE)

import ref_driver  # noqa: E402

sim, ext = ref_driver.import_reference()
import torch  # noqa: E402
import src.utils.fdm as fdm  # noqa: E402

from presets import *  # noqa: F401,F403  (SR, NSYNTH, ALLFIXED, LINEAR, FINEHAMMER*, PRESETS)

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
    # the same physical time (10 ms) on three more grids: observed order of accuracy against the analytic solution
    manufactured_sr12k=dict(preset='linear12', model='pluck', B=1, length=0.01, manufactured=True, full_state=True),
    manufactured_sr24k=dict(preset='linear24', model='pluck', B=1, length=0.01, manufactured=True, full_state=True),
    manufactured_sr48k=dict(preset='linear', model='pluck', B=1, length=0.01, manufactured=True, full_state=True),
    manufactured_sr96k=dict(preset='linear96', model='pluck', B=1, length=0.01, manufactured=True, full_state=True),
    pluck_b24=dict(preset='nsynth', model='pluck', B=24, length=0.004),
    random_b24=dict(preset='nsynth', model='random', B=24, length=0.004, seed=3),
    pluck_b2_long=dict(preset='nsynth', model='pluck', B=2, length=0.1, seed=5),
    random_b4_long=dict(preset='nsynth', model='random', B=4, length=0.05, seed=9),
    # strings below 52 Hz: more than 256 transverse rows (the 32-lane x 20-row kernels).  Run with MKL_NUM_THREADS=1: the threaded
    # MKL getrf of this image fails on matrices of this size ("Parameter 6 was incorrect on entry to DLASWP").
    lowf0_pluck_b2=dict(preset='lowf0', model='pluck', B=2, length=0.002, seed=21, threads=1),
    lowf0_hammer_b2=dict(preset='lowf0', model='hammer', B=2, length=0.002, seed=22, threads=1),
    lowf0_bow_b2=dict(preset='lowf0', model='bow', B=2, length=0.002, seed=23, threads=1),
    # strings that blow up: NaN mask and onset (with the perturbed twin: the reference's own onset shift)
    pluck_hot_b6=dict(preset='hot', model='pluck', B=6, length=0.06, seed=31, long=True, threads=2, keys=('uout', 'zout')),
    # ---- full-length runs of the BASELINE configs (long format: audio outputs only, time-constant curves stored once) ----
    # configs[0]: single plucked string, nsynth-like, 1 s @ 48 kHz (reference ~8 min)
    pluck_b1_1s=dict(preset='nsynth', model='pluck', B=1, length=1.0, long=True, threads=1),
    # configs[1]: one nsynth-like reference batch at full length (reference ~2-3 h); NaN strings included
    pluck_b24_1s=dict(preset='nsynth', model='pluck', B=24, length=1.0, long=True, threads=3, keys=('uout', 'zout')),
    # configs[2]: bowed string, Helmholtz regime, 4 s (reference ~1 h)
    allfixed_bow_b1_4s=dict(preset='allfixed', model='bow', B=1, length=4.0, long=True, threads=1, keys=('uout', 'zout', 'v_r_out')),
    # configs[3]: hammered string, tension modulation, 192 kHz, first 9600 steps (reference ~13 min)
    finehammer192_b1=dict(preset='finehammer192', model='hammer', B=1, length=0.05, long=True, threads=2),
    # sensitivity of the reference itself: the same inputs with state_u *= 1 + 2^-50 (outputs only, merged with --merge-pert)
    pluck_b24_01s=dict(preset='nsynth', model='pluck', B=24, length=0.1, long=True, threads=3, keys=('uout', 'zout')),
)
PERT = 1.0 + 2.0 ** -50


def compact_state(x):
    """(B,Nt,Nx) -> (row indices with any non-zero, rows)"""
    nz = (x != 0).any(dim=2).any(dim=0).nonzero().view(-1)
    return nz.numpy().astype(np.int64), x[:, nz, :].numpy().copy()


def compact_time(x):
    """(B,Nt) curve -> (B,1) when it is constant in time"""
    x = x.numpy()
    if x.ndim == 2 and x.shape[1] > 1 and (x == x[:, :1]).all():
        return x[:, :1].copy()
    return x


def run_case(name, spec, perturb=False):
    p = PRESETS[spec['preset']]
    torch.set_num_threads(spec.get('threads', 2))
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
        if perturb:
            state_u.mul_(PERT)
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
    if perturb:
        path = os.path.join(PERT_DIR, f"{name}_pert.npz")
        os.makedirs(PERT_DIR, exist_ok=True)
        np.savez(path, uout=uout.numpy(), zout=zout.numpy())
        print(f"{name} (perturbed): ref {dt:.1f}s -> {path}", flush=True)
        return
    if spec.get('long'):
        return save_long(name, spec, p, c, res, params, dt)
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


PERT_DIR = os.path.join(ROOT, "gpurun_out", "golden_pert")     # scratch (git-ignored): raw outputs of the perturbed runs


def save_long(name, spec, p, c, res, params, dt):
    """Long format: compact inputs (time-constant curves once, non-zero state rows / u_H columns only), audio outputs."""
    uout, zout, state_u, state_z, v_r, F_H, u_H_o, sig0, sig1 = res
    su_idx, su_rows = compact_state(c['state_u'])
    sz_idx, sz_rows = compact_state(c['state_z'])
    sp = c['string_params']; bp = c['bow_params']; hp = c['hammer_params']
    B, Nt, Nx_t1 = c['state_u'].shape
    uH = hp[2]
    uH_idx = (uH != 0).any(dim=0).nonzero().view(-1)
    outs = dict(uout=uout, zout=zout, v_r_out=v_r, F_H_out=F_H, u_H_out=u_H_o)
    out = dict(
        long_format=True,
        B=B, Nt=Nt, Nx_t1=Nx_t1, Nx_l1=c['state_z'].shape[2], sr=p['sr'],
        chunk_size=c['chunk_size'], consts=np.array(c['consts'], dtype=np.float64),
        relative_order=p['relative_order'], surface_integral=spec.get('surface_integral', True),
        manufactured=spec.get('manufactured', False),
        state_u_idx=su_idx, state_u_rows=su_rows, state_z_idx=sz_idx, state_z_rows=sz_rows,
        kappa=sp[0].numpy(), alpha=sp[1].numpy(), p_a=sp[4].numpy(), f0=compact_time(sp[5]), pos=sp[6].numpy(),
        T60=sp[7].numpy(),
        x_b=compact_time(bp[0]), v_b=compact_time(bp[1]), F_b=compact_time(bp[2]), phi_0=bp[3].numpy(), phi_1=bp[4].numpy(),
        wid=compact_time(bp[5]),
        x_H=hp[0].numpy(), u_H_idx=uH_idx.numpy().astype(np.int64), u_H_cols=uH[:, uH_idx].numpy().copy(),
        w_H=hp[3].numpy(), M_r=hp[4].numpy(), alpha_H=hp[5].numpy(),
        bow_mask=c['bow_mask'].numpy(), hammer_mask=c['hammer_mask'].numpy(),
        sig0=sig0.numpy(), sig1=sig1.numpy(),
        state_u_last=state_u[:, -2:, :].numpy(), state_z_last=state_z[:, -2:, :].numpy(),
        ref_seconds=dt,
    )
    for k in spec.get('keys', tuple(outs)):
        out[k] = outs[k].numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    nan = np.isnan(out['uout']).any(axis=1)
    print(f"{name}: B={B} Nt={Nt} Nx_t1={Nx_t1} Nx_l1={out['Nx_l1']} ref {dt:.1f}s -> "
          f"{os.path.getsize(path) / 1024:.0f} KiB; NaN strings {int(nan.sum())}/{B}", flush=True)


def merge_pert(name, win=480):
    """Adds the reference's own sensitivity to fixture `name`: per string and per window of `win` samples, the distance
    between the reference run and the reference run on state_u * (1 + 2^-50), plus the NaN onsets of the perturbed run."""
    path = os.path.join(HERE, f"{name}.npz")
    g = dict(np.load(path))
    q = np.load(os.path.join(PERT_DIR, f"{name}_pert.npz"))
    for k in ('uout', 'zout'):
        a, b = g[k], q[k]
        n = a.shape[1] // win * win
        d = np.nan_to_num(b[:, :n] - a[:, :n] * PERT, nan=0.0, posinf=0.0, neginf=0.0).reshape(a.shape[0], -1, win)
        r = np.nan_to_num(a[:, :n], nan=0.0, posinf=0.0, neginf=0.0).reshape(a.shape[0], -1, win)
        g[f'pert_{k}_err'] = np.sqrt((d ** 2).sum(-1))            # (B, n_win) absolute L2 distance per window
        g[f'pert_{k}_norm'] = np.sqrt((r ** 2).sum(-1))
        bad = ~np.isfinite(b)
        g[f'pert_{k}_nan_onset'] = np.where(bad.any(1), bad.argmax(1), -1).astype(np.int64)
    g['pert_win'] = win
    np.savez_compressed(path, **g)
    print(f"{name}: merged perturbed-run sensitivity ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_single(name):
    """The reference's `precision: single` arithmetic on the fixture's OWN inputs: the stored (double) inputs are rounded to
    float32 and handed to the unmodified reference's process(); its outputs go into tests/golden/f32/<name>.npz.  They give
    (a) the reference's own fp32-vs-fp64 distance on these strings and (b) the target of the fp32 kernels."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    g = gu.load_golden(name)
    inp = gu.build_inputs(g, dtype=torch.float32)
    torch.set_num_threads(CASES.get(name, {}).get('threads', 2))
    t0 = time.time()
    with torch.no_grad():
        res = sim.process(ref_driver.scratch_root(), inp["state_u"], inp["state_z"], inp["string_params"], inp["bow_params"],
                          inp["hammer_params"], inp["bow_mask"], inp["hammer_mask"], inp["consts"], inp["Nt"],
                          inp["chunk_size"], None, True, inp["relative_order"], inp["surface_integral"], inp["manufactured"])
    uout, zout, state_u, state_z, v_r, F_H, u_H_o, sig0, sig1 = res
    os.makedirs(os.path.join(HERE, "f32"), exist_ok=True)
    path = os.path.join(HERE, "f32", f"{name}.npz")
    np.savez_compressed(path, uout=uout.numpy(), zout=zout.numpy(), v_r_out=v_r.numpy(), F_H_out=F_H.numpy(),
                        u_H_out=u_H_o.numpy(), state_u_last=state_u[:, -2:, :].numpy(), state_z_last=state_z[:, -2:, :].numpy())
    d = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300))
    print(f"{name} (single): ref {time.time() - t0:.1f}s; fp32 vs fp64 reference: uout {d(uout.numpy(), g['uout']):.2e} "
          f"zout {d(zout.numpy(), g['zout']):.2e} -> {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


if __name__ == "__main__":
    argv = sys.argv[1:]
    if argv and argv[0] == "--single":
        for nm in argv[1:]:
            run_single(nm)
    elif argv and argv[0] == "--merge-pert":
        for nm in argv[1:]:
            merge_pert(nm, win=480 if CASES[nm].get('long') else 48)
    elif argv and argv[0] == "--perturbed":
        for nm in argv[1:]:
            run_case(nm, CASES[nm], perturb=True)
    else:
        for nm in (argv or [k for k, v in CASES.items() if not v.get('long')]):
            run_case(nm, CASES[nm])
