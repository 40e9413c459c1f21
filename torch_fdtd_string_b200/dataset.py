"""Dataset generation on top of the B200 stepper: the part of the reference's ``run()`` that sits directly up- and
downstream of the time stepper (reference src/task/simulate.py:219-455, src/utils/misc.py:235-299,
src/utils/audio.py:11-76) -- parameter draws, NaN / silence filtering, l-infinity normalisation, PCM quantisation and the
per-string result layout

    {save_dir}/{id}-{b}/output-u.wav  output-z.wav  output.wav      (PCM_24 for double precision, PCM_16 for single)
    {save_dir}/{id}-{b}/simulation.npz  string_params.npz  hammer_params.npz  bow_params.npz  simulation_config.yaml

with everything per-sample done on the device (``sfdtd_postprocess``: NaN mask, RMS, peak, gain, PCM), so that only the
audio and the arrays of strings that are KEPT cross PCIe.  Whole batches are sharded over ranks
(``parallel.rank_batches``) -- one process per GPU, no collective.

Two parameter sources:
  * ``reference_source`` (the CLI's default): the reference's own draws, RNG-stream compatible (``sampler_ref.py``) --
    ``proc.seed`` gives the reference's dataset.  The stream is sequential over batches, so every rank draws all batches
    and keeps its share.
  * ``native_source``: the compact nsynth-like sampler with one generator per batch index (``sampler.py``).

``full_layout=True`` (reference-faithful, the CLI's default) also stores what makes the reference's files large: the
state histories ``state_u[:, :max N_t + 1]`` / ``state_z[:, :max N_l + 1]`` in ``simulation.npz`` and the (Nt, Nx) initial
displacement / velocity arrays in ``string_params.npz`` (src/task/simulate.py:405-408, src/utils/misc.py:241-251): the
stepper then runs in the reference's (B, Nt, Nx) layout (``SFDTD_SAVE_STATE``), ~20 MB per string-second.  With
``full_layout=False`` the archives hold the audio, the per-step outputs and the parameters only.

    python -m torch_fdtd_string_b200.dataset --save-dir out --num-samples 100 [--batch-size 24] [--excitation pluck]
"""
import argparse
import os
import string as _string

import numpy as np
import torch
import yaml

from . import sampler, sampler_ref
from .forward_fn import Plan, build_args, postprocess as pcm_postprocess, synth_controls
from .parallel import rank_batches
from .wavio import write_wav_pcm

_CHARS = np.array(list(_string.ascii_lowercase + _string.digits))


def postprocess(uout, zout, silence_threshold=-23.0, normalize_output=True):
    """PyTorch restatement of reference src/task/simulate.py:333-335 and src/utils/audio.py:42-76 (kept for tests and for
    callers that want float signals; the generation path uses the fused device kernel ``sfdtd_postprocess``).
    uout, zout: (B, Nt-2) CUDA tensors.  Returns dict(is_nan, is_silent, gain (B,1), u, z, w)."""
    is_nan = torch.isnan(uout.sum(-1))
    u = uout * (~is_nan).unsqueeze(-1)                       # NaN * 0 stays NaN, like the reference's multiply
    rms = u.pow(2).mean(-1, keepdim=True).pow(0.5)
    is_silent = (20 * torch.log10(rms)).le(silence_threshold).squeeze(-1)
    if normalize_output:
        maxv = uout.abs().max(-1).values.unsqueeze(-1)
        gain = torch.where(maxv.eq(0) | torch.isnan(maxv), torch.ones_like(maxv), 1.0 / maxv)
    else:
        gain = torch.ones(uout.size(0), 1, dtype=uout.dtype, device=uout.device)
    un = gain * uout
    zn = gain * zout
    return dict(is_nan=is_nan, is_silent=is_silent, gain=gain, u=un, z=zn, w=un + zn)


def save_simulation_data(directory, excitation_type, simulation_dict, string_dict, hammer_dict, bow_dict, theta_t, lambda_c):
    """File layout of reference src/utils/misc.py:235-299."""
    os.makedirs(directory, exist_ok=True)

    def sample(val):
        v = np.asarray(val)
        return v.reshape(-1)[0].item() if v.size else None

    short = {"excitation_type": excitation_type, "theta_t": float(theta_t), "lambda_c": float(lambda_c),
             "value-string": {k: sample(v) for k, v in string_dict.items()},
             "value-hammer": {k: sample(v) for k, v in hammer_dict.items()},
             "value-bow": {k: sample(v) for k, v in bow_dict.items()}}
    np.savez_compressed(f"{directory}/simulation.npz", **simulation_dict)
    np.savez_compressed(f"{directory}/string_params.npz", **string_dict)
    np.savez_compressed(f"{directory}/hammer_params.npz", **hammer_dict)
    np.savez_compressed(f"{directory}/bow_params.npz", **bow_dict)
    with open(f"{directory}/simulation_config.yaml", "w") as f:
        yaml.dump(short, f, default_flow_style=False)


def _write_string(d, pcm, sr, bits, save, kinds, sim, string_dict, hammer_dict, bow_dict, theta_t, lambda_c):
    os.makedirs(d, exist_ok=True)
    for name, raw in zip(("output-u.wav", "output-z.wav", "output.wav"), pcm):
        write_wav_pcm(f"{d}/{name}", raw, sr, bits)
    if save:
        save_simulation_data(d, kinds, sim, string_dict, hammer_dict, bow_dict, theta_t, lambda_c)


# ---- parameter sources: callables  it -> compact batch (CPU), called for it = 0, 1, 2, ... in order ------------------
def native_source(batch_size, sr, length, excitation, seed, cfg=None):
    """one generator per batch index: the result does not depend on the sharding"""
    def draw(it):
        return sampler.sample_nsynth_like(batch_size, sr=sr, length=length, excitation=excitation, seed=seed + it, cfg=cfg)
    draw.sequential = False
    return draw


def reference_source(batch_size, sr, length, excitation, theta_t, f0_inf, alpha_inf, lambda_c, precision="double",
                     string_kwargs=None, bow_kwargs=None, hammer_kwargs=None, manufactured=False, relative_order=4,
                     redraw_v_H=False):
    """the reference's draws from the GLOBAL torch RNG (seed it like reference run.py:75 first); must be called for every
    batch index in order, on every rank"""
    def draw(it):
        return sampler_ref.sample_reference(batch_size, excitation, sr, length, theta_t, f0_inf, alpha_inf, lambda_c, precision,
                                            string_kwargs, bow_kwargs, hammer_kwargs, manufactured, relative_order,
                                            redraw_v_H=redraw_v_H)
    draw.sequential = True
    return draw


# ---- task.load_config: predefined control curves (reference src/task/simulate.py:164-185, README 1.3) ----------------------
OVERRIDE_KEYS = ("string-f0", "bow-x_b", "bow-v_b", "bow-F_b", "bow-wid", "hammer-v_H", "hammer-u_H")


def load_overrides(path, total_size):
    """``{model}-{param}.npy`` files of a preset directory -> {key: (total_size,) float64 array}, padded with the edge value
    or cut like the reference does (simulate.py:166-172).  Parameters the reference could dump but that are not time curves
    (or re-initialise the state: ``string-plucked``) are refused."""
    import glob
    out = {}
    for f in sorted(glob.glob(os.path.join(path, "*.npy"))):
        val = np.load(f)
        if val.shape[-1] < total_size:
            val = np.pad(val, (0, total_size - val.shape[-1]), mode="edge")
        else:
            val = val[:total_size]
        model, param = os.path.basename(f).split(".")[0].split("-")
        key = f"{model.lower()}-{param}"
        if key not in OVERRIDE_KEYS:
            raise NotImplementedError(f"task.load_config: {os.path.basename(f)} ({key}) is not built; supported: {', '.join(OVERRIDE_KEYS)}")
        out[key] = np.asarray(val, dtype=np.float64)
    return out


def apply_overrides(p, ctl, overrides):
    """Replaces control curves of one batch by the loaded ones, like ``dump_parameter`` of the reference's modules
    (src/model/simulator.py:98-112, 441-446, 555-564): every string of the batch gets the same curve.  p: compact batch,
    ctl: dict of (B,Nt) curves (sampler.expand_controls) on p's device; returns ctl (modified in place)."""
    B, Nt = p["B"], ctl["f0"].size(1)
    dev = ctl["f0"].device
    for key, val in overrides.items():
        v = torch.from_numpy(np.ascontiguousarray(val[:Nt])).to(dev)
        if key == "string-f0":
            # String.dump_parameter casts the dump to float32 and pre-corrects it by the Fletcher factor of each string
            v = v.float().double().view(1, -1)
            w0 = sampler.fletcher_w0(p["kappa"].to(dev)).view(-1, 1)
            f0 = v / w0
            lo = p.get("f0_inf_corrected")
            if lo is not None:
                assert float(f0.min()) >= lo, (float(f0.min()), lo)          # simulator.py:109
            ctl["f0"] = f0.expand(B, Nt).contiguous()
        elif key in ("bow-x_b", "bow-v_b", "bow-F_b", "bow-wid"):
            ctl[key[4:]] = v.view(1, -1).expand(B, Nt).contiguous()
        elif key == "hammer-u_H":
            ctl["u_H"] = v.view(1, -1).expand(B, Nt).contiguous()
        elif key == "hammer-v_H":
            # initialize_velocity(profile): v_H = (a fresh draw) x profile; u_H = M_HD on samples 0, 1 + k v_H (simulator.py:570-578)
            prof = v.float().double().view(1, -1)
            v_H = p["v_H_redrawn"].to(dev).view(-1, 1) * prof
            u_H = torch.zeros(B, Nt, dtype=torch.float64, device=dev)
            u_H[:, :2] += -1e-3
            ctl["u_H"] = u_H + p["k"] * v_H
        else:
            raise NotImplementedError(key)
    return ctl


def generate(save_dir, num_samples, batch_size=24, excitation="pluck", sr=48000, length=1.0, seed=1234,
             precision="double", normalize_output=True, skip_silence=True, silence_threshold=-23.0, save=True,
             randomize_name=False, batches_per_call=64, rank=0, world_size=1, device=None, surface_integral=True,
             sampler_cfg=None, time_log=False, num_workers=4, source=None, full_layout=False, manufactured=False,
             overrides=None):
    """Generates ``num_samples // batch_size`` reference batches (reference run.py:109) and writes the kept strings.
    The per-string files (three wavs, four compressed archives: ~0.3 s of zlib per string on one core) are written by
    ``num_workers`` threads (``proc.num_workers`` of the reference's config; zlib releases the GIL) while the GPU runs the
    next call; at most two calls' host arrays are alive.
    Returns dict(strings, written, nan, silent, seconds_stepper, seconds_total)."""
    import concurrent.futures
    import time
    t_start = time.perf_counter()
    pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, int(num_workers)))
    pending = []
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n_batches = num_samples // batch_size
    mine = list(rank_batches(n_batches, world_size, rank))
    bits = 24 if precision == "double" else 16
    stats = dict(strings=0, written=0, nan=0, silent=0, seconds_stepper=0.0)
    rng = np.random.RandomState(seed + 7919 * rank)
    os.makedirs(save_dir, exist_ok=True)
    if source is None:
        source = native_source(batch_size, sr, length, excitation, seed, sampler_cfg)
    drawn = {}
    if getattr(source, "sequential", False):
        for it in range(n_batches):                    # the stream is sequential: draw every batch, keep this rank's
            q = source(it)
            if it in mine:
                drawn[it] = q
    if full_layout:
        # (B, Nt, Nx) histories on the device: bound a call to ~48 GB of state
        Nt_ = int(sr * length)
        q0 = drawn[mine[0]] if drawn else (source(mine[0]) if mine else None)
        if q0 is not None:
            per_batch = batch_size * Nt_ * (q0["Nx_t1"] + q0["Nx_l1"]) * 8
            batches_per_call = max(1, min(batches_per_call, int(48e9 // max(per_batch, 1))))
            if not drawn:
                drawn[mine[0]] = q0
    calls, names = [], {}
    for c0 in range(0, len(mine), batches_per_call):
        # batches of one call must share the padded state widths (they come from the batch's largest stiffness, like in
        # the reference)
        by_width = {}
        for it in mine[c0:c0 + batches_per_call]:
            q = drawn.pop(it) if it in drawn else source(it)
            by_width.setdefault((q["Nx_t1"], q["Nx_l1"]), []).append((it, q))
        calls += list(by_width.values())
    for group in calls:
        chunk = [it for it, _ in group]
        B = len(chunk) * batch_size
        p_host = sampler.concat([q for _, q in group])
        p = sampler.to_device(p_host, device)
        Nt = p_host["Nt"]
        # `precision: single` (the reference's default preset) runs the fp32 kernels: states, u_H and outputs float32
        f64 = dict(dtype=torch.float32 if precision == "single" else torch.float64, device=device)
        if full_layout:
            su = torch.zeros(B, Nt, p_host["Nx_t1"], **f64); su[:, :2] = p["state_u"]
            sz = torch.zeros(B, Nt, p_host["Nx_l1"], **f64); sz[:, :2] = p["state_z"]
        else:
            su, sz = p["state_u"].to(f64["dtype"], copy=True), p["state_z"].to(f64["dtype"], copy=True)
        # hammer displacement: pre-loaded like the reference's Hammer module (simulator.py:573-578) and updated IN PLACE by the
        # stepper (string.cpp:303) -- the reference saves the updated tensor as hammer_params.npz:u_H
        uH = torch.zeros(B, Nt, **f64)
        uH[:, :2] = -1e-3
        uH[:, 1] += p_host["k"] * p["v_H"]
        ctl_tab = None
        if overrides:
            # task.load_config: the curves are materialised as (B,Nt) tables with the loaded ones substituted (the stepper's
            # table mode); everything else is unchanged
            pq = dict(p); pq["B"] = B
            for kx in ("v_H_redrawn",):
                if kx in p_host:
                    pq[kx] = p_host[kx]
            pq["f0_inf_corrected"] = p_host.get("f0_inf_corrected")
            ctl_tab = apply_overrides(pq, dict(sampler.expand_controls(pq, device)), overrides)
            uH = ctl_tab["u_H"].to(f64["dtype"]).contiguous()
        common = dict(kappa=p["kappa"], alpha=p["alpha"], pos=p["pos"], T60=p["T60"], phi_0=p["phi_0"], phi_1=p["phi_1"],
                      x_H=p["x_H"], w_H=p["w_H"], M_r=p["M_r"], alpha_H=p["alpha_H"], bow_mask=p["bow_mask"],
                      hammer_mask=p["hammer_mask"], k=p_host["k"], theta_t=p_host["theta_t"], lambda_c=p_host["lambda_c"],
                      relative_order=p_host["relative_order"], Nt=Nt, group_size=batch_size, surface_integral=surface_integral,
                      save_state=full_layout, manufactured=manufactured, p_a=p["p_a"])
        if ctl_tab is None:
            args, res, keep_alive = build_args(su, sz, u_H=uH, synth=sampler.synth_dict(p), **common)
        else:
            args, res, keep_alive = build_args(su, sz, u_H=uH, f0=ctl_tab["f0"], x_b=ctl_tab["x_b"], v_b=ctl_tab["v_b"],
                                               F_b=ctl_tab["F_b"], wid=ctl_tab["wid"], **common)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        plan = Plan(args)
        e0.record()
        plan.run(args)
        e1.record()
        plan.close()
        pp = pcm_postprocess(res["uout"], res["zout"], n0=2, silence_db=silence_threshold, normalize=normalize_output, bits=bits)
        is_nan = pp["is_nan"].cpu().numpy().astype(bool); is_silent = pp["is_silent"].cpu().numpy().astype(bool)     # (synchronises)
        stats["seconds_stepper"] += e0.elapsed_time(e1) * 1e-3
        if time_log:
            # reference src/task/simulate.py:327-328 logs one line per batch; here one stepper call covers several batches
            with open(f"{save_dir}/gpu_time.txt", "a") as f:
                for it in chunk:
                    f.write(f"{it}\t{e0.elapsed_time(e1) * 1e-3 / len(chunk):.4f}\n")
        keep = ~is_nan & ~(is_silent & skip_silence)
        stats["strings"] += B; stats["nan"] += int(is_nan.sum()); stats["silent"] += int((is_silent & ~is_nan).sum())
        if not keep.any():
            continue
        kept = np.nonzero(keep)[0]
        idx = torch.from_numpy(kept).to(device)
        row = pp["row_bytes"]
        pcm = {k: pp["pcm_" + k].index_select(0, idx)[:, :row].cpu().numpy() for k in ("u", "z", "w")}   # only kept audio crosses PCIe
        raw = {k: res[k].index_select(0, idx)[:, 2:].cpu().numpy() for k in ("uout", "zout", "v_r", "F_H", "u_H_out")}
        sub = dict(p)
        for kx in sampler.SYNTH_KEYS:
            sub[kx] = p[kx].index_select(0, idx)
        if ctl_tab is None:
            ctl = synth_controls(sampler.synth_dict(sub), len(kept), Nt, device)              # the curves the stepper used
        else:
            ctl = {k: ctl_tab[k].index_select(0, idx) for k in ("f0", "x_b", "v_b", "F_b")}
        ctl_h = {k: ctl[k].cpu().numpy() for k in ("f0", "x_b", "v_b", "F_b")}
        ctl_h["u_H"] = uH.index_select(0, idx).cpu().numpy()
        f0 = ctl_h["f0"]
        sig0 = res["sig0"].cpu().numpy(); sig1 = res["sig1"].cpu().numpy()
        nt_, nl_ = sampler.derived_grid(torch.from_numpy(f0), p_host["kappa"][keep].view(-1, 1), p_host["k"], p_host["theta_t"],
                                        p_host["lambda_c"], p_host["alpha"][keep].view(-1, 1))
        w0 = sampler.fletcher_w0(p_host["kappa"])
        new_jobs = []
        for j, b in enumerate(kept):
            it = chunk[b // batch_size]; bb = b % batch_size
            if it not in names:
                names[it] = "".join(rng.choice(_CHARS, 8)) if randomize_name else str(it)
            dx = names[it]
            d = f"{save_dir}/{dx}-{bb}"
            bow, ham = bool(p_host["bow_mask"][b]), bool(p_host["hammer_mask"][b])
            kinds = (["bow"] if bow else []) + (["hammer"] if ham else []) + (["pluck"] if not (bow or ham) else [])
            sim = dict(uout=raw["uout"][j], zout=raw["zout"][j], v_r_out=raw["v_r"][j], F_H_out=raw["F_H"][j],
                       u_H_out=raw["u_H_out"][j], bow_mask=np.array([[bow]]), hammer_mask=np.array([[ham]]),
                       pluck_mask=np.array([[not (bow or ham)]]), Nx_t=nt_[j].numpy(), Nx_l=nl_[j].numpy(),
                       sig0=sig0[b], sig1=sig1[b])
            T = lambda key: p_host[key][b].numpy()
            u0_row = p_host["state_u"][b, 1].numpy()
            if full_layout:
                # reference layout (src/task/simulate.py:405-408): histories trimmed to the largest grid of the string
                sim["state_u"] = su[b, :, :int(nt_[j].max()) + 1].cpu().numpy()
                sim["state_z"] = sz[b, :, :int(nl_[j].max()) + 1].cpu().numpy()
                u0 = np.zeros((Nt, u0_row.size)); u0[0] = p_host["state_u"][b, 0].numpy()
                v0 = np.zeros_like(u0)
                v_H = np.zeros(Nt); v_H[1] = float(p_host["v_H"][b])
            else:
                u0, v0, v_H = u0_row, np.zeros_like(u0_row), T("v_H")
            tf = f0[j] * float(w0[b]) if p_host.get("target_f0_a") is None else None
            if tf is None:
                q1 = {k: p_host[k][b:b + 1] for k in ("mod_frq", "mod_amp", "vib_t0")}
                q1["f0_a"] = p_host["target_f0_a"][b:b + 1]; q1["f0_b"] = p_host["target_f0_b"][b:b + 1]
                tf = sampler._f0_curve(q1, Nt, p_host["k"], torch.arange(1, Nt + 1, dtype=torch.float64).view(1, -1))[0].numpy()
            string_dict = dict(kappa=T("kappa"), alpha=T("alpha"), u0=u0, v0=v0, p_a=np.array([[float(p_host["p_a"][b])]]),
                               f0=f0[j], pos=T("pos"), T60=T("T60"), target_f0=tf)
            hammer_dict = dict(x_H=T("x_H"), v_H=v_H, u_H=ctl_h["u_H"][j], w_H=T("w_H"), M_r=T("M_r"), alpha=T("alpha_H"))
            bow_dict = dict(x_B=ctl_h["x_b"][j], v_B=ctl_h["v_b"][j], F_B=ctl_h["F_b"][j], phi_0=T("phi_0"), phi_1=T("phi_1"),
                            wid_B=np.full(Nt, float(p_host["wid"][b])) if full_layout else T("wid"))
            new_jobs.append(pool.submit(_write_string, d, (pcm["u"][j], pcm["z"][j], pcm["w"][j]), sr, bits, save,
                                        ",".join(kinds), sim, string_dict, hammer_dict, bow_dict, p_host["theta_t"], p_host["lambda_c"]))
            stats["written"] += 1 if save else 0
        for fut in pending:                     # the previous call's files must be on disk before a third call's arrays pile up
            fut.result()
        pending = new_jobs
        del res, ctl, pp, su, sz, uH, args, keep_alive
    for fut in pending:
        fut.result()
    pool.shutdown()
    stats["seconds_total"] = time.perf_counter() - t_start
    return stats


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--save-dir", required=True)
    ap.add_argument("--num-samples", type=int, default=100)
    ap.add_argument("--batch-size", type=int, default=24)
    ap.add_argument("--excitation", default="pluck")
    ap.add_argument("--sr", type=int, default=48000)
    ap.add_argument("--length", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--precision", default="double", choices=["single", "double"])
    ap.add_argument("--no-normalize", action="store_true")
    ap.add_argument("--keep-silent", action="store_true")
    ap.add_argument("--randomize-name", action="store_true")
    ap.add_argument("--full-layout", action="store_true", help="also store the state histories and (Nt, Nx) initial arrays like the reference")
    ap.add_argument("--num-workers", type=int, default=4, help="file-writer threads (proc.num_workers of the reference's config)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    st = generate(a.save_dir, a.num_samples, a.batch_size, a.excitation, a.sr, a.length, a.seed, a.precision,
                  not a.no_normalize, not a.keep_silent, randomize_name=a.randomize_name, rank=rank, world_size=world,
                  num_workers=a.num_workers, full_layout=a.full_layout)
    print(st)


if __name__ == "__main__":
    main()
