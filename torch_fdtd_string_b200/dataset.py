"""Dataset generation on top of the B200 stepper: the part of the reference's ``run()`` that sits directly
downstream of the time stepper (reference src/task/simulate.py:272-455, src/utils/misc.py:235-299,
src/utils/audio.py:11-76) -- NaN / silence filtering, l-infinity normalisation, and the per-string result layout

    {save_dir}/{id}-{b}/output-u.wav  output-z.wav  output.wav      (PCM_24 for double precision, PCM_16 for single)
    {save_dir}/{id}-{b}/simulation.npz  string_params.npz  hammer_params.npz  bow_params.npz  simulation_config.yaml

with the reduction work (NaN mask, RMS, peak, gain) done on the device so that only the audio that is kept crosses
PCIe.  Parameters come from the compact nsynth-like sampler (``sampler.py``); whole batches are sharded over ranks
(``parallel.rank_batches``) -- one process per GPU, no collective.

    python -m torch_fdtd_string_b200.dataset --save-dir out --num-samples 100 [--batch-size 24] [--excitation pluck]

Differences from the reference, all deliberate: ``simulation.npz`` does not hold ``state_u`` / ``state_z`` (the reference
keeps the full (Nt, Nx) history of every string on the host: 21 MB per string-second; the drop-in ``forward_fn`` /
``process`` path still produces it); no plots; ids are the batch index or 8 random characters per batch
(``--randomize-name``) like the reference.
"""
import argparse
import os
import string as _string

import numpy as np
import torch
import yaml

from . import sampler
from .parallel import rank_batches
from .wavio import write_wav

_CHARS = np.array(list(_string.ascii_lowercase + _string.digits))


def postprocess(uout, zout, silence_threshold=-23.0, normalize_output=True):
    """Device-side reductions of reference src/task/simulate.py:333-335 and src/utils/audio.py:42-76.
    uout, zout: (B, Nt-2) CUDA tensors.  Returns dict(is_nan, is_silent, gain (B,1), u, z, w) -- u/z/w are the
    signals that go to the wav files (normalised by the l-infinity gain of ``uout`` when requested)."""
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


def _write_string(d, wavs, sr, bitrate, save, kinds, sim, string_dict, hammer_dict, bow_dict, theta_t, lambda_c):
    os.makedirs(d, exist_ok=True)
    for name, x in zip(("output-u.wav", "output-z.wav", "output.wav"), wavs):
        write_wav(f"{d}/{name}", x, sr, bitrate)
    if save:
        save_simulation_data(d, kinds, sim, string_dict, hammer_dict, bow_dict, theta_t, lambda_c)


def generate(save_dir, num_samples, batch_size=24, excitation="pluck", sr=48000, length=1.0, seed=1234,
             precision="double", normalize_output=True, skip_silence=True, silence_threshold=-23.0, save=True,
             randomize_name=False, batches_per_call=64, rank=0, world_size=1, device=None, surface_integral=True,
             sampler_cfg=None, time_log=False, num_workers=4):
    """Generates ``num_samples // batch_size`` reference batches (reference run.py:109) and writes the kept strings.
    The per-string files (three wavs, four compressed archives: ~0.3 s of zlib per string on one core, 3000x what the
    stepper needs for that string) are written by ``num_workers`` threads (``proc.num_workers`` of the reference's config; zlib
    releases the GIL) while the GPU runs the next call; at most two calls' host arrays are alive.
    Returns dict(strings, written, nan, silent, seconds_stepper, seconds_total)."""
    import concurrent.futures
    import time
    t_start = time.perf_counter()
    pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, int(num_workers)))
    pending = []
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n_batches = num_samples // batch_size
    mine = list(rank_batches(n_batches, world_size, rank))
    bitrate = "PCM_24" if precision == "double" else "PCM_16"
    stats = dict(strings=0, written=0, nan=0, silent=0, seconds_stepper=0.0)
    rng = np.random.RandomState(seed + 7919 * rank)
    os.makedirs(save_dir, exist_ok=True)
    calls, names = [], {}
    for c0 in range(0, len(mine), batches_per_call):
        # one sampler stream per batch index, so that the result does not depend on the sharding; batches of one call must
        # share the padded state widths (they come from the batch's largest stiffness, like in the reference)
        by_width = {}
        for it in mine[c0:c0 + batches_per_call]:
            q = sampler.sample_nsynth_like(batch_size, sr=sr, length=length, excitation=excitation, seed=seed + it, cfg=sampler_cfg)
            by_width.setdefault((q["Nx_t1"], q["Nx_l1"]), []).append((it, q))
        calls += list(by_width.values())
    for group in calls:
        chunk = [it for it, _ in group]
        B = len(chunk) * batch_size
        p_host = sampler.concat([q for _, q in group])
        p = sampler.to_device(p_host, device)
        ctl = sampler.expand_controls(p, device)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sampler.run_compact(p, batch_size, surface_integral=surface_integral, controls=ctl)
        e1.record()
        uout, zout = res["uout"][:, 2:], res["zout"][:, 2:]
        pp = postprocess(uout, zout, silence_threshold, normalize_output)
        torch.cuda.synchronize()
        stats["seconds_stepper"] += e0.elapsed_time(e1) * 1e-3
        if time_log:
            # reference src/task/simulate.py:327-328 logs one line per batch; here one stepper call covers several batches
            with open(f"{save_dir}/gpu_time.txt", "a") as f:
                for it in chunk:
                    f.write(f"{it}\t{e0.elapsed_time(e1) * 1e-3 / len(chunk):.4f}\n")
        is_nan = pp["is_nan"].cpu().numpy(); is_silent = pp["is_silent"].cpu().numpy()
        keep = ~is_nan & ~(is_silent & skip_silence)
        stats["strings"] += B; stats["nan"] += int(is_nan.sum()); stats["silent"] += int((is_silent & ~is_nan).sum())
        if not keep.any():
            continue
        idx = torch.from_numpy(np.nonzero(keep)[0]).to(device)
        host = {k: pp[k].index_select(0, idx).cpu().numpy() for k in ("u", "z", "w")}       # only kept audio crosses PCIe
        raw = {k: res[k].index_select(0, idx)[:, 2:].cpu().numpy() for k in ("uout", "zout", "v_r", "F_H", "u_H_out")}
        f0 = ctl["f0"].index_select(0, idx).cpu().numpy()
        ctl_h = {k: ctl[k].index_select(0, idx).cpu().numpy() for k in ("x_b", "v_b", "F_b", "u_H")}
        sig0 = res["sig0"].cpu().numpy(); sig1 = res["sig1"].cpu().numpy()
        nt_, nl_ = sampler.derived_grid(torch.from_numpy(f0), p_host["kappa"][keep].view(-1, 1), p_host["k"], p_host["theta_t"],
                                        p_host["lambda_c"], p_host["alpha"][keep].view(-1, 1))
        new_jobs = []
        for j, b in enumerate(np.nonzero(keep)[0]):
            it = chunk[b // batch_size]; bb = b % batch_size
            if it not in names:
                names[it] = "".join(rng.choice(_CHARS, 8)) if randomize_name else str(it)
            dx = names[it]
            d = f"{save_dir}/{dx}-{bb}"
            bow, ham = bool(p_host["bow_mask"][b]), bool(p_host["hammer_mask"][b])
            kinds = (["bow"] if bow else []) + (["hammer"] if ham else []) + (["pluck"] if not (bow or ham) else [])
            sim = dict(uout=raw["uout"][j], zout=raw["zout"][j], v_r_out=raw["v_r"][j], F_H_out=raw["F_H"][j],
                       u_H_out=raw["u_H_out"][j], bow_mask=bow, hammer_mask=ham, pluck_mask=not (bow or ham),
                       Nx_t=nt_[j].numpy(), Nx_l=nl_[j].numpy(), sig0=sig0[b], sig1=sig1[b])
            T = lambda key: p_host[key][b].numpy()
            string_dict = dict(kappa=T("kappa"), alpha=T("alpha"), u0=p_host["state_u"][b, 1].numpy(),
                               v0=np.zeros_like(p_host["state_u"][b, 1].numpy()), p_a=T("p_a"), f0=f0[j], pos=T("pos"),
                               T60=T("T60"), target_f0=f0[j] * float(sampler.fletcher_w0(p_host["kappa"][b])))
            hammer_dict = dict(x_H=T("x_H"), v_H=T("v_H"), u_H=ctl_h["u_H"][j], w_H=T("w_H"), M_r=T("M_r"), alpha=T("alpha_H"))
            bow_dict = dict(x_B=ctl_h["x_b"][j], v_B=ctl_h["v_b"][j], F_B=ctl_h["F_b"][j], phi_0=T("phi_0"), phi_1=T("phi_1"),
                            wid_B=T("wid"))
            new_jobs.append(pool.submit(_write_string, d, (host["u"][j], host["z"][j], host["w"][j]), sr, bitrate, save,
                                        ",".join(kinds), sim, string_dict, hammer_dict, bow_dict, p_host["theta_t"], p_host["lambda_c"]))
            stats["written"] += 1 if save else 0
        for fut in pending:                     # the previous call's files must be on disk before a third call's arrays pile up
            fut.result()
        pending = new_jobs
        del res, ctl, pp
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
    ap.add_argument("--num-workers", type=int, default=4, help="file-writer threads (proc.num_workers of the reference's config)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    st = generate(a.save_dir, a.num_samples, a.batch_size, a.excitation, a.sr, a.length, a.seed, a.precision,
                  not a.no_normalize, not a.keep_silent, randomize_name=a.randomize_name, rank=rank, world_size=world,
                  num_workers=a.num_workers)
    print(st)


if __name__ == "__main__":
    main()
