#!/usr/bin/env python
"""bench.py -- throughput of the batched StringFDTD time stepper (BASELINE.json metric:
simulated string-seconds/sec and grid-point-updates/sec at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, host cores

One "step" = one pass of the hot path over one batch: `--strings` nsynth-like strings (groups of 24
= the reference's batch_size, experiment=nsynth-like), every string simulated for `--length` seconds
at 48 kHz in fp64.  N>1: one process per GPU (torchrun), the batch of independent groups is sharded,
no collective on the data path (weak scaling: per-GPU work fixed).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the stepper launches one kernel per string-size bucket on its own stream: give every stream its own hardware queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

SR = 48000
GROUP = 24                      # task.batch_size of experiment=nsynth-like


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strings", type=int, default=148 * GROUP * 8, help="strings per GPU (multiple of 24); BASELINE configs[4] sweeps 1k..1M; the default (28416) needs 133 GB of HBM at 1 s")
    ap.add_argument("--length", type=float, default=1.0, help="seconds of audio per string")
    ap.add_argument("--excitation", default="pluck")
    ap.add_argument("--group", type=int, default=GROUP, help="strings per reference batch (task.batch_size); diagnostics only")
    ap.add_argument("--skip-aux", action="store_true")
    ap.add_argument("--controls", default="synth", choices=["synth", "table"],
                    help="synth: the stepper evaluates the control curves from per-string scalars; table: (B,Nt) fp64 arrays in HBM")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-nt", type=int, default=26, help="samples per reference-arm step (bounded sample)")
    ap.add_argument("--sweep", default="", help="comma-separated TOTAL string counts (BASELINE configs[4]: 1024,4096,16384,65536,"
                    "262144,1048576): each point is sharded over the ranks and run in waves of <= --strings per call, audio only")
    ap.add_argument("--no-drop-in", action="store_true", help="skip the reference-signature forward_fn leg")
    ap.add_argument("--drop-in-batches", type=int, default=4, help="reference batches of the forward_fn leg (configs[1]: 100 // 24 = 4)")
    ap.add_argument("--async-steps", action="store_true", help="diagnostics: queue all timed steps without synchronising the host in "
                    "between (measured 10-15 %% slower per step on B200: launches queued behind a running call slow it down)")
    ap.add_argument("--p-a-max", type=float, default=None, help="override the pluck amplitude cap (diagnostics only)")
    ap.add_argument("--no-grouped", action="store_true", help="skip the short hammer / bow / random legs")
    ap.add_argument("--no-dataset", action="store_true", help="skip the result-file leg (dataset.generate to a scratch directory)")
    ap.add_argument("--dataset-strings", type=int, default=240)
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32 leg (the reference's `precision: single` preset on the fp32 kernels)")
    ap.add_argument("--fp32-steps", type=int, default=2)
    return ap.parse_args()


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of one GPU during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort(); pw.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_min_mhz=(sm[0] if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    power_w=(pw[len(pw) // 2] if pw else None), samples=len(sm), reasons=sorted(reasons))


def reference_arm(a, rank):
    """times the reference's own CPU implementation (oracle/_ref, else the C port) on the host cores"""
    if rank != 0:
        return
    Nt_s = a.ref_nt
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--B", str(GROUP), "--nt", str(Nt_s),
           "--steps", str(a.steps), "--warmup", str(a.warmup), "--excitation", a.excitation]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=3000, env=host_thread_env())
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(json.dumps({"impl": "reference", "unavailable": (out.stderr.strip().splitlines() or ["no output"])[-1][:200]}))
        return
    r = json.loads(line[-1])
    t = sum(r["sec_per_call"]) / len(r["sec_per_call"])
    val = GROUP * (Nt_s - 2) / SR / t
    sample = (f"{GROUP} nsynth-like {a.excitation} strings (one reference batch), first {Nt_s - 2} of 47998 steps per "
              f"step, fp64, {r['cores']} host threads; steps are homogeneous so the rate extrapolates linearly")
    print(json.dumps({
        "impl": "reference", "metric": "simulated string-seconds/sec", "value": val, "unit": "string-seconds/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(a, a.strings), measured=f"{GROUP} strings x {Nt_s - 2} steps per step (one reference batch)",
                       extrapolated=True),
        "extrapolated": True,
        "cpu_baseline": {"value": val, "unit": "string-seconds/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": val, "unit": "string-seconds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def host_thread_env():
    """environment of the CPU reference: every host core (torchrun exports OMP_NUM_THREADS=1 to its ranks, which would
    make the reference single-threaded exactly when it is launched like the GPU arm)"""
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        env.pop(k, None)
    return env


def workload_config(a, strings):
    return {"workload": f"nsynth-like dataset generation ({a.excitation}), {strings} strings/GPU in reference batches "
                        f"of {GROUP}, {a.length:g} s @ 48 kHz, random string params (experiment=nsynth-like ranges)",
            "strings_per_gpu": strings, "group_size": GROUP, "sr": SR, "length_s": a.length,
            "excitation": a.excitation, "precision": "double", "aux_outputs": "skipped" if a.skip_aux else "reference-faithful",
            "controls": getattr(a, "controls", "synth"),
            "l2": "inputs larger than L2 (f0 + outputs >> 126 MB), no flush needed"}


def algorithmic_work(p, counters, group, n_run):
    """F_step = 60 W_t + 30 W_l + S (110 W_t + 74 W_l) + I (4 (W_t + W_l) + 50) per string-step,
    W = batch-max operator widths of the step; grid-point updates = N_t+1 + N_l+1 per string-step.
    Also the same formula on the rows the kernel actually solves (own grid: N_t + 3, N_l + 3 <= W) -> `F_exec`."""
    import numpy as np
    import torch
    from torch_fdtd_string_b200 import sampler
    k = float(np.float32(p["k"])); th = float(np.float32(p["theta_t"])); lam = float(np.float32(p["lambda_c"]))
    tt1 = float(np.float32(2 * np.float32(th) - 1)); tt2 = float(np.float32(2 * np.float32(tt1)))
    B = p["B"]
    dev = p["kappa"].device
    F_sum = F_exec = gpu_updates = Wt_sum = Wl_sum = 0.0
    gpc = max(1, 4096 // group)                                  # groups per chunk: the temporaries are (strings, Nt) sized
    tt = torch.arange(3, n_run + 1, dtype=torch.float64, device=dev).view(1, -1)      # 1-based sample index of steps 2..n_run-1
    for g0 in range(0, B // group, gpc):
        sl = slice(g0 * group, min(B, (g0 + gpc) * group))
        q = {kx: p[kx][sl] for kx in ("f0_a", "f0_b", "mod_frq", "mod_amp", "vib_t0")}
        f0 = sampler._f0_curve(q, p["Nt"], p["k"], tt)
        G = f0.size(0) // group
        gamma = 2 * f0
        K = gamma * p["kappa"][sl].view(-1, 1)
        h1 = lam * ((gamma ** 2 * k ** 2 + (gamma ** 4 * k ** 4 + 16 * K ** 2 * k ** 2 * tt1).sqrt()) / tt2).sqrt()
        Nt_ = (1 / h1).floor()
        Nl_ = (1 / (lam * gamma * p["alpha"][sl].view(-1, 1) * k)).floor()
        gpu_updates += float((Nt_ + 1 + Nl_ + 1).sum())
        Wt = (Nt_.view(G, group, -1).max(dim=1).values + 1)
        Wl = (Nl_.view(G, group, -1).max(dim=1).values + 1)
        S = (1 + p["bow_mask"][sl].double() + p["hammer_mask"][sl].double()).view(G, group, 1)
        I = (counters[sl, 0].double() / counters[sl, 3].clamp(min=1).double()).view(G, group, 1)
        Wt_ = Wt.unsqueeze(1); Wl_ = Wl.unsqueeze(1)
        F = (60 * Wt_ + 30 * Wl_) + S * (110 * Wt_ + 74 * Wl_) + I * (4 * (Wt_ + Wl_) + 50)
        F_sum += float(F.sum()); Wt_sum += float(Wt.sum()); Wl_sum += float(Wl.sum())
        Rt = torch.minimum(Nt_.view(G, group, -1) + 3, Wt_); Rl = torch.minimum(Nl_.view(G, group, -1) + 3, Wl_)
        Fx = (60 * Rt + 30 * Rl) + S * (110 * Rt + 74 * Rl) + I * (4 * (Rt + Rl) + 50)
        F_exec += float(Fx.sum())
        del gamma, K, h1, Nt_, Nl_, F, Fx, Rt, Rl
    n_steps = n_run - 2
    n_wt = (B // group) * n_steps
    return F_sum, F_exec, gpu_updates, Wt_sum / n_wt, Wl_sum / n_wt


def run_sweep(a, points, rank, world, dev):
    """BASELINE configs[4]: batch sweep.  Every point = `total` nsynth-like strings in reference batches of GROUP, sharded
    over the ranks (whole batches, no collective), run in waves of <= a.strings strings per call with in-kernel control
    synthesis and audio-only outputs, parameters drawn on the GPU.  Kernel time = CUDA events around every wave's call, max
    over ranks; wall time also includes parameter sampling and plan creation."""
    import torch
    import torch.distributed as dist
    from torch_fdtd_string_b200 import sampler
    from torch_fdtd_string_b200.forward_fn import Plan
    rows = []
    wave_cap = (a.strings // GROUP) * GROUP
    for total in points:
        groups = max(1, total // GROUP)
        mine = len(range(rank, groups, world)) * GROUP                 # round-robin over ranks (parallel.rank_batches)
        done, t_k, waves, nan = 0, 0.0, 0, 0
        out = None
        n_waves = max(1, -(-mine // wave_cap))
        per_wave = -(-(mine // GROUP) // n_waves) * GROUP              # equal waves (whole batches) instead of full + remainder
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while done < mine:
            n = min(per_wave, mine - done)
            # parameters drawn on the GPU (sampler device=...): nothing but the plan's small read-back crosses PCIe
            ph = sampler.sample_nsynth_like(n, sr=SR, length=a.length, excitation=a.excitation, seed=50000 + 97 * rank + waves + total,
                                            device=dev)
            p = ph
            Nt = ph["Nt"]
            if out is None or out["uout"].size(0) != n:
                out = {k: torch.empty(n, Nt, dtype=torch.float64, device=dev) for k in ("uout", "zout")}
            args, res, keep = sampler.compact_args(p, GROUP, skip_aux=a.skip_aux, out=out, aux_outputs=False)
            plan = Plan(args)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); plan.run(args); e1.record()
            torch.cuda.synchronize()
            t_k += e0.elapsed_time(e1) * 1e-3
            nan += int(torch.isnan(out["uout"][:, -1]).sum())
            plan.close()
            done += n; waves += 1
        wall = time.perf_counter() - t0
        tt = torch.tensor([t_k, wall, float(nan)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            t_k, wall, nan = float(mx[0]), float(mx[1]), int(sm[2])
        ss = groups * GROUP * (int(SR * a.length) - 2) / SR
        rows.append({"strings_total": groups * GROUP, "strings_per_gpu_max": -(-groups // world) * GROUP, "waves_per_gpu": waves,
                     "value": ss / t_k if t_k > 0 else None, "ms_kernel": t_k * 1e3, "wall_value": ss / wall, "nan_strings": nan})
        if rank == 0:
            print(f"[sweep] {rows[-1]}", file=sys.stderr, flush=True)
    return rows


def run_fp32(a, p, out64, dev, rank, world):
    """The same workload on the fp32 kernels (SFDTD_F32: the reference's `precision: single`, its default preset): states,
    solves and outputs float32, grid sizes from the reference's float32 get_derived_vars, per-step scalar tables in fp64.
    Reported beside the fp64 headline with its own (FP32 FMA) roofline and its distance to the fp64 run of the same strings."""
    import ctypes
    import torch
    import torch.distributed as dist
    from torch_fdtd_string_b200 import sampler, _lib
    from torch_fdtd_string_b200.forward_fn import Plan
    B, Nt = p["B"], p["Nt"]
    out = {n: torch.zeros(B, Nt, dtype=torch.float32, device=dev) for n in ("uout", "zout", "v_r", "F_H", "u_H_out")}
    su = p["state_u"].float(); sz = p["state_z"].float()
    su0, sz0 = su.clone(), sz.clone()
    args, res, keep = sampler.compact_args(p, GROUP, skip_aux=a.skip_aux, counters=True, out=out, su=su, sz=sz, precision="single")
    plan = Plan(args)

    def one_step():
        su.copy_(su0); sz.copy_(sz0); res["status"].zero_(); res["counters"].zero_()
        plan.run(args)

    one_step(); torch.cuda.synchronize()
    counters = res["counters"].clone()
    nan32 = torch.isnan(out["uout"][:, 2:]).any(dim=1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.fp32_steps):
        one_step()
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) * 1e-3 / a.fp32_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    per_step = float(tt)
    plan.close()
    flops, flops_exec, gpu_upd, _, _ = algorithmic_work(p, counters, GROUP, Nt)
    peak = ctypes.c_double(0.0)
    if _lib.load().sfdtd_measure_fma_peak(1, ctypes.byref(peak)) != 0 or peak.value <= 0:
        peak.value = 74.4
    # distance to the fp64 run of the same strings over the first 50 ms (later the chaotic strings of DESIGN "Sensitivity"
    # decorrelate in any arithmetic), strings finite in both runs
    n1 = min(Nt, 2 + int(0.05 * SR))
    a64 = out64["uout"][:, 2:n1]; a32 = out["uout"][:, 2:n1].double()
    ok = torch.isfinite(a64).all(dim=1) & torch.isfinite(a32).all(dim=1) & (a64.norm(dim=1) > 0)
    err = ((a32[ok] - a64[ok]).norm(dim=1) / a64[ok].norm(dim=1))
    q = torch.quantile(err, torch.tensor([0.5, 0.9, 0.99], dtype=torch.float64, device=dev)) if err.numel() else torch.zeros(3)
    ss = world * B * (Nt - 2) / SR
    return {"dtype": "f32", "value": ss / per_step, "unit": "string-seconds/s", "ms_per_step": per_step * 1e3, "steps": a.fp32_steps,
            "grid_point_updates_per_s": world * gpu_upd / per_step,
            "roofline": {"bound": "fp32_fma", "achieved": flops / per_step / 1e12, "peak": peak.value, "unit": "TFLOP/s",
                         "frac": flops / per_step / 1e12 / peak.value, "frac_executed": flops_exec / per_step / 1e12 / peak.value,
                         "peak_source": "measured (sfdtd_measure_fma_peak(1): register-resident FFMA chains, this GPU)"},
            "nan_strings": int(nan32.sum()),
            "mean_sweeps_per_step": float(counters[:, 1].sum()) / max(1.0, float(counters[:, 3].sum())),
            "uout_rel_l2_vs_fp64_first_50ms_p50_p90_p99": [float(x) for x in q],
            "note": "same strings and call as the fp64 headline on the fp32 kernels; the reference's own fp32-vs-fp64 distance on "
                    "its 10 ms fixtures is 4e-5 ... 3e-4 (tests/golden/f32)"}


def run_grouped(a, dev, rank, world, peak64):
    """Grouped mode (BASELINE configs[1] names pluck / hammer): reference batches whose strings are hammered / bowed / mixed run as
    thread-block clusters with any-over-batch votes (string.cpp:252-253, hammer.cpp:51).  Short legs (14 208 strings x 0.05 s per
    GPU, fp64, reference-faithful outputs) so that the driver's record carries them beside the pluck headline."""
    import torch
    import torch.distributed as dist
    from torch_fdtd_string_b200 import sampler
    from torch_fdtd_string_b200.forward_fn import Plan
    rows = {}
    B, length = 4 * 148 * GROUP, 0.05            # four groups per SM: a cluster of a typical batch is 4 CTAs, 3 CTAs fit an SM
    for ex in ("hammer", "bow", "random"):
        ph = sampler.sample_nsynth_like(B, sr=SR, length=length, excitation=ex, seed=4321 + rank)
        p = sampler.to_device(ph, dev)
        Nt = ph["Nt"]
        su = p["state_u"].clone(); sz = p["state_z"].clone()
        args, res, keep = sampler.compact_args(p, GROUP, counters=True, su=su, sz=sz)
        plan = Plan(args)

        def one_step():
            su.copy_(p["state_u"]); sz.copy_(p["state_z"]); res["status"].zero_(); res["counters"].zero_()
            plan.run(args)

        one_step(); torch.cuda.synchronize()
        counters = res["counters"].clone()
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            one_step()
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3 / 2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        per_step = float(tt)
        plan.close()
        flops, flops_exec, gpu_upd, _, _ = algorithmic_work(p, counters, GROUP, Nt)
        rows[ex] = {"value": world * B * (Nt - 2) / SR / per_step, "unit": "string-seconds/s", "ms_per_step": per_step * 1e3,
                    "frac": flops / per_step / 1e12 / peak64, "frac_executed": flops_exec / per_step / 1e12 / peak64,
                    "mean_outer_iters": float(counters[:, 0].sum()) / max(1.0, float(counters[:, 3].sum())),
                    "status_bits": int(res["status"].max())}
        del plan, args, res, keep, su, sz, p
    rows["config"] = f"{B} strings/GPU in reference batches of {GROUP}, {length} s @ 48 kHz, fp64, frac of the FP64 FMA peak ({peak64:.1f} TFLOP/s)"
    return rows


def run_dataset(a, dev):
    """`python -m run experiment=nsynth-like` downstream of the stepper: reference batches -> stepper -> device NaN / silence /
    gain / PCM -> the reference's result files (three wavs, four compressed archives, one yaml per kept string;
    src/task/simulate.py:344-425, src/utils/misc.py:235-299) written to a scratch directory by writer threads while the GPU
    computes.  Reports wall time next to the stepper's own time: the zlib archives, not the stepper, set it."""
    import shutil
    import tempfile
    from torch_fdtd_string_b200 import dataset
    d = tempfile.mkdtemp(prefix="sfdtd_bench_ds_")
    try:
        workers = max(1, (os.cpu_count() or 4) - 2)
        st = dataset.generate(d, num_samples=a.dataset_strings, batch_size=GROUP, excitation=a.excitation, sr=SR, length=a.length,
                              seed=4242, num_workers=workers, device=dev)
        nbytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(d) for f in fs)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return {"strings": st["strings"], "written": st["written"], "nan": st["nan"], "silent": st["silent"],
            "seconds_total": st["seconds_total"], "seconds_stepper": st["seconds_stepper"], "writer_threads": workers,
            "bytes_written": nbytes, "strings_per_s_wall": st["strings"] / st["seconds_total"],
            "note": "compact result layout (no (Nt,Nx) state histories); wall time is the compressed-archive writers'"}


def run_drop_in(a, dev, batches=4):
    """The literal drop-in: the reference's own call -- forward_fn(state_u (B,Nt,Nx), ...) with B = 24 fat tensors, one
    batch per call (src/task/simulate.py:65-76), `batches` calls = BASELINE configs[1] (num_samples=100 -> 4 batches) --
    (a) one after the other like the reference's loop (src/task/simulate.py:272), (b) all in flight on their own streams
    (torch_fdtd_string_b200.deferred_checks).  A single batch is 12 warps on a 148-SM GPU and its time loop is sequential:
    this path is latency-bound whatever the kernel does; the native API above is what fills the GPU."""
    import torch
    from torch_fdtd_string_b200 import sampler, forward_fn, deferred_checks
    Nt = int(SR * a.length)
    calls = []
    for i in range(batches):
        ph = sampler.sample_nsynth_like(GROUP, sr=SR, length=a.length, excitation=a.excitation, seed=900 + i)
        p = sampler.to_device(ph, dev)
        c = sampler.expand_controls(p, dev)
        su = torch.zeros(GROUP, Nt, ph["Nx_t1"], dtype=torch.float64, device=dev); su[:, :2] = p["state_u"]
        sz = torch.zeros(GROUP, Nt, ph["Nx_l1"], dtype=torch.float64, device=dev); sz[:, :2] = p["state_z"]
        u0 = torch.zeros(GROUP, 1, ph["Nx_t1"], dtype=torch.float64, device=dev)
        sp = [p["kappa"], p["alpha"], u0, u0, p["p_a"].view(-1, 1, 1), c["f0"], p["pos"], p["T60"]]
        bp = [c["x_b"], c["v_b"], c["F_b"], p["phi_0"], p["phi_1"], c["wid"].contiguous()]
        hp = [p["x_H"], torch.zeros(GROUP, Nt, dtype=torch.float64, device=dev), c["u_H"], p["w_H"], p["M_r"], p["alpha_H"]]
        calls.append((su, sz, sp, bp, hp, p["bow_mask"].view(-1, 1, 1), p["hammer_mask"].view(-1, 1, 1),
                      [ph["k"], ph["theta_t"], ph["lambda_c"]], float(ph["relative_order"]), True, False, 0, Nt))

    def reset():
        for cl in calls:
            cl[0][:, 2:].zero_(); cl[1][:, 2:].zero_(); cl[4][2][:, 2:].zero_()
        torch.cuda.synchronize()

    forward_fn(*calls[0]); reset()                                      # warm-up
    t0 = time.perf_counter()
    for cl in calls:
        forward_fn(*cl)
        torch.cuda.synchronize()
    t_seq = time.perf_counter() - t0
    reset()
    streams = [torch.cuda.Stream(device=dev) for _ in calls]
    t0 = time.perf_counter()
    with deferred_checks():
        for st, cl in zip(streams, calls):
            with torch.cuda.stream(st):
                forward_fn(*cl)
    t_par = time.perf_counter() - t0
    ss = batches * GROUP * (Nt - 2) / SR
    return {"api": "forward_fn(state_u (B,Nt,Nx), ...) -- the reference's call, B = 24, SAVE_STATE", "batches": batches,
            "sequential": {"value": ss / t_seq, "unit": "string-seconds/s", "ms_per_batch": t_seq / batches * 1e3},
            "in_flight": {"value": ss / t_par, "unit": "string-seconds/s", "ms_total": t_par * 1e3, "streams": batches},
            "note": "wall clock incl. launches and the final synchronise; tensors resident on the device"}


def main():
    global GROUP
    a = parse()
    GROUP = a.group
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        reference_arm(a, rank)
        return
    import torch
    import torch.distributed as dist
    from torch_fdtd_string_b200 import _lib, launch_count
    from torch_fdtd_string_b200 import sampler
    from torch_fdtd_string_b200.forward_fn import build_args, Plan, postprocess

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B = (a.strings // GROUP) * GROUP
    p_host = sampler.sample_nsynth_like(B, sr=SR, length=a.length, excitation=a.excitation, seed=1234 + rank,
                                        cfg=(dict(p_a_max=a.p_a_max) if a.p_a_max else None))
    Nt = p_host["Nt"]
    p = sampler.to_device(p_host, dev)
    f64 = dict(dtype=torch.float64, device=dev)
    out = {n: torch.zeros(B, Nt, **f64) for n in ("uout", "zout", "v_r", "F_H", "u_H_out")}
    su = p["state_u"].clone(); sz = p["state_z"].clone()
    ctl = uH = None
    if a.controls == "synth":
        # control curves synthesised inside the stepper from the compact scalars (no (B,Nt) input arrays)
        args, res, keep = sampler.compact_args(p, GROUP, skip_aux=a.skip_aux, counters=True, out=out, su=su, sz=sz)
    else:
        ctl = sampler.expand_controls(p, dev)
        uH = torch.empty_like(ctl["u_H"])
        args, res, keep = build_args(
            su, sz, kappa=p["kappa"], alpha=p["alpha"], f0=ctl["f0"], pos=p["pos"], T60=p["T60"],
            x_b=ctl["x_b"], v_b=ctl["v_b"], F_b=ctl["F_b"], wid=ctl["wid"], phi_0=p["phi_0"], phi_1=p["phi_1"],
            x_H=p["x_H"], w_H=p["w_H"], M_r=p["M_r"], alpha_H=p["alpha_H"], u_H=uH,
            bow_mask=p["bow_mask"], hammer_mask=p["hammer_mask"], k=p["k"], theta_t=p["theta_t"],
            lambda_c=p["lambda_c"], relative_order=p["relative_order"], Nt=Nt, group_size=GROUP,
            surface_integral=True, save_state=False, skip_aux=a.skip_aux, p_a=p["p_a"], out=out, counters=True)
    torch.cuda.synchronize()
    t_plan = time.perf_counter()
    plan = Plan(args)                             # prepass + one device->host read, outside the timed region
    t_plan = time.perf_counter() - t_plan

    bg = None
    if os.environ.get("SFDTD_BENCH_BG_D2H"):      # diagnostics: a bulk device->host copy beside every step (like the e2e leg's)
        gb = float(os.environ["SFDTD_BENCH_BG_D2H"])
        chunk = int(os.environ.get("SFDTD_BENCH_BG_CHUNK_MB", "256")) << 20
        bg = dict(src=torch.empty(int(gb * 1e9), dtype=torch.uint8, device=dev), stage=[torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(2)],
                  stream=torch.cuda.Stream(device=dev), chunk=chunk)

    def one_step():
        # inputs resident in HBM; nothing here synchronises the host
        if bg is not None:
            with torch.cuda.stream(bg["stream"]):
                for j, o in enumerate(range(0, bg["src"].numel(), bg["chunk"])):
                    n_ = min(bg["chunk"], bg["src"].numel() - o)
                    bg["stage"][j % 2][:n_].copy_(bg["src"][o:o + n_], non_blocking=True)
        su.copy_(p["state_u"]); sz.copy_(p["state_z"])
        if uH is not None:
            uH.copy_(ctl["u_H"])                  # u_H is updated in place by the stepper (string.cpp:303)
        res["status"].zero_(); res["counters"].zero_()
        plan.run(args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(a.warmup, 1)):
        one_step()
    torch.cuda.synchronize()
    status = int(res["status"].max())
    counters = res["counters"].clone()
    nan_strings = int(torch.isnan(out["uout"][:, 2:]).any(dim=1).sum())

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks ----
    clocks = ClockSampler(local_rank); clocks.start()
    l0 = launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    marks = []
    for _ in range(a.steps):
        one_step()
        ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append(ev)
        if not a.async_steps:
            torch.cuda.synchronize()        # the host waits for each step like a caller that consumes its result would
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(len(marks))]
    launches = launch_count() - l0
    clk = clocks.stop()
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt)
    per_step = t_max / a.steps
    string_seconds = world * B * (Nt - 2) / SR
    value = string_seconds / per_step

    flops, flops_exec, gpu_upd, Wt_mean, Wl_mean = algorithmic_work(p, counters, GROUP, Nt)
    n_ctl_reads = 0 if a.controls == "synth" else 6
    bytes_ = B * (Nt - 2) * (n_ctl_reads + 5) * 8 + B * 2 * (p["Nx_t1"] + p["Nx_l1"]) * 8
    plan.close()

    fp32 = None
    if not a.no_fp32 and a.controls == "synth":
        try:
            fp32 = run_fp32(a, p, out, dev, rank, world)
        except Exception as e:                                          # never takes the headline down
            if world > 1:
                raise
            fp32 = {"error": str(e)[:300]}
        torch.cuda.empty_cache()

    # ---- end to end through the public API with HOST buffers ----
    # every step: pinned host compact parameters -> H2D -> plan (prepass, one small D2H) -> stepper (controls synthesised in
    # the kernel) -> device post-processing (NaN / silence flags, l-infinity gain, PCM_24 quantisation: what the reference
    # writes to output-u/-z/.wav, src/task/simulate.py:333-337,416-425) -> D2H of the three PCM streams on a copy stream,
    # overlapping the next step's kernels.  The timed region ends when the last byte is on the host.
    e2e = None
    if not a.no_e2e:
        del ctl, uH
        for n in ("v_r", "F_H", "u_H_out"):
            out.pop(n)
        del args, res, keep
        torch.cuda.empty_cache()
        pin = {kx: p_host[kx].pin_memory() for kx in sampler.TENSOR_KEYS}
        ph = dict(p_host); ph.update(pin)
        ns = Nt - 2
        row = ns * 3
        pitch = (row + 15) // 16 * 16
        pcm = [{kx: torch.empty(B, pitch, dtype=torch.uint8, device=dev) for kx in ("u", "z", "w")} for _ in range(2)]
        chunk_rows = max(1, min(B, (int(os.environ.get("SFDTD_BENCH_STAGE_MB", "256")) << 20) // pitch))   # pinned staging: two slots of <= 256 MiB
        stage = [torch.empty(chunk_rows, pitch, dtype=torch.uint8).pin_memory() for _ in range(2)]
        flags_h = torch.empty(3, B, dtype=torch.float64).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get("SFDTD_BENCH_COPY_PRIO", "0")))
        d2h_bytes = 3 * B * pitch + 3 * B * 8

        def prepare():
            # inputs and plan of a step are prepared once the previous call has finished (its read-back still runs on the copy
            # stream): H2D of the compact parameters, prepass kernel (results through mapped pinned memory: a device->host copy
            # would queue behind the bulk PCM read-back), bucketing.  Measured alternatives (tools/r02_e2e_diag.sh, 0.5 s strings,
            # 1480 ms stepper): prepared beside the running call on a side stream the stepper takes 1554-1638 ms and a step
            # 1626-1727 ms; prepared in the gap 1493 / 1537 ms.
            q = sampler.to_device(ph, dev, non_blocking=True)
            a_, r_, k_ = sampler.compact_args(q, GROUP, skip_aux=a.skip_aux, out=out, aux_outputs=False)
            pl = Plan(a_)
            return (q, a_, r_, k_, pl)           # (kept alive by the caller until the step's kernels have completed)

        def launch(i, prep):
            q, a_, r_, k_, pl = prep
            m0 = torch.cuda.Event(enable_timing=True); m1 = torch.cuda.Event(enable_timing=True); m2 = torch.cuda.Event(enable_timing=True)
            m0.record()
            th = time.perf_counter()
            pl.run(a_)
            host_run.append((time.perf_counter() - th) * 1e3)
            m1.record()
            pp = postprocess(out["uout"], out["zout"], n0=2, bits=24, out=pcm[i % 2])
            m2.record()
            marks_e2e.append((m0, m1, m2))
            fl = torch.stack([pp["is_nan"].double(), pp["is_silent"].double(), pp["gain"]])
            ev = torch.cuda.Event(); ev.record()
            pl.close()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev)
                flags_h.copy_(fl, non_blocking=True)
                j = 0
                for kx in (() if os.environ.get("SFDTD_BENCH_NO_D2H") else ("u", "z", "w")):      # (diagnostics knob)
                    src = pcm[i % 2][kx]
                    for r0 in range(0, B, chunk_rows):
                        r1 = min(B, r0 + chunk_rows)
                        stage[j % 2][: r1 - r0].copy_(src[r0:r1], non_blocking=True)
                        j += 1
                done = torch.cuda.Event(enable_timing=True); done.record(copy_stream)
            fl.record_stream(copy_stream)
            arrived.append(done)
            return ev

        arrived = []                             # per step: event after the last byte of its results reached the host buffers
        marks_e2e = []
        host_run = []
        prep = prepare()
        ev = launch(0, prep)
        barrier()
        arrived.clear(); marks_e2e.clear()
        # the last step's read-back is not hidden: amortised over the pipelined steps (a dataset run pipelines hundreds)
        n_e2e = max(1, min(a.steps, 12)) if a.length < 0.5 else max(6, min(a.steps, 12))
        clocks_e2e = ClockSampler(local_rank); clocks_e2e.start()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            prep = prepare()
            ev = launch(i + 1, prep)             # stepper + post-processing of step i+1 on the main stream, read-back on the copy stream
            ev.synchronize()                     # one call in flight: a launch queued behind a running call slows it down (--async-steps)
        barrier()
        clk_e2e = clocks_e2e.stop()
        # value: wall clock of the n pipelined steps incl. the un-hidden read-back of the last one (cold-start + drain);
        # steady state: interval between the arrival of the first and of the last step's results on the host
        steady = arrived[0].elapsed_time(arrived[-1]) * 1e-3 / max(1, len(arrived) - 1) if len(arrived) > 1 else float("nan")
        te = torch.tensor([(time.perf_counter() - t0) / n_e2e, steady], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": string_seconds / float(te[0]), "unit": "string-seconds/s",
               "h2d_bytes_per_step": world * sampler.compact_nbytes(p_host), "d2h_bytes_per_step": world * d2h_bytes,
               "ms_per_step": float(te[0]) * 1e3, "pipelined_steps": n_e2e,
               "ms_stepper": sum(m[0].elapsed_time(m[1]) for m in marks_e2e) / len(marks_e2e),
               "ms_postprocess": sum(m[1].elapsed_time(m[2]) for m in marks_e2e) / len(marks_e2e),
               "ms_stepper_steps": [round(m[0].elapsed_time(m[1]), 1) for m in marks_e2e],
               "ms_launch_host": [round(x, 2) for x in host_run[1:]],
               "clocks": clk_e2e,
               "steady_state": {"value": string_seconds / float(te[1]), "ms_per_step": float(te[1]) * 1e3,
                                "note": "interval between the host arrival of consecutive steps' results (a dataset run pipelines hundreds of steps: the drain of the last one vanishes)"},
               "note": "all ranks, every step: pinned host compact parameters -> H2D -> plan -> stepper "
                       "(in-kernel control synthesis) -> device NaN/silence/gain + PCM_24 quantisation -> D2H of output-u/-z/sum PCM (copy "
                       "stream, overlapping the next step)"}

    sweep = None
    if a.sweep:
        del out
        torch.cuda.empty_cache()
        sweep = run_sweep(a, [int(x) for x in a.sweep.split(",") if x], rank, world, dev)
    grouped = None
    if not a.no_grouped and a.excitation == "pluck":
        try:
            import ctypes as _ct
            pk = _ct.c_double(0.0)
            if lib.sfdtd_measure_fma_peak(0, _ct.byref(pk)) != 0 or pk.value <= 0:
                pk.value = 37.2
            grouped = run_grouped(a, dev, rank, world, pk.value)
        except Exception as e:
            if world > 1:
                raise
            grouped = {"error": str(e)[:300]}
    drop_in = None
    if rank == 0 and not a.no_drop_in:
        try:
            drop_in = run_drop_in(a, dev, a.drop_in_batches)
        except Exception as e:                                          # diagnostics leg: never takes the headline down
            drop_in = {"error": str(e)[:200]}
    ds = None
    if rank == 0 and not a.no_dataset:
        try:
            ds = run_dataset(a, dev)
        except Exception as e:                                          # diagnostics leg: never takes the headline down
            ds = {"error": str(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: FP64 FMA pipe (measured on this GPU), HBM as the secondary line ----
    import ctypes
    peak = ctypes.c_double(0.0)
    peak_src = "measured (sfdtd_measure_fma_peak: register-resident DFMA chains, this GPU, this run)"
    if lib.sfdtd_measure_fma_peak(0, ctypes.byref(peak)) != 0 or peak.value <= 0:
        peak.value = 37.2; peak_src = "nominal 148 SM x 64 lanes x 2 x 1.965 GHz"
    achieved = flops / per_step / 1e12
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; hbm_src = "of measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak = 6650.0; hbm_src = "of fallback"
    traffic = None
    try:   # DRAM bytes per string-step of the dominant kernel from the committed ncu --set full capture, scaled to this call
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02f_f64.json")))
        traffic = tr["dram_bytes_per_string_step"] * B * (Nt - 2)
    except Exception:
        pass
    roofline = {"bound": "fp64_fma", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                "traffic": traffic, "traffic_unit": "DRAM bytes per call (ncu dram__bytes_read+write per string-step of the largest bucket kernel x string-steps of the call; profiles/ncu_traffic_r02f_f64.json)",
                "peak_source": peak_src, "peak_nominal": 37.2, "frac_of_nominal": achieved / 37.2,
                "flops_per_string_step": flops / (B * (Nt - 2)),
                # the same formula on the rows the kernel actually solves (own grid N_t+3 / N_l+3 instead of the batch-max
                # operator widths the reference solves; the ghost rows are folded into one pivot exactly)
                "frac_executed": flops_exec / per_step / 1e12 / peak.value,
                "flops_executed_per_string_step": flops_exec / (B * (Nt - 2)),
                "note": "algorithmic flops per SURVEY.md 8(d) of all strings of the call / CUDA-event time of the call (one kernel launch per string-size bucket, all overlapped on side streams); tensor cores unused (no dense contraction)"}
    roofline_hbm = {"bound": "hbm", "achieved": bytes_ / per_step / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_ / per_step / 1e9 / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                    "algorithmic_bytes_per_string_step": (n_ctl_reads + 5) * 8,
                    "note": f"{n_ctl_reads} control reads + 5 output writes per string-step (fp64)"}

    cpu = None
    if not a.no_cpu_baseline:
        try:
            o = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--B", str(GROUP), "--nt", "62",
                                "--steps", "1", "--warmup", "0", "--excitation", a.excitation],
                               capture_output=True, text=True, timeout=1200, env=host_thread_env())
            r = json.loads([l for l in o.stdout.splitlines() if l.startswith("{")][-1])
            t = r["sec_per_call"][0]
            cpu = {"value": GROUP * 60 / SR / t, "unit": "string-seconds/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": f"one reference batch ({GROUP} strings) x 60 steps, fp64, {r['cores']} host threads "
                             f"({r['host_cpus']} cpus), {t:.1f} s wall; steps homogeneous, rate extrapolates linearly"}
        except Exception as e:
            cpu = {"value": None, "unit": "string-seconds/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    print(json.dumps({
        "metric": "simulated string-seconds/sec", "value": value, "unit": "string-seconds/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, B),
        "grid_point_updates_per_s": world * gpu_upd / per_step,
        "mean_operator_widths": {"W_t": Wt_mean, "W_l": Wl_mean},
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "e2e": e2e,
        "fp32": fp32, "grouped": grouped, "drop_in": drop_in, "dataset": ds, "sweep": sweep,
        "plan_create_ms": t_plan * 1e3, "gpu_launches": int(launches), "clocks": clk, "step_ms": [round(x, 2) for x in step_ms],
        "peak_device_memory_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1),
        "health": {"status_bits": status, "nan_strings": nan_strings,
                   "mean_outer_iters": float(counters[:, 0].sum()) / max(1.0, float(counters[:, 3].sum())),
                   "mean_sweeps_per_step": float(counters[:, 1].sum()) / max(1.0, float(counters[:, 3].sum())),
                   "per_string_mean_sweeps_p50_p99_max": [float(x) for x in torch.quantile(
                       counters[:, 1].double() / counters[:, 3].clamp(min=1).double(),
                       torch.tensor([0.5, 0.99, 1.0], dtype=torch.float64, device=dev))]},
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
