"""diagnostics: host time per asynchronous forward_fn call and GPU-side overlap of several reference batches in flight"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from torch_fdtd_string_b200 import sampler, forward_fn, deferred_checks

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
SR, GROUP, Nt = 48000, 24, 48000
def make(i):
    ph = sampler.sample_nsynth_like(GROUP, sr=SR, length=1.0, excitation="pluck", seed=900 + i)
    p = sampler.to_device(ph, dev); c = sampler.expand_controls(p, dev)
    su = torch.zeros(GROUP, Nt, ph["Nx_t1"], dtype=torch.float64, device=dev); su[:, :2] = p["state_u"]
    sz = torch.zeros(GROUP, Nt, ph["Nx_l1"], dtype=torch.float64, device=dev); sz[:, :2] = p["state_z"]
    u0 = torch.zeros(GROUP, 1, ph["Nx_t1"], dtype=torch.float64, device=dev)
    sp = [p["kappa"], p["alpha"], u0, u0, p["p_a"].view(-1, 1, 1), c["f0"], p["pos"], p["T60"]]
    bp = [c["x_b"], c["v_b"], c["F_b"], p["phi_0"], p["phi_1"], c["wid"].contiguous()]
    hp = [p["x_H"], torch.zeros(GROUP, Nt, dtype=torch.float64, device=dev), c["u_H"], p["w_H"], p["M_r"], p["alpha_H"]]
    return (su, sz, sp, bp, hp, p["bow_mask"].view(-1, 1, 1), p["hammer_mask"].view(-1, 1, 1), [ph["k"], ph["theta_t"], ph["lambda_c"]], 4.0, True, False, 0, Nt)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
calls = [make(i) for i in range(n + 1)]
forward_fn(*calls[n]); torch.cuda.synchronize()
streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]; ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
base = torch.cuda.Event(enable_timing=True); base.record(); torch.cuda.synchronize()
host = []
t00 = time.perf_counter()
with deferred_checks():
    for i in range(n):
        with torch.cuda.stream(streams[i]):
            t0 = time.perf_counter()
            ev0[i].record()
            forward_fn(*calls[i])
            ev1[i].record()
            host.append((time.perf_counter() - t0) * 1e3)
print("host ms per call", [round(h, 1) for h in host], "total wall", round((time.perf_counter() - t00) * 1e3, 1))
for i in range(n):
    print(f"call {i}: starts at {base.elapsed_time(ev0[i]):.0f} ms, ends at {base.elapsed_time(ev1[i]):.0f} ms")
