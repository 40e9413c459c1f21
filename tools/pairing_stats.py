"""Diagnostics (GPU): how well do strings that share a warp match in sweep count?  Compares the a-priori estimate with a
short pilot run as the sort key; lower bound = sorting by the full run's own mean sweeps."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch_fdtd_string_b200 import sampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 14208
length = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
dev = torch.device("cuda")
p_host = sampler.sample_nsynth_like(B, length=length, excitation="pluck", seed=1234)
p = sampler.to_device(p_host, dev)
ctl = sampler.expand_controls(p, dev)
def sweeps(n_run):
    c = {k: (v[:, :n_run] if v.size(1) >= n_run else v) for k, v in ctl.items()}
    res = sampler.run_compact(p, 24, counters=True, controls=c, n_run=n_run)
    cc = res["counters"].double().cpu()
    return cc[:, 1], cc[:, 3]
Nt = p_host["Nt"]
s_full, n_full = sweeps(Nt)
sw = s_full / n_full
print("mean sweeps/step", float(sw.mean()))
keys = {}
for n_run in (18, 34, 66, 130, 514):
    s, n = sweeps(n_run)
    keys[f"pilot{n_run - 2}"] = s / n
k = p_host["k"]
f0 = torch.minimum(p_host["f0_a"], p_host["f0_b"])
nt, nl = sampler.derived_grid(f0, p_host["kappa"], k, p_host["theta_t"], 1.0, p_host["alpha"])
phi = (2 * f0) ** 2 * k ** 2 * (p_host["alpha"] ** 2 - 1) / 4
u1 = p_host["state_u"][:, 1]
dm = (u1[:, 1:] - u1[:, :-1]).abs().max(dim=1).values
keys["est"] = phi * nt ** 4 * dm ** 2
keys["actual"] = sw
# only strings of the bulk bucket matter most, but report over all strings: pairs of consecutive strings in sort order
for name, e in keys.items():
    e = torch.nan_to_num(e, nan=1e30)
    order = torch.argsort(e, descending=True)
    s = sw[order]
    n2 = (B // 2) * 2
    pm = s[:n2].view(-1, 2).max(dim=1).values.mean()
    print(f"{name:10s} E[max over 2 strings of mean sweeps] {float(pm):.3f}   (mean {float(sw.mean()):.3f})")
