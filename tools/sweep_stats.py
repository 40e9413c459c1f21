"""Diagnostics (GPU): per-string sweep counts of the block iteration vs a-priori difficulty estimates."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch_fdtd_string_b200 import sampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3552
length = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
dev = torch.device("cuda")
p_host = sampler.sample_nsynth_like(B, length=length, excitation="pluck", seed=1234)
p = sampler.to_device(p_host, dev)
res = sampler.run_compact(p, 24, counters=True)
c = res["counters"].double().cpu()
sw = c[:, 1] / c[:, 3]
print("mean sweeps/step", float(sw.mean()), "p50/p90/p99/max", [float(x) for x in torch.quantile(sw, torch.tensor([.5, .9, .99, 1.], dtype=torch.float64))])
k = p_host["k"]
f0 = torch.minimum(p_host["f0_a"], p_host["f0_b"])
gamma = 2 * f0
nt, nl = sampler.derived_grid(f0, p_host["kappa"], k, p_host["theta_t"], 1.0, p_host["alpha"])
phi = gamma ** 2 * k ** 2 * (p_host["alpha"] ** 2 - 1) / 4
u1 = p_host["state_u"][:, 1]
dm = (u1[:, 1:] - u1[:, :-1]).abs().max(dim=1).values
est = phi * nt ** 4 * dm ** 2
est2 = est * (phi * nl * nt)          # variants
for name, e in (("est", est), ("phi", phi), ("est*phi", est * phi), ("p_a*alpha", p_host["p_a"] * p_host["alpha"]), ("actual", sw)):
    order = torch.argsort(e, descending=True)
    s = sw[order]
    n4 = (B // 4) * 4
    wmax = s[:n4].view(-1, 4).max(dim=1).values.mean()
    rank_corr = float(torch.corrcoef(torch.stack([torch.argsort(torch.argsort(e)).double(), torch.argsort(torch.argsort(sw)).double()]))[0, 1])
    print(f"{name:10s} warp-max mean (4/warp) {float(wmax):.2f}  rank-corr {rank_corr:.3f}")
perm = torch.randperm(B)
print("random     warp-max mean", float(sw[perm][: (B // 4) * 4].view(-1, 4).max(dim=1).values.mean()))
nan = torch.isnan(res["uout"][:, 2:]).any(dim=1).cpu()
print("nan strings", int(nan.sum()), "their mean sweeps", float(sw[nan].mean()) if nan.any() else None)
