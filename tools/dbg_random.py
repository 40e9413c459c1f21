"""Diagnostics (GPU): which strings of a mixed-excitation workload keep the fixed-point loop spinning."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch_fdtd_string_b200 import sampler
B = int(sys.argv[1]); Nt = int(sys.argv[2])
p_host = sampler.sample_nsynth_like(B, length=1.0, excitation="random", seed=1234)
p = sampler.to_device(p_host, torch.device("cuda"))
res = sampler.run_compact(p, 24, counters=True, n_run=Nt)
c = res["counters"].double().cpu(); st = res["status"].cpu()
outer = c[:, 0] / c[:, 3]; sw = c[:, 1] / c[:, 3]; ham = c[:, 2] / c[:, 3]
typ = torch.where(p_host["bow_mask"], 1, torch.where(p_host["hammer_mask"], 2, 0))
for t, nm in ((0, "pluck"), (1, "bow"), (2, "hammer")):
    m = typ == t
    print(nm, int(m.sum()), "outer mean %.2f max %.1f" % (outer[m].mean(), outer[m].max()), "sweeps mean %.2f max %.1f" % (sw[m].mean(), sw[m].max()),
          "status bits", sorted(set(st[m].tolist())))
g_outer = outer.view(-1, 24).max(dim=1).values
worst = torch.argsort(g_outer, descending=True)[:5]
print("worst groups", worst.tolist(), g_outer[worst].tolist())
g = int(worst[0])
sl = slice(g * 24, g * 24 + 24)
nan = torch.isnan(res["uout"][sl, 2:]).any(dim=1).cpu()
amax = res["uout"][sl, 2:].abs().nan_to_num(0).max(dim=1).values.cpu()
for s in range(24):
    i = g * 24 + s
    print(s, ["pluck", "bow", "hammer"][int(typ[i])], "outer %.2f sweeps %.2f ham %.2f st %d nan %d amax %.2e" % (outer[i], sw[i], ham[i], int(st[i]), int(nan[s]), float(amax[s])),
          "alpha %.1f p_a %.4f f0 %.0f" % (p_host["alpha"][i], p_host["p_a"][i], p_host["f0_a"][i]))
