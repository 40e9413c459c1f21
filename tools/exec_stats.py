"""Diagnostics (GPU): sweeps a string needs (counted until its own convergence) vs sweeps its warp executes (lockstep +
predicted minimum), per longitudinal-size class.  Run twice: default library, then SFDTD_LIB=.../lib_cntexec.so."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch_fdtd_string_b200 import sampler
B = 14208; length = 0.05
dev = torch.device("cuda")
p_host = sampler.sample_nsynth_like(B, length=length, excitation="pluck", seed=1234)
p = sampler.to_device(p_host, dev)
res = sampler.run_compact(p, 24, counters=True)
c = res["counters"].double().cpu()
sw = c[:, 1] / c[:, 3]
f0 = torch.minimum(p_host["f0_a"], p_host["f0_b"])
nt, nl = sampler.derived_grid(f0, p_host["kappa"], p_host["k"], p_host["theta_t"], 1.0, p_host["alpha"])
for lo, hi in ((0, 12), (12, 28), (28, 60), (60, 1000)):
    m = (nl >= lo) & (nl < hi)
    print(f"N_l in [{lo},{hi}): {int(m.sum())} strings, mean sweeps {float(sw[m].mean()):.3f}")
print("all", float(sw.mean()))
