"""Debug (GPU): compare one sampled group against the oracle step by step."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import sfdtd_oracle, golden_util as gu, sampler_util as su
from torch_fdtd_string_b200 import sampler
seed, Btot, g, Nt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
full = len(sys.argv) > 5 and sys.argv[5] == "full"
GROUP = 24
p = sampler.sample_nsynth_like(Btot, length=1.0, excitation="pluck", seed=seed)
sl = slice(g * GROUP, (g + 1) * GROUP)
ref = gu.run_process(sfdtd_oracle.forward_fn, su.reference_inputs(p, sl, Nt))
if full:
    q = sampler.to_device(p, torch.device("cuda"))
    res = sampler.run_compact(q, GROUP, counters=True, n_run=Nt)
    uo = res["uout"][sl, 2:].cpu().numpy(); cnt = res["counters"][sl].cpu().numpy(); st = res["status"][sl].cpu().numpy()
else:
    sub = {k: (v[sl] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.size(0) == Btot else v) for k, v in p.items()}
    sub["B"] = GROUP
    q = sampler.to_device(sub, torch.device("cuda"))
    res = sampler.run_compact(q, GROUP, counters=True, n_run=Nt)
    uo = res["uout"][:, 2:].cpu().numpy(); cnt = res["counters"].cpu().numpy(); st = res["status"].cpu().numpy()
for s in range(GROUP):
    r = ref["uout"][s].numpy()
    e = gu.rel_l2(uo[s], r)
    d = np.abs(uo[s] - r) / (np.abs(r).max() + 1e-300)
    first = int(np.argmax(d > 1e-9)) if (d > 1e-9).any() else -1
    print(s, "err %.2e" % e, "first step >1e-9:", first, "sweeps/step %.2f" % (cnt[s, 1] / max(1, cnt[s, 3])), "status", st[s])
