"""BASELINE configs 1/3/4 (single-string latency): runs the golden fixtures' parameter sets for the full duration on the GPU
through the drop-in forward path and reports seconds of GPU time per simulated second (no oracle: parity on the prefix is
covered by tests/test_gpu_parity.py)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import golden_util as gu
from torch_fdtd_string_b200.forward_fn import step_strings

def run(name, seconds, sr=None):
    g = gu.load_golden(name)
    inp = gu.build_inputs(g, device="cuda")
    B = int(g["B"]); sr = int(g["sr"]) if sr is None else sr
    Nt = int(seconds * sr)
    sp, bp, hp = inp["string_params"], inp["bow_params"], inp["hammer_params"]
    rep = lambda t: t[:, :1].expand(B, Nt) if t.dim() == 2 else t           # fixtures with constant controls: extend in time
    su = inp["state_u"][:, :2].contiguous(); sz = inp["state_z"][:, :2].contiguous()
    uH = torch.zeros(B, Nt, dtype=torch.float64, device="cuda"); n0 = min(Nt, hp[2].size(1)); uH[:, :n0] = hp[2][:, :n0]
    args = dict(kappa=sp[0], alpha=sp[1], f0=rep(sp[5]), pos=sp[6], T60=sp[7], x_b=rep(bp[0]), v_b=rep(bp[1]), F_b=rep(bp[2]),
                wid=rep(bp[5]), phi_0=bp[3], phi_1=bp[4], x_H=hp[0], w_H=hp[3], M_r=hp[4], alpha_H=hp[5], u_H=uH,
                bow_mask=inp["bow_mask"], hammer_mask=inp["hammer_mask"], k=inp["consts"][0], theta_t=inp["consts"][1],
                lambda_c=inp["consts"][2], relative_order=inp["relative_order"], Nt=Nt, group_size=B,
                surface_integral=inp["surface_integral"], save_state=False, counters=True)
    for it in range(2):
        s1, z1, u1 = su.clone(), sz.clone(), uH.clone()
        a = dict(args); a["u_H"] = u1
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = step_strings(s1, z1, **a)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    c = res["counters"][0].tolist()
    u = res["uout"][:, 2:]
    print(f"{name}: B={B} Nt={Nt} ({seconds} s @ {sr}): {dt:.3f} s GPU wall -> {B * seconds / dt:.2f} string-seconds/s; outer/step {c[0] / c[3]:.2f} "
          f"sweeps/step {c[1] / c[3]:.2f}; finite={bool(torch.isfinite(u).all())} max|u|={float(u.abs().max()):.3e} status={int(res['status'].max())}")

run("pluck_b1", 1.0)
run("allfixed_bow_b1", 4.0)
run("finehammer_b1", 2.0)
