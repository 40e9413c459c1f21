"""TEST / BASELINE INFRASTRUCTURE ONLY -- times the reference's own CPU implementation of the
hot path (oracle/_ref/forward_fn.so = the unmodified reference extension; falls back to the C
restatement when that .so was not built) on the host cores.  Executed by bench.py in a
subprocess with CUDA hidden (the reference picks CUDA whenever it is visible, misc.cpp:13-15).

Prints one JSON line: {"kind", "cores", "B", "Nt", "steps", "sec_per_call": [...], ...}
"""
import argparse
import json
import os
import sys
import time

os.environ["CUDA_VISIBLE_DEVICES"] = "-1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def reference_layout(p, Nt_s):
    """compact sampler output -> the argument lists of the reference's forward_fn for the first Nt_s samples"""
    from torch_fdtd_string_b200.sampler import expand_controls
    B = p["B"]
    q = dict(p); q["Nt_full"] = p["Nt"]
    c = expand_controls(p, "cpu")
    su = torch.zeros(B, Nt_s, p["Nx_t1"], dtype=torch.float64); su[:, :2] = p["state_u"]
    sz = torch.zeros(B, Nt_s, p["Nx_l1"], dtype=torch.float64)
    u0 = torch.zeros(B, 1, p["Nx_t1"], dtype=torch.float64)
    sl = lambda t: t[:, :Nt_s].contiguous()
    string_params = [p["kappa"], p["alpha"], u0, u0.clone(), p["p_a"].view(-1, 1, 1), sl(c["f0"]), p["pos"], p["T60"]]
    bow_params = [sl(c["x_b"]), sl(c["v_b"]), sl(c["F_b"]), p["phi_0"], p["phi_1"], sl(c["wid"])]
    hammer_params = [p["x_H"], torch.zeros(B, Nt_s, dtype=torch.float64), sl(c["u_H"]), p["w_H"], p["M_r"], p["alpha_H"]]
    return su, sz, string_params, bow_params, hammer_params, p["bow_mask"].view(-1, 1, 1), p["hammer_mask"].view(-1, 1, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=24)
    ap.add_argument("--nt", type=int, default=42)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--excitation", default="pluck")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    from torch_fdtd_string_b200.sampler import sample_nsynth_like
    kind = "reference"
    try:
        import build_ref
        if not os.path.exists(build_ref.ref_so_path()):
            raise FileNotFoundError(build_ref.ref_so_path())
        fwd = build_ref.load_ref().forward_fn
    except Exception as e:  # the C restatement of the same algorithm
        import sfdtd_oracle
        sfdtd_oracle.build()
        fwd = sfdtd_oracle.forward_fn
        kind = "port"
    p = sample_nsynth_like(a.B, excitation=a.excitation, seed=a.seed)
    times = []
    for it in range(a.warmup + a.steps):
        su, sz, sp, bp, hp, bm, hm = reference_layout(p, a.nt)
        t0 = time.perf_counter()
        with torch.no_grad():
            fwd(su, sz, sp, bp, hp, bm, hm, [p["k"], p["theta_t"], p["lambda_c"]], float(p["relative_order"]),
                True, False, 0, a.nt)
        dt = time.perf_counter() - t0
        if it >= a.warmup:
            times.append(dt)
    print(json.dumps(dict(kind=kind, cores=torch.get_num_threads() if kind == "reference" else 1, B=a.B, Nt=a.nt,
                          sr=p["sr"], steps=a.steps, sec_per_call=times, host_cpus=os.cpu_count())))


if __name__ == "__main__":
    main()
