"""Helpers shared by the parity tests: load a golden fixture (made by
tests/golden/make_golden.py from the unmodified reference) back into the exact
argument list of the reference's ``process()`` / ``forward_fn``."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


LONG = ("pluck_b1_1s", "pluck_b24_1s", "allfixed_bow_b1_4s", "finehammer192_b1", "pluck_b24_01s", "pluck_hot_b6")   # full-length fixtures


def golden_names(long=False):
    """short fixtures (all outputs + states), or the full-length ones of the BASELINE configs (audio only)"""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    names = [n for n in names if not n.startswith("run_")]             # run_*: files of the reference's run() (test_gpu_dataset.py)
    return [n for n in names if (n in LONG) == long]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))


def build_inputs(g, dtype=torch.float64, device="cpu"):
    """-> dict(state_u, state_z, string_params[8], bow_params[6], hammer_params[6],
    bow_mask, hammer_mask, consts, Nt, chunk_size, relative_order, surface_integral, manufactured)"""
    B, Nt, Nx_t1, Nx_l1 = int(g["B"]), int(g["Nt"]), int(g["Nx_t1"]), int(g["Nx_l1"])

    def T(a):
        return torch.from_numpy(np.array(a)).to(dtype).to(device)

    state_u = torch.zeros(B, Nt, Nx_t1, dtype=dtype, device=device)
    state_z = torch.zeros(B, Nt, Nx_l1, dtype=dtype, device=device)
    if g["state_u_idx"].size:
        state_u[:, torch.from_numpy(g["state_u_idx"]).to(device), :] = T(g["state_u_rows"])
    if g["state_z_idx"].size:
        state_z[:, torch.from_numpy(g["state_z_idx"]).to(device), :] = T(g["state_z_rows"])
    u0 = torch.zeros(B, 1, Nx_t1, dtype=dtype, device=device)   # never read by the stepper
    v_H = torch.zeros(B, Nt, dtype=dtype, device=device)        # never read by the stepper

    def C(a):   # (B,Nt) control curve; the long format stores time-constant curves as (B,1)
        t = T(a)
        return t.expand(B, Nt).contiguous() if (t.dim() == 2 and t.size(1) == 1 and Nt > 1) else t

    if "u_H" in g:
        u_H = T(g["u_H"])
    else:                                                       # long format: non-zero columns only
        u_H = torch.zeros(B, Nt, dtype=dtype, device=device)
        if g["u_H_idx"].size:
            u_H[:, torch.from_numpy(g["u_H_idx"]).to(device)] = T(g["u_H_cols"])
    string_params = [T(g["kappa"]), T(g["alpha"]), u0, u0.clone(), T(g["p_a"]), C(g["f0"]), T(g["pos"]), T(g["T60"])]
    bow_params = [C(g["x_b"]), C(g["v_b"]), C(g["F_b"]), T(g["phi_0"]), T(g["phi_1"]), C(g["wid"])]
    hammer_params = [T(g["x_H"]), v_H, u_H, T(g["w_H"]), T(g["M_r"]), T(g["alpha_H"])]
    return dict(
        state_u=state_u, state_z=state_z, string_params=string_params, bow_params=bow_params,
        hammer_params=hammer_params,
        bow_mask=torch.from_numpy(g["bow_mask"]).to(device), hammer_mask=torch.from_numpy(g["hammer_mask"]).to(device),
        consts=[float(x) for x in g["consts"]], Nt=Nt, chunk_size=int(g["chunk_size"]),
        relative_order=float(g["relative_order"]), surface_integral=bool(g["surface_integral"]),
        manufactured=bool(g["manufactured"]))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    if den == 0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / den)


def run_process(forward_fn, inp):
    """The reference's chunk driver (reference src/task/simulate.py:38-119), restated
    for the tests so that any ``forward_fn`` implementation can be driven through it."""
    state_u, state_z = inp["state_u"], inp["state_z"]
    Nt, chunk_size = inp["Nt"], inp["chunk_size"]

    def chunk(x, n, size):
        if isinstance(x, torch.Tensor) and x.dim() > 1 and x.size(1) > 2:
            return x.narrow(1, n, size)
        return x

    cn = 0
    tot = [[] for _ in range(5)]
    sig0 = sig1 = None
    while cn < Nt - 2:
        size = min(chunk_size, state_u.size(1) - cn)
        outs = forward_fn(
            chunk(state_u, cn, size), chunk(state_z, cn, size),
            [chunk(p, cn, size) for p in inp["string_params"]],
            [chunk(p, cn, size) for p in inp["bow_params"]],
            [chunk(p, cn, size) for p in inp["hammer_params"]],
            inp["bow_mask"], inp["hammer_mask"], inp["consts"], inp["relative_order"],
            inp["surface_integral"], inp["manufactured"], cn, size)
        uout, zout, c_su, c_sz, v_r, F_H, u_H, sig0, sig1 = outs
        state_u[:, cn + 2:cn + size, :] = c_su[:, 2:2 + size, :]
        state_z[:, cn + 2:cn + size, :] = c_sz[:, 2:2 + size, :]
        for lst, t in zip(tot, (uout, zout, v_r, F_H, u_H)):
            lst.append(t.narrow(1, 2, size - 2))
        cn += chunk_size - 2
    uout, zout, v_r, F_H, u_H = [torch.cat(x, dim=1) for x in tot]
    return dict(uout=uout, zout=zout, state_u=state_u, state_z=state_z, v_r_out=v_r, F_H_out=F_H,
                u_H_out=u_H, sig0=sig0, sig1=sig1)
