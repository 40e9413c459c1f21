"""Test helper: expands the compact nsynth-like sampler output (torch_fdtd_string_b200.sampler)
into the reference-shaped argument list of ``process()`` / ``forward_fn`` so that the C oracle
(or any other ``forward_fn``) can be run on the same strings the throughput workloads use."""
import torch

from torch_fdtd_string_b200 import sampler


def device_controls(p, Nt=None):
    """the (B,Nt) control curves exactly as the stepper synthesises them from the compact description (needs a GPU),
    as CPU tensors -- the oracle must see the same doubles (floor(1/h(f0)) decides grid sizes)"""
    from torch_fdtd_string_b200 import synth_controls
    Nt = p["Nt"] if Nt is None else Nt
    dev = torch.device("cuda")
    c = synth_controls(sampler.synth_dict(sampler.to_device(p, dev)), p["B"], Nt, dev)
    return {k: v.cpu() for k, v in c.items()}


def reference_inputs(p, sl=None, Nt=None, controls=None):
    """p: sampler.sample_nsynth_like(...) (CPU).  sl: slice of strings (one reference batch).  Nt: prefix length.
    controls: (B,>=Nt) curves to use instead of the host-side expansion."""
    sl = slice(None) if sl is None else sl
    Nt = p["Nt"] if Nt is None else Nt
    c = controls if controls is not None else sampler.expand_controls(p, torch.device("cpu"))
    B = p["kappa"][sl].numel()
    su = torch.zeros(B, Nt, p["Nx_t1"], dtype=torch.float64)
    sz = torch.zeros(B, Nt, p["Nx_l1"], dtype=torch.float64)
    su[:, :2] = p["state_u"][sl]
    sz[:, :2] = p["state_z"][sl]
    u0 = torch.zeros(B, 1, p["Nx_t1"], dtype=torch.float64)
    cut = lambda t: t[sl, :Nt].contiguous()
    string_params = [p["kappa"][sl], p["alpha"][sl], u0, u0.clone(), p["p_a"][sl].view(-1, 1, 1), cut(c["f0"]),
                     p["pos"][sl], p["T60"][sl]]
    bow_params = [cut(c["x_b"]), cut(c["v_b"]), cut(c["F_b"]), p["phi_0"][sl], p["phi_1"][sl], cut(c["wid"])]
    hammer_params = [p["x_H"][sl], torch.zeros(B, Nt, dtype=torch.float64), cut(c["u_H"]), p["w_H"][sl], p["M_r"][sl],
                     p["alpha_H"][sl]]
    return dict(state_u=su, state_z=sz, string_params=string_params, bow_params=bow_params, hammer_params=hammer_params,
                bow_mask=p["bow_mask"][sl].view(-1, 1, 1), hammer_mask=p["hammer_mask"][sl].view(-1, 1, 1),
                consts=[p["k"], p["theta_t"], p["lambda_c"]], Nt=Nt, chunk_size=Nt,
                relative_order=float(p["relative_order"]), surface_integral=True, manufactured=False)
