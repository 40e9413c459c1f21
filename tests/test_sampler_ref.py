"""CPU: the RNG-stream-compatible compact sampler (torch_fdtd_string_b200/sampler_ref.py) reproduces what the unmodified
reference's String / Bow / Hammer modules handed to process() for the same seed -- checked against the INPUT arrays of the
golden fixtures (captured from the reference by tests/golden/make_golden.py; seeds and sampler arguments in
tests/golden/make_golden.py CASES / tests/golden/presets.py)."""
import os
import sys

import numpy as np
import pytest
import torch

import golden_util as gu

sys.path.insert(0, gu.GOLDEN_DIR)
import presets  # noqa: E402

from torch_fdtd_string_b200 import sampler, sampler_ref  # noqa: E402

# fixture -> (preset, model, B, length, seed[, manufactured])
CASES = {
    "pluck_b3": ("nsynth", "pluck", 3, 0.01, 1234),
    "hammer_b3": ("nsynth", "hammer", 3, 0.01, 1234),
    "bow_b3": ("nsynth", "bow", 3, 0.01, 1234),
    "random_b6": ("nsynth", "random", 6, 0.01, 7),
    "pluck_b24": ("nsynth", "pluck", 24, 0.004, 1234),
    "random_b24": ("nsynth", "random", 24, 0.004, 3),
    "random_b4_long": ("nsynth", "random", 4, 0.05, 9),
    "allfixed_bow_b1": ("allfixed", "bow", 1, 0.01, 1234),
    "finehammer_b1": ("finehammer", "hammer", 1, 0.002, 1234),
    "manufactured_b1": ("linear", "pluck", 1, 0.005, 1234, True),
}


def draw(preset, model, B, length, seed, manufactured=False):
    p = presets.PRESETS[preset]
    _, kap, f0m = p["theta"]
    theta_t = sampler.get_theta(kap, f0m, p["sr"])
    torch.manual_seed(seed)
    return sampler_ref.sample_reference(B, model, p["sr"], length, theta_t, p["f0_inf"], p["alpha_inf"], p["lambda_c"],
                                        precision="double", string_kwargs=p["string_kwargs"], bow_kwargs=p["bow_kwargs"],
                                        hammer_kwargs=p["hammer_kwargs"], manufactured=manufactured,
                                        relative_order=p["relative_order"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_same_seed_gives_the_reference_parameters(name):
    g = gu.load_golden(name)
    q = draw(*CASES[name])
    B, Nt = int(g["B"]), int(g["Nt"])
    assert (q["B"], q["Nt"], q["Nx_t1"], q["Nx_l1"]) == (B, Nt, int(g["Nx_t1"]), int(g["Nx_l1"]))
    assert np.array_equal(q["bow_mask"].numpy(), g["bow_mask"].reshape(-1).astype(bool))
    assert np.array_equal(q["hammer_mask"].numpy(), g["hammer_mask"].reshape(-1).astype(bool))
    # per-string scalars: bit-identical
    for k in ("kappa", "alpha", "pos", "phi_0", "phi_1", "x_H", "w_H", "M_r", "alpha_H"):
        assert np.array_equal(q[k].numpy().reshape(-1), g[k].reshape(-1)), k
    assert np.array_equal(q["T60"].numpy(), g["T60"])
    assert np.array_equal(q["p_a"].numpy().reshape(-1), g["p_a"].reshape(-1))
    # initial state rows: bit-identical (same torch ops on the plucked time slice)
    rows = dict(zip(g["state_u_idx"].tolist(), range(len(g["state_u_idx"]))))
    for r in (0, 1):
        ref = g["state_u_rows"][:, rows[r], :] if r in rows else np.zeros((B, int(g["Nx_t1"])))
        assert np.array_equal(q["state_u"][:, r].numpy(), ref), ("state_u row", r)
    # control curves: to rounding (the reference divides the whole f0 curve by the Fletcher factor, F.interpolate's lerp)
    c = sampler.expand_controls(q, torch.device("cpu"))
    for k, tol in (("f0", 1e-14), ("x_b", 1e-14), ("v_b", 1e-14), ("F_b", 1e-13), ("wid", 0.0), ("u_H", 1e-16)):
        ref = g[k]
        err = float(np.abs(c[k].numpy() - ref).max() / max(np.abs(ref).max(), 1e-300))
        assert err <= tol, (k, err)


def test_successive_batches_continue_the_stream():
    """run.py seeds once and loops over batches (reference run.py:75, src/task/simulate.py:272): the second batch of a
    two-batch run differs from the first and does not depend on anything but the stream position."""
    torch.manual_seed(1234)
    p = presets.PRESETS["nsynth"]
    theta_t = sampler.get_theta(0.03, 98.0, 48000)
    kw = dict(string_kwargs=p["string_kwargs"], bow_kwargs=p["bow_kwargs"], hammer_kwargs=p["hammer_kwargs"])
    a = sampler_ref.sample_reference(3, "pluck", 48000, 0.01, theta_t, 98.0, 1, 1, **kw)
    b = sampler_ref.sample_reference(3, "pluck", 48000, 0.01, theta_t, 98.0, 1, 1, **kw)
    g = gu.load_golden("pluck_b3")
    assert np.array_equal(a["kappa"].numpy(), g["kappa"].reshape(-1))
    assert not np.array_equal(a["kappa"].numpy(), b["kappa"].numpy())


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference checkout")
def test_load_config_overrides_match_the_reference_modules():
    """task.load_config (reference src/task/simulate.py:164-185): the reference's String / Bow / Hammer modules after
    dump_parameter('f0' | 'v_b' | 'v_H', curve) against dataset.apply_overrides on the RNG-compatible compact batch."""
    sys.path.insert(0, os.path.join(os.path.dirname(gu.GOLDEN_DIR), "..", "oracle"))
    import ref_driver
    ref_driver.import_reference()
    import src.model.simulator as rs
    from torch_fdtd_string_b200 import dataset
    p = presets.PRESETS["nsynth"]
    sr, length, B = p["sr"], 0.01, 3
    Nt = int(sr * length)
    theta_t = sampler.get_theta(0.03, 98.0, sr)
    rng = np.random.RandomState(0)
    over = {"string-f0": 200.0 + 20.0 * np.linspace(0, 1, Nt) + rng.rand(Nt), "bow-v_b": 0.3 + 0.05 * rng.rand(Nt),
            "hammer-v_H": np.concatenate([np.zeros(3), np.ones(2), np.zeros(Nt - 5)])}
    # the reference: modules, then the dumps in simulate()'s order (after all three are constructed)
    torch.manual_seed(77)
    bm, hm = torch.zeros(B, dtype=torch.bool), torch.ones(B, dtype=torch.bool)
    pm = ~(bm | hm)
    string = rs.String(1 / sr, theta_t, p["lambda_c"], sr, length, p["f0_inf"], p["alpha_inf"], B, "double", False, pm, hm, "batch", False,
                       **p["string_kwargs"])
    bow = rs.Bow(sr, length, B, "double", "batch", **p["bow_kwargs"])
    hammer = rs.Hammer(sr, length, B, "double", 1 / sr, "batch", **p["hammer_kwargs"])
    string.dump_parameter("f0", over["string-f0"]); bow.dump_parameter("v_b", over["bow-v_b"]); hammer.dump_parameter("v_H", over["hammer-v_H"])
    with torch.no_grad():
        sp, bp, hp = string(), bow(), hammer()
    f0_ref, vb_ref, uH_ref = sp[7].numpy(), bp[1].numpy(), hp[2].numpy()          # after the two state tensors: kappa alpha u0 v0 p_a f0
    # this package
    torch.manual_seed(77)
    q = sampler_ref.sample_reference(B, "hammer", sr, length, theta_t, p["f0_inf"], p["alpha_inf"], p["lambda_c"], precision="double",
                                     string_kwargs=p["string_kwargs"], bow_kwargs=p["bow_kwargs"], hammer_kwargs=p["hammer_kwargs"],
                                     redraw_v_H=True)
    ctl = dataset.apply_overrides(q, dict(sampler.expand_controls(q, torch.device("cpu"))), over)
    assert np.abs(ctl["f0"].numpy() - f0_ref).max() <= 1e-12 * np.abs(f0_ref).max()
    assert np.array_equal(ctl["v_b"].numpy(), np.broadcast_to(over["bow-v_b"], (B, Nt))) and np.array_equal(vb_ref, ctl["v_b"].numpy())
    assert np.abs(ctl["u_H"].numpy() - uH_ref).max() <= 1e-18
    # the other draws are untouched by the dumps
    assert np.array_equal(q["kappa"].numpy(), sp[2].detach().numpy().reshape(-1))
