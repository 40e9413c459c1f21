"""GPU: the round-2 additions to the C ABI -- in-kernel control synthesis (sfdtd_synth / sfdtd_synth_controls), plans
(sfdtd_plan_create / sfdtd_forward_plan / sfdtd_plan_destroy), audio-only outputs and device post-processing
(sfdtd_postprocess) -- against the table-driven path and against numpy restatements of the reference's host code."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GROUP = 24
B = 20 * GROUP


@pytest.fixture(scope="module", params=["pluck", "random"])
def batch(request):
    from torch_fdtd_string_b200 import sampler
    return sampler.sample_nsynth_like(B, length=0.25, excitation=request.param, seed=77)


def _dev(p):
    from torch_fdtd_string_b200 import sampler
    return sampler.to_device(p, torch.device("cuda"))


def test_synth_controls_match_host_expansion(batch):
    """sfdtd_synth_controls (what the stepper evaluates) vs the sampler's PyTorch expansion of the same compact scalars
    (reference src/utils/control.py:5-45, src/model/simulator.py:210-235,419-484,573-578)."""
    from torch_fdtd_string_b200 import sampler, synth_controls
    p = _dev(batch)
    Nt = batch["Nt"]
    c = synth_controls(sampler.synth_dict(p), B, Nt, p["kappa"].device)
    ref = sampler.expand_controls(p, p["kappa"].device)
    for k in ("f0", "x_b", "v_b", "F_b", "u_H", "wid"):
        err = float((c[k] - ref[k]).abs().max() / ref[k].abs().max().clamp(min=1e-300))
        assert err < 1e-13, (k, err)


def test_synth_equals_table_driven_bitwise(batch):
    """The stepper with synthesised controls == the stepper reading the same curves from (B,Nt) arrays, bit for bit
    (every kernel evaluates the curves with the same explicitly rounded operations)."""
    from torch_fdtd_string_b200 import sampler, synth_controls
    p = _dev(batch)
    Nt = 1202
    a = sampler.run_compact(p, GROUP, n_run=Nt, counters=True)
    c = synth_controls(sampler.synth_dict(p), B, Nt, p["kappa"].device)
    b = sampler.run_compact(p, GROUP, n_run=Nt, counters=True, controls={k: v.clone() for k, v in c.items()})
    torch.cuda.synchronize()
    for k in ("uout", "zout", "v_r", "F_H", "u_H_out", "sig0", "sig1"):
        x, y = a[k], b[k]
        same = (x == y) | (torch.isnan(x) & torch.isnan(y))
        assert bool(same.all()), (k, int((~same).sum()))
    assert torch.equal(a["counters"], b["counters"])


def test_plan_runs_repeat_and_audio_only(batch):
    """A plan is reusable: two queued runs give identical audio; audio-only calls (v_r / F_H / u_H_out NULL) give the same
    audio as full calls."""
    from torch_fdtd_string_b200 import sampler, Plan
    p = _dev(batch)
    Nt = 962
    full = sampler.run_compact(p, GROUP, n_run=Nt)
    a, res, keep = sampler.compact_args(p, GROUP, n_run=Nt, aux_outputs=False)
    plan = Plan(a)
    plan.run(a)
    torch.cuda.synchronize()
    u1 = res["uout"].clone(); z1 = res["zout"].clone()
    # (the state rows were advanced in place: the second run gets fresh ones, same plan)
    keep2 = sampler.compact_args(p, GROUP, n_run=Nt, aux_outputs=False, out={"uout": res["uout"], "zout": res["zout"]})
    plan.run(keep2[0])
    plan.close()
    torch.cuda.synchronize()
    for x, y in ((u1, res["uout"]), (z1, res["zout"]), (u1, full["uout"]), (z1, full["zout"])):
        same = (x[:, 2:] == y[:, 2:]) | (torch.isnan(x[:, 2:]) & torch.isnan(y[:, 2:]))
        assert bool(same.all())
    assert set(res) >= {"uout", "zout", "sig0", "sig1", "status"} and "v_r" not in res


def test_postprocess_matches_host_restatement(batch):
    """sfdtd_postprocess vs numpy: NaN mask, silence test, l-infinity gain (reference src/task/simulate.py:333-337,
    src/utils/audio.py:42-48,72-76) and the PCM_24 / PCM_16 samples a wav writer would store."""
    from torch_fdtd_string_b200 import sampler, postprocess
    p = _dev(batch)
    Nt = 1502
    res = sampler.run_compact(p, GROUP, n_run=Nt, aux_outputs=False)
    u = res["uout"].clone(); z = res["zout"]
    u[3, 100] = float("nan")                          # a NaN string
    u[5, 2:] *= 1e-4                                  # a silent one
    u[7, 2:] = 0.0                                    # an all-zero one (gain 1)
    for bits in (24, 16):
        pp = postprocess(u, z, n0=2, bits=bits)
        torch.cuda.synchronize()
        un = u[:, 2:].cpu().numpy(); zn = z[:, 2:].cpu().numpy()
        nan = np.isnan(un.sum(-1))
        uz = un * ~nan[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            silent = 20 * np.log10(np.sqrt(np.mean(np.square(np.nan_to_num(uz)), -1))) <= -23.0
        mx = np.abs(un).max(-1)
        gain = np.where((mx == 0) | np.isnan(mx), 1.0, 1.0 / np.where(mx == 0, 1, mx))
        assert np.array_equal(pp["is_nan"].cpu().numpy().astype(bool), nan)
        assert np.array_equal(pp["is_silent"].cpu().numpy().astype(bool)[~nan], silent[~nan])
        np.testing.assert_allclose(pp["gain"].cpu().numpy()[~nan], gain[~nan], rtol=1e-15)
        sc = float(1 << (bits - 1))
        by = bits // 8
        for key, sig in (("pcm_u", gain[:, None] * un), ("pcm_z", gain[:, None] * zn), ("pcm_w", gain[:, None] * un + gain[:, None] * zn)):
            raw = pp[key].cpu().numpy()[:, : (Nt - 2) * by].reshape(B, Nt - 2, by).astype(np.int64)
            v = sum(raw[:, :, k] << (8 * k) for k in range(by))
            v = np.where(v >= 1 << (bits - 1), v - (1 << bits), v)
            want = np.clip(np.rint(np.nan_to_num(sig) * sc), -sc, sc - 1).astype(np.int64)
            assert np.array_equal(v[~nan], want[~nan]), (bits, key, int((v[~nan] != want[~nan]).sum()))
    assert bool(pp["is_nan"][3]) and bool(pp["is_silent"][5]) and float(pp["gain"][7]) == 1.0


def test_tensors_on_a_non_current_device_are_followed():
    """The C ABI switches to the device the state lives on (kernels were launched on the current device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from torch_fdtd_string_b200 import sampler
    p_host = sampler.sample_nsynth_like(GROUP, length=0.01, seed=3)
    a = sampler.run_compact(sampler.to_device(p_host, torch.device("cuda:0")), GROUP)
    with torch.cuda.device(0):
        b = sampler.run_compact(sampler.to_device(p_host, torch.device("cuda:1")), GROUP)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    assert torch.equal(a["uout"].cpu(), b["uout"].cpu())
