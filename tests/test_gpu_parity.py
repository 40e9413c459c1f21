"""GPU: parity of the CUDA stepper (through the C ABI, via the reference-shaped forward_fn /
process) against the golden vectors of the unmodified reference, and against the C oracle on
seeded inputs.  Tolerance (BASELINE.json north_star): max relative L2 <= 1e-6 in fp64 on the
output waveform; sample counts bit-exact."""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

KEYS = ["uout", "zout", "v_r_out", "F_H_out", "u_H_out"]
# Gate of the short fixtures: measured on B200 (round 2) every key of every fixture is <= 2.4e-10 except on the two
# fixtures that contain strings the reference itself is sensitive on (below) -> 3e-8 (~100 x measured; north_star asks 1e-6).
TOL_U = 3e-8
# `pluck_b24` / `pluck_b2_long` contain strings whose dynamics amplify 1-ulp perturbations in the reference itself; they
# carry the reference-vs-perturbed-reference distance (`pert_*`), and uout / zout are checked per string and window against
# 300 x that distance (floor 3e-8); their other keys are checked on the strings that are calm in the reference.
SENSITIVE = ("pluck_b24", "pluck_b2_long")
NOT_BUILT = set()


def run_cuda(g, chunk=None):
    from torch_fdtd_string_b200 import process
    inp = gu.build_inputs(g, device="cuda")
    if chunk is not None:
        inp["chunk_size"] = chunk
    out = process("unused", inp["state_u"], inp["state_z"], inp["string_params"], inp["bow_params"],
                  inp["hammer_params"], inp["bow_mask"], inp["hammer_mask"], inp["consts"], inp["Nt"],
                  inp["chunk_size"], None, True, inp["relative_order"], inp["surface_integral"], inp["manufactured"])
    names = ["uout", "zout", "state_u", "state_z", "v_r_out", "F_H_out", "u_H_out", "sig0", "sig1"]
    return dict(zip(names, out)), inp


@pytest.mark.parametrize("name", [n for n in gu.golden_names() if n not in NOT_BUILT])
def test_cuda_matches_reference_golden(name):
    g = gu.load_golden(name)
    out, inp = run_cuda(g)
    B = int(g["B"])
    for k in KEYS:
        assert tuple(out[k].shape) == g[k].shape, (k, out[k].shape, g[k].shape)
    calm = np.ones(B, dtype=bool)
    if name in SENSITIVE and "pert_win" in g:
        for key in ("uout", "zout"):
            x = out[key].cpu().numpy(); r = g[key]
            for b in range(B):
                worst, tot, sens = _check_against_reference_sensitivity(g, key, b, x[b], r[b], floor=TOL_U)
                print(f"{name} {key}[{b}] rel.L2 {tot:.2e} (reference vs itself {sens:.2e}) worst window err/bound {worst:.2f}")
                assert worst <= 1.0, (name, key, b, worst)
                if sens > 1e-12:
                    calm[b] = False
        assert calm.sum() >= B // 2
    elif name in SENSITIVE:
        pytest.skip("perturbed-reference data not merged into the fixture yet")
    errs = {}
    sel = np.nonzero(calm)[0]
    for k in KEYS:
        errs[k] = gu.rel_l2(out[k].cpu().numpy()[sel], g[k][sel])
    errs["state_u_last"] = gu.rel_l2(out["state_u"][:, -2:, :].cpu().numpy()[sel], g["state_u_last"][sel])
    errs["state_z_last"] = gu.rel_l2(out["state_z"][:, -2:, :].cpu().numpy()[sel], g["state_z_last"][sel])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < TOL_U, (name, k, v)
    np.testing.assert_allclose(out["sig0"].cpu().numpy().ravel(), g["sig0"].ravel(), rtol=1e-10)
    assert gu.rel_l2(inp["hammer_params"][2].cpu().numpy()[sel], g["u_H_inplace"][sel]) < TOL_U
    if "state_u_full" in g:
        assert gu.rel_l2(out["state_u"].cpu().numpy()[sel], g["state_u_full"][sel]) < TOL_U


# ---- the BASELINE configs at their stated length (fixtures: audio only, made by tests/golden/make_golden.py) ----------
# Gate: 1e-6 relative L2 (north_star) over the whole run -- except where the REFERENCE ITSELF is more sensitive than that.
# Its own sensitivity is measured, not assumed: the fixtures `pluck_b1_1s` / `pluck_b24_1s` carry the distance between the
# unmodified reference and the unmodified reference run on state_u * (1 + 2^-50) (`pert_*`, per string and 10 ms window;
# tests/golden/make_golden.py --perturbed / --merge-pert).  For the nsynth-like string of configs[0] that distance is
# 4e-13 after 10 ms, 1e-6 after 60 ms and 1e-3 after 120 ms: the scheme amplifies one ulp by ~30x per 10 ms, so no
# implementation (including the reference on another BLAS) can agree with it to 1e-6 over a second.  There the bound per
# prefix of the run is 300 x the reference's own distance (see SENS_FACTOR).
FULL_LENGTH = {
    "finehammer192_b1": ("uout", "zout", "v_r_out", "F_H_out", "u_H_out"),     # configs[3] at 192 kHz: 9 598 samples, N_t = 237
    "allfixed_bow_b1_4s": ("uout", "zout", "v_r_out"),                         # configs[2]: 191 998 samples
}
# measured on B200 this round (rel. L2 vs the reference): finehammer192_b1 uout 2.7e-10 zout 1.7e-10;
# allfixed_bow_b1_4s see profiles/parity_r02.md -> gates at <= 100 x measured, never looser than 1e-6
FULL_LENGTH_GATE = {"finehammer192_b1": 3e-8, "allfixed_bow_b1_4s": 1e-6}


def _have(name):
    import os
    return os.path.exists(os.path.join(gu.GOLDEN_DIR, f"{name}.npz"))


@pytest.mark.parametrize("name", sorted(FULL_LENGTH))
def test_full_length_config_matches_reference(name):
    if not _have(name):
        pytest.skip(f"fixture {name} not generated")
    g = gu.load_golden(name)
    out, inp = run_cuda(g)
    errs = {}
    for k in FULL_LENGTH[name]:
        assert tuple(out[k].shape) == g[k].shape, (k, out[k].shape, g[k].shape)      # bit-exact sample counts (Nt - 2)
        errs[k] = gu.rel_l2(out[k].cpu().numpy(), g[k])
    errs["state_u_last"] = gu.rel_l2(out["state_u"][:, -2:, :].cpu().numpy(), g["state_u_last"])
    errs["state_z_last"] = gu.rel_l2(out["state_z"][:, -2:, :].cpu().numpy(), g["state_z_last"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert not np.isnan(g["uout"]).any()
    for k, v in errs.items():
        assert v < FULL_LENGTH_GATE[name], (name, k, v)


# The reference's distance comes from ONE perturbation of 1 ulp (2^-50) of the initial state; the block iteration here
# stops at a predicted relative error of 1e-13 (~1000 ulp) in EVERY step.  Measured on B200 over all sensitive strings of all
# fixtures: up to 173 x the reference's own distance (string 5 of pluck_b24_1s; pluck_b2_long 97 x) -> factor 300.
SENS_FACTOR = 300.0


def _check_against_reference_sensitivity(g, key, b, x, r, factor=SENS_FACTOR, floor=1e-6):
    """x, r: one string's samples (CUDA, reference).  For every prefix of the run, in steps of one 10 ms window: relative L2
    distance over the prefix <= max(floor, factor x the reference's own distance to its perturbed run over the same prefix).
    (Prefixes, not single windows: two decorrelating trajectories cross and re-separate, so the distance inside one window
    fluctuates by an order of magnitude in the reference's own pair of runs.)
    Returns (worst ratio err / bound over the prefixes, total rel. L2, total bound)."""
    win = int(g["pert_win"])
    nw = min(len(r) // win, g[f"pert_{key}_err"].shape[1])
    worst = 0.0
    ce = cs = cn = 0.0
    used = 0
    for w in range(nw):
        sl = slice(w * win, (w + 1) * win)
        nr = np.linalg.norm(r[sl])
        if not np.isfinite(nr) or not np.isfinite(x[sl]).all():
            break
        ce += float(np.sum((x[sl] - r[sl]) ** 2)); cn += float(nr ** 2); cs += float(g[f"pert_{key}_err"][b, w]) ** 2
        used = w + 1
        if cn == 0:
            continue
        worst = max(worst, np.sqrt(ce / cn) / max(floor, factor * np.sqrt(cs / cn)))
    if used == 0 or cn == 0:
        return 0.0, 0.0, 0.0
    return worst, float(np.sqrt(ce / cn)), float(np.sqrt(cs / cn))


def test_single_string_full_length_within_reference_sensitivity():
    """BASELINE configs[0]: single plucked nsynth-like string, 1 s @ 48 kHz, fp64, 47 998 samples."""
    name = "pluck_b1_1s"
    if not _have(name):
        pytest.skip(f"fixture {name} not generated")
    g = gu.load_golden(name)
    out, _ = run_cuda(g)
    for key in ("uout", "zout"):
        x = out[key].cpu().numpy()[0]; r = g[key][0]
        assert x.shape == r.shape and np.isfinite(x).all() and np.isfinite(r).all()
        worst, tot, sens = _check_against_reference_sensitivity(g, key, 0, x, r)
        first = np.linalg.norm(x[:480] - r[:480]) / np.linalg.norm(r[:480])
        print(name, key, f"first 10 ms {first:.2e}; whole second {tot:.2e} (reference vs its own 1-ulp perturbation: {sens:.2e}); "
              f"worst window err/bound {worst:.2f}")
        assert first < 1e-9, (key, first)           # before the amplification sets in the paths agree to ~1e-12
        assert worst <= 1.0, (key, worst)
    for k in ("F_H_out", "u_H_out"):               # not fed by the chaotic displacement of an un-hammered string
        assert gu.rel_l2(out[k].cpu().numpy(), g[k]) < 1e-9, k


def _nan_mask_and_per_string_parity(name, need_nan=0):
    """Per string of fixture `name` (reference run + its 1-ulp-perturbed twin): NaN set and onset, parity on run prefixes."""
    g = gu.load_golden(name)
    out, _ = run_cuda(g)
    rep = []
    n_nan = 0
    for key in ("uout", "zout"):
        x = out[key].cpu().numpy(); r = g[key]
        assert x.shape == r.shape
        bad_x, bad_r = ~np.isfinite(x), ~np.isfinite(r)
        nan_x, nan_r = bad_x.any(1), bad_r.any(1)
        on_x = np.where(nan_x, bad_x.argmax(1), -1); on_r = np.where(nan_r, bad_r.argmax(1), -1)
        for b in range(x.shape[0]):
            on_p = int(g[f"pert_{key}_nan_onset"][b])                 # NaN onset of the reference's own perturbed run (-1: none)
            if nan_r[b] or nan_x[b] or on_p >= 0:
                # NaN strings: the reference's set; onset within the reference's own sensitivity (its perturbed run's onset) or
                # 10 ms.  A string that blows up in only one of the reference's two runs (unperturbed / 1-ulp perturbed) is
                # undecided in the reference itself: either outcome is accepted for it, the samples before the earliest onset
                # are still compared.
                undecided = bool(nan_r[b]) != (on_p >= 0)
                if not undecided:
                    assert nan_r[b] and nan_x[b], (key, b, "NaN mask differs", int(on_x[b]), int(on_r[b]))
                    # (the last stretch before the overflow is a diverging run-away in which the reference's dense inverse and
                    # the capped block iteration here no longer compute the same thing: measured onset distances on
                    # `pluck_hot_b6` are 7 ... 560 samples against 7 ... 386 between the reference's own two runs -> 20 ms)
                    slack = max(960, 3 * abs(on_p - int(on_r[b])))
                    assert abs(int(on_x[b]) - int(on_r[b])) <= slack, (key, b, int(on_x[b]), int(on_r[b]), slack)
                    n_nan += key == "uout"
                    print(f"{name} {key}[{b}] NaN onset: reference {int(on_r[b])}, its perturbed run {on_p}, CUDA {int(on_x[b])}")
                onsets = [int(o) for o, f in ((on_x[b], nan_x[b]), (on_r[b], nan_r[b]), (on_p, on_p >= 0)) if f]
                n_ok = max(0, min(onsets) - 480)
            else:
                n_ok = x.shape[1]
            worst, tot, sens = _check_against_reference_sensitivity(g, key, b, x[b, :n_ok], r[b, :n_ok])
            rep.append((key, b, "nan" if (nan_r[b] or nan_x[b] or on_p >= 0) else "ok", float(tot), float(sens), float(worst)))
    for row in rep:
        print("%s[%2d] %-3s rel.L2 %.2e  (reference vs itself %.2e)  worst window err/bound %.2f" % row)
    assert max(r[5] for r in rep) <= 1.0
    # strings that are NOT sensitive in the reference must meet the plain 1e-6 bound over the whole run
    calm = [r for r in rep if r[2] == "ok" and r[4] < 1e-8]
    assert all(r[3] < 1e-6 for r in calm), [r for r in calm if r[3] >= 1e-6]
    assert n_nan >= need_nan, (n_nan, need_nan)


def test_nsynth_batch_full_length_nan_mask_and_per_string_parity():
    """BASELINE configs[1]: one nsynth-like reference batch (24 strings, 1 s @ 48 kHz, fp64) at full length.
    * the set of strings that blow up to NaN is the reference's, and they do so at (nearly) the same sample
      (reference src/task/simulate.py:91-93,333-334 drops them);
    * every other string matches the reference per string and on every prefix of the run: <= 1e-6 relative L2, except where
      the reference itself is more sensitive (see above): there <= 300 x the reference's own distance."""
    if not _have("pluck_b24_1s"):
        pytest.skip("fixture pluck_b24_1s not generated")
    _nan_mask_and_per_string_parity("pluck_b24_1s")


def test_nan_onset_of_strings_that_blow_up():
    """`pluck_hot_b6`: six strings with alpha 18-25 and p_a = 0.02 (the corner of the nsynth-like ranges where the scheme
    diverges), 60 ms from the unmodified reference and from its 1-ulp-perturbed twin: the strings that go NaN are the
    reference's, at the reference's sample (within three times its own onset shift, at least 20 ms of slack)."""
    if not _have("pluck_hot_b6"):
        pytest.skip("fixture pluck_hot_b6 not generated")
    _nan_mask_and_per_string_parity("pluck_hot_b6", need_nan=1)


def test_chunked_equals_unchunked_on_gpu():
    g = gu.load_golden("hammer_b2_chunked")
    a, _ = run_cuda(g)
    b, _ = run_cuda(g, chunk=int(g["Nt"]))
    for k in KEYS:
        assert gu.rel_l2(a[k].cpu().numpy(), b[k].cpu().numpy()) < 1e-12, k


@pytest.mark.parametrize("name,B", [("random_b24", 24), ("random_b24", 20), ("pluck_b24", 21)])
def test_cuda_matches_oracle_on_seeded_groups(oracle, name, B):
    """Several groups in one launch (native API), compact state, vs the C oracle per group; B = 20 / 21 leave a short
    last group (4 / 5 of 8 strings) in the grouped (random) and in the independent (pluck) mode."""
    from torch_fdtd_string_b200 import step_strings
    g = gu.load_golden(name)
    inp = gu.build_inputs(g)                      # CPU tensors
    Nt = 96
    cutB = lambda t: t[:B] if (isinstance(t, torch.Tensor) and t.dim() > 0 and t.size(0) == int(g["B"])) else t
    inp = {k: ([cutB(x) for x in v] if isinstance(v, list) and k.endswith("_params") else cutB(v)) for k, v in inp.items()}
    # oracle: three groups of 8 strings, each its own reference batch
    ref = {k: [] for k in KEYS}
    for g0 in range(0, B, 8):
        sl = slice(g0, min(g0 + 8, B))
        sub = dict(inp)
        sub["state_u"] = inp["state_u"][sl, :Nt].clone(); sub["state_z"] = inp["state_z"][sl, :Nt].clone()
        sub["string_params"] = [p[sl, :Nt].clone() if (p.dim() > 1 and p.size(1) > 2) else p[sl].clone() for p in inp["string_params"]]
        sub["bow_params"] = [p[sl, :Nt].clone() if p.dim() > 1 else p[sl].clone() for p in inp["bow_params"]]
        sub["hammer_params"] = [p[sl, :Nt].clone() if p.dim() > 1 else p[sl].clone() for p in inp["hammer_params"]]
        sub["bow_mask"] = inp["bow_mask"][sl]; sub["hammer_mask"] = inp["hammer_mask"][sl]
        sub["Nt"] = Nt; sub["chunk_size"] = Nt
        o = gu.run_process(oracle.forward_fn, sub)
        for k in KEYS:
            ref[k].append(o[k])
    ref = {k: torch.cat(v, 0).numpy() for k, v in ref.items()}
    d = lambda t: t.cuda()
    sp, bp, hp = inp["string_params"], inp["bow_params"], inp["hammer_params"]
    su = inp["state_u"][:, :2].contiguous().cuda(); sz = inp["state_z"][:, :2].contiguous().cuda()
    uH = hp[2][:, :Nt].contiguous().cuda()
    res = step_strings(su, sz, kappa=d(sp[0]), alpha=d(sp[1]), f0=d(sp[5][:, :Nt]), pos=d(sp[6]), T60=d(sp[7]),
                       x_b=d(bp[0][:, :Nt]), v_b=d(bp[1][:, :Nt]), F_b=d(bp[2][:, :Nt]), wid=d(bp[5][:, :Nt]),
                       phi_0=d(bp[3]), phi_1=d(bp[4]), x_H=d(hp[0]), w_H=d(hp[3]), M_r=d(hp[4]), alpha_H=d(hp[5]),
                       u_H=uH, bow_mask=inp["bow_mask"], hammer_mask=inp["hammer_mask"],
                       k=inp["consts"][0], theta_t=inp["consts"][1], lambda_c=inp["consts"][2],
                       relative_order=inp["relative_order"], Nt=Nt, group_size=8, surface_integral=True,
                       save_state=False, counters=True)
    assert int(res["status"].abs().max()) == 0
    m = dict(uout="uout", zout="zout", v_r_out="v_r", F_H_out="F_H", u_H_out="u_H_out")
    for k in KEYS:
        err = gu.rel_l2(res[m[k]][:, 2:].cpu().numpy(), ref[k])
        print(k, f"{err:.2e}")
        assert err < (TOL_U if name != "pluck_b24" else 1e-5), (k, err)      # (pluck_b24: vs the ORACLE, which sits 1e-6 from the reference on its sensitive strings)
    cnt = res["counters"].cpu().numpy()
    assert (cnt[:, 3] == Nt - 2).all()


def test_single_precision_tensors_are_widened(monkeypatch):
    """SFDTD_WIDEN_F32=1: float32 tensors (the reference's `precision: single` preset) are computed in fp64 and returned
    as float32 (same dtype as the reference would), equal to the fp64 run up to float32 rounding; the in-place state /
    u_H updates land in the caller's float32 tensors.  (Default: the fp32 kernels, tests/test_gpu_fp32.py.)"""
    from torch_fdtd_string_b200 import forward_fn
    monkeypatch.setenv("SFDTD_WIDEN_F32", "1")
    g = gu.load_golden("random_b6")
    i64 = gu.build_inputs(g, device="cuda")
    i32 = gu.build_inputs(g, dtype=torch.float32, device="cuda")
    args = lambda i: (i["state_u"], i["state_z"], i["string_params"], i["bow_params"], i["hammer_params"], i["bow_mask"],
                      i["hammer_mask"], i["consts"], i["relative_order"], i["surface_integral"], False, 0, i["Nt"])
    o64 = forward_fn(*args(i64))
    o32 = forward_fn(*args(i32))
    assert o32[0].dtype == torch.float32 and o32[2].dtype == torch.float32
    assert o32[2].data_ptr() == i32["state_u"].data_ptr()
    for a, b in zip(o32[:2], o64[:2]):
        ref = b[:, 2:].float()
        assert float((a[:, 2:] - ref).norm() / ref.norm()) < 5e-5      # inputs themselves were rounded to float32
    assert float(i32["state_u"][:, 5].abs().max()) > 0                  # history written in place
