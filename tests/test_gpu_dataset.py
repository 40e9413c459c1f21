"""GPU: value test of the dataset driver against what the UNMODIFIED reference's ``run()`` wrote for the same
seed (fixture tests/golden/run_nsynth_pluck_b3.npz, made by tests/golden/make_run_golden.py: experiment=nsynth-like,
batch_size 3, 10 ms, double): every array of simulation.npz (incl. the trimmed state histories, src/task/simulate.py:405-408),
string_params.npz, hammer_params.npz, bow_params.npz, the yaml summary, and the three wav files (the reference's float
signals quantised to PCM_24 vs the bytes this package's device post-processing produced)."""
import os
import sys
import wave

import numpy as np
import pytest
import torch
import yaml

import golden_util as gu

sys.path.insert(0, gu.GOLDEN_DIR)
import presets  # noqa: E402

pytestmark = pytest.mark.gpu


def read_wav24(path):
    w = wave.open(str(path))
    assert w.getsampwidth() == 3 and w.getnchannels() == 1
    raw = np.frombuffer(w.readframes(w.getnframes()), dtype=np.uint8).reshape(-1, 3).astype(np.int64)
    v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
    return np.where(v >= 1 << 23, v - (1 << 24), v), w.getframerate()


def test_files_match_the_reference_run(tmp_path):
    from torch_fdtd_string_b200 import dataset, sampler
    g = dict(np.load(os.path.join(gu.GOLDEN_DIR, "run_nsynth_pluck_b3.npz"), allow_pickle=False))
    p = presets.PRESETS["nsynth"]
    theta_t = sampler.get_theta(0.03, 98.0, p["sr"])
    torch.manual_seed(int(g["seed"]))                                         # reference run.py:75
    src = dataset.reference_source(3, p["sr"], 0.01, "pluck", theta_t, p["f0_inf"], p["alpha_inf"], p["lambda_c"], "double",
                                   p["string_kwargs"], p["bow_kwargs"], p["hammer_kwargs"], False, p["relative_order"])
    st = dataset.generate(str(tmp_path), 3, 3, "pluck", p["sr"], 0.01, precision="double", normalize_output=True,
                          skip_silence=True, silence_threshold=-23.0, source=src, full_layout=True, surface_integral=True)
    ref_dirs = [str(d) for d in g["dirs"]]
    assert st["written"] == len(ref_dirs)
    for rd in ref_dirs:
        b = rd.rsplit("-", 1)[1]
        d = tmp_path / f"0-{b}"
        assert sorted(os.listdir(d)) == ["bow_params.npz", "hammer_params.npz", "output-u.wav", "output-z.wav", "output.wav",
                                         "simulation.npz", "simulation_config.yaml", "string_params.npz"]
        for arch in ("simulation", "string_params", "hammer_params", "bow_params"):
            z = np.load(d / f"{arch}.npz")
            want = {k.split("/")[2]: v for k, v in g.items() if k.startswith(f"{rd}/{arch}/")}
            assert sorted(z.files) == sorted(want), (arch, sorted(z.files), sorted(want))
            for k, r in want.items():
                x = z[k]
                assert x.shape == r.shape, (arch, k, x.shape, r.shape)
                if r.dtype == bool:
                    assert np.array_equal(x, r), (arch, k)
                elif k in ("uout", "zout", "v_r_out", "F_H_out", "u_H_out", "state_u", "state_z", "u_H"):     # (u_H: updated in place)
                    assert gu.rel_l2(x, r) < 3e-8, (arch, k, gu.rel_l2(x, r))
                elif k in ("f0", "target_f0", "x_B", "v_B", "F_B", "v_H", "sig0", "sig1"):
                    np.testing.assert_allclose(x, r, rtol=1e-12, atol=1e-300, err_msg=f"{arch}/{k}")
                else:
                    assert np.array_equal(x, r), (arch, k, x, r)
        y, yr = yaml.safe_load(open(d / "simulation_config.yaml")), yaml.safe_load(str(g[f"{rd}/yaml"]))
        assert y["excitation_type"] == yr["excitation_type"] and sorted(y) == sorted(yr)
        for sec in ("value-string", "value-hammer", "value-bow"):
            assert sorted(y[sec]) == sorted(yr[sec])
            for k in y[sec]:
                assert y[sec][k] == pytest.approx(yr[sec][k], rel=1e-12, abs=1e-300), (sec, k)
        for wname in ("output-u.wav", "output-z.wav", "output.wav"):
            v, sr = read_wav24(d / wname)
            ref = g[f"{rd}/{wname}"]
            assert str(g[f"{rd}/{wname}/subtype"]) == "PCM_24" and sr == int(g[f"{rd}/{wname}/sr"]) and v.shape == ref.shape
            q = np.clip(np.rint(ref * 8388608.0), -8388608, 8388607).astype(np.int64)
            assert np.abs(v - q).max() <= 1, (wname, int(np.abs(v - q).max()))          # a 1e-12 difference can flip one rounding
            assert (v != q).mean() < 0.01


def test_load_config_overrides_reach_the_stepper_and_the_files(tmp_path):
    """task.load_config (reference README 1.3, src/task/simulate.py:164-185): predefined f0 / hammer-velocity curves as npy
    files -> the stepper's table mode; the saved f0 is the loaded curve pre-corrected by each string's Fletcher factor, the
    saved hammer displacement starts from the loaded strike profile, and the run differs from the un-conditioned one."""
    from torch_fdtd_string_b200 import dataset, sampler
    p = presets.PRESETS["nsynth"]
    sr, length, B = p["sr"], 0.02, 6
    Nt = int(sr * length)
    theta_t = sampler.get_theta(0.03, 98.0, sr)
    pre = tmp_path / "preset"; pre.mkdir()
    f0 = 220.0 + 10.0 * np.sin(np.linspace(0, 3, Nt // 2))                  # shorter than the run: edge-padded like the reference
    prof = np.zeros(Nt); prof[5:8] = 1.0
    np.save(pre / "string-f0.npy", f0); np.save(pre / "hammer-v_H.npy", prof)
    over = dataset.load_overrides(str(pre), Nt)
    assert sorted(over) == ["hammer-v_H", "string-f0"] and over["string-f0"].shape == (Nt,) and over["string-f0"][-1] == f0[-1]
    outs = {}
    for tag, ov in (("plain", None), ("cond", over)):
        torch.manual_seed(5)
        src = dataset.reference_source(B, sr, length, "hammer", theta_t, p["f0_inf"], p["alpha_inf"], p["lambda_c"], "double",
                                       p["string_kwargs"], p["bow_kwargs"], p["hammer_kwargs"], False, p["relative_order"],
                                       redraw_v_H=bool(ov))
        d = tmp_path / tag
        st = dataset.generate(str(d), B, B, "hammer", sr, length, precision="double", source=src, full_layout=True, overrides=ov,
                              skip_silence=False)
        assert st["strings"] == B and st["written"] >= B - 1
        outs[tag] = d
    d = outs["cond"] / sorted(os.listdir(outs["cond"]))[0]
    sp = np.load(d / "string_params.npz"); hp = np.load(d / "hammer_params.npz"); sim = np.load(d / "simulation.npz")
    w0 = float(sampler.fletcher_w0(torch.tensor([float(sp["kappa"])], dtype=torch.float64))[0])
    np.testing.assert_allclose(sp["f0"], over["string-f0"].astype(np.float32).astype(np.float64) / w0, rtol=1e-12)
    assert np.isfinite(sim["uout"]).all() and np.abs(sim["uout"]).max() > 0
    # the hammer is thrown at samples 5..7 instead of sample 1: nothing touches the string before
    assert np.abs(sim["F_H_out"][:3]).max() == 0.0
    d0 = outs["plain"] / sorted(os.listdir(outs["plain"]))[0]
    assert gu.rel_l2(np.load(d0 / "simulation.npz")["uout"], sim["uout"]) > 1e-3
