"""GPU: the fp32 kernels (SFDTD_F32) -- the reference's `precision: single`, its default preset
(src/configs/experiment/nsynth-like.yaml:16, src/task/simulate.py:131-135).

Stated bound (BASELINE.json north_star: "a stated bound for an fp32 mode"): on every fixture and output the fp32 kernels are
within 4 x the distance the reference's OWN float32 run of the same inputs keeps from its fp64 run (relative L2 over the
fixture; floor 2e-5), and per string within 1e-3 on uout wherever the reference's own float32 run is within 2.5e-4.  Float32 round-off is a random walk over the steps, so two float32
implementations land at comparable, not identical, distances: measured on B200 the kernels are between 17 x closer
(pluck_b1 uout 1.6e-5 vs the reference's 2.8e-4) and 2.7 x further (pluck_b3_pickup 1.7e-4 vs 5.5e-5) than the reference's
float32 run.  Those runs are committed as tests/golden/f32/<name>.npz (tests/golden/make_golden.py --single: the fixture's
inputs rounded to float32 and handed to the unmodified reference's process()).  Sample counts and shapes are bit-exact;
dtypes are float32 like the reference's."""
import os

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

F32_DIR = os.path.join(gu.GOLDEN_DIR, "f32")
NAMES = sorted(os.path.basename(p)[:-4] for p in (os.listdir(F32_DIR) if os.path.isdir(F32_DIR) else []) if p.endswith(".npz"))
MARGIN = 4.0
FLOOR = 2e-5


def run_cuda32(g):
    from torch_fdtd_string_b200 import process
    inp = gu.build_inputs(g, dtype=torch.float32, device="cuda")
    out = process("unused", inp["state_u"], inp["state_z"], inp["string_params"], inp["bow_params"],
                  inp["hammer_params"], inp["bow_mask"], inp["hammer_mask"], inp["consts"], inp["Nt"],
                  inp["chunk_size"], None, True, inp["relative_order"], inp["surface_integral"], inp["manufactured"])
    names = ["uout", "zout", "state_u", "state_z", "v_r_out", "F_H_out", "u_H_out", "sig0", "sig1"]
    return dict(zip(names, out)), inp


@pytest.mark.parametrize("name", NAMES)
def test_fp32_kernels_within_the_references_own_fp32_distance(name):
    g = gu.load_golden(name)
    r32 = dict(np.load(os.path.join(F32_DIR, f"{name}.npz")))
    out, inp = run_cuda32(g)
    line = []
    for k in ("uout", "zout", "v_r_out", "F_H_out", "u_H_out"):
        x = out[k]
        assert x.dtype == torch.float32 and tuple(x.shape) == g[k].shape, (k, x.dtype, x.shape)
        x = x.double().cpu().numpy()
        if not np.isfinite(g[k]).all() or np.linalg.norm(g[k]) == 0:
            continue
        mine = gu.rel_l2(x, g[k])                      # fp32 kernels vs the fp64 reference
        ref = gu.rel_l2(r32[k], g[k])                  # the reference's float32 run vs its fp64 run
        both = gu.rel_l2(x, r32[k].astype(np.float64))     # fp32 kernels vs the reference's float32 run
        line.append(f"{k} {mine:.1e} (ref32 {ref:.1e}; to ref32 {both:.1e})")
        # (the auxiliary outputs of strings that are not bowed / hammered amplify the state's round-off through the contact
        # nonlinearity: twice the margin)
        assert mine <= (MARGIN if k in ("uout", "zout") else 2 * MARGIN) * max(ref, FLOOR), (name, k, mine, ref)
        if k == "uout":
            # per string: <= 1e-3 wherever the reference's own float32 run stays within 2.5e-4 of its fp64 run (the string of
            # `pluck_b24` that amplifies 1-ulp perturbations in fp64 -- DESIGN.md "Sensitivity" -- is 5e-2 away after 4 ms in the
            # reference's float32 run and 4.5e-2 here: float32 round-off is 1e8 ulp of a double)
            for b in range(x.shape[0]):
                mb, rb = gu.rel_l2(x[b], g[k][b]), gu.rel_l2(r32[k][b], g[k][b])
                if rb > 2.5e-4 or mb > 2.5e-4:
                    print(f"  {name} uout[{b}]: kernels {mb:.1e}, reference float32 {rb:.1e}, kernels to reference float32 "
                          f"{gu.rel_l2(x[b], r32[k][b].astype(np.float64)):.1e}")
                if rb <= 2.5e-4:
                    assert mb <= 1e-3, (name, b, mb, rb)
    print(name, "; ".join(line))
    # in-place side effects land in the caller's float32 tensors (string.cpp:264-265, 303)
    assert out["state_u"].data_ptr() == inp["state_u"].data_ptr() and inp["state_u"].dtype == torch.float32
    assert gu.rel_l2(out["state_u"][:, -2:, :].double().cpu().numpy(), g["state_u_last"]) <= 2 * MARGIN * max(
        gu.rel_l2(r32["state_u_last"], g["state_u_last"]), FLOOR)


def test_fp32_workload_tracks_fp64():
    """Throughput-workload shape (compact API, in-kernel synthesis): fp32 vs fp64 kernels on the same 3552 strings."""
    from torch_fdtd_string_b200 import sampler
    B, Nt = 148 * 24, 482
    ph = sampler.sample_nsynth_like(B, length=1.0, excitation="pluck", seed=77)
    p = sampler.to_device(ph, torch.device("cuda"))
    r64 = sampler.run_compact(p, 24, counters=True, n_run=Nt)
    r32 = sampler.run_compact(p, 24, counters=True, n_run=Nt, precision="single")
    torch.cuda.synchronize()
    assert r32["uout"].dtype == torch.float32
    assert not (int(r32["status"].max()) & ~1)
    a, b = r64["uout"][:, 2:], r32["uout"][:, 2:].double()
    ok = torch.isfinite(a).all(dim=1) & torch.isfinite(b).all(dim=1) & (a.norm(dim=1) > 0)
    assert int(ok.sum()) > 0.95 * B
    err = (a[ok] - b[ok]).norm(dim=1) / a[ok].norm(dim=1)
    q = torch.quantile(err, torch.tensor([0.5, 0.9, 0.99], dtype=torch.float64, device=err.device))
    sw = float(r32["counters"][:, 1].sum()) / float(r32["counters"][:, 3].sum())
    print("fp32 vs fp64 (10 ms, uout): median %.1e q90 %.1e q99 %.1e max %.1e; fp32 sweeps/step %.2f" % (float(q[0]), float(q[1]), float(q[2]), float(err.max()), sw))
    # A string whose grid size changes inside the window (glissando / vibrato move f0 across a floor(1/h) cliff) switches a
    # few samples earlier or later in float32 than in fp64 -- exactly like the reference's own single-precision run -- and
    # is percents away afterwards: ~10 % of the strings within 10 ms (measured q90 1.8e-2); the median string is at 1e-4.
    assert float(q[0]) < 3e-4 and float(q[1]) < 6e-2
    assert int(r32["counters"][:, 3].min()) == Nt - 2


def test_fp32_hammer_bow_groups():
    from torch_fdtd_string_b200 import sampler
    for ex in ("hammer", "bow", "random"):
        ph = sampler.sample_nsynth_like(24 * 24, length=1.0, excitation=ex, seed=78)
        p = sampler.to_device(ph, torch.device("cuda"))
        Nt = 482
        r64 = sampler.run_compact(p, 24, counters=True, n_run=Nt)
        r32 = sampler.run_compact(p, 24, counters=True, n_run=Nt, precision="single")
        torch.cuda.synchronize()
        assert int((r32["status"] & (2 | 4 | 8 | 16)).max()) == 0, ex
        a, b = r64["uout"][:, 2:], r32["uout"][:, 2:].double()
        ok = torch.isfinite(a).all(dim=1) & torch.isfinite(b).all(dim=1) & (a.norm(dim=1) > 0)
        err = (a[ok] - b[ok]).norm(dim=1) / a[ok].norm(dim=1)
        q = torch.quantile(err, torch.tensor([0.5, 0.9], dtype=torch.float64, device=err.device))
        outer = float(r32["counters"][:, 0].sum()) / float(r32["counters"][:, 3].sum())
        print(ex, "fp32 vs fp64: median %.1e q90 %.1e; outer iterations/step %.2f" % (float(q[0]), float(q[1]), outer))
        assert float(q[0]) < 1e-3 and 1.5 < outer < 4.0, (ex, float(q[0]), outer)


def test_fp32_dataset_files(tmp_path):
    """`task.precision: single`: the dataset driver runs the fp32 kernels and writes what the reference writes in that mode --
    PCM_16 wavs (src/task/simulate.py:416-425) and float32 archives."""
    import wave
    from torch_fdtd_string_b200 import dataset
    st = dataset.generate(str(tmp_path), num_samples=24, batch_size=24, excitation="pluck", length=0.05, seed=5, precision="single")
    assert st["strings"] == 24 and st["written"] > 12
    d = tmp_path / sorted(os.listdir(tmp_path))[0]
    w = wave.open(str(d / "output-u.wav"))
    assert (w.getframerate(), w.getnframes(), w.getsampwidth()) == (48000, 2400 - 2, 2)
    v = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.int64)
    assert abs(np.abs(v).max() - 32767) <= 1                                   # l-infinity normalised
    sim = np.load(d / "simulation.npz")
    assert sim["uout"].dtype == np.float32 and sim["uout"].shape == (2398,) and np.isfinite(sim["uout"]).all()
