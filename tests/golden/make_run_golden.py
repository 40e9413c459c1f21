"""Generates tests/golden/run_*.npz: what the UNMODIFIED reference's ``run()`` (reference src/task/simulate.py:219-455)
writes for one small batch -- every array of simulation.npz / string_params.npz / hammer_params.npz / bow_params.npz, the
yaml and the three wav signals (soundfile is not installed: ``sf.write`` is intercepted and the float data recorded) --
so that the drop-in CLI's files can be compared value by value on the GPU box.  Runs only where /root/reference exists.

    python tests/golden/make_run_golden.py
"""
import os
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_driver  # noqa: E402

sim, ext = ref_driver.import_reference()
import torch  # noqa: E402
from torch_fdtd_string_b200 import hydra_lite as H  # noqa: E402

CASES = dict(
    run_nsynth_pluck_b3=["experiment=nsynth-like", "task.batch_size=3", "task.num_samples=3", "task.length=0.01",
                         "task.precision=double"],
    run_nsynth_random_b4=["experiment=nsynth-like", "task.batch_size=4", "task.num_samples=4", "task.length=0.01",
                          "task.precision=double", "model.excitation=random", "task.skip_silence=false"],
)


def run_case(name, overrides):
    wavs = {}
    sim.sf.write = lambda path, data, sr, subtype=None: wavs.__setitem__(os.path.relpath(path, save_dir), (np.asarray(data, dtype=np.float64), sr, subtype))
    tmp = tempfile.mkdtemp(prefix="sfdtd_run_")
    cfg, _ = H.compose("/root/reference/src/configs", "config.yaml", overrides + ["task.result_dir=ref", f"task.root_dir={tmp}",
                                                                                 "proc.cpu=true", "task.plot=false", "task.plot_state=false"])
    args = H.to_namespace(cfg)
    args.cwd = ref_driver.scratch_root()            # reference run.py:77 (root_dir of the JIT build, patched away below)
    save_dir = os.path.join(tmp, "ref")
    os.makedirs(save_dir, exist_ok=True)
    sim.process.__globals__["cpp_load"] = lambda **kw: ext
    torch.manual_seed(cfg["proc"]["seed"])
    model_name = cfg["model"]["excitation"] or "random"
    n = cfg["task"]["num_samples"] // cfg["task"]["batch_size"]
    with torch.no_grad():
        sim.run(args, save_dir, model_name, n)
    out = {"overrides": np.array(overrides), "seed": cfg["proc"]["seed"]}
    dirs = sorted(d for d in os.listdir(save_dir) if os.path.isdir(os.path.join(save_dir, d)))
    out["dirs"] = np.array(dirs)
    for d in dirs:
        for f in ("simulation", "string_params", "hammer_params", "bow_params"):
            z = np.load(os.path.join(save_dir, d, f + ".npz"))
            for k in z.files:
                out[f"{d}/{f}/{k}"] = z[k]
        out[f"{d}/yaml"] = np.array(open(os.path.join(save_dir, d, "simulation_config.yaml")).read())
        for w in ("output-u.wav", "output-z.wav", "output.wav"):
            data, sr, subtype = wavs[f"{d}/{w}"]
            out[f"{d}/{w}"] = data; out[f"{d}/{w}/subtype"] = np.array(subtype); out[f"{d}/{w}/sr"] = sr
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, dirs, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        run_case(nm, CASES[nm])
