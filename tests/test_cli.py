"""Hydra-compatible CLI: config composition without Hydra, the mapping onto the sampler, and (GPU) an
end-to-end `python -m run`-style invocation writing the reference's result layout (reference run.py:54-112)."""
import datetime
import os
import subprocess
import sys

import pytest
import yaml

from torch_fdtd_string_b200 import hydra_lite as H
from torch_fdtd_string_b200 import run as cli

HERE = os.path.dirname(os.path.abspath(__file__))
MINI = os.path.join(HERE, "configs")
REF = "/root/reference/src/configs"
NOW = datetime.datetime(2026, 1, 2, 3, 4, 5)


def test_defaults_list_and_packages():
    cfg, hydra = H.compose(MINI, "config.yaml", ["experiment=tiny", "task.result_dir=out"], now=NOW)
    assert cfg["task"]["_name_"] == "simulate" and cfg["model"] == {"_name_": "fdtd", "excitation": "pluck"}
    assert cfg["task"]["batch_size"] == 2 and cfg["task"]["sr"] == 48000            # experiment wins over task/simulate
    assert cfg["proc"]["num_workers"] == 4 and cfg["proc"]["seed"] == 1234            # experiment wins over config.yaml (_self_ first)
    assert cfg["task"]["string_condition"][0] == {"f0_min": 98.0}                     # lists are replaced, not merged
    assert cfg["callbacks"]["timer"] == {"step": True}                                 # group file lands in its group package
    assert hydra == {"run": {"dir": "./results/out"}}
    assert cfg["proc"]["port"] == "0405"


def test_default_experiment_and_interpolation_errors():
    # experiment=base has no task group: the result_dir interpolation cannot resolve, as in the reference
    with pytest.raises(H.ConfigError, match="task._name_"):
        H.compose(MINI, "config.yaml", [], now=NOW)
    cfg, hydra = H.compose(MINI, "config.yaml", ["experiment=tiny"], now=NOW)
    assert cfg["task"]["result_dir"] == "simulate-fdtd-20260102-030405"
    assert hydra["run"]["dir"] == "./results/simulate-fdtd-20260102-030405"


def test_overrides():
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=tiny", "task.result_dir=r", "proc.gpus=[1,2]", "model.excitation=hammer",
                                             "task.length=0.5", "+task.extra.deep=3", "~task.plot", "task.theta_t=null"], now=NOW)
    assert cfg["proc"]["gpus"] == [1, 2] and cfg["model"]["excitation"] == "hammer" and cfg["task"]["length"] == 0.5
    assert cfg["task"]["extra"] == {"deep": 3} and "plot" not in cfg["task"] and cfg["task"]["theta_t"] is None
    with pytest.raises(H.ConfigError, match="not in the config"):
        H.compose(MINI, "config.yaml", ["experiment=tiny", "task.no_such_key=1"], now=NOW)
    with pytest.raises(H.ConfigError, match="not found"):
        H.compose(MINI, "config.yaml", ["experiment=nope"], now=NOW)
    with pytest.raises(H.ConfigError, match="no match in the defaults list"):
        H.compose(MINI, "config.yaml", ["experiment=base", "task=simulate"], now=NOW)      # base has no /task entry


def test_defaults_keywords():
    """optional / override / group@package entries and a nested group option (model/excitation)"""
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=keywords", "task.result_dir=r"], now=NOW)
    assert cfg["model"]["excitation"] == {"kind": "strike", "force": 2.5} and cfg["model"]["_name_"] == "fdtd"
    assert cfg["task"]["more"] == {"x": 1, "y": 48000}                 # key-level package, interpolation through the root
    assert "extra" not in cfg                                            # the optional entry does not exist: skipped
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=keywords", "task.result_dir=r", "model/excitation=null"], now=NOW)
    assert "excitation" not in cfg["model"] or cfg["model"]["excitation"] is None


def test_same_group_include_and_filter():
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=tiny", "task.result_dir=r", "model=pluck"], now=NOW)
    assert cfg["model"]["_name_"] == "fdtd" and cfg["model"]["excitation"] == "pluck"
    assert "__scratch" in cfg["callbacks"]
    assert "__scratch" not in H.filter_keys(cfg, lambda k: not k.startswith("__"))["callbacks"]
    ns = H.to_namespace(cfg)
    assert ns.task.batch_size == 2 and ns["task"]["batch_size"] == 2 and ns.task.string_condition[1] == {"f0_max": 440.0}


def test_plan_and_sampler_mapping():
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=tiny", "task.result_dir=my"], now=NOW)
    p = cli.plan(cfg, "/work")
    assert p["save_dir"] == "/work/./results/my" and p["model_name"] == "pluck" and p["n_batches"] == 2
    c = cli.sampler_config(cfg["task"])
    assert (c["f0_min"], c["f0_max"], c["kappa_max"], c["alpha_max"], c["p_a_max"], c["p_x_max"]) == (98.0, 440.0, 0.03, 25.0, 0.02, 0.5)
    assert c["p_a_min"] == 0.001 and c["pos_min"] == 0.3 and c["alpha_H"] == 3.0 and c["M_r_max"] == 10.0      # module defaults kept
    assert c["sampling_T60"] == "random" and c["relative_order"] == 4 and c["theta_t"] is None
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=fixed", "task.result_dir=debug", "model.excitation=null"], now=NOW)
    p = cli.plan(cfg, "/work")
    assert p["save_dir"].endswith("/debug") and cfg["proc"]["debug"] is True and p["model_name"] == "random"
    c = cli.sampler_config(cfg["task"])
    assert c["f0_min"] == c["f0_max"] == 55.0 and c["f0_mod_max"] == 0.0 and c["kappa_min"] == c["kappa_max"] == 0.08
    assert c["alpha_min"] == c["alpha_max"] == 20.0 and c["pos_min"] == c["pos_max"] == 0.5 and c["p_x_min"] == c["p_x_max"] == 0.2
    assert c["sampling_T60"] == "fix" and c["t60_fixed"] == 20.0 and c["relative_order"] == 8
    cfg["task"]["sampling_f0"] = "equidist"
    with pytest.raises(NotImplementedError):
        cli.sampler_config(cfg["task"])


def test_fixed_sampler_draw():
    from torch_fdtd_string_b200 import sampler
    cfg, _ = H.compose(MINI, "config.yaml", ["experiment=fixed", "task.result_dir=r"], now=NOW)
    q = sampler.sample_nsynth_like(3, sr=48000, length=0.02, excitation="pluck", seed=1, cfg=cli.sampler_config(cfg["task"]))
    assert (q["kappa"] == 0.08).all() and (q["alpha"] == 20.0).all() and (q["pos"] == 0.5).all()
    assert (q["T60"][:, 0, 0] == 1000).all() and (q["T60"][:, 1, 0] == 100).all() and (q["T60"][:, :, 1] == 20.0).all()
    assert (q["f0_a"] == q["f0_b"]).all() and (q["mod_amp"] == 0).all() and (q["p_a"] == 0.02).all()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_config_tree():
    """the reference's own configs, composed the way `python -m run experiment=nsynth-like ...` (README.md:43) composes them"""
    cfg, hydra = H.compose(REF, "config.yaml", ["experiment=nsynth-like", "task.num_samples=100", "task.result_dir=my_fdtd_simulation"], now=NOW)
    t = cfg["task"]
    assert (t["num_samples"], t["batch_size"], t["precision"], t["length"], t["sr"]) == (100, 24, "single", 1.0, 48000)
    assert cfg["model"] == {"_name_": "fdtd", "excitation": "pluck"} and cfg["proc"]["num_workers"] == 4
    assert hydra["run"]["dir"] == "./results/my_fdtd_simulation"
    c = cli.sampler_config(t)
    from torch_fdtd_string_b200.sampler import NSYNTH
    for k, v in NSYNTH.items():      # the compact sampler's built-in nsynth-like table is exactly what the reference's preset composes to
        if k in c and k not in ("theta_t",):
            assert c[k] == v, (k, c[k], v)
    p = cli.plan(cfg, "/w")
    assert p["n_batches"] == 4 and p["model_name"] == "pluck"
    cfg, _ = H.compose(REF, "config.yaml", ["experiment=all-fixed", "model.excitation=bow", "task.length=4.0", "task.result_dir=x"], now=NOW)
    c = cli.sampler_config(cfg["task"])
    assert c["f0_min"] == 55.0 and c["x_b_min"] == c["x_b_max"] == 0.2 and c["F_b_min"] == 90 and c["wid_max"] == 4 and c["phi_0_min"] == 9.0
    assert cfg["proc"]["cpu"] is True and cfg["task"]["relative_order"] == 8


def test_cli_refuses_cpu(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    r = subprocess.run([sys.executable, "-m", "torch_fdtd_string_b200.run", "--config-dir", MINI, "experiment=tiny",
                        "task.result_dir=t", f"task.root_dir={tmp_path}", "proc.cpu=true"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr
    assert os.path.isfile(tmp_path / "t" / ".hydra" / "config.yaml") and os.path.isfile(tmp_path / "t" / "config_tree.txt")
    assert yaml.safe_load(open(tmp_path / "t" / ".hydra" / "overrides.yaml"))[0] == "experiment=tiny"


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    r = subprocess.run([sys.executable, "-m", "torch_fdtd_string_b200.run", "--config-dir", MINI, "experiment=tiny",
                        "task.result_dir=gen", f"task.root_dir={tmp_path}"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    root = tmp_path / "gen"
    assert os.path.isfile(root / "config_tree.txt") and os.path.isdir(root / "codes" / "torch_fdtd_string_b200")
    assert os.path.isfile(root / "gpu_time.txt")
    dirs = sorted(d for d in os.listdir(root) if d[0].isdigit())
    assert dirs == ["0-0", "0-1", "1-0", "1-1"], dirs
    for d in dirs:
        for f in ("output.wav", "output-u.wav", "output-z.wav", "simulation.npz", "string_params.npz", "hammer_params.npz",
                  "bow_params.npz", "simulation_config.yaml"):
            assert os.path.isfile(root / d / f), (d, f)
