"""``python -m run experiment=nsynth-like task.num_samples=100 task.result_dir=my_fdtd_simulation`` without Hydra.

Drop-in for the reference's entry point (reference run.py:54-148) for the simulation task: the reference's own
``src/configs`` tree is composed by ``hydra_lite`` (same defaults lists, same overrides syntax), the run directory is
``${task.root_dir}/${task.result_dir}`` (``hydra.run.dir``) with ``.hydra/{config,overrides}.yaml``, ``config_tree.txt``
and the ``codes/`` backup written into it, and the strings land in ``{task.root_dir}/{result_dir}/{id}-{b}/`` in the
reference's result-file layout (``dataset.py``).

    python -m torch_fdtd_string_b200.run --config-dir /path/to/StringFDTD-Torch/src/configs \\
        experiment=nsynth-like task.num_samples=100 task.result_dir=my_fdtd_simulation

``--config-dir`` defaults to ``$SFDTD_CONFIG_DIR``, then ``./src/configs`` (running from a checkout of the reference).
Under ``torchrun`` every rank takes its share of the batches (``proc.gpus`` is ignored then); otherwise the first entry of
``proc.gpus`` selects the device.  What the reference's CLI can do and this one refuses, loudly: ``proc.cpu=true`` (there is
no CPU path), ``proc.{evaluate,summarize,process_training_data,train,test}`` (not part of the time-stepping path),
``task.plot`` / ``plot_state`` (ignored with a notice).  ``task.load_config`` (predefined f0 / bow / hammer curves as npy files)
is honoured through the stepper's table mode (``dataset.load_overrides`` / ``apply_overrides``).

Parameters are drawn by ``sampler_ref`` -- the reference's String / Bow / Hammer draws restated RNG-stream compatibly, so
``proc.seed`` reproduces the reference's dataset (all sampling modes, pluck profiles, the manufactured initial condition).
``SFDTD_SAMPLER=native`` selects the per-batch-seeded compact sampler instead; ``SFDTD_COMPACT_RESULTS=1`` leaves the state
histories and the (Nt, Nx) initial arrays out of the archives (the reference always stores them: ~20 MB per string-second).
"""
import os
import shutil
import sys

import yaml

from . import hydra_lite as H

# keyword defaults of the reference's String / Hammer / Bow modules (reference src/model/simulator.py:123-135,419-425,531-537)
STRING_DEFAULTS = dict(
    f0_min=27.50, f0_max=440.0, f0_diff_max=50.0, f0_mod_max=0.02, f0_fixed=20.0, kappa_min=0.0, kappa_max=0.08,
    kappa_fixed=0.08, alpha_min=1.0, alpha_max=25.0, alpha_fixed=3.0, pos_min=0.3, pos_max=0.7, pos_fixed=0.5, lossless=False,
    t60_min_1=20.0, t60_max_1=30.0, t60_min_2=30.0, t60_max_2=30.0, t60_fixed=20.0, t60_diff_max=5.0,
    sampling_p_a="random", sampling_p_x="random", p_a_min=0.001, p_a_max=0.01, p_a_fixed=0.01, p_x_min=0.1, p_x_max=0.9,
    p_x_fixed=0.5, pluck_profile=None)
HAMMER_DEFAULTS = dict(x_H_min=0.1, x_H_max=0.9, v_H_min=0.5, v_H_max=5.0, M_r_min=1.0, M_r_max=10.0, w_H_min=1000.0,
                       w_H_max=3000.0, alpha_fixed=3.0)
BOW_DEFAULTS = dict(x_b_min=0.2, x_b_max=0.5, x_b_maxdiff=0.2, v_b_min=0.3, v_b_max=0.4, F_b_min=80.0, F_b_max=100.0,
                    F_b_maxdiff=10.0, phi_0_min=2.0, phi_0_max=6.0, phi_1_min=0.0, phi_1_max=0.5, wid_min=3.0, wid_max=6.0,
                    do_pulloff=None)


def _conditions(lst, defaults, what):
    """list of one-key dicts (the reference's ``*_condition`` format, src/task/simulate.py:243-266) -> kwargs"""
    out = dict(defaults)
    for item in lst or []:
        (key, val), = item.items()
        if val is None:
            continue
        if key not in defaults:
            raise H.ConfigError(f"task.{what}: unknown key '{key}'")
        out[key] = val
    return out


def sampler_config(task):
    """task node of the composed config -> the compact sampler's ``cfg`` (torch_fdtd_string_b200/sampler.py)"""
    s = _conditions(task.get("string_condition"), STRING_DEFAULTS, "string_condition")
    s = _conditions(task.get("pluck_condition"), s, "pluck_condition")
    h = _conditions(task.get("hammer_condition"), HAMMER_DEFAULTS, "hammer_condition")
    b = _conditions(task.get("bow_condition"), BOW_DEFAULTS, "bow_condition")
    mode = {k: ("random" if task.get(k) is None else task.get(k))
            for k in ("sampling_f0", "sampling_kappa", "sampling_alpha", "sampling_pickup", "sampling_T60")}
    mode["sampling_p_a"], mode["sampling_p_x"] = s["sampling_p_a"], s["sampling_p_x"]
    for k, v in mode.items():
        if v not in ("random", "fix"):
            raise NotImplementedError(f"task.{k}={v!r}: only 'random' and 'fix' are built")
    if s["pluck_profile"] not in (None, "triangular"):
        raise NotImplementedError(f"pluck_profile={s['pluck_profile']!r}: only the triangular pluck is built")
    if task.get("precorrect") is False:
        raise NotImplementedError("task.precorrect=false is not built (f0 is always pre-corrected for stiffness)")
    if b["do_pulloff"] not in (None,):
        raise NotImplementedError("bow_condition.do_pulloff is not built (pull-off is drawn with probability 1/2)")
    c = {k: v for k, v in s.items() if k.endswith(("_min", "_max", "_min_1", "_max_1", "_min_2", "_max_2", "diff_max", "mod_max"))}

    def fix(name, fixed):
        c[f"{name}_min"] = c[f"{name}_max"] = float(fixed)

    if mode["sampling_f0"] == "fix":
        fix("f0", s["f0_fixed"]); c["f0_diff_max"] = 0.0; c["f0_mod_max"] = 0.0
    if mode["sampling_kappa"] == "fix":
        fix("kappa", s["kappa_fixed"])
    if mode["sampling_alpha"] == "fix":
        fix("alpha", max(float(s["alpha_fixed"]), float(task["alpha_inf"])))        # simulator.py:304
    if mode["sampling_pickup"] == "fix":
        fix("pos", s["pos_fixed"])
    if mode["sampling_p_a"] == "fix":
        fix("p_a", s["p_a_fixed"])
    if mode["sampling_p_x"] == "fix":
        fix("p_x", s["p_x_fixed"])
    c.update(sampling_T60=mode["sampling_T60"], lossless=bool(s["lossless"]), t60_fixed=float(s["t60_fixed"]))
    c.update({k: float(v) for k, v in h.items() if k != "alpha_fixed"}); c["alpha_H"] = float(h["alpha_fixed"])
    c.update({k: float(v) for k, v in b.items() if k != "do_pulloff"})
    c.update(f0_inf=float(task["f0_inf"]), alpha_inf=float(task["alpha_inf"]), lambda_c=float(task["lambda_c"]),
             relative_order=task["relative_order"], theta_t=task.get("theta_t"))
    return c


def reference_kwargs(task):
    """task node -> (string_kwargs, hammer_kwargs, bow_kwargs, theta_t) exactly as reference run() builds them
    (src/task/simulate.py:224-267)"""
    def first(lst, key):
        for item in lst or []:
            if key in item:
                return item[key]
        raise H.ConfigError(f"Specify '{key}' for task.string_condition")
    sc = task.get("string_condition")
    kappa_max = first(sc, "kappa_fixed") if task.get("sampling_kappa") == "fix" else first(sc, "kappa_max")
    f0_min = first(sc, "f0_fixed") if task.get("sampling_f0") == "fix" else first(sc, "f0_min")
    from . import sampler
    theta_t = sampler.get_theta(kappa_max, f0_min, task["sr"]) if task.get("theta_t") is None else task["theta_t"]
    sk = {k: ("random" if task.get(k) is None else task.get(k))
          for k in ("sampling_f0", "sampling_kappa", "sampling_alpha", "sampling_pickup", "sampling_T60", "precorrect")}

    def upd(dst, lst):
        for item in lst or []:
            (key, val), = item.items()
            if val is not None:
                dst[key] = val
        return dst
    upd(sk, sc); upd(sk, task.get("pluck_condition"))
    return sk, upd({}, task.get("hammer_condition")), upd({}, task.get("bow_condition")), theta_t


def print_config(cfg, path="config_tree.txt"):
    """reference src/utils/config.py:166-196: a rich tree of the resolved config on stdout and in config_tree.txt"""
    try:
        import rich
        import rich.syntax
        import rich.tree
        tree = rich.tree.Tree("CONFIG", style="dim", guide_style="dim")
        for field, section in cfg.items():
            branch = tree.add(field, style="dim", guide_style="dim")
            text = yaml.dump(section, sort_keys=False) if isinstance(section, dict) else str(section)
            branch.add(rich.syntax.Syntax(text, "yaml"))
        rich.print(tree)
        with open(path, "w") as fp:
            rich.print(tree, file=fp)
    except ImportError:
        text = yaml.dump(cfg, sort_keys=False)
        print(text)
        with open(path, "w") as fp:
            fp.write(text)


def backup_code(dst="codes"):
    """reference run.py:31-52 copies the working tree's sources next to the results; here: this package + include/"""
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for src, name in ((here, os.path.basename(here)), (os.path.join(root, "include"), "include")):
        if not os.path.isdir(src):
            continue
        for dirpath, dirnames, filenames in os.walk(src):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            rel = os.path.relpath(dirpath, src)
            out = os.path.normpath(os.path.join(dst, name, rel))
            os.makedirs(out, exist_ok=True)
            for f in filenames:
                if os.path.splitext(f)[1] in (".so", ".pyc", ".o", ".npz", ".pt", ".png", ".jpg"):
                    continue
                shutil.copyfile(os.path.join(dirpath, f), os.path.join(out, f))


def plan(cfg, launch_cwd):
    """Everything reference run.py:77-112 decides before ``simulate.run``: save directory, flags, model name, batch count."""
    task, proc = cfg["task"], cfg["proc"]
    if task.get("save_name") is not None:
        save_dir_name = task["save_name"]
    elif proc.get("debug") or task.get("result_dir") == "debug":
        proc["debug"] = True
        save_dir_name = "debug"
    else:
        save_dir_name = task["result_dir"]
    root_dir = task["root_dir"]
    if not os.path.isabs(root_dir):
        root_dir = os.path.join(launch_cwd, root_dir)
    if task.get("measure_time"):
        task["plot"] = False; task["save"] = False; task["plot_state"] = False
    excitation = cfg.get("model", {}).get("excitation")
    model_name = "random" if excitation is None else excitation
    for key in ("load_config",):
        if task.get(key) == H.MISSING:
            raise H.ConfigError(f"task.{key} is missing (???): choose an experiment that sets it")
    return dict(save_dir=f"{root_dir}/{save_dir_name}", model_name=model_name,
                n_batches=int(task["num_samples"]) // int(task["batch_size"]), root_dir=root_dir)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    config_dir = os.environ.get("SFDTD_CONFIG_DIR", os.path.join(os.getcwd(), "src", "configs"))
    config_name = "config.yaml"
    overrides = []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("--config-dir", "--config-path", "-cd", "-cp"):
            config_dir = argv[i + 1]; i += 2; continue
        if a in ("--config-name", "-cn"):
            config_name = argv[i + 1]; i += 2; continue
        if a in ("-h", "--help"):
            print(__doc__); return 0
        overrides.append(a); i += 1
    launch_cwd = os.getcwd()
    cfg, hydra = H.compose(config_dir, config_name, overrides)
    cfg = H.filter_keys(cfg, lambda k: not str(k).startswith("__"))            # src/utils/config.py:140
    proc, task = cfg["proc"], cfg["task"]

    # hydra.run.dir: created, entered (Hydra's job.chdir), .hydra/ bookkeeping
    run_dir = (hydra.get("run") or {}).get("dir") or "."
    run_dir = run_dir if os.path.isabs(run_dir) else os.path.join(launch_cwd, run_dir)
    os.makedirs(os.path.join(run_dir, ".hydra"), exist_ok=True)
    os.chdir(run_dir)
    with open(".hydra/config.yaml", "w") as f:
        yaml.dump(cfg, f, sort_keys=False)
    with open(".hydra/overrides.yaml", "w") as f:
        yaml.dump(overrides, f)
    print_config(cfg)

    for key in ("evaluate", "summarize", "process_training_data", "train", "test"):
        if proc.get(key):
            raise NotImplementedError(f"proc.{key}=true is outside the time-stepping path this package replaces; "
                                      f"run the reference's own `python -m run` for it")
    if proc.get("cpu"):
        raise RuntimeError("proc.cpu=true: this package has no CPU path (the stepper is CUDA-only); use the reference for CPU runs")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if "LOCAL_RANK" in os.environ:
        device_index = int(os.environ["LOCAL_RANK"])
    else:
        gpus = proc.get("gpus") or [0]
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", ",".join(str(g) for g in gpus))      # reference run.py:63-64
        device_index = 0
    import torch
    torch.manual_seed(proc["seed"])                                                          # reference run.py:75
    p = plan(cfg, launch_cwd)
    if proc.get("simulate"):
        if rank == 0:
            backup_code()
        for key in ("plot", "plot_state"):
            if task.get(key):
                print(f"[run] task.{key}=true ignored: plotting is outside the time-stepping path")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: this package has no CPU path")
        torch.cuda.set_device(device_index)
        from . import dataset
        # parameter draws: the reference's own, RNG-stream compatible (proc.seed gives the reference's dataset); set
        # SFDTD_SAMPLER=native for the per-batch-seeded compact sampler
        source = None
        overrides = None
        if task.get("load_config") is not None:
            # predefined conditions (reference README 1.3, src/task/simulate.py:164-185): {model}-{param}.npy curves
            overrides = dataset.load_overrides(str(task["load_config"]), int(float(task["length"]) * int(task["sr"])))
            print(f"[run] task.load_config: {sorted(overrides)}")
        if os.environ.get("SFDTD_SAMPLER", "reference") == "reference":
            sk, hk, bk, theta_t = reference_kwargs(task)
            source = dataset.reference_source(
                int(task["batch_size"]), int(task["sr"]), float(task["length"]), p["model_name"], theta_t, task["f0_inf"],
                task["alpha_inf"], task["lambda_c"], task["precision"], sk, bk, hk, bool(task.get("manufactured")),
                task["relative_order"], redraw_v_H=bool(overrides and "hammer-v_H" in overrides))
        stats = dataset.generate(
            p["save_dir"], int(task["num_samples"]), int(task["batch_size"]), p["model_name"], int(task["sr"]),
            float(task["length"]), int(proc["seed"]), task["precision"], bool(task["normalize_output"]),
            bool(task["skip_silence"]), float(task["silence_threshold"]), save=bool(task.get("save", True)),
            randomize_name=bool(task["randomize_name"]), rank=rank, world_size=world,
            surface_integral=bool(task["surface_integral"]), sampler_cfg=(sampler_config(task) if source is None else None),
            time_log=True, num_workers=int(proc.get("num_workers") or 1), source=source,
            full_layout=os.environ.get("SFDTD_COMPACT_RESULTS", "0") != "1", manufactured=bool(task.get("manufactured")),
            overrides=overrides)
        print(f"[run] rank {rank}/{world}: {stats}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
