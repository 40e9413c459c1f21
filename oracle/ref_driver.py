"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference in THIS container.

Imports the reference's Python samplers and its ``src.task.simulate`` from
/root/reference (never copied), with the optional plotting/audio deps stubbed,
and points its ``cpp_load`` at the prebuilt ``oracle/_ref/forward_fn.so``.
Only usable where /root/reference exists (not on the GPU box); used by
tests/golden/make_golden.py to create the committed fixtures.
"""
import os
import sys
import types
import tempfile

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (simulate_module, forward_fn_module)."""
    if not os.path.isdir(REF_ROOT):
        raise FileNotFoundError(REF_ROOT)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "-1")
    sys.path.insert(0, HERE)
    import build_ref
    ext = build_ref.load_ref()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # deps of src.task.simulate / src.utils.{plot,audio} that are not installed
    for name in ["soundfile", "librosa", "librosa.display", "librosa.filters",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.animation",
                 "matplotlib.cm", "matplotlib.colors", "matplotlib.gridspec",
                 "matplotlib.ticker", "matplotlib.patches", "matplotlib.lines",
                 "crepe", "wandb", "seaborn", "torchaudio", "torchaudio.functional",
                 "torchaudio.transforms"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    import importlib
    try:
        sim = importlib.import_module("src.task.simulate")
    except Exception:
        # fall back: stub the plotting/audio helper modules wholesale
        _stub("src.utils.plot")
        import torch

        def dB_RMS(x):
            return 10 * torch.log10(x.pow(2).mean(-1) + 1e-30)
        _stub("src.utils.audio", dB_RMS=dB_RMS)
        sim = importlib.import_module("src.task.simulate")
    sim.cpp_load = lambda **kw: ext
    return sim, ext


def scratch_root():
    d = tempfile.mkdtemp(prefix="sfdtd_ref_")
    os.makedirs(os.path.join(d, "src/model/cpp"), exist_ok=True)
    return d
