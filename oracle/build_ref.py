"""TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference time-stepper.

Compiles /root/reference/src/model/cpp/*.cpp (where they lie; nothing is copied
into this repository) into ``oracle/_ref/forward_fn.so`` -- the same pybind
module the reference JIT-builds at src/task/simulate.py:28-36 -- with one
force-included shim (oracle/oracle_shim.h) for ``torch::linalg::inv``.

``oracle/_ref/`` is git-ignored but travels to the GPU box with gpurun, so the
prebuilt .so is usable there (where /root/reference does not exist).

Usage:  python oracle/build_ref.py            (no-op when already built)
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CPP = "/root/reference/src/model/cpp"
OUT = os.path.join(HERE, "_ref")


def ref_so_path():
    return os.path.join(OUT, "forward_fn.so")


def build(verbose=False):
    """Build oracle/_ref/forward_fn.so if the reference sources are present."""
    if os.path.exists(ref_so_path()):
        return ref_so_path()
    if not os.path.isdir(REF_CPP):
        raise FileNotFoundError(
            f"{REF_CPP} not present and {ref_so_path()} not prebuilt")
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    load(
        name="forward_fn",
        sources=sorted(glob.glob(f"{REF_CPP}/*.cpp")),
        extra_cflags=["-O2", "-include", os.path.join(HERE, "oracle_shim.h")],
        build_directory=OUT,
        verbose=verbose,
    )
    return ref_so_path()


def load_ref():
    """Import the prebuilt reference module (pybind); torch must be importable."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    path = ref_so_path()
    if not os.path.exists(path):
        build()
    spec = importlib.util.spec_from_file_location("forward_fn", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
