"""B200-native drop-in for StringFDTD-Torch's batched time stepper.

Public surface (mirrors the reference for this path only):
  forward_fn(...)            -- reference src/model/cpp/simulator.cpp:14-27 (the pybind entry point)
  process(...)               -- reference src/task/simulate.py:16-119 (the chunk driver)
  step_strings(...)          -- native compact-input API (no (B,Nt,Nx) tensors; controls as curves or synthesised)
  Plan / build_args          -- plan once, then queue calls without host synchronisation
  postprocess(...)           -- device-side NaN / silence / gain / PCM quantisation of the audio
"""
import os as _os

# The stepper launches one kernel per string-size bucket, each on its own stream; with the default of 8 hardware
# queues those streams alias and serialise.  Only effective when set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .forward_fn import (forward_fn, step_strings, build_args, Plan, synth_controls, postprocess, deferred_checks,  # noqa: F401
                         make_xax, launch_count)
from .simulate import process  # noqa: F401

__all__ = ["forward_fn", "process", "step_strings", "build_args", "Plan", "synth_controls", "postprocess", "deferred_checks", "make_xax",
           "launch_count"]
