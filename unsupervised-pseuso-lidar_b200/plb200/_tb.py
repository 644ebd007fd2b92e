"""Loader of the thin torch C++ binding (_plb200_torch.so, built by plb200/build.py::build_torch from
csrc/torch_binding.cpp).  `mod` is None when it has not been built; ops.fused_losses then goes through the ctypes
binding - the same C ABI and the same kernels, only more host time per call.  PLB200_TORCH_BINDING=0 forces that."""
import importlib.util
import os

import torch  # noqa: F401  (its shared libraries must be loaded first)

from . import _lib  # noqa: F401  (libplb200.so: fails loudly when missing)

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_plb200_torch.so")
mod = None
if os.environ.get("PLB200_TORCH_BINDING", "1") != "0" and os.path.isfile(PATH):
    _spec = importlib.util.spec_from_file_location("_plb200_torch", PATH)
    mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(mod)
