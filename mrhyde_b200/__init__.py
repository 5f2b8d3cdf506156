"""mrhyde_b200: B200-native replacement of MrHyDE's element residual/Jacobian assembly path.

The product is the C-ABI shared library (include/mrhyde_b200.h, mrhyde_b200/csrc); this package is
the thin Python plumbing around it (ctypes binding, synthetic inline meshes, problem set-up helpers).
"""
from .capi import AssemblyPlan, MrhydeB200Error, TimeSpec, lib  # noqa: F401
