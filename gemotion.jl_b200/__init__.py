"""gemotion.jl_b200 -- B200-native assembly of GeMotion's Navier-Stokes-Fourier residual + Jacobian.

The product is the C-ABI CUDA library built from `csrc/` (declared in include/gemotion_b200.h);
the Python modules here are the host side that sits where the reference's Julia host code sits
(/root/reference/src/Solver.jl:80-195): FE tables (`fe`), the ctypes binding (`capi`), the
Gridap-style operator plug-in (`operator`), the Gmsh reader (`gmsh`) and the Newton / trust-region drivers (`solver`).

The directory name contains a dot, so it is loaded through `_pkg.py` at the repo root
(`from _pkg import gm`), which registers it as module `gemotion_jl_b200`.
"""
from . import fe  # noqa: F401

__all__ = ["fe", "gmsh", "capi", "operator", "solver"]


def __getattr__(name):
    if name in ("capi", "operator", "solver", "gmsh"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
