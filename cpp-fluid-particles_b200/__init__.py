"""cpp-fluid-particles_b200 -- B200-native SPH particle engine behind the CPP-Fluid-Particles API.

The product is native: libsphk.so (hand-written sm_100a CUDA kernels behind the C-ABI of include/sphk.h)
and libsphhost.so (the reference-shaped C++ classes SPHSystem / BaseSolver / SPHParticles / DArray over
that C-ABI).  This python package is plumbing only: scene generation (scene.py), ctypes bindings
(capi.py), a torch-tensor front end of the C-ABI (engine.py) and the multi-GPU slab driver (slabs.py).

The directory name contains a hyphen; import it through `pkgload.load()` at the repository root, which
registers it as `cpp_fluid_particles_b200`.
"""
from . import scene  # noqa: F401
from . import capi  # noqa: F401

__all__ = ["scene", "capi"]
