"""martini_b200 -- B200-native particle->datacube projection behind MARTINI's Python API.

Scope: the work of ``Martini.insert_source_in_cube`` (prune, per-particle spectra, SPH kernel
pixel integrals, per-pixel accumulation) as hand-written sm_100a CUDA kernels behind a C ABI
(``include/martini_b200.h``).  See DESIGN.md.  There is no CPU fallback.

    from martini_b200 import Martini, DataCube, SPHSource
    from martini_b200.sph_kernels import WendlandC2Kernel
    from martini_b200.spectral_models import GaussianSpectrum
"""

__version__ = "0.1.0"

from . import sph_kernels, spectral_models  # noqa: F401
from .beams import GaussianBeam  # noqa: F401
from .datacube import DataCube  # noqa: F401
from .engine import Engine, KernelTable  # noqa: F401
from .martini import GlobalProfile, Martini, demo  # noqa: F401
from .noise import GaussianNoise  # noqa: F401
from .sources import L_coords, PixelSource, SPHSource, demo_source  # noqa: F401

__all__ = ["Martini", "GlobalProfile", "DataCube", "SPHSource", "PixelSource", "L_coords",
           "demo", "demo_source", "Engine", "KernelTable", "sph_kernels", "spectral_models",
           "GaussianBeam", "GaussianNoise"]
