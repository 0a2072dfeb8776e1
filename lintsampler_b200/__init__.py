"""B200-native linear-interpolant sampler — drop-in for ``lintsampler``'s
``LintSampler(domain, pdf=..., vectorizedpdf=..., seed=..., qmc=...).sample(N)``
over ``DensityGrid`` and ``DensityTree`` domains.

Python host code (this package) calls hand-written sm_100a CUDA kernels through
the C ABI of ``include/lintsampler_b200.h``; PyTorch supplies device memory and
streams.  There is no CPU fallback.
"""
from .base import DensityStructure
from .grid import DensityGrid
from .tree import DensityTree
from .pdfs import DoublePowerLaw, GaussianMixture
from .sampler import LintSampler

__version__ = "0.1.0"
__all__ = ["LintSampler", "DensityGrid", "DensityTree", "DensityStructure", "GaussianMixture",
           "DoublePowerLaw"]
