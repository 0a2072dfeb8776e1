"""The domain plug-in interface of the reference, unchanged.

Mirrors ``DensityStructure`` (reference density_structures/base.py:4-81): user
subclasses written against the reference keep working; ``LintSampler`` drives
them through their own ``choose_cells`` on the host and runs only the in-cell
stage on the device (``lint_incell_sample_u``).
"""
from abc import ABC, abstractmethod


class DensityStructure(ABC):
    """Abstract domain: anything that can pick cells from uniforms and report
    each picked cell's bounding box and 2^k corner densities."""

    @abstractmethod
    def choose_cells(self, u):
        """``u``: (N,) uniforms -> ``(mins (N,k), maxs (N,k), corners)`` where
        ``corners`` is a 2^k-tuple of (N,) arrays in corner order f00..0,
        f00..1, ... (dim 0 = most significant bit)."""
        pass

    @property
    @abstractmethod
    def dim(self):
        """Dimensionality k of the domain."""
        pass

    @property
    @abstractmethod
    def mins(self):
        """Length-k coordinates of the domain's first corner."""
        pass

    @property
    @abstractmethod
    def maxs(self):
        """Length-k coordinates of the domain's last corner."""
        pass

    @property
    @abstractmethod
    def total_mass(self):
        """Total (trapezoid) probability mass summed over the domain's cells."""
        pass
