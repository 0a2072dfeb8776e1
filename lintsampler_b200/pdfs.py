"""Built-in analytic densities with fused device evaluation.

The reference evaluates an arbitrary user callable on a materialised
``(npts, k)`` vertex mesh (density_structures/grid.py:384-398).  A
``GaussianMixture`` passed as ``pdf`` is recognised by ``DensityGrid`` and
evaluated by the ``lint_grid_eval_gmm`` kernel straight from the edge arrays;
called like a function it evaluates on the host with NumPy, so it also works
anywhere a plain callable does (``DensityTree``, the reference itself).
"""
import ctypes as C

import numpy as np

from . import _capi


class GaussianMixture:
    """``sum_j w_j N(x; mu_j, Sigma_j)`` in k dimensions.

    Parameters
    ----------
    weights : (K,) array
    means : (K, k) array (or (K,) for k = 1)
    covs : (K, k, k) full covariances, (K,) isotropic sigmas**2 are NOT implied:
        pass ``sigmas=`` for isotropic components instead.
    sigmas : (K,) isotropic standard deviations (alternative to ``covs``).
    """

    def __init__(self, weights, means, covs=None, sigmas=None):
        self.weights = np.asarray(weights, dtype=np.float64).reshape(-1)
        means = np.asarray(means, dtype=np.float64)
        if means.ndim == 1:
            means = means[:, None]
        self.means = means
        K, k = means.shape
        if self.weights.shape != (K,):
            raise ValueError("GaussianMixture: weights and means disagree on the number of components")
        if (covs is None) == (sigmas is None):
            raise ValueError("GaussianMixture: give exactly one of covs= or sigmas=")
        if sigmas is not None:
            s = np.asarray(sigmas, dtype=np.float64).reshape(-1)
            if s.shape != (K,) or np.any(s <= 0):
                raise ValueError("GaussianMixture: sigmas must be K positive numbers")
            covs = s[:, None, None] ** 2 * np.eye(k)[None]
        covs = np.asarray(covs, dtype=np.float64)
        if covs.shape != (K, k, k):
            raise ValueError(f"GaussianMixture: covs must have shape {(K, k, k)}")
        if K > _capi.LINT_GMM_MAX_COMP or k > _capi.LINT_MAX_DIM:
            raise ValueError("GaussianMixture: at most 16 components and 6 dimensions")
        if np.any(self.weights < 0):
            raise ValueError("GaussianMixture: weights must be non-negative")
        self.covs = covs
        self.dim = k
        self.ncomp = K
        # kernel convention: Sigma^-1 = U^T U with U upper triangular, so the
        # Mahalanobis term is |U (x - mu)|^2
        L = np.linalg.cholesky(covs)
        prec = np.linalg.inv(covs)
        self._U = np.ascontiguousarray(np.swapaxes(np.linalg.cholesky(prec), 1, 2))
        logdet = 2.0 * np.log(np.diagonal(L, axis1=1, axis2=2)).sum(axis=1)
        with np.errstate(divide="ignore"):
            self._log_norm = np.ascontiguousarray(
                np.log(self.weights) - 0.5 * logdet - 0.5 * k * np.log(2.0 * np.pi))

    def __call__(self, x):
        """Host evaluation at ``x`` of shape (..., k), or any shape when k = 1."""
        x = np.asarray(x, dtype=np.float64)
        if self.dim == 1:
            x = x[..., None]
        d = x[..., None, :] - self.means                     # (..., K, k)
        y = np.einsum('jrc,...jc->...jr', self._U, d)
        q = np.sum(y * y, axis=-1)
        return np.exp(self._log_norm - 0.5 * q).sum(axis=-1)

    def _as_struct(self):
        """(LintGmm, keepalive) for the C ABI (host arrays)."""
        s = _capi.LintGmm()
        s.dim, s.ncomp = self.dim, self.ncomp
        ln = np.ascontiguousarray(self._log_norm)
        mu = np.ascontiguousarray(self.means)
        U = np.ascontiguousarray(self._U)
        dp = C.POINTER(C.c_double)
        s.log_norm = ln.ctypes.data_as(dp)
        s.mean = mu.ctypes.data_as(dp)
        s.prec_chol = U.ctypes.data_as(dp)
        return s, (ln, mu, U)
