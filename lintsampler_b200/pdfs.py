"""Built-in analytic densities with fused device evaluation.

The reference evaluates an arbitrary user callable on a materialised
``(npts, k)`` vertex mesh (density_structures/grid.py:384-398).  A
``GaussianMixture`` or ``DoublePowerLaw`` passed as ``pdf`` is recognised by
``DensityGrid`` and evaluated by a fused kernel (``lint_grid_eval_gmm`` /
``lint_grid_eval_halos``) straight from the edge arrays; called like a function
it evaluates on the host with NumPy, so it also works anywhere a plain callable
does (``DensityTree``, the reference itself).

Protocol of a fused pdf (what ``DensityGrid`` uses): ``dim``, ``_peak_bound(edges)``
(an upper bound of the density on the grid vertices, for the fp32 storage scale) and
``_eval_on_grid(descriptor, scale, stream)``.
"""
import ctypes as C

import numpy as np

from . import _capi


def _eval_points(self, x, device=None):
    """The pdf at points ``x`` (n, k) — or (n,) when k = 1 — evaluated by the device code that also
    fills grid vertices and tree corners (``lint_points_eval``); returns a NumPy array (n,).
    A vectorised pdf with exactly the library's arithmetic, e.g. to cross-check a device-built
    tree against the host builder."""
    import torch
    from . import _device
    dev = _device.require_cuda(device)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, self.dim))
    with torch.cuda.device(dev):
        pts = torch.as_tensor(x, device=dev)
        out = torch.empty(len(x), dtype=torch.float64, device=dev)
        st, keep = self._as_struct()
        gmm = C.byref(st) if isinstance(st, _capi.LintGmm) else None
        halos = C.byref(st) if isinstance(st, _capi.LintHalos) else None
        _capi.call("lint_points_eval", self.dim, gmm, halos, _device.ptr(pts), len(x), _device.ptr(out),
                   _device.stream_ptr(dev))
        del keep
        return out.cpu().numpy()


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

    def _peak_bound(self, edges):
        return float(np.exp(self._log_norm).sum())

    def _eval_on_grid(self, desc, scale, stream):
        gs, keep = self._as_struct()
        _capi.call("lint_grid_eval_gmm", C.byref(desc), C.byref(gs), scale, stream)
        del keep

    eval_points = _eval_points

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


class DoublePowerLaw:
    """Sum of spherical double-power-law halos,

    ``rho(x) = sum_h rho0_h / ((q + eps)**gamma * (1 + q**alpha)**((beta - gamma) / alpha))``,
    ``q = |x - centre_h| / r_s_h``, with a halo contributing nothing beyond ``q > q_cut``.

    ``DoublePowerLaw.nfw(...)`` is the softened, truncated NFW profile of the reference's
    dark-matter example (docs/source/example_notebooks/3_dark_matter.ipynb, cells 11 and 16:
    ``rho_0 / ((q + 0.1) * (1 + q)**2)``, zero for ``q > 10``).
    """

    def __init__(self, rho0, centres, scale_radii, alpha=1.0, beta=3.0, gamma=1.0, eps=0.0, q_cut=None):
        self.rho0 = np.atleast_1d(np.asarray(rho0, dtype=np.float64))
        centres = np.asarray(centres, dtype=np.float64)
        if centres.ndim == 1:
            centres = centres[None, :] if self.rho0.size == 1 else centres[:, None]
        self.centres = np.ascontiguousarray(centres)
        H, k = self.centres.shape
        self.scale_radii = np.broadcast_to(np.asarray(scale_radii, dtype=np.float64), (H,)).copy()
        if self.rho0.shape != (H,):
            raise ValueError("DoublePowerLaw: rho0 and centres disagree on the number of halos")
        if H > _capi.LINT_GMM_MAX_COMP or k > _capi.LINT_MAX_DIM:
            raise ValueError("DoublePowerLaw: at most 16 halos and 6 dimensions")
        if np.any(self.rho0 < 0) or np.any(self.scale_radii <= 0):
            raise ValueError("DoublePowerLaw: rho0 must be non-negative and scale radii positive")
        if not alpha > 0 or eps < 0:
            raise ValueError("DoublePowerLaw: needs alpha > 0 and eps >= 0")
        self.alpha, self.beta, self.gamma = float(alpha), float(beta), float(gamma)
        self.eps = float(eps)
        self.q_cut = 0.0 if q_cut is None else float(q_cut)
        self.dim = k
        self.ncomp = H

    @classmethod
    def nfw(cls, rho0, centres, scale_radii, eps=0.1, q_cut=10.0):
        return cls(rho0, centres, scale_radii, alpha=1.0, beta=3.0, gamma=1.0, eps=eps, q_cut=q_cut)

    def _profile(self, q):
        with np.errstate(divide="ignore"):
            rho = 1.0 / ((q + self.eps) ** self.gamma
                         * (1.0 + q ** self.alpha) ** ((self.beta - self.gamma) / self.alpha))
        if self.q_cut > 0:
            rho = np.where(q > self.q_cut, 0.0, rho)
        return rho

    def __call__(self, x):
        """Host evaluation at ``x`` of shape (..., k), or any shape when k = 1."""
        x = np.asarray(x, dtype=np.float64)
        if self.dim == 1:
            x = x[..., None]
        q = np.linalg.norm(x[..., None, :] - self.centres, axis=-1) / self.scale_radii     # (..., H)
        return (self.rho0 * self._profile(q)).sum(axis=-1)

    def _peak_bound(self, edges):
        """Each halo is largest at the vertex nearest its centre (the profile decreases
        with q for gamma >= 0, beta >= gamma); the sum of those maxima bounds the sum."""
        total = 0.0
        for h in range(self.ncomp):
            d2 = sum(float(np.min((np.asarray(e, dtype=np.float64) - self.centres[h, d]) ** 2))
                     for d, e in enumerate(edges))
            total += self.rho0[h] * float(self._profile(np.asarray(np.sqrt(d2) / self.scale_radii[h])))
        return total

    def _as_struct(self):
        """(LintHalos, keepalive) for the C ABI (host arrays)."""
        s = _capi.LintHalos()
        s.dim, s.ncomp = self.dim, self.ncomp
        dp = C.POINTER(C.c_double)
        keep = (np.ascontiguousarray(self.rho0), np.ascontiguousarray(self.centres),
                np.ascontiguousarray(self.scale_radii))
        s.rho0, s.centre, s.r_s = (a.ctypes.data_as(dp) for a in keep)
        s.alpha, s.beta, s.gamma, s.eps, s.q_cut = self.alpha, self.beta, self.gamma, self.eps, self.q_cut
        return s, keep

    def _eval_on_grid(self, desc, scale, stream):
        s, keep = self._as_struct()
        _capi.call("lint_grid_eval_halos", C.byref(desc), C.byref(s), scale, stream)
        del keep

    eval_points = _eval_points


FUSED_PDFS = (GaussianMixture, DoublePowerLaw)
