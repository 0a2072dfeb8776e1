"""Deterministic test densities shared by tests/golden/make_golden.py (which
feeds them to the real reference) and by the parity tests (which feed them to
the B200 build).  Plain NumPy so they behave identically on both sides."""
import numpy as np


def gmm_params(k, K=4, seed=0, smin=0.3, smax=0.8, full_cov=False):
    """Synthetic mixture in the style of SURVEY.md §8(d): means U(-2,2)^k,
    isotropic sigmas U(smin,smax) (or full covariances A A^T + 0.05 I),
    Dirichlet(1) weights, all drawn from default_rng(seed) in that order."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-2, 2, size=(K, k))
    if full_cov:
        A = rng.uniform(-0.5, 0.5, size=(K, k, k))
        covs = A @ np.swapaxes(A, 1, 2) + 0.05 * np.eye(k)
    else:
        sig = rng.uniform(smin, smax, size=K)
        covs = sig[:, None, None] ** 2 * np.eye(k)[None]
    w = rng.dirichlet(np.ones(K))
    return w, means, covs


class GMM:
    """Vectorised k-D Gaussian mixture; accepts (..., k), or any array of
    scalar positions when k=1 (the reference passes 1-D grids that way,
    grid.py:388)."""

    def __init__(self, w, means, covs):
        self.w = np.asarray(w, dtype=np.float64)
        self.means = np.asarray(means, dtype=np.float64)
        self.covs = np.asarray(covs, dtype=np.float64)
        self.k = self.means.shape[1]
        self.prec = np.linalg.inv(self.covs)
        self.norm = self.w / np.sqrt((2 * np.pi) ** self.k * np.linalg.det(self.covs))

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        if self.k == 1:
            x = x[..., None]
        d = x[..., None, :] - self.means
        q = np.einsum('...ji,jil,...jl->...j', d, self.prec, d)
        return (self.norm * np.exp(-0.5 * q)).sum(axis=-1)


def readme_gmm_1d(x):
    """The README quickstart density (reference README.md:42-49)."""
    mu = np.array([-3.0, 0.5, 2.5])
    sig = np.array([1.0, 0.25, 0.75])
    w = np.array([0.4, 0.25, 0.35])
    return np.sum([w[i] * np.exp(-0.5 * ((x - mu[i]) / sig[i]) ** 2)
                   / (sig[i] * np.sqrt(2 * np.pi)) for i in range(3)], axis=0)


def bump_with_holes(x):
    """Non-negative density that is exactly zero on large regions, to exercise
    flat CDF runs (zero-mass cells are never selected, utils.py:197)."""
    x = np.asarray(x, dtype=np.float64)
    r2 = (x ** 2).sum(axis=-1)
    return np.where(r2 < 4.0, (4.0 - r2) ** 2, 0.0) * (x[..., 0] > -0.5)


def flat_pdf(x):
    """Constant density: every cell hits the f0 == f1 branch (sampling.py:32)."""
    x = np.asarray(x)
    return np.ones(x.shape[:-1] if x.ndim > 1 else x.shape)
