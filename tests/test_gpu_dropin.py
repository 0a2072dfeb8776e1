"""GPU tests at the drop-in boundary: ``LintSampler(...).sample(N)`` against the
real reference's outputs (golden) and the reference's own behavioural contract
(shapes, determinism, statistics — modelled on the reference's tests/)."""
import warnings

import numpy as np
import pytest
from scipy import stats
from scipy.stats.qmc import Halton, Sobol

import pdfs
from conftest import load_golden, golden_edges
from oracle import lint_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lintsampler_b200
    return lintsampler_b200


PDFS = {
    "c1_readme_1d": (pdfs.readme_gmm_1d, False),
    "g1d_nonuniform": (pdfs.GMM(*pdfs.gmm_params(1, seed=3)), True),
    "g2d_fullcov": (pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)), True),
    "g3d_iso": (pdfs.GMM(*pdfs.gmm_params(3, seed=0)), True),
    "g3d_holes": (pdfs.bump_with_holes, True),
    "g2d_flat": (pdfs.flat_pdf, True),
    "g2d_onecell": (pdfs.GMM(*pdfs.gmm_params(2, seed=4)), True),
    "g4d_iso": (pdfs.GMM(*pdfs.gmm_params(4, seed=2, smin=0.5, smax=0.9)), True),
    "g6d_iso": (pdfs.GMM(*pdfs.gmm_params(6, seed=6, smin=0.7, smax=1.1)), True),
    "g5d_sobol": (pdfs.GMM(*pdfs.gmm_params(5, seed=0, smin=0.5, smax=0.9)), True),
    "g2d_sobol": (pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)), True),
}


def _domain(g):
    e = golden_edges(g)
    return tuple(e) if len(e) > 1 else e[0]


# ----------------------------------------------------------------------------- exact reproduction of the reference
@pytest.mark.parametrize("name", list(PDFS))
def test_sample_reproduces_reference_exactly(lb, name):
    """Same pdf, domain and seed as the golden run of the real reference; host
    PCG64 stream (or scipy-identical device Sobol), NumPy-order CDF: the final,
    shuffled output of sample(N) is bit-identical."""
    g = load_golden(name)
    pdf, vec = PDFS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = lb.LintSampler(_domain(g), pdf=pdf, vectorizedpdf=vec, seed=int(g["seed"]),
                           qmc=bool(g["qmc"]), rng_backend="numpy", cdf_mode="sequential")
        x = s.sample(int(g["N"]))
    assert np.array_equal(np.atleast_1d(x).reshape(g["x_final"].shape), g["x_final"])


@pytest.mark.parametrize("name,pdf,kw,ref", [
    ("t3d_open3", pdfs.GMM(*pdfs.gmm_params(3, seed=0)), dict(min_openings=3), None),
    ("t3d_refined", pdfs.GMM(*pdfs.gmm_params(3, K=8, seed=0, smin=0.1, smax=0.4)),
     dict(min_openings=3), dict(tree_tol=2e-2, leaf_tol=2e-2)),
    ("t2d_refined", pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)), dict(min_openings=2),
     dict(tree_tol=1e-3)),
    ("t1d_refined", pdfs.GMM(*pdfs.gmm_params(1, seed=3)), dict(min_openings=2),
     dict(tree_tol=1e-4, leaf_tol=1e-3)),
])
def test_tree_sample_reproduces_reference_exactly(lb, name, pdf, kw, ref):
    g = load_golden(name)
    mins, maxs = g["mins"], g["maxs"]
    if len(mins) == 1:
        mins, maxs = float(mins[0]), float(maxs[0])
    tree = lb.DensityTree(mins, maxs, pdf, vectorizedpdf=True, cdf_mode="sequential", **kw)
    if ref:
        tree.refine(**ref)
    x = lb.LintSampler(tree, seed=int(g["seed"]), rng_backend="numpy").sample(int(g["N"]))
    assert np.array_equal(np.atleast_1d(x).reshape(g["x_final"].shape), g["x_final"])
    # the choose_cells return contract [tree.py:314-355] against the reference's leaf table:
    # (leaf.x, leaf.x + leaf.dx, leaf.corner_densities) of the leaf the oracle's search picks
    u = np.random.default_rng(11).random(3000)
    u[:3] = [0.0, 0.5, 1.0 - 2.0 ** -53]
    tmins, tmaxs, tcorners = tree.choose_cells(u)
    idx = O.choose_index(O.cdf_from_masses(g["leaf_mass"]), u)
    assert np.array_equal(tmins, g["leaf_x"][idx])
    assert np.array_equal(tmaxs, g["leaf_x"][idx] + g["leaf_dx"][idx])
    assert len(tcorners) == g["leaf_corners"].shape[1]
    for j, c in enumerate(tcorners):
        assert np.array_equal(c, g["leaf_corners"][idx, j])


@pytest.mark.parametrize("name,qmc", [("m2d_quadrants", False), ("m1d_halves", False),
                                      ("m2d_quadrants_sobol", True)])
def test_multi_domain_reproduces_reference_exactly(lb, name, qmc):
    g = load_golden(name)
    k, nd = int(g["k"]), int(g["ndomains"])
    pdf = (pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)) if k == 2
           else pdfs.GMM(*pdfs.gmm_params(1, seed=3)))
    doms = []
    for i in range(nd):
        e = [g[f"dom{i}_edges{d}"] for d in range(k)]
        doms.append(tuple(e) if k > 1 else e[0])
    s = lb.LintSampler(doms, pdf=pdf, vectorizedpdf=True, seed=int(g["seed"]), qmc=qmc,
                       rng_backend="numpy", cdf_mode="sequential")
    assert np.array_equal(np.array([gr.total_mass for gr in s.grids]), g["domain_masses"])
    x = s.sample(int(g["N"]))
    assert np.array_equal(np.atleast_1d(x).reshape(g["x_final"].shape), g["x_final"])


# ----------------------------------------------------------------------------- statistics of the device stream
def _ks_all_dims(x, y):
    return min(stats.ks_2samp(x[:, d], y[:, d]).pvalue for d in range(x.shape[1]))


@pytest.mark.parametrize("qmc", [False, True])
def test_list_of_grids_single_launch_equals_the_per_domain_path(lb, qmc, monkeypatch):
    """A list of grids is drawn by ONE launch (lint_multi_grid_sample_*: grid choice, the u-rescaling
    of lintsampler.py:306 and the cell choice fused); regrouped by grid it must be, bit for bit, what
    the per-domain path (mask, rescale, one launch per grid, concatenate) returns — for host PCG64
    uniforms and for the device Sobol stream — and the device Philox stream must keep the mixture."""
    import warnings
    pdf = pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True))
    xe, ye = np.linspace(-10, 10, 41), np.linspace(-5, 5, 25)
    doms = [(xe[:21], ye[:13]), (xe[20:], ye[:13]), (xe[:21], ye[12:]), (xe[20:], ye[12:])]
    outs = []
    for fused in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s = lb.LintSampler(doms, pdf=pdf, vectorizedpdf=True, seed=11, qmc=qmc,
                               rng_backend="numpy" if not qmc else "philox", cdf_mode="sequential")
            if not fused:
                monkeypatch.setattr(type(s), "_fused_list", lambda self: False)
            outs.append(s.sample(5000))
            monkeypatch.undo()
    assert np.array_equal(outs[0], outs[1])
    # device Philox rows of the fused launch: same mixture as the reference-stream rows
    x = lb.LintSampler(doms, pdf=pdf, vectorizedpdf=True, seed=5).sample(200_000)
    y = lb.LintSampler(doms, pdf=pdf, vectorizedpdf=True, seed=5, rng_backend="numpy").sample(200_000)
    for d in range(2):
        assert stats.ks_2samp(x[:, d], y[:, d]).pvalue > 1e-4
    frac = [np.mean((x[:, 0] < 0) & (x[:, 1] < 0)), np.mean((y[:, 0] < 0) & (y[:, 1] < 0))]
    assert abs(frac[0] - frac[1]) < 5e-3


@pytest.mark.parametrize("name", ["g1d_nonuniform", "g2d_fullcov", "g3d_iso", "g4d_iso"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_philox_samples_match_reference_distribution(lb, name, dtype):
    """KS per coordinate and chi^2 over cells against samples drawn by the oracle
    (== the reference) with an independent NumPy stream."""
    g = load_golden(name)
    k = int(g["k"])
    pdf, vec = PDFS[name]
    N = 200_000
    grid = lb.DensityGrid(_domain(g), pdf, vectorizedpdf=vec, dtype=dtype)
    x = lb.LintSampler(grid, seed=123).sample(N).reshape(N, k).astype(np.float64)
    u = np.random.default_rng(999).random((N, k + 1))
    x_ref = O.grid_sample(golden_edges(g), g["vertex_densities"], u)
    assert _ks_all_dims(x, x_ref) > 1e-4
    # chi^2 of cell occupancy against the exact cell probabilities
    edges = golden_edges(g)
    idx = [np.clip(np.searchsorted(edges[d], x[:, d], side="right") - 1, 0, len(edges[d]) - 2)
           for d in range(k)]
    counts = np.bincount(np.ravel_multi_index(idx, tuple(len(e) - 1 for e in edges)),
                         minlength=g["masses"].size)
    p = (g["masses"] / g["masses"].sum()).ravel()
    big = p * N >= 20
    obs = np.append(counts[big], counts[~big].sum())
    exp = np.append(p[big], p[~big].sum()) * N
    keep = exp > 0
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4


def test_gaussian_moments_grid_and_tree(lb):
    """Model: reference tests/test_lintsampler.py:569-657 (moments to 1 decimal)."""
    mu, cov = np.array([1.5, -0.5]), np.array([[1.0, 1.2], [1.2, 1.8]])
    pdf = stats.multivariate_normal(mean=mu, cov=cov).pdf
    e = (np.linspace(-12, 12, 513), np.linspace(-12, 12, 513))
    x = lb.LintSampler(e, pdf=pdf, vectorizedpdf=True, seed=42).sample(2 ** 18)
    assert np.all(np.round(x.mean(axis=0), 1) == mu)
    assert np.all(np.round(np.cov(x.T), 1) == cov)
    tree = lb.DensityTree([-12, -12], [12, 12], pdf, vectorizedpdf=True, min_openings=4)
    tree.refine(1e-3)
    x = lb.LintSampler(tree, seed=42).sample(2 ** 18)
    # the tree's coarser interpolant inflates the variance slightly (as in the
    # reference); tolerance 0.1 absolute on the moments
    assert np.allclose(x.mean(axis=0), mu, atol=0.05)
    assert np.allclose(np.cov(x.T), cov, atol=0.1)
    xq = lb.LintSampler(e, pdf=pdf, vectorizedpdf=True, seed=42, qmc=True).sample(2 ** 18)
    assert np.all(np.round(xq.mean(axis=0), 1) == mu)


def test_two_component_mixture_on_disjoint_grids(lb):
    """Model: reference tests/test_lintsampler.py:737-830 (weights by domain)."""
    def pdf(x):
        return 0.3 * stats.norm.pdf(x, -5, 1.0) + 0.7 * stats.norm.pdf(x, 5, 0.5)
    doms = [np.linspace(-10, -1, 257), np.linspace(1, 10, 257)]
    for qmc, order in ((False, "draw"), (True, "draw"), (False, "cell")):
        s = lb.LintSampler(doms, pdf=pdf, vectorizedpdf=True, seed=7, qmc=qmc, order=order)
        x = s.sample(2 ** 17)
        assert round((x < 0).mean(), 2) == 0.3
        assert round(x[x < 0].mean(), 1) == -5.0 and round(x[x > 0].std(), 1) == 0.5
        if order == "draw":        # rows grouped by domain are permuted on the device (reference: rng.shuffle)
            assert abs(np.mean(x[: 2 ** 12] < 0) - 0.3) < 0.05
        assert not np.array_equal(x, s.sample(2 ** 17))


# ----------------------------------------------------------------------------- shapes, determinism, state
@pytest.mark.parametrize("qmc", [False, True])
def test_output_shapes(lb, qmc):
    """Model: reference tests/test_lintsampler.py:356-514 and SURVEY §8b quirks."""
    g1 = np.linspace(-10, 10, 65)
    g2 = (np.linspace(-10, 10, 65), np.linspace(-5, 5, 33))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s1 = lb.LintSampler(g1, pdf=stats.norm.pdf, vectorizedpdf=True, qmc=qmc)
        s2 = lb.LintSampler(g2, pdf=stats.multivariate_normal(np.zeros(2), np.eye(2)).pdf,
                            vectorizedpdf=True, qmc=qmc)
        assert isinstance(s1.sample(), float)
        assert s1.sample(16).shape == (16,)
        assert s1.sample(1).shape == ()
        assert s2.sample().shape == (2,)
        assert s2.sample(16).shape == (16, 2)
        assert s2.sample(1).shape == (1, 2)
        t1 = lb.DensityTree(-10, 10, stats.norm.pdf, vectorizedpdf=True, min_openings=3)
        assert lb.LintSampler(t1, qmc=qmc).sample(8).shape == (8,)


@pytest.mark.parametrize("qmc", [False, True])
@pytest.mark.parametrize("backend", ["philox", "numpy"])
def test_same_seed_same_samples(lb, qmc, backend):
    """Model: reference tests/test_lintsampler.py:520-541."""
    g2 = (np.linspace(-10, 10, 65), np.linspace(-5, 5, 33))
    pdf = stats.multivariate_normal(np.zeros(2), np.eye(2)).pdf
    a = lb.LintSampler(g2, pdf=pdf, vectorizedpdf=True, seed=42, qmc=qmc, rng_backend=backend)
    b = lb.LintSampler(g2, pdf=pdf, vectorizedpdf=True, seed=42, qmc=qmc, rng_backend=backend)
    xa, xb = a.sample(1024), b.sample(1024)
    assert np.all(xa == xb)
    # the stream continues rather than restarts
    assert not np.array_equal(xa, a.sample(1024))
    a = lb.LintSampler(g2, pdf=pdf, vectorizedpdf=True, seed=np.random.default_rng(42), qmc=qmc,
                       rng_backend=backend)
    b = lb.LintSampler(g2, pdf=pdf, vectorizedpdf=True, seed=np.random.default_rng(42), qmc=qmc,
                       rng_backend=backend)
    assert np.all(a.sample(1024) == b.sample(1024))


def test_user_qmc_engine_reset_replays(lb):
    """Model: reference tests/test_lintsampler.py:544-552; a Halton engine goes
    through the host path, a 32-bit Sobol engine through the device path."""
    g2 = (np.linspace(-10, 10, 65), np.linspace(-5, 5, 33))
    pdf = stats.multivariate_normal(np.zeros(2), np.eye(2)).pdf
    for eng in (Halton(d=3, seed=1), Sobol(d=3, bits=32, seed=1), Sobol(d=3, seed=1)):
        s = lb.LintSampler(g2, pdf=pdf, vectorizedpdf=True, qmc=True, qmc_engine=eng)
        x1 = s.sample(256)
        s.rng = np.random.default_rng(0)
        eng.reset()
        s.rng = np.random.default_rng(0)
        x2 = s.sample(256)
        assert sorted(map(tuple, x1)) == sorted(map(tuple, x2))


def test_reset_domain_keeps_random_state(lb):
    g1 = np.linspace(-10, 10, 65)
    s = lb.LintSampler(g1, pdf=stats.norm.pdf, vectorizedpdf=True, seed=3)
    s.sample(100)
    off = s._philox_offset
    s.reset_domain(np.linspace(-5, 5, 33))
    assert s._philox_offset == off and s.grids[0].ncells == 32
    assert s.sample(10).shape == (10,)


def test_sobol_non_power_of_two_warns(lb):
    """Model: reference tests/test_lintsampler.py:177-181."""
    g1 = np.linspace(-10, 10, 65)
    s = lb.LintSampler(g1, pdf=stats.norm.pdf, vectorizedpdf=True, qmc=True)
    with pytest.warns(UserWarning):
        s.sample(10)


def test_user_defined_density_structure(lb):
    """A user subclass of the ABC keeps working: its choose_cells runs on the
    host, the in-cell stage on the device (reference base.py:4-81)."""
    class TwoCells(lb.DensityStructure):
        dim, mins, maxs, total_mass = 1, np.array([0.0]), np.array([2.0]), 1.0

        def choose_cells(self, u):
            right = u >= 0.25
            lo = np.where(right, 1.0, 0.0)[:, None]
            return lo, lo + 1.0, (np.where(right, 1.0, 0.0), np.where(right, 1.0, 1.0))

    x = lb.LintSampler(TwoCells(), seed=1).sample(20000)
    assert x.shape == (20000,) and x.min() >= 0 and x.max() <= 2
    assert abs((x >= 1).mean() - 0.75) < 0.02


def test_device_tensor_return(lb):
    import torch
    g3 = tuple(np.linspace(-4, 4, 33) for _ in range(3))
    gm = lb.GaussianMixture(*_gm3())
    s = lb.LintSampler(g3, pdf=gm, vectorizedpdf=True, seed=5, dtype=np.float32, return_tensor=True)
    x = s.sample(100000)
    assert isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.shape == (100000, 3)
    assert float(x.min()) >= -4 and float(x.max()) <= 4


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_double_power_law_matches_the_callable_path(lb, dtype):
    """A DoublePowerLaw pdf is evaluated on the vertices by lint_grid_eval_halos; the result
    must agree with passing the same object as an ordinary vectorised callable (host NumPy
    evaluation on the materialised mesh, grid.py:384-398) — densities, masses and samples."""
    centres = [np.array([0.5, -2.5, 3.0]), np.array([-2.5, -2.0, -5.5])]
    edges = (np.linspace(-6, 6, 49), np.linspace(-7, 5, 41), np.linspace(-9, 7, 57))
    for dpl in (lb.DoublePowerLaw.nfw([1.0, 2.0], centres, [0.5, 0.5]),
                lb.DoublePowerLaw([3e-7, 1e-7], centres, [0.7, 1.3], alpha=2.0, beta=5.0, gamma=0.5, eps=0.05)):
        fused = lb.DensityGrid(edges, dpl, vectorizedpdf=True, dtype=dtype)
        plain = lb.DensityGrid(edges, lambda x: dpl(x), vectorizedpdf=True, dtype=dtype)
        tol = 1e-12 if dtype == np.float64 else 3e-7
        np.testing.assert_allclose(fused.vertex_densities, plain.vertex_densities, rtol=tol,
                                   atol=tol * plain.vertex_densities.max())
        assert fused.total_mass == pytest.approx(plain.total_mass, rel=10 * tol)
        a = lb.LintSampler(fused, seed=3).sample(200_000)
        b = lb.LintSampler(plain, seed=3).sample(200_000)
        # same uniforms; vertex values differ in the last bits only, so all but a handful of
        # draws (those on a CDF boundary) land on the same point
        close = np.all(np.abs(a - b) <= (1e-9 if dtype == np.float64 else 1e-3), axis=1)
        assert close.mean() > 0.999
    with pytest.raises(ValueError):
        lb.DensityGrid(edges[:2], dpl, vectorizedpdf=True)


def _gm3():
    w, mu, cov = pdfs.gmm_params(3, seed=0)
    return w, mu, cov


@pytest.mark.parametrize("factor", [1e-42, 1e-30, 1.0, 1e+30, 1e+42])
def test_fp32_mode_is_scale_invariant(lb, factor):
    """Unnormalised densities in physical units must not flush to zero or overflow in
    fp32 storage: the same seed gives the same samples whatever the overall factor,
    and masses/densities are reported unscaled."""
    g = load_golden("g3d_iso")
    edges = tuple(golden_edges(g))
    V = g["vertex_densities"]
    ref = lb.LintSampler(lb.DensityGrid(edges, lambda p: V.reshape(-1), vectorizedpdf=True,
                                        dtype=np.float32), seed=4).sample(50_000)
    grid = lb.DensityGrid(edges, lambda p: V.reshape(-1) * factor, vectorizedpdf=True, dtype=np.float32)
    x = lb.LintSampler(grid, seed=4).sample(50_000)
    assert np.allclose(x, ref, rtol=0, atol=1e-5)
    assert grid.total_mass == pytest.approx(float(g["total_mass"]) * factor, rel=1e-6)
    np.testing.assert_allclose(grid.vertex_densities, V * factor, rtol=1e-6, atol=1e-12 * V.max() * factor)
    pdf, _ = PDFS["g3d_iso"]
    w, mu, cov = pdfs.gmm_params(3, seed=0)
    gm = lb.GaussianMixture(w * factor, mu, cov)
    g2 = lb.DensityGrid(edges, gm, vectorizedpdf=True, dtype=np.float32)
    assert g2.total_mass == pytest.approx(float(g["total_mass"]) * factor, rel=1e-5)


def test_sorted_sobol_draw_is_the_same_point_set(lb):
    """With the default (device) shuffle a fresh power-of-two QMC draw uses the sorted
    enumeration of the Sobol points; the samples are, as a set, exactly those of the
    direct enumeration (= the reference's)."""
    g = load_golden("g5d_sobol")
    grid = lb.DensityGrid(_domain(g), PDFS["g5d_sobol"][0], vectorizedpdf=True, cdf_mode="sequential")
    N = int(g["N"])
    s = lb.LintSampler(grid, seed=int(g["seed"]), qmc=True)
    assert s._sobol_sorted_ok(N)
    x = s.sample(N)
    assert s.qmc_engine.num_generated == N
    ref = g["x_pre"]                                        # reference's pre-shuffle samples
    assert np.array_equal(x[np.lexsort(x.T)], ref[np.lexsort(ref.T)])
    assert not np.array_equal(x, ref)                       # but in permuted order
    # a second call continues the sequence through the direct enumeration
    x2 = s.sample(N)
    assert x2.shape == ref.shape and s.qmc_engine.num_generated == 2 * N
    direct = lb.LintSampler(grid, seed=int(g["seed"]), qmc=True, shuffle=False).sample(N)
    assert np.array_equal(direct, ref)


def test_device_resident_inputs_and_updates(lb):
    """pdf_on_device (torch callable), from_densities (NumPy / pinned / device tensors) and
    update_densities all reach the same tables as the host-evaluated construction."""
    import torch
    g = load_golden("g3d_iso")
    edges = tuple(golden_edges(g))
    V = g["vertex_densities"]
    ref = lb.DensityGrid(edges, PDFS["g3d_iso"][0], vectorizedpdf=True, cdf_mode="sequential")
    w, mu, cov = pdfs.gmm_params(3, seed=0)
    sig2 = torch.tensor([c[0, 0] for c in cov], device="cuda")
    wt, mut = torch.tensor(w, device="cuda"), torch.tensor(mu, device="cuda")

    def torch_pdf(x):                                   # isotropic mixture on the device
        d2 = ((x[:, None, :] - mut[None]) ** 2).sum(-1)
        return (wt * torch.exp(-0.5 * d2 / sig2) / (2 * torch.pi * sig2) ** 1.5).sum(-1)

    a = lb.DensityGrid(edges, torch_pdf, vectorizedpdf=True, pdf_on_device=True, cdf_mode="sequential")
    np.testing.assert_allclose(a.vertex_densities, V, rtol=1e-10)
    for dens in (V, torch.from_numpy(V.copy()).pin_memory(), torch.from_numpy(V.copy()).cuda()):
        b = lb.DensityGrid.from_densities(edges, dens, cdf_mode="sequential")
        assert np.array_equal(b.masses, ref.masses) and b.total_mass == ref.total_mass
        assert torch.equal(b._cdf, ref._cdf)
    b.update_densities(2.0 * V)                         # same shape of density, twice the mass
    assert b.total_mass == pytest.approx(2 * ref.total_mass, rel=1e-14)
    assert torch.equal(b._cdf, ref._cdf)
    with pytest.raises(ValueError):
        b.update_densities(-V)
    with pytest.raises(ValueError):
        lb.DensityGrid.from_densities(edges, V[:-1])


# ---------------------------------------------------------------------------
# device tree builder: lint_tree_open against the host builder
# ---------------------------------------------------------------------------
def _fused_pdf(lb, kind, k):
    rng = np.random.default_rng(7 + k)
    if kind == "gmm":
        return lb.GaussianMixture(rng.dirichlet(np.ones(3)), rng.uniform(-2, 2, (3, k)),
                                  sigmas=rng.uniform(0.15, 0.6, 3))
    return lb.DoublePowerLaw.nfw(rng.uniform(0.5, 2.0, 2), rng.uniform(-2, 2, (2, k)) + 0.01234,
                                 rng.uniform(0.3, 0.8, 2))


@pytest.mark.parametrize("kind,k,openings", [("gmm", 1, 3), ("gmm", 2, 3), ("gmm", 3, 2), ("nfw", 2, 3),
                                             ("nfw", 3, 2), ("gmm", 4, 1)])
def test_device_tree_build_equals_host_builder(lb, kind, k, openings):
    """A tree built in the device cell table (lint_tree_open: corner densities, trapezoid masses,
    Romberg chains) and refined by the array version of `refine` is, bit for bit, the tree the host
    builder makes from the same densities (the pdf evaluated through lint_points_eval): same
    leaves in the same order, same corner tables, raw and Romberg masses, Romberg chains."""
    pdf = _fused_pdf(lb, kind, k)
    mins, maxs = np.full(k, -4.0), np.full(k, 4.25)
    dev = lb.DensityTree(mins, maxs, pdf, min_openings=openings)
    host = lb.DensityTree(mins, maxs, pdf.eval_points, vectorizedpdf=True, min_openings=openings)
    assert dev._dt is not None and host._dt is None
    for stage in range(2):
        a, b = dev.leaf_table(), host.leaf_table()
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y)
        la, lb_ = dev.leaves, host.leaves
        assert [c.idx for c in la] == [c.idx for c in lb_] and [c.level for c in la] == [c.level for c in lb_]
        assert all(np.array_equal(p.romberg_estimates, q.romberg_estimates) for p, q in zip(la, lb_))
        assert all(p.mass_romberg == q.mass_romberg and p.vol == q.vol for p, q in zip(la, lb_))
        assert dev.total_mass == host.total_mass
        if stage == 0:
            dev.refine(2e-3, leaf_tol=2e-2)
            host.refine(2e-3, leaf_tol=2e-2)
            assert len(dev.leaves) > len(la)                       # the refinement did open cells
    # the same refinement with stage 2 cut into launches of three openings, and with the host-side
    # stage-2 loop over the device table (verbose mode): the same leaves again
    for variant in ("batches", "verbose"):
        other = lb.DensityTree(mins, maxs, pdf, min_openings=openings, build="device")
        if variant == "batches":
            other._dt.room_override = 3
        other.refine(2e-3, leaf_tol=2e-2, verbose=variant == "verbose")
        assert np.array_equal(other._leaf_ids, dev._leaf_ids)
        for x, y in zip(other.leaf_table(), dev.leaf_table()):
            assert np.array_equal(x, y)
    pos = np.full(k, 0.3)
    p, q = dev.get_leaf_at_pos(pos), host.get_leaf_at_pos(pos)
    assert p.idx == q.idx and np.array_equal(p.x, q.x) and np.array_equal(p.corner_densities, q.corner_densities)
    # and it samples like any other tree
    s1 = lb.LintSampler(dev, seed=3, rng_backend="numpy").sample(500)
    s2 = lb.LintSampler(host, seed=3, rng_backend="numpy").sample(500)
    assert np.array_equal(s1, s2)


def test_device_tree_build_checks_densities(lb):
    """tree.py:602-605 on the device path: an unsoftened halo centred on a lattice vertex is infinite there."""
    halo = lb.DoublePowerLaw(1.0, [[0.0, 0.0]], 0.5, eps=0.0)
    with pytest.raises(ValueError, match="non-finite"):
        lb.DensityTree([-4.0, -4.0], [4.0, 4.0], halo, min_openings=2)
    with pytest.raises(ValueError):
        lb.DensityTree([-4.0, -4.0], [4.0, 4.0], lambda x: x.sum(-1), vectorizedpdf=True, build="device")
