"""CPU-only tests: host-side mirror of the reference interface (argument
validation, exception and warning types — modelled on the reference's
tests/test_lintsampler.py:76-350 and tests/test_densitytree.py:10-194), the
tree builder against golden leaf tables from the real reference, and the C-ABI
library's exported symbols.  No kernel is launched here."""
import os
import re
import warnings

import numpy as np
import pytest
from scipy.stats import norm, multivariate_normal
from scipy.stats.qmc import Sobol

import pdfs
from conftest import load_golden, ROOT
import lintsampler_b200 as lb
from lintsampler_b200 import _capi

X = np.linspace(-10, 10, 65)
Y = np.linspace(-5, 5, 33)
PDF2 = multivariate_normal(np.zeros(2), np.eye(2)).pdf


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


# ----------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lintsampler_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|size_t|const char \*)\s*\*?(lint_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    lib = _capi.load()                       # resolves each symbol or raises
    assert lib.lint_abi_version() == 2
    assert lib.lint_cdf_scratch_bytes(1 << 24) > 0


def test_abi_rejects_bad_arguments_without_a_gpu():
    import ctypes as C
    lib = _capi.load()
    assert lib.lint_grid_masses(None, None, None, None) == 1          # LINT_ERR_BAD_ARG
    assert b"NULL" in lib.lint_last_error()
    g = _capi.LintGrid()
    g.dim = 9
    assert lib.lint_grid_sample_u(C.byref(g), None, 0, None, None, None) == 1
    assert b"dim" in lib.lint_last_error()
    with pytest.raises(_capi.LintError):
        _capi.call("lint_guide_build", None, 10, 4, None, None)


def test_product_fails_loudly_without_a_device():
    if not _no_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lb.DensityGrid(X, norm.pdf, vectorizedpdf=True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lintsampler_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), fn


# ----------------------------------------------------------------------------- density validation (before any device work)
def _neg1(x):
    return -norm.pdf(x)


def _neg2(x):
    return -multivariate_normal(mean=np.ones(2), cov=np.eye(2)).pdf(x)


@pytest.mark.parametrize("pdf,cells", [(_neg1, X), (_neg2, (X, Y))])
@pytest.mark.parametrize("vec", [True, False])
def test_negative_density_raises(pdf, cells, vec):
    with pytest.raises(ValueError):
        lb.LintSampler(domain=cells, pdf=pdf, vectorizedpdf=vec)


def _nan_vec_1d(x):
    f = np.ones_like(x)
    f[x < 0] = np.nan
    return f


def _nan_vec_kd(x):
    f = np.ones(x.shape[:-1])
    f[x[..., 0] < 0] = np.nan
    return f


@pytest.mark.parametrize("pdf,cells,vec", [
    (lambda x: np.nan if x < 0 else 1.0, X, False),
    (_nan_vec_1d, X, True),
    (lambda x: np.nan if x[0] < 0 else 1.0, (X, Y), False),
    (_nan_vec_kd, (X, Y), True),
])
def test_nonfinite_density_raises(pdf, cells, vec):
    with pytest.raises(ValueError):
        lb.LintSampler(domain=cells, pdf=pdf, vectorizedpdf=vec)


@pytest.mark.parametrize("pdf,cells,vec", [
    (lambda x: np.ones(10), X, False),
    (lambda x: 1.0, X, True),
    (lambda x: np.ones((len(x), 2)), X, True),
    (lambda x: np.ones(2), (X, Y), False),
    (lambda x: 1.0, (X, Y), True),
    (lambda x: np.ones((len(x), 2)), (X, Y), True),
])
def test_bad_density_shape_raises(pdf, cells, vec):
    with pytest.raises(ValueError):
        lb.LintSampler(domain=cells, pdf=pdf, vectorizedpdf=vec)


def test_pdf_argument_checks():
    with pytest.raises(TypeError):
        lb.LintSampler(X, pdf=3.0)
    with pytest.raises(ValueError):
        lb.LintSampler(X)
    with pytest.raises(ValueError):
        lb.LintSampler([X[:33], X[32:]])
    with pytest.raises(ValueError), pytest.warns(UserWarning):
        lb.LintSampler(X, vectorizedpdf=True)


# ----------------------------------------------------------------------------- domain validation
@pytest.mark.parametrize("bad", [
    np.array([1.0, 1.5, 1.5, 2.0]), np.ones(10), np.array([9.0, 8.0, 10.0, 12.0]),
    np.array([1.0, 1.5, np.nan]), np.array([1.0, 1.5, np.inf]),
])
def test_bad_edges_raise(bad):
    with pytest.raises(ValueError):
        lb.LintSampler(bad, pdf=norm.pdf, vectorizedpdf=True)
    with pytest.raises(ValueError):
        lb.LintSampler((X, bad), pdf=PDF2, vectorizedpdf=True)
    with pytest.raises(ValueError):
        lb.DensityGrid((bad, Y), PDF2, vectorizedpdf=True)


def test_nonsense_domains_raise():
    with pytest.raises(TypeError):
        lb.DensityGrid(17.0, norm.pdf)
    with pytest.raises(TypeError):
        lb.DensityGrid([[X, Y]], PDF2)
    with pytest.raises(TypeError):
        lb.LintSampler([X, (X, Y)], pdf=norm.pdf, vectorizedpdf=True)


def test_seed_and_N_validation_messages():
    with pytest.raises(TypeError):
        np.random.default_rng(42.5)          # the reference relies on NumPy for this (lintsampler.py:458)


# ----------------------------------------------------------------------------- tree: validation + structure (host only)
def test_tree_argument_validation():
    with pytest.raises(ValueError):
        lb.DensityTree([0, 0], [1, 1, 1], PDF2)
    with pytest.raises(ValueError):
        lb.DensityTree([[0, 0]], [[1, 1]], PDF2)
    with pytest.raises(ValueError):
        lb.DensityTree([0, 2], [1, 1], PDF2)
    with pytest.raises(ValueError):
        lb.DensityTree([0, 0], [1, np.inf], PDF2)
    with pytest.raises(ValueError):
        lb.DensityTree(-10, 10, _neg1, vectorizedpdf=True)
    with pytest.raises(ValueError):
        lb.DensityTree(-10, 10, _nan_vec_1d, vectorizedpdf=True)
    with pytest.raises(ValueError):
        lb.DensityTree(-10, 10, lambda x: np.ones((len(x), 2)), vectorizedpdf=True)
    with pytest.warns(UserWarning):
        lb.DensityTree([-5, -5], [5, 5], PDF2, vectorizedpdf=True, batch=True, usecache=True)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("openings", [0, 1, 2, 3])
@pytest.mark.parametrize("batch", [False, True])
def test_tree_leaf_count(dim, openings, batch):
    pdf = multivariate_normal(np.zeros(dim), np.eye(dim)).pdf
    t = lb.DensityTree(-5 * np.ones(dim), 5 * np.ones(dim), pdf, vectorizedpdf=True,
                       min_openings=openings, usecache=not batch, batch=batch)
    assert len(t.leaves) == 2 ** (dim * openings)


def test_tree_refine_needs_two_levels_and_verbose_prints(capfd):
    t = lb.DensityTree([-5, -5], [5, 5], PDF2, vectorizedpdf=True, min_openings=0)
    with pytest.raises(RuntimeError):
        t.refine(1e-2)
    t = lb.DensityTree([-5, -5], [5, 5], PDF2, vectorizedpdf=True, min_openings=2)
    t.refine(1e-2, verbose=True)
    assert "Pre-loop" in capfd.readouterr().out


def test_tree_get_leaf_at_pos():
    t = lb.DensityTree([-5, -5], [5, 5], PDF2, vectorizedpdf=True, min_openings=2)
    t.refine(1e-2)
    with pytest.raises(ValueError):
        t.get_leaf_at_pos(np.array([0.0, 0.0, 0.0]))
    with pytest.raises(ValueError):
        t.get_leaf_at_pos(np.array([0.0, 7.0]))
    pos = np.array([1.02, -3.74])
    leaf = t.get_leaf_at_pos(pos)
    assert leaf.children is None and leaf in t.leaves
    assert np.all(pos >= leaf.x) and np.all(pos <= leaf.x + leaf.dx)
    t1 = lb.DensityTree(-10, 10, norm.pdf, vectorizedpdf=True, min_openings=3)
    assert t1.get_leaf_at_pos(0.3).level == 3


CASES = {
    "t3d_open3": (pdfs.GMM(*pdfs.gmm_params(3, seed=0)), dict(min_openings=3), None),
    "t3d_refined": (pdfs.GMM(*pdfs.gmm_params(3, K=8, seed=0, smin=0.1, smax=0.4)),
                    dict(min_openings=3), dict(tree_tol=2e-2, leaf_tol=2e-2)),
    "t2d_refined": (pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)), dict(min_openings=2),
                    dict(tree_tol=1e-3)),
    "t2d_batch": (pdfs.GMM(*pdfs.gmm_params(2, seed=2)), dict(min_openings=3, batch=True, usecache=False),
                  dict(tree_tol=5e-3)),
    "t1d_refined": (pdfs.GMM(*pdfs.gmm_params(1, seed=3)), dict(min_openings=2),
                    dict(tree_tol=1e-4, leaf_tol=1e-3)),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("vectorized", [True, False])
def test_tree_builder_reproduces_reference_leaf_table(name, vectorized):
    """Leaf ORDER, levels, origins, corner densities, trapezoid and Romberg masses
    are bit-identical to the real reference's tree (golden), whichever way the
    pdf is evaluated."""
    pdf, kw, ref = CASES[name]
    if kw.get("batch") and not vectorized:
        pytest.skip("batch mode always calls the pdf vectorised")
    g = load_golden(name)
    mins, maxs = g["mins"], g["maxs"]
    if len(mins) == 1:
        mins, maxs = float(mins[0]), float(maxs[0])
    t = lb.DensityTree(mins, maxs, pdf, vectorizedpdf=vectorized, **kw)
    if ref:
        t.refine(**ref)
    x, dx, cd, m = t.leaf_table()
    assert np.array_equal(np.array([leaf.level for leaf in t.leaves]), g["leaf_level"])
    assert np.array_equal(np.array([leaf.idx for leaf in t.leaves]), g["leaf_idx"])
    assert np.array_equal(x, g["leaf_x"]) and np.array_equal(dx, g["leaf_dx"])
    # a vectorised einsum and a per-point call may differ in the last bit
    if vectorized:
        assert np.array_equal(cd, g["leaf_corners"]) and np.array_equal(m, g["leaf_mass"])
        assert np.array_equal(np.array([leaf.mass_romberg for leaf in t.leaves]), g["leaf_mass_romberg"])
    else:
        np.testing.assert_allclose(cd, g["leaf_corners"], rtol=1e-12)
    assert t.total_mass == pytest.approx(float(g["total_mass"]), rel=1e-12)


def test_tree_cache_saves_evaluations():
    calls = [0]

    def pdf(x):
        calls[0] += len(x)
        return PDF2(x)
    lb.DensityTree([-5, -5], [5, 5], pdf, vectorizedpdf=True, min_openings=4, usecache=True)
    assert calls[0] == 17 * 17                # every lattice vertex exactly once
    calls[0] = 0
    lb.DensityTree([-5, -5], [5, 5], pdf, vectorizedpdf=True, min_openings=4, usecache=False)
    assert calls[0] > 17 * 17


# ----------------------------------------------------------------------------- GaussianMixture host evaluation
@pytest.mark.parametrize("k,full", [(1, False), (2, True), (3, True), (6, False)])
def test_gaussian_mixture_host_call_matches_scipy(k, full):
    w, mu, cov = pdfs.gmm_params(k, K=3, seed=k, full_cov=full)
    gm = lb.GaussianMixture(w, mu, cov)
    x = np.random.default_rng(0).normal(size=(50, k))
    ref = sum(w[j] * multivariate_normal(mu[j], cov[j]).pdf(x) for j in range(3))
    np.testing.assert_allclose(gm(x if k > 1 else x[:, 0]), ref, rtol=1e-11)
    with pytest.raises(ValueError):
        lb.GaussianMixture(w, mu)


def test_double_power_law_host_call_matches_the_notebook_nfw():
    """DoublePowerLaw.nfw == the two-halo pdf of the reference's dark-matter example
    (docs/source/example_notebooks/3_dark_matter.ipynb, cells 11 and 16), restated here."""
    def nfw_density_vectorized(X, rho_0, r_s):
        r = np.linalg.norm(X, axis=-1)
        q = r / r_s
        rho = rho_0 / ((q + 1e-1) * (1 + q) ** 2)
        rho[q > 10] = 0
        return rho
    centres = [np.array([0.5, -2.5, 3.0]), np.array([-2.5, -2.0, -5.5])]
    rho_0s, scale_radii = [1.0, 2.0], [0.5, 0.5]
    X = np.random.default_rng(0).uniform(-8, 8, size=(4000, 3))
    ref = np.sum([nfw_density_vectorized(X - centres[h], rho_0s[h], scale_radii[h]) for h in range(2)], axis=0)
    dpl = lb.DoublePowerLaw.nfw(rho_0s, centres, scale_radii)
    np.testing.assert_allclose(dpl(X), ref, rtol=1e-13)
    assert (ref == 0).any() and np.array_equal(dpl(X) == 0, ref == 0)          # the q > 10 truncation
    # general exponents (Hernquist: alpha 1, beta 4, gamma 1) and the 1-D calling convention
    h = lb.DoublePowerLaw(2.0, [0.25], 0.5, alpha=1, beta=4, gamma=1, eps=0.01)
    x = np.linspace(-3, 3, 41)
    q = np.abs(x - 0.25) / 0.5
    np.testing.assert_allclose(h(x), 2.0 / ((q + 0.01) * (1 + q) ** 3), rtol=1e-13)
    # the bound used for the fp32 storage scale really bounds the vertex values
    edges = (np.linspace(-6, 6, 30), np.linspace(-6, 6, 31), np.linspace(-7, 7, 32))
    mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1)
    assert dpl(mesh).max() <= dpl._peak_bound(edges) * (1 + 1e-12)
    with pytest.raises(ValueError):
        lb.DoublePowerLaw([1.0, 2.0], centres, [0.5, -0.5])
    with pytest.raises(ValueError):
        lb.DoublePowerLaw([1.0], centres, 0.5)


@pytest.mark.parametrize("d,m,seed", [(2, 5, 3), (4, 10, 1), (6, 12, 7)])
def test_sobol_sorted_enumeration_construction(d, m, seed):
    """The GF(2) inverse used by the sorted Sobol mode: enumerating y = 0..2^m-1 yields
    exactly scipy's first 2^m points, in increasing order of the last coordinate."""
    from lintsampler_b200.sampler import sobol_sorted_inverse
    e = Sobol(d=d, bits=32, seed=seed)
    sv, sh = e._sv.copy(), e._shift.copy()
    U = e.random(2 ** m)
    k = d - 1
    ainv = sobol_sorted_inverse(sv[k], m)
    top = int(sh[k]) >> (32 - m)
    X = np.empty((2 ** m, d), dtype=np.uint64)
    for y in range(2 ** m):
        t, g = y ^ top, 0
        for r in range(m):
            if (t >> r) & 1:
                g ^= ainv[r]
        x = [int(v) for v in sh]
        for c in range(m):
            if (g >> c) & 1:
                x = [x[j] ^ int(sv[j, c]) for j in range(d)]
        assert x[k] >> (32 - m) == y
        X[y] = x
    assert np.array_equal(U[np.argsort(U[:, k])], X.astype(np.float64) * 2.0 ** -32)


def test_ptrs_fp32_screen_never_decides_wrongly():
    """The Poisson count sampler settles most PTRS acceptance tests in fp32 (csrc/lint_cellmajor.cuh);
    scripts/ptrs_fp32_screen_check.py replays that screen in float32 with the worst-case error of the
    device intrinsics added: its error must stay well inside the tolerance band, for every lambda
    regime the kernel uses it in (10 <= lambda < 1e5)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                        "ptrs_fp32_screen_check.py")
    spec = importlib.util.spec_from_file_location("ptrs_fp32_screen_check", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.check(40_000, seed=1)
    assert res["proposals"] > 100_000
    assert res["wrong_decisions"] == 0
    assert res["max_err_over_tol"] < 0.5


def test_sobol_advance_equals_fast_forward():
    """The O(bits) state jump used after a device Sobol draw leaves the scipy engine exactly where
    ``fast_forward`` would (next points identical, including across several jumps)."""
    import warnings
    from lintsampler_b200.sampler import sobol_advance
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n1, n2, m in [(1000, 0, 7), (1, 5, 3), (12345, 678, 11), (2 ** 20, 3, 5), (100_000_000, 1, 4)]:
            a, b = Sobol(d=4, bits=32, seed=7), Sobol(d=4, bits=32, seed=7)
            if n1 <= 2 ** 20:
                a.random(n1)
                if n2:
                    a.random(n2)
            else:
                a.fast_forward(n1 + n2)
            sobol_advance(b, n1)
            if n2:
                sobol_advance(b, n2)
            assert a.num_generated == b.num_generated
            assert np.array_equal(a.random(m), b.random(m)), (n1, n2)


def test_tree_builder_choice_without_cuda():
    """Which builder a DensityTree uses: the device table only for the library's own pdf families
    and only where a CUDA device exists; explicit requests are validated."""
    import torch
    from lintsampler_b200.tree import DensityTree
    gmm = lb.GaussianMixture([0.5, 0.5], [[-1.0, 0.0], [1.0, 0.5]], sigmas=[0.5, 0.4])
    assert DensityTree._wants_device_build(gmm, (), {}, "host") is False
    assert DensityTree._wants_device_build(lambda x: x, (), {}, None) is False
    assert DensityTree._wants_device_build(gmm, (1,), {}, None) is False          # extra pdf arguments: host callable
    assert DensityTree._wants_device_build(gmm, (), {}, "device") is True
    with pytest.raises(ValueError):
        DensityTree._wants_device_build(lambda x: x, (), {}, "device")
    with pytest.raises(ValueError):
        DensityTree._wants_device_build(gmm, (), {}, "gpu")
    assert DensityTree._wants_device_build(gmm, (), {}, None) is torch.cuda.is_available()
    if not torch.cuda.is_available():
        # the host builder takes over: same class, same API, the pdf evaluated by its NumPy __call__
        t = DensityTree([-4.0, -4.0], [4.0, 4.0], gmm, vectorizedpdf=True, min_openings=3)
        assert t._dt is None and len(t.leaves) == 64 and abs(t.total_mass - 1.0) < 0.2
        t.refine(1e-2, leaf_tol=5e-2)
        assert len(t.leaves) > 64
        with pytest.raises(RuntimeError, match="CUDA"):
            DensityTree([-4.0, -4.0], [4.0, 4.0], gmm, build="device")


def test_chunk_start_ownership_rule():
    """The rule by which the scans of the grouped path write chunk starts (scan_compact_kernel):
    a segment of rows [lo, hi) owns the chunks [ceil(lo/R), ceil(hi/R)) when the run starts on an even
    global row, and [(lo+R)//R, (hi+R)//R) — plus chunk 0 if lo = 0 — when it starts on an odd one
    (the first row of chunk t is then max(t R, 1) - 1).  Restated here and checked against the
    definition 'the segment that holds the chunk's first row' on random segment lists."""
    rng = np.random.default_rng(0)
    for R in (2048, 4096):
        for trial in range(20):
            counts = rng.integers(1, 3 * R, size=rng.integers(1, 60))
            if trial % 4 == 0:
                counts[rng.integers(len(counts))] = rng.integers(5 * R, 40 * R)      # one heavy segment
            off = np.concatenate([[0], np.cumsum(counts)])
            total = int(off[-1])
            for shift in (0, 1):
                nchunks = -(-(total + shift) // R)
                table = np.full(nchunks + 2, -1)
                for j, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
                    lo, hi = int(lo), int(hi)
                    if shift == 0:
                        table[-(-lo // R):-(-hi // R)] = j
                    else:
                        if lo == 0:
                            table[0] = j
                        table[(lo + R) // R:(hi + R) // R] = j
                for t in range(nchunks):
                    first = max(t * R, shift) - shift
                    owner = int(np.searchsorted(off, first, side="right")) - 1
                    assert table[t] == owner, (R, shift, t)
