"""Generate the golden fixtures in this directory by running the REAL reference
(aneeshnaik/lintsampler, mounted read-only at /root/reference) in the build
container.  The reference cannot travel to the GPU box, so its inputs/outputs
are committed as small .npz files together with this script.

Run from the repo root:   python tests/golden/make_golden.py

The reference's ``__init__`` asks importlib.metadata for its installed version
(src/lintsampler/__init__.py:15); a three-line stub ``.dist-info`` in a temp dir
satisfies that without installing or modifying anything.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True

_stub = tempfile.mkdtemp()
os.makedirs(os.path.join(_stub, "lintsampler-0.0.0+ref.dist-info"))
with open(os.path.join(_stub, "lintsampler-0.0.0+ref.dist-info", "METADATA"), "w") as fh:
    fh.write("Metadata-Version: 2.1\nName: lintsampler\nVersion: 0.0.0+ref\n")
sys.path[:0] = [_stub, "/root/reference/src"]

import lintsampler  # noqa: E402  (the reference)
from lintsampler import LintSampler, DensityGrid, DensityTree  # noqa: E402
from lintsampler.sampling import _grid_sample, _unitsample_kd, _unitsample_1d  # noqa: E402
from lintsampler.utils import _choice  # noqa: E402
from scipy.stats.qmc import Sobol  # noqa: E402

import pdfs  # noqa: E402  (tests/pdfs.py)


def ref_cdf(masses):
    """utils.py:195-196 applied to grid.py:431."""
    p = (masses / masses.sum()).flatten()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf


def grid_case(name, edges, pdf, N, seed, vectorized=True, qmc=False):
    """One DensityGrid domain: store everything the parity tests compare."""
    grid = DensityGrid(edges, pdf, vectorizedpdf=vectorized)
    k = grid.dim
    sampler = LintSampler(grid, seed=seed, qmc=qmc)
    if qmc:
        eng = Sobol(d=k + 1, bits=32, seed=seed)
        sv, shift = eng._sv.copy(), eng._shift.copy()
        u = eng.random(N)
        rng = np.random.default_rng(seed)
    else:
        sv = shift = np.zeros(0)
        rng = np.random.default_rng(seed)
        u = rng.random((N, k + 1))
    x_pre = _grid_sample(grid, u)
    cells = (grid._choose_indices(u[:, -1]) if grid.ncells > 1
             else np.zeros((N, k), dtype=np.int64))
    flat = np.ravel_multi_index(tuple(cells.T), grid.shape)
    x_final = x_pre.copy()
    rng.shuffle(x_final, axis=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x_api = sampler.sample(N)
    assert np.array_equal(np.atleast_1d(x_api).reshape(x_final.shape), x_final), name
    out = dict(
        k=k, N=N, seed=seed, qmc=int(qmc),
        vertex_densities=grid.vertex_densities, masses=grid.masses,
        total_mass=grid.total_mass,
        cdf=ref_cdf(grid.masses) if grid.ncells > 1 else np.ones(1),
        u=u, flat=flat.astype(np.int64), x_pre=x_pre, x_final=x_final,
        sobol_sv=sv, sobol_shift=shift,
    )
    for d, e in enumerate(grid.edgearrays):
        out[f"edges{d}"] = np.asarray(e, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: k={k} cells={grid.ncells} N={N}")


def leaf_table(tree):
    k = tree.dim
    L = len(tree.leaves)
    x = np.array([leaf.x for leaf in tree.leaves]).reshape(L, k)
    dx = np.array([leaf.dx for leaf in tree.leaves]).reshape(L, k)
    cd = np.array([leaf.corner_densities for leaf in tree.leaves]).reshape(L, 2 ** k)
    m = np.array([leaf.mass_raw for leaf in tree.leaves])
    mr = np.array([leaf.mass_romberg for leaf in tree.leaves])
    lev = np.array([leaf.level for leaf in tree.leaves], dtype=np.int64)
    idx = np.array([leaf.idx for leaf in tree.leaves], dtype=np.int64)
    return x, dx, cd, m, mr, lev, idx


def tree_case(name, mins, maxs, pdf, N, seed, min_openings, refine=None,
              batch=False, usecache=True):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = DensityTree(mins, maxs, pdf, vectorizedpdf=True,
                           min_openings=min_openings, batch=batch, usecache=usecache)
    if refine is not None:
        tree.refine(**refine)
    k = tree.dim
    x, dx, cd, m, mr, lev, idx = leaf_table(tree)
    rng = np.random.default_rng(seed)
    u = rng.random((N, k + 1))
    x_pre = _grid_sample(tree, u)
    leaf_choice = _choice(m / np.sum(m), u[:, -1])
    x_final = x_pre.copy()
    rng.shuffle(x_final, axis=0)
    x_api = LintSampler(tree, seed=seed).sample(N)
    assert np.array_equal(np.atleast_1d(x_api).reshape(x_final.shape), x_final), name
    ref = dict(refine) if refine else {}
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        k=k, N=N, seed=seed, mins=np.atleast_1d(mins).astype(float),
        maxs=np.atleast_1d(maxs).astype(float),
        min_openings=min_openings, batch=int(batch), usecache=int(usecache),
        refined=int(refine is not None),
        tree_tol=ref.get("tree_tol", np.nan), leaf_tol=ref.get("leaf_tol", 0.01),
        leaf_mass_contribution_tol=ref.get("leaf_mass_contribution_tol", 0.01),
        leaf_x=x, leaf_dx=dx, leaf_corners=cd, leaf_mass=m, leaf_mass_romberg=mr,
        leaf_level=lev, leaf_idx=idx, total_mass=tree.total_mass,
        u=u, leaf_choice=leaf_choice.astype(np.int64), x_pre=x_pre, x_final=x_final,
    )
    print(f"{name}: k={k} leaves={len(m)} levels={lev.min()}..{lev.max()} N={N}")


def multi_case(name, domains, pdf, N, seed, qmc=False):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sampler = LintSampler(domains, pdf=pdf, vectorizedpdf=True, seed=seed, qmc=qmc)
        x_api = sampler.sample(N)
    k = sampler.dim
    out = dict(k=k, N=N, seed=seed, qmc=int(qmc), ndomains=len(domains),
               x_final=np.atleast_1d(x_api).reshape(N, k),
               domain_masses=np.array([g.total_mass for g in sampler.grids]))
    for i, dom in enumerate(domains):
        dom = dom if isinstance(dom, tuple) else (dom,)
        for d, e in enumerate(dom):
            out[f"dom{i}_edges{d}"] = np.asarray(e, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: k={k} domains={len(domains)} N={N}")


def kernels_case():
    """Known-answer vectors for the in-cell routines alone (sampling.py:6-94),
    including the ill-conditioned f0 ~ f1 cases of SURVEY.md §7.3-3 and the
    worked example of docs/source/theory/linear_interpolant.md:319-320."""
    rng = np.random.default_rng(123)
    out = {}
    for k in (1, 2, 3, 4, 5, 6):
        n = 512
        corners = rng.random((2 ** k, n)) * np.exp(rng.normal(size=(1, n)) * 3)
        # sprinkle exact ties and near-ties
        corners[:, ::7] = corners[0, ::7]
        corners[1, 3::11] = corners[0, 3::11] * (1 + 1e-10)
        u = rng.random((n, k))
        out[f"corners{k}"] = corners
        out[f"u{k}"] = u
        out[f"z{k}"] = _unitsample_kd(*corners, u=u)
    f0 = np.array([6.0, 1.0, 0.0, 3.0, 2.0, 1e-30])
    f1 = np.array([3.0, 1.0, 5.0, 0.0, 2.0 * (1 + 1e-13), 2e-30])
    uu = np.array([0.3, 0.9, 0.25, 0.75, 0.5, 0.5])
    out["q_f0"], out["q_f1"], out["q_u"] = f0, f1, uu
    out["q_z"] = _unitsample_1d(f0, f1, uu)
    np.savez_compressed(os.path.join(HERE, "incell_kernels.npz"), **out)
    print("incell_kernels")


def main():
    # C1: README quickstart (README.md:42-55), non-vectorised scalar pdf
    grid_case("c1_readme_1d", np.linspace(-7, 7, 100), pdfs.readme_gmm_1d,
              N=10000, seed=42, vectorized=False)
    # 1-D, non-uniform edges
    e = np.sort(np.concatenate([[-6.0, 6.0], np.random.default_rng(5).uniform(-6, 6, 61)]))
    grid_case("g1d_nonuniform", e, pdfs.GMM(*pdfs.gmm_params(1, seed=3)), N=2048, seed=7)
    # 2-D full-covariance mixture, non-uniform edges (C2 shape in miniature)
    r = np.random.default_rng(11)
    ex = np.sort(np.concatenate([[-4.0, 4.0], r.uniform(-4, 4, 39)]))
    ey = np.linspace(-4, 4, 29)
    grid_case("g2d_fullcov", (ex, ey), pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)),
              N=4096, seed=1)
    # 3-D isotropic mixture (C3 shape in miniature)
    grid_case("g3d_iso", tuple(np.linspace(-4, 4, n) for n in (17, 13, 11)),
              pdfs.GMM(*pdfs.gmm_params(3, seed=0)), N=4096, seed=1)
    # 3-D with exactly-zero regions (flat CDF runs) and u == 0 rows
    grid_case("g3d_holes", tuple(np.linspace(-3, 3, n) for n in (13, 12, 11)),
              pdfs.bump_with_holes, N=2048, seed=9)
    # flat density: every 1-D solve hits the f0 == f1 branch
    grid_case("g2d_flat", (np.linspace(0, 1, 6), np.linspace(2, 5, 4)), pdfs.flat_pdf,
              N=1024, seed=2)
    # single-cell grid (grid.py:427-428 shortcut)
    grid_case("g2d_onecell", (np.array([0.0, 2.0]), np.array([-1.0, 1.0])),
              pdfs.GMM(*pdfs.gmm_params(2, seed=4)), N=512, seed=3)
    # 4-D, 5-D (C4 shape in miniature, Sobol) and 6-D
    grid_case("g4d_iso", tuple(np.linspace(-4, 4, n) for n in (6, 5, 7, 4)),
              pdfs.GMM(*pdfs.gmm_params(4, seed=2, smin=0.5, smax=0.9)), N=2048, seed=5)
    grid_case("g5d_sobol", tuple(np.linspace(-4, 4, n) for n in (6, 5, 4, 5, 4)),
              pdfs.GMM(*pdfs.gmm_params(5, seed=0, smin=0.5, smax=0.9)), N=2048, seed=1,
              qmc=True)
    grid_case("g6d_iso", tuple(np.linspace(-4, 4, n) for n in (4, 3, 4, 3, 3, 4)),
              pdfs.GMM(*pdfs.gmm_params(6, seed=6, smin=0.7, smax=1.1)), N=1024, seed=8)
    # 2-D Sobol (power-of-two N)
    grid_case("g2d_sobol", (np.linspace(-4, 4, 33), np.linspace(-4, 4, 17)),
              pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)), N=4096, seed=3, qmc=True)

    # trees: full openings only; refined; batch mode; 1-D and 2-D
    tree_case("t3d_open3", [-4.0] * 3, [4.0] * 3, pdfs.GMM(*pdfs.gmm_params(3, seed=0)),
              N=2048, seed=1, min_openings=3)
    tree_case("t3d_refined", [-4.0] * 3, [4.0] * 3,
              pdfs.GMM(*pdfs.gmm_params(3, K=8, seed=0, smin=0.1, smax=0.4)),
              N=4096, seed=1, min_openings=3, refine=dict(tree_tol=2e-2, leaf_tol=2e-2))
    tree_case("t2d_refined", [-10.0, -5.0], [10.0, 5.0],
              pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)),
              N=2048, seed=4, min_openings=2, refine=dict(tree_tol=1e-3))
    tree_case("t2d_batch", [-4.0, -4.0], [4.0, 4.0], pdfs.GMM(*pdfs.gmm_params(2, seed=2)),
              N=1024, seed=6, min_openings=3, refine=dict(tree_tol=5e-3), batch=True,
              usecache=False)
    tree_case("t1d_refined", -8.0, 8.0, pdfs.GMM(*pdfs.gmm_params(1, seed=3)),
              N=1024, seed=2, min_openings=2, refine=dict(tree_tol=1e-4, leaf_tol=1e-3))

    # list of domains (lintsampler.py:281-307): four quadrants of one mixture
    x0, x1 = np.linspace(-4, 0, 17), np.linspace(0, 4, 21)
    y0, y1 = np.linspace(-4, 0, 13), np.linspace(0, 4, 9)
    pdf2 = pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True))
    multi_case("m2d_quadrants", [(x0, y0), (x0, y1), (x1, y0), (x1, y1)], pdf2, N=4096, seed=5)
    multi_case("m2d_quadrants_sobol", [(x0, y0), (x0, y1), (x1, y0), (x1, y1)], pdf2,
               N=4096, seed=5, qmc=True)
    multi_case("m1d_halves", [np.linspace(-6, 0, 33), np.linspace(0, 6, 41)],
               pdfs.GMM(*pdfs.gmm_params(1, seed=3)), N=2048, seed=11)

    kernels_case()
    print("reference version:", lintsampler.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
