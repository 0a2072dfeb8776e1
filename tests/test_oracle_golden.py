"""Pin the CPU oracle (oracle/lint_oracle.py) against golden vectors produced by
the real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, golden_edges
from oracle import lint_oracle as O

GRID_CASES = ["c1_readme_1d", "g1d_nonuniform", "g2d_fullcov", "g3d_iso", "g3d_holes",
              "g2d_flat", "g2d_onecell", "g4d_iso", "g5d_sobol", "g6d_iso", "g2d_sobol"]
TREE_CASES = ["t3d_open3", "t3d_refined", "t2d_refined", "t2d_batch", "t1d_refined"]
# in-cell coordinates: the oracle fixes an explicit reduction order, the
# reference uses NumPy multi-axis reduces; they agree to rounding
COORD_RTOL = 0.0


@pytest.mark.parametrize("name", GRID_CASES)
def test_grid_build_bit_exact(name):
    g = load_golden(name)
    edges = golden_edges(g)
    m = O.cell_masses(edges, g["vertex_densities"])
    assert np.array_equal(m, g["masses"])
    assert O.total_mass(m) == g["total_mass"]
    if m.size > 1:
        assert np.array_equal(O.cdf_from_masses(m), g["cdf"])


@pytest.mark.parametrize("name", GRID_CASES)
def test_grid_sample_matches_reference(name):
    g = load_golden(name)
    edges = golden_edges(g)
    x, flat = O.grid_sample(edges, g["vertex_densities"], g["u"], return_cells=True)
    assert np.array_equal(flat, g["flat"])                   # bit-exact cell indices
    assert np.array_equal(x, g["x_pre"])
    # the reference's own stream + final shuffle (lintsampler.py:311,517)
    if not int(g["qmc"]):
        u, rng = O.reference_uniforms(int(g["seed"]), int(g["N"]), int(g["k"]))
        assert np.array_equal(u, g["u"])
        rng.shuffle(x, axis=0)
        assert np.array_equal(x, g["x_final"])


@pytest.mark.parametrize("name", ["g5d_sobol", "g2d_sobol"])
def test_sobol_direct_index_bit_exact(name):
    g = load_golden(name)
    N = int(g["N"])
    u = O.sobol_direct(g["sobol_sv"], g["sobol_shift"], 0, N)
    assert np.array_equal(u, g["u"])
    # continuation from an offset reproduces the tail of the stream
    assert np.array_equal(O.sobol_direct(g["sobol_sv"], g["sobol_shift"], 1000, N - 1000),
                          g["u"][1000:])


@pytest.mark.parametrize("name", TREE_CASES)
def test_tree_leaf_sampling_matches_reference(name):
    g = load_golden(name)
    x, idx = O.cells_sample(g["leaf_x"], g["leaf_dx"], g["leaf_corners"], g["leaf_mass"],
                            g["u"], return_cells=True)
    assert np.array_equal(idx, g["leaf_choice"])
    assert np.array_equal(x, g["x_pre"])
    assert np.sum(g["leaf_mass"]) == g["total_mass"]


@pytest.mark.parametrize("name", ["m2d_quadrants", "m1d_halves"])
def test_multi_domain_matches_reference(name):
    import pdfs
    g = load_golden(name)
    k, N, nd = int(g["k"]), int(g["N"]), int(g["ndomains"])
    pdf = (pdfs.GMM(*pdfs.gmm_params(2, seed=1, full_cov=True)) if k == 2
           else pdfs.GMM(*pdfs.gmm_params(1, seed=3)))
    doms = []
    for i in range(nd):
        edges = [g[f"dom{i}_edges{d}"] for d in range(k)]
        mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1) if k > 1 else edges[0]
        doms.append((edges, pdf(mesh)))
    masses = [O.total_mass(O.cell_masses(e, V)) for e, V in doms]
    assert np.array_equal(np.array(masses), g["domain_masses"])
    u, rng = O.reference_uniforms(int(g["seed"]), N, k)
    x = O.multi_domain_sample(masses, lambda i, us: O.grid_sample(doms[i][0], doms[i][1], us), u, k)
    rng.shuffle(x, axis=0)
    assert np.array_equal(x, g["x_final"])


def test_incell_known_answers():
    g = load_golden("incell_kernels")
    for k in range(1, 7):
        z = O.unit_cube_sample(g[f"corners{k}"], g[f"u{k}"])
        # near-tie columns amplify rounding: compare those loosely
        np.testing.assert_allclose(z, g[f"z{k}"], rtol=1e-5, atol=1e-6)
        tight = np.ones(z.shape[0], dtype=bool)
        tight[3::11] = False
        np.testing.assert_allclose(z[tight], g[f"z{k}"][tight], rtol=1e-11, atol=1e-13)
    z = O.quantile_1d(g["q_f0"], g["q_f1"], g["q_u"])
    assert np.array_equal(z, g["q_z"])
    # worked example docs/source/theory/linear_interpolant.md:319-320: f0=6, f1=3
    assert abs(O.quantile_1d(np.array([6.0]), np.array([3.0]), np.array([0.3]))[0]
               - (2 - np.sqrt(4 - 3 * 0.3))) < 1e-15


def test_pairwise_sum_emulation_matches_numpy():
    rng = np.random.default_rng(0)
    for n in [1, 5, 8, 9, 127, 128, 129, 1000, 4097, 100003]:
        a = rng.random(n) * np.exp(rng.normal(size=n) * 5)
        assert O.numpy_pairwise_sum_emulated(a) == np.sum(a)
