"""GPU parity tests, kernel level: every C-ABI entry point against the CPU oracle
and the golden vectors produced by the real reference.  Bit-exact unless a
tolerance is written in the test."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, golden_edges
from oracle import lint_oracle as O

pytestmark = pytest.mark.gpu

GRID_CASES = ["c1_readme_1d", "g1d_nonuniform", "g2d_fullcov", "g3d_iso", "g3d_holes",
              "g2d_flat", "g2d_onecell", "g4d_iso", "g5d_sobol", "g6d_iso", "g2d_sobol"]
TREE_CASES = ["t3d_open3", "t3d_refined", "t2d_refined", "t2d_batch", "t1d_refined"]


@pytest.fixture(scope="module")
def lb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lintsampler_b200
    return lintsampler_b200


def _grid_from_golden(lb, g, **kw):
    """DensityGrid over the golden vertex densities (pdf = table lookup)."""
    edges = golden_edges(g)
    V = g["vertex_densities"]
    dom = tuple(edges) if len(edges) > 1 else edges[0]
    return lb.DensityGrid(dom, lambda pts: V.reshape(-1), vectorizedpdf=True, **kw)


def _dev(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


# ----------------------------------------------------------------------------- build
@pytest.mark.parametrize("name", GRID_CASES)
def test_masses_total_cdf_bit_exact(lb, name):
    g = load_golden(name)
    grid = _grid_from_golden(lb, g, cdf_mode="sequential")
    assert np.array_equal(grid.masses, g["masses"])
    assert grid.total_mass == g["total_mass"]
    assert np.array_equal(grid._cdf.cpu().numpy(), g["cdf"])


@pytest.mark.parametrize("n", [1, 5, 8, 9, 127, 128, 129, 255, 256, 257, 1000, 4097, 65536,
                               100003, 1 << 20, (1 << 22) + 12345])
def test_sum_numpy_bit_exact(lb, n):
    import torch
    from lintsampler_b200 import _capi, _device
    rng = np.random.default_rng(n)
    a = rng.random(n) * np.exp(rng.normal(size=n) * 5)
    x = _dev(a)
    nbytes = _capi.load().lint_cdf_scratch_bytes(n)
    scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    _capi.call("lint_sum_numpy_f64", _device.ptr(x), n, _device.ptr(out), _device.ptr(scratch), nbytes,
               _device.stream_ptr(x.device))
    assert out.item() == np.sum(a)
    if n <= 4097:
        assert out.item() == O.numpy_pairwise_sum_emulated(a)


@pytest.mark.parametrize("n", [1, 2, 31, 2047, 2048, 2049, 70000, 1 << 21])
@pytest.mark.parametrize("mode", ["sequential", "parallel"])
def test_cdf_modes(lb, n, mode):
    from lintsampler_b200 import _device
    rng = np.random.default_rng(n)
    m = rng.random(n) * np.exp(rng.normal(size=n) * 4)
    m[rng.random(n) < 0.3] = 0.0                       # zero-mass cells: flat CDF runs
    if m.sum() == 0:
        m[0] = 1.0
    cdf, guide, bits, total = _device.build_cdf(_dev(m), mode, _dev(m).device)
    cdf = cdf.cpu().numpy()
    ref = O.cdf_from_masses(m)
    if mode == "sequential":
        assert np.array_equal(cdf, ref)
        assert total == np.sum(m)
    else:
        # tiled scan rounds differently from a sequential one; tolerance 1e-12
        assert np.max(np.abs(cdf - ref)) < 1e-12
        assert abs(total / np.sum(m) - 1) < 1e-13
    assert cdf[-1] == 1.0
    assert np.all(np.diff(cdf) >= 0)
    zero = np.flatnonzero(m == 0)
    zero = zero[zero > 0]
    assert np.array_equal(cdf[zero], cdf[zero - 1])    # zero-mass cells stay unselectable
    # guide table == searchsorted of the same table at j / G
    G = 1 << bits
    raw = guide.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    gt, same = raw & 0x7FFFFFFF, (raw >> 31).astype(bool)
    want = np.minimum(cdf.searchsorted(np.arange(G + 1) / G, side="right"), n - 1)
    assert np.array_equal(gt, want)
    assert np.array_equal(same[:-1], want[:-1] == want[1:])   # bit 31: bin wholly inside one cell


def test_parallel_cdf_is_deterministic(lb):
    from lintsampler_b200 import _device
    rng = np.random.default_rng(0)
    m = _dev(rng.random(3_000_001))
    a = _device.build_cdf(m, "parallel", m.device)[0].cpu().numpy()
    for _ in range(3):
        assert np.array_equal(a, _device.build_cdf(m, "parallel", m.device)[0].cpu().numpy())


def test_density_validity_flags(lb):
    e = (np.linspace(0, 1, 9), np.linspace(0, 1, 7))
    import torch

    def neg(p):
        f = torch.ones(p.shape[0], dtype=torch.float64, device=p.device)
        f[5] = -1.0
        return f

    def nan(p):
        f = torch.ones(p.shape[0], dtype=torch.float64, device=p.device)
        f[-1] = float("nan")
        return f

    with pytest.raises(ValueError):
        lb.DensityGrid(e, neg, vectorizedpdf=True, pdf_on_device=True)
    with pytest.raises(ValueError):
        lb.DensityGrid(e, nan, vectorizedpdf=True, pdf_on_device=True)
    with pytest.raises(ValueError):          # all-zero density: documented deviation
        lb.DensityGrid(e, lambda p: np.zeros(len(p)), vectorizedpdf=True)


@pytest.mark.parametrize("k,full", [(1, False), (2, True), (3, False), (3, True), (5, False), (6, True)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_gmm_vertex_eval(lb, k, full, dtype):
    import pdfs
    w, mu, cov = pdfs.gmm_params(k, K=5, seed=k, full_cov=full, smin=0.5, smax=0.9)
    rng = np.random.default_rng(k)
    edges = tuple(np.sort(rng.uniform(-4, 4, n)) for n in (9, 7, 8, 5, 6, 4)[:k])
    dom = edges if k > 1 else edges[0]
    grid = lb.DensityGrid(dom, lb.GaussianMixture(w, mu, cov), vectorizedpdf=True, dtype=dtype)
    mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1)
    ref = O.gmm_pdf(mesh if k > 1 else edges[0], w, mu, cov)
    # exp() differs from NumPy's by <= 1 ulp: tolerance 1e-12 (fp64) / fp32 rounding 1e-6
    if dtype == np.float64:
        np.testing.assert_allclose(grid.vertex_densities, ref, rtol=1e-12)
    else:       # below fp32's normal range (1.2e-38) values are denormal: absolute floor
        np.testing.assert_allclose(grid.vertex_densities, ref, rtol=1e-6, atol=1e-36)


# ----------------------------------------------------------------------------- sampling, fed the reference's u
@pytest.mark.parametrize("name", GRID_CASES)
def test_grid_sample_u_bit_exact(lb, name):
    g = load_golden(name)
    grid = _grid_from_golden(lb, g, cdf_mode="sequential")
    x, cells = grid._sample_u(_dev(g["u"]), want_cells=True)
    assert np.array_equal(cells.cpu().numpy(), g["flat"])
    assert np.array_equal(x.cpu().numpy(), g["x_pre"])


@pytest.mark.parametrize("name", ["g2d_fullcov", "g3d_iso", "g3d_holes", "g5d_sobol"])
def test_grid_sample_u_parallel_cdf_mismatches_explained(lb, name):
    """With the GPU-built parallel CDF a pick may differ from the reference only
    when u lies between the two tables' values of a boundary between the picks
    (parity mode B); unexplained mismatches must be zero."""
    g = load_golden(name)
    grid = _grid_from_golden(lb, g, cdf_mode="parallel")
    cdf_gpu = grid._cdf.cpu().numpy()
    cdf_ref = g["cdf"]
    u = g["u"][:, -1]
    _, cells = grid._sample_u(_dev(g["u"]), want_cells=True)
    cells = cells.cpu().numpy()
    assert np.array_equal(cells, cdf_gpu.searchsorted(u, side="right"))   # exact on its own table
    bad = np.flatnonzero(cells != g["flat"])
    unexplained = 0
    for i in bad:
        lo, hi = sorted((cells[i], g["flat"][i]))
        b = np.arange(lo, hi)
        inside = (np.minimum(cdf_gpu[b], cdf_ref[b]) <= u[i]) & (u[i] < np.maximum(cdf_gpu[b], cdf_ref[b]))
        unexplained += not inside.any()
    assert unexplained == 0
    assert len(bad) <= max(1, len(u) // 1000)


@pytest.mark.parametrize("name", ["g1d_nonuniform", "g2d_fullcov", "g3d_iso", "g4d_iso", "g5d_sobol", "g6d_iso"])
def test_grid_sample_u_fp32(lb, name):
    """fp32 mode: densities stored in fp32, fp64 CDF/cell uniform.  Cell picks are
    exact against the oracle run on the same fp32-rounded densities; coordinates
    agree to 1e-4 of the cell width (north_star tolerance for fp32)."""
    g = load_golden(name)
    edges = golden_edges(g)
    V32 = g["vertex_densities"].astype(np.float32)
    dom = tuple(edges) if len(edges) > 1 else edges[0]
    grid = lb.DensityGrid(dom, lambda p: V32.reshape(-1).astype(np.float64), vectorizedpdf=True,
                          dtype=np.float32, cdf_mode="sequential")
    x_ref, flat_ref = O.grid_sample(edges, V32.astype(np.float64), g["u"], return_cells=True)
    x, cells = grid._sample_u(_dev(g["u"]), want_cells=True)
    assert x.dtype.is_floating_point and x.element_size() == 4
    assert np.array_equal(cells.cpu().numpy(), flat_ref)
    multi = np.unravel_index(flat_ref, tuple(len(e) - 1 for e in edges))
    width = np.stack([np.diff(edges[d])[multi[d]] for d in range(len(edges))], axis=-1)
    err = np.abs(x.cpu().numpy().astype(np.float64) - x_ref) / width
    assert err.max() < 1e-4, err.max()


@pytest.mark.parametrize("name", TREE_CASES)
def test_cells_sample_u_bit_exact(lb, name):
    from lintsampler_b200.tree import _CellTable
    g = load_golden(name)
    table = _CellTable(g["leaf_x"], g["leaf_dx"], g["leaf_corners"], g["leaf_mass"], np.float64,
                       None, "sequential")
    assert table.total_mass == g["total_mass"]
    x, cells = table.sample_u(_dev(g["u"]), want_cells=True)
    assert np.array_equal(cells.cpu().numpy(), g["leaf_choice"])
    assert np.array_equal(x.cpu().numpy(), g["x_pre"])


def test_incell_kernel_known_answers(lb):
    import torch
    from lintsampler_b200 import _capi, _device
    g = load_golden("incell_kernels")
    for k in range(1, 7):
        corners, u = g[f"corners{k}"], g[f"u{k}"]
        n = u.shape[0]
        ufull = np.concatenate([u, np.zeros((n, 1))], axis=1)
        out = torch.empty((n, k), dtype=torch.float64, device="cuda")
        lo, hi, cd, ud = _dev(np.zeros((n, k))), _dev(np.ones((n, k))), _dev(corners), _dev(ufull)
        _capi.call("lint_incell_sample_u", k, n, _device.ptr(lo), _device.ptr(hi), _device.ptr(cd),
                   _device.ptr(ud), _device.ptr(out), _device.stream_ptr(out.device))
        # mins=0, maxs=1: x = 0 + (1-0)*z = z exactly
        assert np.array_equal(out.cpu().numpy(), g[f"z{k}"])


# ----------------------------------------------------------------------------- device uniform streams
def _philox4x32_10(ctr, key):
    """NumPy Philox4x32-10 (Salmon et al. 2011) — test-side restatement of the
    stream definition in include/lintsampler_b200.h."""
    c = [np.asarray(x, dtype=np.uint64) for x in ctr]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask,
             ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def test_philox_known_answer_vectors():
    # Random123 kat_vectors, philox4x32-10
    r = _philox4x32_10([0, 0, 0, 0], [0, 0])
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = _philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = _philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _philox_rows(k, dtype, N, seed, offset):
    i = np.arange(offset, offset + N, dtype=np.uint64)
    lo, hi = i & np.uint64(0xFFFFFFFF), i >> np.uint64(32)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    z = np.zeros(N, dtype=np.uint64)

    def u53(a, b):
        return ((((b << np.uint64(32)) | a) >> np.uint64(11)).astype(np.float64)) * 2.0 ** -53
    u = np.empty((N, k + 1))
    if dtype == np.float64:
        w = []
        for q in range((k + 2) // 2):
            w += _philox4x32_10([lo, hi, z + np.uint64(q), z], key)
        u[:, k] = u53(w[0], w[1])
        for d in range(k):
            u[:, d] = u53(w[2 + 2 * d], w[3 + 2 * d])
    else:
        r = _philox4x32_10([lo, hi, z, z], key)
        s = _philox4x32_10([lo, hi, z + np.uint64(1), z], key)
        u[:, k] = u53(r[0], r[1])
        S = 2.0 ** -24
        cols = [(r[2] >> np.uint64(8)), (r[3] >> np.uint64(8)),
                ((r[2] & np.uint64(255)) << np.uint64(16)) | ((r[3] & np.uint64(255)) << np.uint64(8))
                | (r[0] & np.uint64(255)),
                s[0] >> np.uint64(8), s[1] >> np.uint64(8), s[2] >> np.uint64(8)]
        for d in range(k):
            u[:, d] = cols[d].astype(np.float64) * S
    return u


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_philox_uniform_stream_definition(lb, k, dtype):
    import torch
    from lintsampler_b200 import _capi, _device
    N, seed, offset = 4099, 0x1234567890ABCDEF, (1 << 32) - 1000     # counter carries into the high word
    u = torch.empty((N, k + 1), dtype=torch.float64, device="cuda")
    _capi.call("lint_philox_uniforms", k, _device.lint_dtype(dtype), N, C.c_uint64(seed),
               C.c_uint64(offset), _device.ptr(u), _device.stream_ptr(u.device))
    assert np.array_equal(u.cpu().numpy(), _philox_rows(k, dtype, N, seed, offset))


@pytest.mark.parametrize("name", ["g1d_nonuniform", "g2d_fullcov", "g3d_iso", "g4d_iso", "g6d_iso"])
def test_philox_sampler_equals_oracle_on_its_own_uniforms(lb, name):
    import torch
    from lintsampler_b200 import _capi, _device
    g = load_golden(name)
    k = int(g["k"])
    grid = _grid_from_golden(lb, g, cdf_mode="sequential")
    N, seed, offset = 50000, 77, 123456789
    x, cells = grid._sample_philox(N, seed, offset, want_cells=True)
    u = torch.empty((N, k + 1), dtype=torch.float64, device="cuda")
    _capi.call("lint_philox_uniforms", k, _capi.LINT_F64, N, C.c_uint64(seed), C.c_uint64(offset),
               _device.ptr(u), _device.stream_ptr(u.device))
    x_ref, flat_ref = O.grid_sample(golden_edges(g), g["vertex_densities"], u.cpu().numpy(),
                                    cdf=g["cdf"], return_cells=True)
    assert np.array_equal(cells.cpu().numpy(), flat_ref)
    assert np.array_equal(x.cpu().numpy(), x_ref)
    # stream continuation: two half calls == one call
    xa = grid._sample_philox(N // 2, seed, offset)
    xb = grid._sample_philox(N - N // 2, seed, offset + N // 2)
    assert torch.equal(torch.cat([xa, xb]), x)


@pytest.mark.parametrize("name", ["g5d_sobol", "g2d_sobol"])
def test_sobol_stream_and_sampler_bit_exact(lb, name):
    import torch
    from lintsampler_b200 import _capi, _device
    g = load_golden(name)
    k, N = int(g["k"]), int(g["N"])
    s = _capi.LintSobol()
    for j in range(k + 1):
        for b in range(32):
            s.sv[j][b] = int(g["sobol_sv"][j, b])
        s.shift[j] = int(g["sobol_shift"][j])
    s.first_index = 0
    u = torch.empty((N, k + 1), dtype=torch.float64, device="cuda")
    _capi.call("lint_sobol_uniforms", k, N, C.byref(s), _device.ptr(u), _device.stream_ptr(u.device))
    assert np.array_equal(u.cpu().numpy(), g["u"])          # == scipy Sobol.random(N)
    grid = _grid_from_golden(lb, g, cdf_mode="sequential")
    x, cells = grid._sample_sobol(N, s, want_cells=True)
    assert np.array_equal(cells.cpu().numpy(), g["flat"])
    assert np.array_equal(x.cpu().numpy(), g["x_pre"])
    s.first_index = 1500                                     # continuation
    x2 = grid._sample_sobol(N - 1500, s)
    assert np.array_equal(x2.cpu().numpy(), g["x_pre"][1500:])


@pytest.mark.parametrize("name", ["g2d_fullcov", "g3d_holes"])
def test_grid_choose_cells_contract_equals_oracle(lb, name):
    """DensityGrid.choose_cells returns the reference's triple (mins, maxs, corners)
    [grid.py:325-356] — compared with the oracle's restatement element for element."""
    g = load_golden(name)
    edges = golden_edges(g)
    grid = _grid_from_golden(lb, g, cdf_mode="sequential")
    u = np.random.default_rng(3).random(5000)
    u[:3] = [0.0, 0.5, 1.0 - 2.0 ** -53]
    mins, maxs, corners = grid.choose_cells(u)
    _, rmins, rmaxs, rcorners = O.grid_choose_cells(edges, g["vertex_densities"], u)
    assert np.array_equal(mins, rmins) and np.array_equal(maxs, rmaxs)
    assert len(corners) == len(rcorners) == 2 ** len(edges)
    for a, b in zip(corners, rcorners):
        assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("name", ["g1d_nonuniform", "g2d_fullcov", "g3d_iso", "g3d_holes", "g4d_iso", "g6d_iso"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_hierarchical_grid_build(lb, name, dtype):
    """lint_grid_build (group totals from the vertex table, exact scan inside each group, masses
    never stored): the CDF is non-decreasing, repeats its predecessor exactly on zero-mass cells,
    ends at 1.0, lies within 1e-13 of the NumPy-sequential table of the same masses, and its total
    agrees with np.sum(masses) to rounding; a build restricted to a stratum writes exactly the
    same values on its slab, which contains every cell the stratum touches."""
    import torch
    g = load_golden(name)
    grid = _grid_from_golden(lb, g, dtype=dtype)                  # hierarchical build, whole domain
    n = int(grid.ncells)
    full = grid._cdf.cpu().numpy()
    m = grid._masses.cpu().numpy()                                 # lint_grid_masses (bit-exact vs the oracle elsewhere)
    assert np.all(np.diff(full) >= 0) and full[-1] == 1.0 and full[0] >= 0
    prev = np.concatenate([[0.0], full[:-1]])
    assert np.all(full[m == 0] == prev[m == 0])                    # zero-mass cells stay unselectable
    ref = O.cdf_from_masses(m)
    assert np.max(np.abs(full - ref)) < 1e-13
    assert abs(float(grid._total.item()) - m.sum()) <= 1e-13 * m.sum()
    assert grid._range.tolist()[:2] == [0, n]
    again = _grid_from_golden(lb, g, dtype=dtype)
    assert torch.equal(again._cdf, grid._cdf) and torch.equal(again._guide, grid._guide)     # deterministic
    for lo, hi in [(0.0, 0.3), (0.3, 0.55), (0.55, 1.0), (0.999, 1.0)]:
        part = _grid_from_golden(lb, g, dtype=dtype)
        part._cdf.fill_(-1.0)
        part._enqueue_build(stratum=(lo, hi))
        c0, c1, g0, g1 = part._range.tolist()
        got = part._cdf.cpu().numpy()
        assert 0 <= c0 < c1 <= n and g0 < g1
        assert np.array_equal(got[c0:c1], full[c0:c1])
        assert np.all(got[:max(c0 - 1, 0)] == -1.0) and np.all(got[c1:] == -1.0)      # nothing else was touched
        if c0 > 0:
            assert abs(got[c0 - 1] - full[c0 - 1]) < 1e-15         # the entry before the slab: the group's offset
        touched = np.flatnonzero((full > lo) & (prev < hi))       # cells whose CDF interval meets the stratum
        assert touched.min() >= c0 and touched.max() < c1
        assert float(part._total.item()) == float(grid._total.item())
        # draws from the stratum are the same rows whichever build made the tables
        N = 20_000
        a = grid._sample_philox(N, 5, 0, stratum=(lo, hi))
        b = part._sample_philox(N, 5, 0, stratum=(lo, hi))
        assert torch.equal(a, b)
        a = grid._sample_cellmajor(N, 5, 0, stratum=(lo, hi))
        b = part._sample_cellmajor(N, 5, 0, stratum=(lo, hi))
        assert torch.equal(a, b)
        with pytest.raises(ValueError):
            part._sample_philox(10, 5, 0, stratum=(0.0, 1.0))


@pytest.mark.parametrize("shape", [(1024,), (5, 512), (3, 4, 256), (2, 3, 2, 256)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_hierarchical_build_masses_on_wide_rows(lb, shape, dtype):
    """Grids whose last dimension is a multiple of the scan's 256 threads (every round of a CTA
    lies in one cell row — the benchmark's shape): the masses lint_grid_build computes on the way
    must be the very masses of the per-cell kernel (lint_grid_masses, pinned bit for bit to the
    oracle elsewhere), and the CDF the NumPy-sequential table of those masses to 1e-13 — on
    non-uniform edges, so that every width matters."""
    import torch
    import ctypes as C
    from lintsampler_b200 import _capi, _device
    rng = np.random.default_rng(len(shape))
    edges = tuple(np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, n))]) for n in shape)
    dens = rng.uniform(0.0, 2.0, tuple(n + 1 for n in shape))
    dens[tuple(slice(0, 2) for _ in shape)] = 0.0                  # a zero-mass corner of the domain
    grid = lb.DensityGrid.from_densities(edges if len(shape) > 1 else edges[0], dens, dtype=dtype)
    n = int(grid.ncells)
    ref = grid._masses.clone()                                     # lint_grid_masses
    got = torch.empty(n, dtype=torch.float64, device="cuda")
    cdf = torch.empty(n, dtype=torch.float64, device="cuda")
    desc = grid._descriptor()
    _capi.call("lint_grid_build", C.byref(desc), C.c_double(0.0), C.c_double(1.0), _device.ptr(cdf),
               _device.ptr(grid._total), _device.ptr(grid._guide), 0, _device.ptr(got), _device.ptr(grid._flags),
               _device.ptr(grid._range), _device.ptr(grid._scratch), grid._scratch.numel(),
               _device.stream_ptr(grid.device))
    assert torch.equal(got, ref)
    assert torch.equal(cdf, grid._cdf)
    m = ref.cpu().numpy()
    V = dens.astype(dtype).astype(np.float64) * grid._scale
    assert np.array_equal(m, O.cell_masses(edges, V).ravel())
    assert np.max(np.abs(cdf.cpu().numpy() - O.cdf_from_masses(m))) < 1e-13


@pytest.mark.parametrize("name", ["g2d_fullcov", "g3d_holes", "g1d_nonuniform"])
def test_grid_without_a_guide_table(lb, name):
    """guide_bits=0: no guide table is built; the u-driven kernels search the whole CDF (same cells
    as NumPy's searchsorted), the grouped draw of a stratum-restricted build brackets its top-up
    rows by the slab the build reports, and u-driven draws from such a build are refused."""
    import torch
    g = load_golden(name)
    grid = _grid_from_golden(lb, g, guide_bits=0)
    ref = _grid_from_golden(lb, g)
    assert grid._guide_bits == 0 and torch.equal(grid._cdf, ref._cdf)
    k = len(grid.shape)
    u = np.random.default_rng(3).random((5000, k + 1))
    xa, a = grid._sample_u(torch.as_tensor(u, device="cuda"), want_cells=True)
    xb, b = ref._sample_u(torch.as_tensor(u, device="cuda"), want_cells=True)
    assert torch.equal(a, b) and torch.equal(xa, xb)
    assert np.array_equal(a.cpu().numpy(), np.searchsorted(grid._cdf.cpu().numpy(), u[:, -1], side="right"))
    # grouped draw on a stratum: rows (top-up rows included) stay in cells the stratum touches
    grid._enqueue_build(stratum=(0.25, 0.75))
    cdf = ref._cdf.cpu().numpy()
    c_lo = int(np.searchsorted(cdf, 0.25, side="right"))          # first / last cell the stratum touches
    c_hi = int(np.searchsorted(cdf, 0.75, side="left"))
    N = 200_000
    x, cells = grid._sample_cellmajor(N, 11, 0, want_cells=True, stratum=(0.25, 0.75))
    cells = cells.cpu().numpy()
    assert cells.min() >= c_lo and cells.max() <= c_hi
    # (the top-up tail is non-empty at this N: 8 sqrt(N) rows are held back)
    counts = np.bincount(cells, minlength=len(cdf)).astype(np.float64)
    p = np.diff(np.concatenate([[0.25], np.clip(cdf, 0.25, 0.75)])) / 0.5     # share of the stratum in each cell
    big = p * N >= 20
    if big.sum() > 1:
        from scipy import stats
        obs = np.append(counts[big], counts[~big].sum())
        exp = np.append(p[big], p[~big].sum()) * N
        keep = exp > 0
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4
    with pytest.raises(ValueError):
        grid._sample_philox(10, 5, 0, stratum=(0.3, 0.6))


@pytest.mark.parametrize("name", ["g2d_fullcov", "g3d_iso", "g3d_holes", "g4d_iso"])
@pytest.mark.parametrize("bits", [1, 3, 5])
def test_small_guide_tables_by_bin_search_equal_the_cell_walk(lb, name, bits):
    """A guide table with far fewer bins than cells is built by one binary search per bin
    (lint_grid_build); it must be the very table the cell-walking lint_guide_build writes, on the
    whole domain and, restricted to a stratum, on every bin a cell of the slab owns."""
    import torch
    from lintsampler_b200 import _capi, _device
    g = load_golden(name)
    grid = _grid_from_golden(lb, g, guide_bits=bits)
    n = int(grid.ncells)
    if (1 << bits) * 16 > n:
        pytest.skip("this grid is too small for the bin-search path at this table size")
    walk = torch.empty_like(grid._guide)
    _capi.call("lint_guide_build", _device.ptr(grid._cdf), n, bits, _device.ptr(walk), _device.stream_ptr(grid.device))
    assert torch.equal(walk, grid._guide)
    cdf = grid._cdf.cpu().numpy()
    part = _grid_from_golden(lb, g, guide_bits=bits)
    part._guide.fill_(-7)
    part._enqueue_build(stratum=(0.3, 0.6))
    c0, c1 = part._range.tolist()[:2]
    got, ref = part._guide.cpu().numpy(), walk.cpu().numpy()
    owner = ref & 0x7fffffff
    mine = (owner >= c0) & (owner < c1)
    assert mine.any() and np.array_equal(got[mine], ref[mine]) and np.all(got[~mine] == -7)
    # and the u-driven sampler finds the same cells through either table
    u = np.random.default_rng(1).random((4000, len(grid.shape) + 1))
    u[:, -1] = 0.3 + 0.3 * u[:, -1]
    _, a = grid._sample_u(torch.as_tensor(u, device="cuda"), want_cells=True)
    _, b = part._sample_u(torch.as_tensor(u, device="cuda"), want_cells=True)
    assert torch.equal(a, b) and np.array_equal(a.cpu().numpy(), np.searchsorted(cdf, u[:, -1], side="right"))


# ----------------------------------------------------------------------------- cell-major throughput mode
def grouped_runs(cells, N):
    """Row layout of a cell-major draw: rows [0, T_f) flat parts grouped by cell, rows [T_f, T)
    the rest grouped by cell, rows [T, N) the u-driven top-up in draw order (at most
    16 sqrt(N) + 64 rows).  Returns (descents inside the grouped part, T)."""
    tail = int(16 * np.sqrt(N)) + 64
    down = np.flatnonzero(np.diff(cells) < 0)
    early = down[down < N - tail - 1]            # a descent this early is the flat -> rest boundary
    assert len(early) <= 1, early[:5]
    rest = down[len(early):]
    T = int(rest[0]) + 1 if len(rest) else N     # the grouped part ends at the next descent
    assert N - tail <= T <= N
    return len(early), T


@pytest.mark.parametrize("name,dtype", [
    ("g1d_nonuniform", np.float64), ("g2d_fullcov", np.float64), ("g3d_iso", np.float64),
    ("g3d_holes", np.float64), ("g4d_iso", np.float64), ("g6d_iso", np.float64),
    ("g1d_nonuniform", np.float32), ("g2d_fullcov", np.float32), ("g3d_iso", np.float32),
    ("g3d_holes", np.float32), ("g4d_iso", np.float32), ("g6d_iso", np.float32),
])
def test_cellmajor_structure_and_distribution(lb, name, dtype, monkeypatch):
    """Rows grouped by cell in C order (two runs: the flat parts, then the rest), every row inside its cell, counts sum to N and
    follow Multinomial(N, p) (chi^2), zero-mass cells never drawn, per-coordinate KS
    against oracle samples, determinism."""
    from scipy import stats
    g = load_golden(name)
    edges = golden_edges(g)
    k = int(g["k"])
    V = g["vertex_densities"]
    dom = tuple(edges) if k > 1 else edges[0]
    grid = lb.DensityGrid(dom, lambda p: V.reshape(-1), vectorizedpdf=True, dtype=dtype)
    N = 300_000
    x, cells = grid._sample_cellmajor(N, 11, 0, want_cells=True)
    x2, cells2 = grid._sample_cellmajor(N, 11, 0, want_cells=True)
    import torch
    assert torch.equal(x, x2) and torch.equal(cells, cells2)                 # deterministic
    x3 = grid._sample_cellmajor(N, 11, N)
    assert not torch.equal(x, x3)                                            # offset decorrelates calls
    x, cells = x.cpu().numpy().astype(np.float64), cells.cpu().numpy()
    assert x.shape == (N, k)
    _, T = grouped_runs(cells, N)                                            # grouped, C order, two runs
    shape = tuple(len(e) - 1 for e in edges)
    multi = np.unravel_index(cells, shape)
    tol = 1e-6 if dtype == np.float32 else 0.0
    for d in range(k):
        lo, hi = edges[d][multi[d]], edges[d][multi[d] + 1]
        assert np.all(x[:, d] >= lo - tol * (hi - lo)) and np.all(x[:, d] <= hi + tol * (hi - lo))
    p = (g["masses"] / g["masses"].sum()).ravel()
    counts = np.bincount(cells, minlength=p.size)
    assert counts.sum() == N and np.all(counts[p == 0] == 0)
    big = p * N >= 20
    obs = np.append(counts[big], counts[~big].sum())
    exp = np.append(p[big], p[~big].sum()) * N
    keep = exp > 0
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4
    u = np.random.default_rng(5).random((N, k + 1))
    if name == "g3d_holes":
        # On this grid's exactly symmetric cells g0 and g1 differ by rounding noise only and
        # the reference's literal quantile (sampling.py:37) collapses z onto the cell faces
        # (SURVEY §7.3-3).  The cell-major path draws most rows of such a cell from the flat
        # part of the density and the rest with the cancellation-free form of the quantile,
        # so it follows the interpolant itself: compare with that form.
        def conjugate_quantile(f0, f1, uu):
            f0, f1, uu = (np.asarray(a, dtype=np.float64) for a in (f0, f1, uu))
            den = f0 + np.sqrt(f0 * f0 + (f1 * f1 - f0 * f0) * uu)
            return np.where(den > 0, (f0 + f1) * uu / np.where(den > 0, den, 1.0), uu)
        monkeypatch.setattr(O, "quantile_1d", conjugate_quantile)
        # the fp64 top-up rows are the u-driven kernel's (literal formula): drop them
        x = x[:T]
    x_ref = O.grid_sample(edges, V, u)
    for d in range(k):
        assert stats.ks_2samp(x[:, d], x_ref[:, d]).pvalue > 1e-4


@pytest.mark.parametrize("k,dtype", [(1, np.float64), (2, np.float32), (3, np.float32), (3, np.float64),
                                     (4, np.float32)])
def test_cellmajor_flat_residual_split_keeps_the_incell_distribution(lb, k, dtype):
    """Few cells, a million rows each: every cell is split into its flat part (plain uniforms)
    and its residual (inverse CDF on corners - min).  The mixture must be the cell's
    interpolant: joint 2-D histograms inside each cell against the exact bilinear masses
    (chi^2), and per-coordinate KS against oracle samples."""
    from scipy import stats
    rng = np.random.default_rng(k)
    shape = (2,) * k
    edges = [np.array([0.0, 0.7, 2.0]) + d for d in range(k)]
    V = rng.uniform(0.2, 1.0, size=tuple(n + 1 for n in shape))
    V[(0,) * k] = 3.0                                  # one steep cell, the others nearly flat
    dom = tuple(edges) if k > 1 else edges[0]
    grid = lb.DensityGrid(dom, lambda p: V.reshape(-1), vectorizedpdf=True, dtype=dtype)
    N = 2_000_000
    x, cells = grid._sample_cellmajor(N, 3, 0, want_cells=True)
    x, cells = x.cpu().numpy().astype(np.float64).reshape(N, k), cells.cpu().numpy()
    u = np.random.default_rng(9).random((N, k + 1))
    x_ref = O.grid_sample(edges, V, u).reshape(N, k)
    for d in range(k):
        assert stats.ks_2samp(x[:, d], x_ref[:, d]).pvalue > 1e-4
    if k < 2:
        return
    # joint distribution of (z0, z_last) in the steep cell: 8 x 8 bins against the oracle's draws
    # of the same cell (a product of correct marginals would pass the KS tests above)
    ref_cells = np.ravel_multi_index(tuple(np.searchsorted(edges[d], x_ref[:, d], side="right") - 1
                                           for d in range(k)), shape)
    for c in (0, int(np.prod(shape)) - 1):
        m, mr = cells == c, ref_cells == c
        idx = np.unravel_index(c, shape)
        def bins(pts):
            z0 = (pts[:, 0] - edges[0][idx[0]]) / np.diff(edges[0])[idx[0]]
            z1 = (pts[:, -1] - edges[-1][idx[-1]]) / np.diff(edges[-1])[idx[-1]]
            return np.histogram2d(z0, z1, bins=8, range=[[0, 1], [0, 1]])[0].ravel()
        a, b = bins(x[m]), bins(x_ref[mr])
        assert a.sum() > 20000 and b.sum() > 20000
        _, p, _, _ = stats.chi2_contingency(np.stack([a, b]))
        assert p > 1e-4


def test_cellmajor_counts_are_exactly_multinomial(lb):
    """Dispersion test of the per-cell counts (Poisson counts conditioned by the
    u-driven top-up to total N) in both Poisson regimes (inversion for small means,
    PTRS for large): z-scores of repeated draws must have zero mean and unit
    variance, and counts of different cells must be negatively correlated as in a
    multinomial (independent Poisson counts would show zero correlation and an
    un-conditioned total)."""
    import torch
    from scipy import stats
    e = np.linspace(0, 1, 9)
    dens = np.array([1, 3, 0.2, 5, 0.01, 2, 8, 1, 4.0])
    grid = lb.DensityGrid(e, lambda p: dens, vectorizedpdf=True)
    p = grid.masses / grid.masses.sum()
    for N in (200, 20_000_000):
        reps = 200
        z = np.empty((reps, 8))
        for r in range(reps):
            _, cells = grid._sample_cellmajor(N, 3, r * N, want_cells=True)
            c = torch.bincount(cells, minlength=8).cpu().numpy()
            assert c.sum() == N
            z[r] = (c - N * p) / np.sqrt(N * p * (1 - p))
        for j in range(8):
            s2 = (z[:, j] ** 2).sum()                     # ~ chi^2(reps) if mean 0 and variance 1
            assert stats.chi2.cdf(s2, reps) > 1e-4 and stats.chi2.sf(s2, reps) > 1e-4, (N, j, s2 / reps)
        rho = np.corrcoef(z[:, 3], z[:, 6])[0, 1]
        expect = -np.sqrt(p[3] * p[6] / ((1 - p[3]) * (1 - p[6])))
        assert abs(rho - expect) < 0.25, (rho, expect)


@pytest.mark.parametrize("lam", [12, 60, 130, 1000, 10_000, 90_000])
@pytest.mark.parametrize("zigzag", [False, True])
def test_cellmajor_poisson_counts_in_every_regime(lb, lam, zigzag):
    """4096 equal-mass cells, N = lam * 4096 rows: the per-cell counts of ONE draw are 4096
    (weakly dependent) Binomial(N, 1/4096) variates.  lam = 12 is the inversion branch of the
    Poisson sampler, 60 PTRS on cells that are not split, 130 ... 9e4 PTRS on split cells (fp32
    screen active below 1e5 — exactly where the 256^3 / 1e9 benchmark lives).  `zigzag`: vertex
    densities 1, 3, 1, 3, ...: the same masses, but half of every cell's mass is residual (q = 1/2),
    so the flat and the residual counts are drawn with means lam / 2 each; flat density: q = 1."""
    import torch
    from scipy import stats
    C_ = 4096
    e = np.linspace(0.0, 1.0, C_ + 1)
    dens = np.where(np.arange(C_ + 1) % 2 == 0, 1.0, 3.0) if zigzag else np.ones(C_ + 1)
    grid = lb.DensityGrid(e, lambda p: dens, vectorizedpdf=True, dtype=np.float32)
    N = lam * C_
    _, cells = grid._sample_cellmajor(N, 17, 0, want_cells=True)
    counts = torch.bincount(cells, minlength=C_).cpu().numpy().astype(np.float64)
    del cells
    assert counts.sum() == N
    p = 1.0 / C_
    var = N * p * (1 - p)
    # mean is exact by construction; dispersion: sum of squared z-scores ~ chi^2(C - 1)
    z2 = ((counts - lam) ** 2 / var).sum()
    assert stats.chi2.cdf(z2, C_ - 1) > 1e-5 and stats.chi2.sf(z2, C_ - 1) > 1e-5, z2 / (C_ - 1)
    # shape: 32 equiprobable bins of the Binomial(N, p) pmf
    edges_q = stats.binom.ppf(np.linspace(0, 1, 33)[1:-1], N, p)
    edges_q = np.unique(edges_q)
    cdf = np.concatenate([[0.0], stats.binom.cdf(edges_q, N, p), [1.0]])
    expected = np.diff(cdf) * C_
    observed = np.histogram(counts, bins=np.concatenate([[-0.5], edges_q + 0.5, [N + 0.5]]))[0]
    keep = expected > 5
    chi2 = ((observed[keep] - expected[keep]) ** 2 / expected[keep]).sum()
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-5, (chi2, keep.sum())
    # third moment (skewness of a binomial ~ 1/sqrt(lam)): 5 sigma of its sampling error
    skew = (((counts - lam) / np.sqrt(var)) ** 3).mean()
    assert abs(skew - (1 - 2 * p) / np.sqrt(var)) < 5 * np.sqrt(6.0 / C_)


@pytest.mark.parametrize("name", ["t3d_refined", "t2d_refined", "t1d_refined"])
def test_cellmajor_tree_leaves(lb, name):
    from scipy import stats
    from lintsampler_b200.tree import _CellTable
    g = load_golden(name)
    table = _CellTable(g["leaf_x"], g["leaf_dx"], g["leaf_corners"], g["leaf_mass"], np.float64,
                       None, "parallel")
    N = 200_000
    x, cells = table.sample_cellmajor(N, 5, 0, want_cells=True)
    x, cells = x.cpu().numpy(), cells.cpu().numpy()
    grouped_runs(cells, N)
    assert np.all(x >= g["leaf_x"][cells]) and np.all(x <= g["leaf_x"][cells] + g["leaf_dx"][cells])
    p = g["leaf_mass"] / g["leaf_mass"].sum()
    counts = np.bincount(cells, minlength=p.size)
    big = p * N >= 20
    obs = np.append(counts[big], counts[~big].sum())
    exp = np.append(p[big], p[~big].sum()) * N
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, len(obs) - 1) > 1e-4
    u = np.random.default_rng(8).random((N, x.shape[1] + 1))
    x_ref = O.cells_sample(g["leaf_x"], g["leaf_dx"], g["leaf_corners"], g["leaf_mass"], u)
    for d in range(x.shape[1]):
        assert stats.ks_2samp(x[:, d], x_ref[:, d]).pvalue > 1e-4


# ----------------------------------------------------------------------------- device row permutation
@pytest.mark.parametrize("N", [1, 2, 3, 31, 1000, 65536, 1_000_003])
@pytest.mark.parametrize("k,dtype", [(1, np.float64), (3, np.float32), (6, np.float64)])
def test_permute_rows_is_a_bijection(lb, N, k, dtype):
    """lint_permute_rows == the reference's rng.shuffle as an operation: the same
    rows, each exactly once, in a seed-determined pseudo-random order."""
    import torch
    from lintsampler_b200 import _capi, _device
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    x = torch.arange(N * k, dtype=tdt, device="cuda").reshape(N, k)
    outs = []
    for seed in (1, 1, 2):
        out = torch.empty_like(x)
        _capi.call("lint_permute_rows", _device.ptr(x), _device.ptr(out), N, k, _device.lint_dtype(dtype),
                   C.c_uint64(seed), _device.stream_ptr(x.device))
        outs.append(out.cpu().numpy())
    a, b, c = outs
    assert np.array_equal(a, b)                                   # deterministic in the seed
    src = (a[:, 0] / k).astype(np.int64)
    assert np.array_equal(np.sort(src), np.arange(N))             # every row exactly once
    assert np.array_equal(a, x.cpu().numpy()[src])                # rows kept intact
    if N >= 1000:
        assert not np.array_equal(a, c)                           # another seed, another order
        assert np.mean(src == np.arange(N)) < 0.01
    if N >= 65536:
        # positions look uniform: where the first tenth of the rows ends up
        dest = np.flatnonzero(src < N // 10)
        from scipy import stats
        assert stats.kstest(dest / N, "uniform").pvalue > 1e-4
        assert abs(np.corrcoef(src[:-1], src[1:])[0, 1]) < 0.02   # neighbours are unrelated (5 sigma)


def test_cell_order_with_device_shuffle_is_exchangeable(lb):
    from scipy import stats
    g = load_golden("g3d_iso")
    edges = golden_edges(g)
    V = g["vertex_densities"]
    grid = lb.DensityGrid(tuple(edges), lambda p: V.reshape(-1), vectorizedpdf=True, dtype=np.float32)
    N = 400_000
    grouped = lb.LintSampler(grid, seed=3, order="cell").sample(N)
    mixed = lb.LintSampler(grid, seed=3, order="cell", shuffle=True).sample(N)
    assert sorted(map(tuple, grouped[:2000])) != sorted(map(tuple, mixed[:2000]))
    assert np.array_equal(np.sort(grouped[:, 0]), np.sort(mixed[:, 0]))          # same multiset
    # grouped rows drift with the row index, shuffled rows do not
    assert abs(stats.spearmanr(np.arange(N), grouped[:, 0])[0]) > 0.5
    assert abs(stats.spearmanr(np.arange(N), mixed[:, 0])[0]) < 0.01
    assert stats.ks_2samp(mixed[: N // 20, 0], mixed[N // 2:, 0]).pvalue > 1e-4  # any slice is representative


@pytest.mark.parametrize("N", [1, 2, 7, 33, 1000, 4097])
def test_cellmajor_edge_cases(lb, N):
    """Tiny draws (everything comes from the top-up), a single-cell grid, and a grid
    whose whole mass sits in one cell."""
    import torch
    e = (np.linspace(0, 1, 5), np.linspace(0, 2, 4))
    dens = np.zeros((5, 4))
    dens[2:4, 1:3] = 1.0                       # only cell (2,1) has four non-zero corners... and neighbours partial
    for grid in (lb.DensityGrid(e, lambda p: dens.reshape(-1), vectorizedpdf=True),
                 lb.DensityGrid((np.array([0.0, 1.0]), np.array([0.0, 2.0])), lambda p: np.ones(4), vectorizedpdf=True),
                 lb.DensityGrid(e, lambda p: dens.reshape(-1), vectorizedpdf=True, dtype=np.float32)):
        x, cells = grid._sample_cellmajor(N, 9, 0, want_cells=True)
        x, cells = x.cpu().numpy(), cells.cpu().numpy()
        assert x.shape == (N, 2) and np.all(np.isfinite(x))
        m = grid.masses.ravel()
        assert np.all(m[cells] > 0)                                  # zero-mass cells never drawn
        multi = np.unravel_index(cells, grid.shape)
        for d in range(2):
            ed = grid.edgearrays[d]
            assert np.all(x[:, d] >= ed[multi[d]] - 1e-6) and np.all(x[:, d] <= ed[multi[d] + 1] + 1e-6)


def test_sample_to_host_pipeline(lb):
    import torch
    g = load_golden("g3d_iso")
    edges = tuple(golden_edges(g))
    V = g["vertex_densities"]
    for order in ("draw", "cell"):
        grid = lb.DensityGrid(edges, lambda p: V.reshape(-1), vectorizedpdf=True, dtype=np.float32)
        s = lb.LintSampler(grid, seed=2, order=order)
        out = s.sample_to_host(300_000, chunk=70_000)                # 5 chunks, ragged tail
        torch.cuda.synchronize()
        assert out.shape == (300_000, 3) and out.is_pinned() and not out.is_cuda
        a = out.numpy()
        assert np.all(a >= -4) and np.all(a <= 4) and np.all(np.isfinite(a))
        # same stream position as five successive device calls
        s2 = lb.LintSampler(grid, seed=2, order=order, return_tensor=True)
        parts = [s2.sample(n).cpu().numpy() for n in (70_000, 70_000, 70_000, 70_000, 20_000)]
        assert np.array_equal(a, np.concatenate(parts))
    with pytest.raises(TypeError):
        s.sample_to_host(3.5)
    with pytest.raises(ValueError):
        s.sample_to_host(0)


def test_cellmajor_draws_beyond_the_per_call_limit_are_chunked(lb, monkeypatch):
    from scipy import stats
    from lintsampler_b200 import _device
    g = load_golden("g2d_fullcov")
    edges = golden_edges(g)
    V = g["vertex_densities"]
    grid = lb.DensityGrid(tuple(edges), lambda p: V.reshape(-1), vectorizedpdf=True)
    monkeypatch.setattr(_device, "CELLMAJOR_MAX_ROWS", 100_000)
    N = 250_000
    x, cells = grid._sample_cellmajor(N, 4, 0, want_cells=True)
    cells = cells.cpu().numpy()
    p = (g["masses"] / g["masses"].sum()).ravel()
    counts = np.bincount(cells, minlength=p.size)
    big = p * N >= 20
    obs = np.append(counts[big], counts[~big].sum())
    exp = np.append(p[big], p[~big].sum()) * N
    assert counts.sum() == N and stats.chi2.sf(((obs - exp) ** 2 / exp).sum(), len(obs) - 1) > 1e-4
    # three pieces, each grouped by cell on its own (the last ~16 sqrt(n) rows of a piece are
    # its u-driven top-up, in draw order)
    assert np.sum(np.diff(cells[:94_000]) < 0) <= 1 and np.sum(np.diff(cells[100_000:194_000]) < 0) <= 1
    assert cells[100_000] < cells[99_000 - int(16 * np.sqrt(100_000))]
