"""CPU oracle for the linear-interpolant sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lintsampler_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline, never as the product path.

This is a NumPy *restatement* (not a copy) of the algorithm implemented by
aneeshnaik/lintsampler; every function cites the reference ``file:line`` it
follows (paths relative to ``/root/reference/src/lintsampler``).  The reference
is pure Python/NumPy, so the third-party arithmetic on the path is NumPy's
(`cumsum`, `searchsorted`, pairwise `sum`, PCG64 `Generator.random`) and SciPy's
scrambled Sobol engine; both are present in this image and on the GPU box.

Parity pinning: ``tests/golden/make_golden.py`` imports the *real* reference in
the build container and stores its inputs/outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against them
(cell masses, CDF, cell indices AND in-cell coordinates bit-exactly: the
reduction orders stated in each docstring are the ones NumPy executes for the
reference's expressions, established empirically and pinned by the fixtures).
"""
import numpy as np


# --------------------------------------------------------------------------
# corner ordering
# --------------------------------------------------------------------------
def corner_bits(k):
    """(2^k, k) int array: offset of corner ``c`` along dim ``d`` is bit
    ``k-1-d`` of ``c`` (dim 0 is the most significant bit).

    Reference: density_structures/grid.py:456-459 (``np.binary_repr(i, width=k)``)
    and utils.py:37-74.
    """
    c = np.arange(2 ** k)[:, None]
    shifts = (k - 1 - np.arange(k))[None, :]
    return (c >> shifts) & 1


# --------------------------------------------------------------------------
# grid build: cell masses, total mass, CDF
# --------------------------------------------------------------------------
def cell_masses(edges, vertex_densities):
    """Trapezoid cell masses of a rectilinear grid.

    ``edges``: list of k 1-D arrays; ``vertex_densities``: (n0+1, ..., nk-1+1).
    Reference: grid.py:469-498 (corner sum in corner order starting from zeros,
    one division by 2^k), grid.py:501-518 (volumes as the left-to-right outer
    product of the per-dimension widths) and grid.py:411 (their product).
    """
    V = np.asarray(vertex_densities, dtype=np.float64)
    k = V.ndim
    shape = tuple(n - 1 for n in V.shape)
    acc = np.zeros(shape)
    for bits in corner_bits(k):
        window = tuple(slice(int(b), int(b) + n) for b, n in zip(bits, shape))
        acc += V[window]
    mean_density = acc / 2 ** k
    widths = [np.diff(np.asarray(e)) for e in edges]
    vol = widths[0]
    for w in widths[1:]:
        vol = np.multiply.outer(vol.ravel(), w).ravel()
    return mean_density * vol.reshape(shape)


def total_mass(masses):
    """NumPy pairwise sum of the masses.  Reference: grid.py:412."""
    return np.sum(masses)


def numpy_pairwise_sum_emulated(a):
    """Pure-Python emulation of NumPy's float64 pairwise summation of a
    contiguous 1-D array (numpy/_core/src/umath/loops_utils.h.src,
    ``DOUBLE_pairwise_sum``; third-party code on the path via grid.py:412,431).
    Small inputs only; used to pin the CUDA ``lint_sum_numpy_f64`` kernel's
    summation tree independently of ``np.sum``.
    """
    a = [float(x) for x in np.asarray(a).ravel()]

    def rec(lo, n):
        if n < 8:
            r = -0.0
            for i in range(lo, lo + n):
                r += a[i]
            return r
        if n <= 128:
            r = a[lo:lo + 8]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] += a[lo + i + j]
                i += 8
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
            while i < n:
                res += a[lo + i]
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return rec(lo, n2) + rec(lo + n2, n - n2)

    return rec(0, len(a))


def cdf_from_masses(masses):
    """Normalised cell CDF exactly as the reference builds it on every
    ``choose_cells`` call: ``p = (m / m.sum()).flatten()`` (grid.py:431) then
    ``cdf = p.cumsum(); cdf /= cdf[-1]`` (utils.py:195-196).  Also used for tree
    leaves (tree.py:335) and for the list-of-domains choice
    (lintsampler.py:286,290 — there via ``_choice`` on ``p``).
    """
    m = np.asarray(masses, dtype=np.float64)
    p = (m / m.sum()).ravel()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf


def choose_index(cdf, u):
    """First index ``i`` with ``cdf[i] > u``.  Reference: utils.py:197."""
    return cdf.searchsorted(u, side='right')


# --------------------------------------------------------------------------
# in-cell inverse CDF
# --------------------------------------------------------------------------
def quantile_1d(f0, f1, u):
    """Inverse CDF of the linear density between f0 (x=0) and f1 (x=1).

    Literal restatement of sampling.py:32-37: exact-equality branch returns
    ``u``; otherwise ``(-f0 + sqrt(f0^2 + (f1^2 - f0^2) u)) / (f1 - f0)``.
    """
    f0 = np.asarray(f0, dtype=np.float64)
    f1 = np.asarray(f1, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    z = u.copy()
    ne = f0 != f1
    a = f0[ne]
    b = f1[ne]
    z[ne] = (-a + np.sqrt(a * a + (b * b - a * a) * u[ne])) / (b - a)
    return z


def unit_cube_sample(corners, u):
    """k-D multilinear-interpolant sample in the unit cube.

    ``corners``: (2^k, N) densities in corner order (dim 0 = MSB of the corner
    index); ``u``: (N, k).  Reference: sampling.py:41-94 and the spec in
    docs/source/theory/linear_interpolant.md:255-304: for d = 0..k-1 the
    density is averaged over dims > d (sampling.py:80), multilinearly
    interpolated at the already-drawn z_i for dims < d (sampling.py:84-89), and
    the resulting (g0, g1) pair is inverted in 1-D (sampling.py:92).

    Arithmetic order (mirrored operation-for-operation by the fp64 CUDA
    kernels, and chosen to coincide with what NumPy executes for the
    reference's whole-array expressions, so results are bit-identical to the
    reference): the average over later dims is a left-to-right sum over the
    later-dim corner bits in corner order followed by one division by
    2^(k-1-d); each retained corner value is then multiplied by its weights
    ``(1 - z_i)`` / ``z_i`` for i = 0..d-1 in that order; the weighted values
    are summed left-to-right over the earlier-dim corner bits in corner order.
    """
    corners = np.asarray(corners, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    N, k = u.shape
    z = np.empty((N, k))
    for d in range(k):
        nlater = 2 ** (k - 1 - d)          # corner bits of dims > d (low bits)
        nearly = 2 ** d                    # corner bits of dims < d (high bits)
        g = [None, None]
        for b in (0, 1):                   # offset along dim d
            total = None
            for hi in range(nearly):
                base = (hi * 2 + b) * nlater
                acc = corners[base].copy()
                for lo in range(1, nlater):
                    acc += corners[base + lo]
                if nlater > 1:
                    acc = acc / nlater
                for i in range(d):
                    bit = (hi >> (d - 1 - i)) & 1
                    acc = acc * (z[:, i] if bit else (1 - z[:, i]))
                total = acc if total is None else total + acc
            g[b] = total
        z[:, d] = quantile_1d(g[0], g[1], u[:, d])
    return z


# --------------------------------------------------------------------------
# DensityGrid path
# --------------------------------------------------------------------------
def grid_choose_cells(edges, vertex_densities, u_cell, cdf=None):
    """(flat_idx, mins, maxs, corners) for a rectilinear grid.

    Reference: grid.py:325-356 (``choose_cells``), :414-438 (``_choose_indices``
    incl. the single-cell shortcut at :427-428 and C-order ``unravel_index`` at
    :437), :440-467 (corner gather).
    """
    V = np.asarray(vertex_densities, dtype=np.float64)
    k = V.ndim
    shape = tuple(n - 1 for n in V.shape)
    ncells = int(np.prod(shape))
    u_cell = np.asarray(u_cell, dtype=np.float64)
    if ncells == 1:
        flat = np.zeros(len(u_cell), dtype=np.int64)
    else:
        if cdf is None:
            cdf = cdf_from_masses(cell_masses(edges, V))
        flat = choose_index(cdf, u_cell)
    multi = np.unravel_index(flat, shape)
    mins = np.stack([np.asarray(edges[d])[multi[d]] for d in range(k)], axis=-1)
    maxs = np.stack([np.asarray(edges[d])[multi[d] + 1] for d in range(k)], axis=-1)
    corners = np.stack([V[tuple(multi[d] + int(b[d]) for d in range(k))]
                        for b in corner_bits(k)])
    return flat, mins, maxs, corners


def grid_sample(edges, vertex_densities, u, cdf=None, return_cells=False):
    """Pre-shuffle samples for uniforms ``u`` of shape (N, k+1): columns 0..k-1
    are the in-cell uniforms, column k selects the cell (lintsampler.py:269-270).
    Reference: sampling.py:97-127.
    """
    u = np.asarray(u, dtype=np.float64)
    flat, mins, maxs, corners = grid_choose_cells(edges, vertex_densities, u[:, -1], cdf)
    z = unit_cube_sample(corners, u[:, :-1])
    x = mins + (maxs - mins) * z
    return (x, flat) if return_cells else x


# --------------------------------------------------------------------------
# cell-list path (DensityTree leaves, tree.py:314-355)
# --------------------------------------------------------------------------
def cells_sample(cell_mins, cell_widths, cell_corners, cell_masses_, u,
                 return_cells=False):
    """Sample from a flat list of L hyperbox cells (tree leaves in list order).

    ``cell_mins``/``cell_widths``: (L, k) (``leaf.x``/``leaf.dx``);
    ``cell_corners``: (L, 2^k) (``leaf.corner_densities``); ``cell_masses_``:
    (L,) (``leaf.mass_raw``).  Reference: tree.py:334-355 — ``p = masses /
    np.sum(masses)``, ``_choice``, then per sample mins = x, maxs = x + dx and
    the 2^k corner densities; followed by sampling.py:123-126.
    """
    u = np.asarray(u, dtype=np.float64)
    m = np.asarray(cell_masses_, dtype=np.float64)
    p = m / np.sum(m)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    idx = choose_index(cdf, u[:, -1])
    mins = np.asarray(cell_mins, dtype=np.float64)[idx]
    maxs = mins + np.asarray(cell_widths, dtype=np.float64)[idx]
    corners = np.asarray(cell_corners, dtype=np.float64)[idx].T
    z = unit_cube_sample(corners, u[:, :-1])
    x = mins + (maxs - mins) * z
    return (x, idx) if return_cells else x


# --------------------------------------------------------------------------
# list of domains (lintsampler.py:281-307)
# --------------------------------------------------------------------------
def multi_domain_sample(domain_masses, per_domain_sampler, u, dim):
    """Choose a domain per row with ``_choice`` on column k, rescale that
    column into the domain's CDF interval (lintsampler.py:306) and sample each
    domain; rows come back grouped by domain (lintsampler.py:300-307).

    ``per_domain_sampler(i, usub)`` returns the (n_i, k) samples for domain i.
    """
    u = np.asarray(u, dtype=np.float64)
    gm = np.asarray(domain_masses, dtype=np.float64)
    p = gm / gm.sum()
    c = p.cumsum()
    cn = c / c[-1]
    choice = cn.searchsorted(u[:, -1], side='right')
    bounds = np.append(0, p.cumsum())
    out = np.empty((0, dim))
    for i in range(len(gm)):
        sel = choice == i
        if sel.any():
            usub = u[sel]
            usub[:, -1] = (usub[:, -1] - bounds[i]) / (bounds[i + 1] - bounds[i])
            out = np.vstack((out, per_domain_sampler(i, usub)))
    return out


# --------------------------------------------------------------------------
# uniform streams
# --------------------------------------------------------------------------
def reference_uniforms(seed, N, k):
    """The reference's pseudo-random stream: ``default_rng(seed).random((N,k+1))``
    (lintsampler.py:458,517).  Returns (u, rng) so the caller can continue the
    stream / apply the final ``rng.shuffle`` (lintsampler.py:311)."""
    rng = np.random.default_rng(seed)
    return rng.random((N, k + 1)), rng


def sobol_direct(sv, shift, n0, N):
    """Scrambled-Sobol points n0..n0+N-1 by direct Gray-code index:
    ``x_n[j] = (shift[j] ^ XOR_{b in bits(n ^ n>>1)} sv[j,b]) * 2^-32``.
    Equals ``scipy.stats.qmc.Sobol(d, bits=32).random(N)`` when ``sv``/``shift``
    are the engine's ``_sv``/``_shift`` (lintsampler.py:465,515; scipy
    ``_sobol._draw``)."""
    sv = np.asarray(sv, dtype=np.uint64)
    shift = np.asarray(shift, dtype=np.uint64)
    n = np.arange(n0, n0 + N, dtype=np.uint64)
    gray = n ^ (n >> np.uint64(1))
    d, nbits = sv.shape
    x = np.tile(shift[None, :], (N, 1))
    for b in range(nbits):
        on = ((gray >> np.uint64(b)) & np.uint64(1)).astype(bool)
        x[on] ^= sv[None, :, b]
    return x.astype(np.float64) * 2.0 ** -32


# --------------------------------------------------------------------------
# built-in analytic pdf (new in the B200 build; SciPy-equivalent formula)
# --------------------------------------------------------------------------
def gmm_pdf(x, weights, means, covs):
    """Gaussian-mixture density ``sum_j w_j N(x; mu_j, Sigma_j)`` at points
    ``x`` (..., k).  ``covs`` is (K, k, k).  Used to check the fused CUDA
    vertex-evaluation kernel (the reference evaluates user pdfs such as
    ``scipy.stats.multivariate_normal.pdf`` on the vertex mesh, grid.py:384-393).
    """
    x = np.asarray(x, dtype=np.float64)
    means = np.asarray(means, dtype=np.float64)
    k = means.shape[1]
    if x.ndim == 1 and k == 1:
        x = x[:, None]
    out = np.zeros(x.shape[:-1])
    for w, mu, S in zip(weights, means, np.asarray(covs, dtype=np.float64)):
        L = np.linalg.cholesky(S)
        y = np.linalg.solve(L, (x - mu).reshape(-1, k).T)
        q = (y * y).sum(axis=0).reshape(x.shape[:-1])
        logdet = 2.0 * np.log(np.diag(L)).sum()
        out += w * np.exp(-0.5 * q - 0.5 * logdet - 0.5 * k * np.log(2 * np.pi))
    return out
