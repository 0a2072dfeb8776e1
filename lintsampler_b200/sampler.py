"""``LintSampler``: the user-facing sampler, drop-in for the reference class
(reference lintsampler.py:11-518): same constructor, ``sample(N)`` /
``reset_domain`` semantics, attributes, exceptions, warnings and return shapes.

All per-sample work runs in fused CUDA kernels through the C ABI.  What differs
from the reference is confined to keyword-only options whose defaults keep the
reference's observable behaviour:

``rng_backend``
    ``"philox"`` (default) draws the uniforms on the device (Philox4x32-10 keyed
    by the seed, counter advanced per call); ``"numpy"`` draws them on the host
    with the very ``Generator`` calls of the reference (lintsampler.py:517, 311)
    so that ``sample(N)`` reproduces the reference bit for bit.
``dtype``
    precision for domains this sampler builds itself from ``pdf`` + edges.
``return_tensor``
    return the device-resident ``torch`` tensor instead of a NumPy array.
``order``
    ``"draw"`` (default): rows in draw order, exchangeable like the reference's
    shuffled output.  ``"cell"``: the throughput mode — the same i.i.d. sample as a
    multiset (exact multinomial cell counts), but rows are emitted grouped by
    cell; several times faster on large grids.  Only for the device Philox stream.
``shuffle``
    the reference ends every multi-sample draw with ``rng.shuffle`` (lintsampler.py:
    309-311).  ``"auto"`` (default) keeps that wherever it changes the distribution
    of row order: replayed on the host with ``rng_backend="numpy"`` (bit parity),
    done on the device (``lint_permute_rows``) for quasi-random rows and for rows
    grouped by domain, skipped for i.i.d. device rows of a single domain (already
    exchangeable) and for ``order="cell"`` (grouped by design).  ``True`` forces
    the device permutation, ``False`` disables it.
"""
import ctypes as C
from warnings import warn

import numpy as np
from scipy.stats.qmc import QMCEngine, Sobol

from . import _capi, _device
from .base import DensityStructure
from .grid import DensityGrid, _is_flat_sequence


def sobol_advance(engine, n):
    """``engine.fast_forward(n)`` for a scipy Sobol engine in O(bits) instead of O(n): after ``t``
    points scipy keeps the integer coordinates of point ``t - 1`` in ``_quasi``, and point ``i`` is
    ``shift ^ XOR_{b in gray(i)} sv[:, b]``, so the state can be written directly.
    (``fast_forward`` walks the sequence: 0.3 s per 1e8 points, longer than the draw itself.)"""
    try:
        total = int(engine.num_generated) + int(n)
        if total > 0:
            g = (total - 1) ^ ((total - 1) >> 1)
            x = engine._shift.copy()
            b = 0
            while g:
                if g & 1:
                    x ^= engine._sv[:, b]
                g >>= 1
                b += 1
            engine._quasi[...] = x
        engine.num_generated = total
    except (AttributeError, TypeError, ValueError, IndexError):
        engine.fast_forward(n)


def sobol_sorted_inverse(sv_cell, m):
    """Columns (as ints) of A^-1 over GF(2), A = top m x m block of the generator matrix
    whose columns are the 32-bit direction numbers ``sv_cell``: with it,
    ``gray(n) = XOR_{r in bits(y)} col[r]`` is the point of the first 2^m whose
    coordinate has top m bits ``y ^ top_m(shift)``."""
    a = [int(sv_cell[c]) >> (32 - m) for c in range(m)]
    M = np.array([[(a[c] >> r) & 1 for c in range(m)] + [int(i == r) for i in range(m)]
                  for r in range(m)], dtype=np.uint8)
    for c in range(m):                                  # Gauss-Jordan elimination mod 2
        p = next(r for r in range(c, m) if M[r, c])     # exists: the block is invertible
        if p != c:
            M[[c, p]] = M[[p, c]]
        for r in range(m):
            if r != c and M[r, c]:
                M[r] ^= M[c]
    inv = M[:, m:]
    return [sum(int(inv[c, r]) << c for c in range(m)) for r in range(m)]


def _boxes_overlap(a_lo, a_hi, b_lo, b_hi):
    """Axis-aligned boxes overlap unless separated along some axis [utils.py:156-176]."""
    return not np.any((a_hi <= b_lo) | (b_hi <= a_lo))


class LintSampler:
    """Linear-interpolant sampler for density functions on grids / trees.

    Parameters (as the reference, lintsampler.py:201-205)
    ----------
    domain : ``DensityStructure`` | list of them | edges | tuple of edges |
        list of edges / tuples of edges
    pdf, vectorizedpdf, pdf_args, pdf_kwargs : density callable and its options
    seed : None | int | ``numpy.random.Generator``
    qmc : use quasi-Monte Carlo (scrambled Sobol by default)
    qmc_engine : optional ``scipy.stats.qmc.QMCEngine`` of dimension k+1
    """

    def __init__(self, domain, pdf=None, vectorizedpdf=False, pdf_args=(), pdf_kwargs={},
                 seed=None, qmc=False, qmc_engine=None,
                 *, rng_backend="philox", dtype=np.float64, device=None, cdf_mode="parallel",
                 return_tensor=False, order="draw", shuffle="auto"):
        if pdf:                                                 # truthiness, then callable
            if not callable(pdf):                               # [lintsampler.py:208-213]
                raise TypeError("LintSampler.__init__: Given PDF is not callable.")
        elif vectorizedpdf or pdf_args or pdf_kwargs:
            warn("LintSampler.__init__: PDF configuration setting given but no PDF provided.")
        if rng_backend not in ("philox", "numpy"):
            raise ValueError("LintSampler: rng_backend must be 'philox' or 'numpy'")
        if order not in ("draw", "cell"):
            raise ValueError("LintSampler: order must be 'draw' or 'cell'")
        if order == "cell" and (qmc or rng_backend != "philox"):
            raise ValueError("LintSampler: order='cell' needs the device Philox stream (no qmc)")
        if shuffle not in ("auto", True, False):
            raise ValueError("LintSampler: shuffle must be 'auto', True or False")
        self.order = order
        self.shuffle = shuffle
        self.pdf = pdf
        self.vectorizedpdf = vectorizedpdf
        self.pdf_args = pdf_args
        self.pdf_kwargs = pdf_kwargs
        self.rng_backend = rng_backend
        self.return_tensor = return_tensor
        self._dtype = _device.check_dtype(dtype)
        self._device_arg = device
        self._cdf_mode = cdf_mode
        self._set_grids(domain)
        self._set_random_state(seed, qmc, qmc_engine)

    # ------------------------------------------------------------------ domains
    def reset_domain(self, domain):
        """Swap the domain(s), keeping pdf and random state [lintsampler.py:323-336]."""
        self._set_grids(domain)

    def _make_grid(self, edges):
        return DensityGrid(edges, self.pdf, vectorizedpdf=self.vectorizedpdf, pdf_args=self.pdf_args,
                           pdf_kwargs=self.pdf_kwargs, dtype=self._dtype, device=self._device_arg,
                           cdf_mode=self._cdf_mode)

    def _set_grids(self, domain):
        """Normalise ``domain`` into ``self.grids`` [lintsampler.py:339-436]."""
        prebuilt_msg = ("LintSampler.__set_grids: Pre-constructed DensityStructure provided, "
                        "so `pdf` parameter is redundant.")
        no_pdf_msg = "LintSampler._set_grids: No PDF provided and no pre-constructed DensityStructure."
        mixed_msg = "LintSampler._set_grids: List members of different types"
        if isinstance(domain, DensityStructure):
            grids = [domain]
            if self.pdf:
                warn(prebuilt_msg)
        elif isinstance(domain, list) and all(isinstance(g, DensityStructure) for g in domain):
            if not all(isinstance(g, type(domain[0])) for g in domain):
                raise TypeError(mixed_msg)
            grids = domain
            if self.pdf:
                warn(prebuilt_msg)
        elif isinstance(domain, list) and not _is_flat_sequence(domain):
            if not self.pdf:
                raise ValueError(no_pdf_msg)
            if not all(isinstance(g, type(domain[0])) for g in domain):
                raise TypeError(mixed_msg)
            grids = [self._make_grid(edges) for edges in domain]
        else:
            if not self.pdf:
                raise ValueError(no_pdf_msg)
            grids = [self._make_grid(domain)]
        dim = grids[0].dim
        for g in grids[1:]:
            if g.dim != dim:
                raise ValueError(
                    f"LintSampler._set_grids: Grids have mismatched dimensions: {g.dim} and {dim}")
        for i in range(len(grids) - 1):
            for j in range(i + 1, len(grids)):
                if _boxes_overlap(grids[i].mins, grids[i].maxs, grids[j].mins, grids[j].maxs):
                    raise ValueError(f"LintSampler._set_grids: Grids {i} and {j} spatially overlapping.")
        self.grids = grids
        self.ngrids = len(grids)
        self.dim = dim

    # ------------------------------------------------------------------ random state
    def _set_random_state(self, seed, qmc, qmc_engine):
        """[lintsampler.py:438-498]; plus the device Philox key/counter."""
        self.qmc = qmc
        self.rng = np.random.default_rng(seed)               # TypeError for e.g. seed=42.5
        if qmc:
            if qmc_engine is None:
                qmc_engine = Sobol(d=self.dim + 1, bits=32, seed=seed)
            elif isinstance(qmc_engine, QMCEngine):
                if qmc_engine.d != self.dim + 1:
                    raise ValueError(
                        f"LintSampler.__init__: qmc_engine inconsistent: expected {self.dim + 1}.")
                if seed is not None:
                    warn("LintSampler.__init__: pre-condigured qmc_engine provided, so given random "
                         "seed won't be used except for final array shuffle.")
            else:
                raise TypeError("qmc_engine must be QMCEngine instance or None")
        elif qmc_engine is not None:
            warn("LintSampler.__init__: provided qmc_engine won't be used as qmc switched off.")
        self.qmc_engine = qmc_engine
        # device stream: 64-bit Philox key.  An int seed is the key; a caller's
        # Generator is shared and consumed, as the reference shares it
        # (lintsampler.py:72-77,458); None draws fresh entropy.
        if isinstance(seed, (int, np.integer)) and 0 <= int(seed) < 2 ** 64:
            self._philox_key = int(seed)
        elif isinstance(seed, np.random.Generator) and self.rng_backend == "philox":
            self._philox_key = int(seed.integers(0, 2 ** 64, dtype=np.uint64))
        else:
            ss = seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(
                seed if isinstance(seed, (int, np.integer, type(None))) else None)
            self._philox_key = int(ss.generate_state(1, np.uint64)[0])
        self._philox_offset = 0

    # ------------------------------------------------------------------ uniform sources
    def _device(self):
        """The device every buffer of a draw lives on: the (first) domain's own device when it
        has one, else the `device=` argument / the current CUDA device."""
        if self._device_arg is None:
            for g in getattr(self, "grids", ()):
                dev = getattr(g, "device", None)
                if dev is not None:
                    return dev
        return _device.require_cuda(self._device_arg)

    def _sobol_on_device(self):
        """A scipy Sobol engine with 32-bit output can be replayed on the device
        by direct index; anything else is drawn on the host."""
        e = self.qmc_engine
        return (type(e) is Sobol and getattr(e, "bits", None) == 32
                and getattr(e, "_sv", None) is not None and e._sv.dtype == np.uint32
                and e._sv.shape == (self.dim + 1, 32))

    def _sobol_precheck(self, n):
        """The checks scipy's ``Sobol.random`` would make for this draw (the
        reference's tests expect its non-power-of-two UserWarning)."""
        e = self.qmc_engine
        if e.num_generated + n > e.maxn:
            raise ValueError(
                f"At most 2**{e.bits}={e.maxn} distinct points can be generated. "
                f"{e.num_generated} points have been previously generated, then: "
                f"n={e.num_generated}+{n}={e.num_generated + n}. Consider increasing `bits`.")
        if e.num_generated == 0 and n & (n - 1):
            warn("The balance properties of Sobol' points require n to be a power of 2.")

    def _sobol_struct(self, sorted_n=None):
        """Engine state for the C ABI.  ``sorted_n`` = 2^m asks for the first 2^m points
        enumerated with the cell-selecting coordinate increasing (see
        ``UniformsSobol`` in csrc/lint_device.cuh): needs the inverse over GF(2) of the
        top m x m block of that coordinate's generator matrix."""
        e = self.qmc_engine
        s = _capi.LintSobol()
        for j in range(self.dim + 1):
            for b in range(32):
                s.sv[j][b] = int(e._sv[j, b])
            s.shift[j] = int(e._shift[j])
        s.first_index = int(e.num_generated)
        s.sorted_bits = 0
        if sorted_n:
            m = int(sorted_n).bit_length() - 1
            for r, col in enumerate(sobol_sorted_inverse(e._sv[self.dim], m)):
                s.sorted_ainv[r] = col
            s.sorted_bits = m
        return s

    def _sobol_sorted_ok(self, n):
        """Sorted enumeration applies to a fresh engine, N a power of two, and only when
        the rows are permuted afterwards anyway (the reference's final shuffle)."""
        return (self.qmc_engine.num_generated == 0 and n >= 2 and n & (n - 1) == 0
                and n <= 2 ** 32 and self._shuffle_mode() == "device")

    def _host_uniforms(self, n):
        """[lintsampler.py:500-518] on the host, uploaded."""
        import torch
        if self.qmc:
            u = self.qmc_engine.random(n)
        else:
            u = self.rng.random((n, self.dim + 1))
        return torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64), device=self._device())

    def _device_uniforms(self, n):
        import torch
        with torch.cuda.device(self._device()):
            return self._device_uniforms_impl(n)

    def _device_uniforms_impl(self, n):
        """The (n,k+1) matrix the fused device samplers would consume, materialised
        (needed only for lists of domains, where rows must be routed by value)."""
        import torch
        dev = self._device()
        u = torch.empty((n, self.dim + 1), dtype=torch.float64, device=dev)
        if self.qmc:
            self._sobol_precheck(n)
            s = self._sobol_struct()
            _capi.call("lint_sobol_uniforms", self.dim, n, C.byref(s), _device.ptr(u),
                       _device.stream_ptr(dev))
            sobol_advance(self.qmc_engine, n)
        else:
            _capi.call("lint_philox_uniforms", self.dim, _device.lint_dtype(self._domain_dtype()), n,
                       C.c_uint64(self._philox_key), C.c_uint64(self._philox_offset), _device.ptr(u),
                       _device.stream_ptr(dev))
            self._philox_offset += n
        return u

    def _domain_dtype(self):
        return np.dtype(getattr(self.grids[0], "dtype", np.float64))

    def _uses_host_uniforms(self):
        if self.qmc:
            return not self._sobol_on_device()
        return self.rng_backend == "numpy"

    # ------------------------------------------------------------------ sampling
    def _sample_domain_u(self, grid, u_t):
        import torch
        with torch.cuda.device(self._device()):
            return self._sample_domain_u_impl(grid, u_t)

    def _sample_domain_u_impl(self, grid, u_t):
        """One domain, uniforms given on the device  [sampling.py:97-127]."""
        import torch
        if hasattr(grid, "_sample_u"):
            return grid._sample_u(u_t)
        # generic DensityStructure: its own choose_cells on the host, then the
        # in-cell stage on the device
        u_host = u_t.cpu().numpy()
        mins, maxs, corners = grid.choose_cells(u_host[:, -1])
        dev = self._device()
        n = u_t.shape[0]
        k = self.dim
        mins_t = torch.as_tensor(np.ascontiguousarray(mins, dtype=np.float64).reshape(n, k), device=dev)
        maxs_t = torch.as_tensor(np.ascontiguousarray(maxs, dtype=np.float64).reshape(n, k), device=dev)
        corn_t = torch.as_tensor(np.ascontiguousarray(np.stack(corners), dtype=np.float64), device=dev)
        out = torch.empty((n, k), dtype=torch.float64, device=dev)
        _capi.call("lint_incell_sample_u", k, n, _device.ptr(mins_t), _device.ptr(maxs_t),
                   _device.ptr(corn_t), _device.ptr(u_t), _device.ptr(out), _device.stream_ptr(dev))
        return out

    def _sample_single(self, grid, n):
        """Single domain: fused kernel with device-generated uniforms where possible."""
        native = hasattr(grid, "_sample_philox")
        if self._uses_host_uniforms() or not native:
            u = self._host_uniforms(n) if self._uses_host_uniforms() else self._device_uniforms(n)
            return self._sample_domain_u(grid, u)
        if self.qmc:
            self._sobol_precheck(n)
            s = self._sobol_struct(sorted_n=n if self._sobol_sorted_ok(n) else None)
            x = grid._sample_sobol(n, s)
            sobol_advance(self.qmc_engine, n)
            return x
        if self.order == "cell" and hasattr(grid, "_sample_cellmajor"):
            x = grid._sample_cellmajor(n, self._philox_key, self._philox_offset)
        else:
            x = grid._sample_philox(n, self._philox_key, self._philox_offset)
        self._philox_offset += n
        return x

    def _fused_list(self):
        """True when the list of domains can be drawn by one launch (lint_multi_grid_sample_*):
        native grids of one dtype on one device, at most 16 of them."""
        from .grid import DensityGrid
        g0 = self.grids[0]
        return (len(self.grids) <= 16 and all(type(g) is DensityGrid for g in self.grids)
                and all(g.dtype == g0.dtype and g.device == g0.device for g in self.grids)
                and all(g._stratum == (0.0, 1.0) for g in self.grids))

    def _sample_multi_fused(self, n, p, mode, u=None):
        """One launch over the whole list  [lintsampler.py:281-307].  Returns (rows in draw order,
        grid index of every row)."""
        import torch
        g0 = self.grids[0]
        with torch.cuda.device(g0.device):
            nd = len(self.grids)
            descs = (_capi.LintGrid * nd)(*[g._descriptor() for g in self.grids])
            c = p.cumsum()
            choice = (C.c_double * nd)(*(c / c[-1]))
            bounds = (C.c_double * (nd + 1))(*np.append(0, p.cumsum()))
            nbytes = int(_capi.load().lint_multi_scratch_bytes(self.dim, nd))
            scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=g0.device)
            out = torch.empty((n, self.dim), dtype=_device.torch_dtype(g0.dtype), device=g0.device)
            cells = torch.empty(n, dtype=torch.int64, device=g0.device)
            st = _device.stream_ptr(g0.device)
            if mode == "u":
                _capi.call("lint_multi_grid_sample_u", descs, nd, choice, bounds, _device.ptr(u), n,
                           _device.ptr(scratch), nbytes, _device.ptr(out), _device.ptr(cells), st)
            elif mode == "philox":
                _capi.call("lint_multi_grid_sample_philox", descs, nd, choice, bounds, n,
                           C.c_uint64(self._philox_key), C.c_uint64(self._philox_offset), _device.ptr(scratch),
                           nbytes, _device.ptr(out), _device.ptr(cells), st)
                self._philox_offset += n
            else:
                self._sobol_precheck(n)
                sob = self._sobol_struct()
                _capi.call("lint_multi_grid_sample_sobol", descs, nd, choice, bounds, n, C.byref(sob),
                           _device.ptr(scratch), nbytes, _device.ptr(out), _device.ptr(cells), st)
                sobol_advance(self.qmc_engine, n)
            ends = torch.as_tensor(np.cumsum([int(g.ncells) for g in self.grids]), device=g0.device)
            return out, torch.bucketize(cells, ends, right=True)

    def _sample_multi(self, n):
        """List of domains  [lintsampler.py:281-307]: domain chosen by the same
        uniform that then (rescaled) chooses the cell; rows grouped by domain."""
        import torch
        masses = np.array([g.total_mass for g in self.grids])
        p = masses / masses.sum()
        device_rows = (not self.qmc and self.rng_backend == "philox"
                       and all(hasattr(g, "_sample_philox") for g in self.grids))
        if self._fused_list() and not (device_rows and self.order == "cell"):
            if device_rows:
                # i.i.d. device rows in draw order are exchangeable as they are: no regrouping
                x, _ = self._sample_multi_fused(n, p, "philox")
                return x.to(torch.float64)
            if self._uses_host_uniforms():
                x, dom = self._sample_multi_fused(n, p, "u", self._host_uniforms(n))
            else:
                x, dom = self._sample_multi_fused(n, p, "sobol")
            # the reference's vstack order: rows of grid 0 in draw order, then grid 1, ...
            order = torch.sort(dom, stable=True).indices
            return x[order].to(torch.float64)
        if device_rows:
            # i.i.d. device rows: choosing a domain per row with probability p (what `_choice`
            # does row by row, lintsampler.py:290) is a Multinomial(n, p) split of the row
            # count; each domain then draws its rows with one fused kernel, no uniforms
            # are materialised or routed
            split = np.random.Generator(np.random.Philox(key=self._philox_key,
                                                         counter=self._philox_offset))
            counts = split.multinomial(n, p)
            parts = []
            for grid, ni in zip(self.grids, counts):
                if ni:
                    fn = grid._sample_cellmajor if self.order == "cell" else grid._sample_philox
                    parts.append(fn(int(ni), self._philox_key, self._philox_offset).to(torch.float64))
                    self._philox_offset += int(ni)
            return torch.cat(parts, dim=0)
        u = self._host_uniforms(n) if self._uses_host_uniforms() else self._device_uniforms(n)
        c = p.cumsum()
        cdf_t = torch.as_tensor(c / c[-1], device=u.device)
        bounds = np.append(0, p.cumsum())
        choice = torch.searchsorted(cdf_t, u[:, -1].contiguous(), right=True)
        parts = []
        for i, grid in enumerate(self.grids):
            rows = (choice == i).nonzero(as_tuple=True)[0]
            if rows.numel():
                usub = u[rows].clone()
                usub[:, -1] = (usub[:, -1] - bounds[i]) / (bounds[i + 1] - bounds[i])
                parts.append(self._sample_domain_u(grid, usub).to(torch.float64))
        return torch.cat(parts, dim=0)

    def sample(self, N=None):
        """Draw ``N`` samples (one if ``N`` is None)  [lintsampler.py:233-321].

        Returns a scalar / (k,) for ``N=None``; (N,) in 1-D; (N,k) otherwise.
        """
        if N is not None:
            if not isinstance(N, int):
                raise TypeError(f"LintSampler.sample: Expected int N, got {type(N)}")
            if N <= 0:
                raise ValueError(f"LintSampler.sample: Expected positive N, got {N}")
        n = N if N else 1
        if self.ngrids == 1:
            X = self._sample_single(self.grids[0], n)
        else:
            X = self._sample_multi(n)
        mode = self._shuffle_mode() if N else None
        if mode == "device":
            X = self._permute_rows(X)
        if self.return_tensor:
            if not N:
                return X[0, 0] if self.dim == 1 else X[0]
            if mode == "host":
                # rng.shuffle(X, axis=0) [lintsampler.py:310-311] replayed on the device rows:
                # shuffling arange(N) consumes the generator exactly as shuffling X would and
                # yields the source row of every output row
                import torch
                perm = np.arange(n)
                self.rng.shuffle(perm)
                X = X[torch.as_tensor(perm, device=X.device)]
            return X.squeeze() if self.dim == 1 else X
        X = X.cpu().numpy()
        if mode == "host":
            self.rng.shuffle(X, axis=0)                         # [lintsampler.py:310-311]
        if not N and self.dim == 1:
            return X.item()
        if not N:
            return X[0]
        if self.dim == 1:
            return X.squeeze()
        return X

    def sample_to_host(self, N, out=None, chunk=1 << 26):
        """Draw ``N`` samples straight into host memory, overlapping the sampling
        kernel of one chunk with the device->host copy of the previous one.

        ``out``: optional pinned host ``torch`` tensor of shape (N, k) in the
        domain's dtype; allocated (pinned) when omitted.  Returns the host tensor
        (``.numpy()`` gives a zero-copy NumPy view).  Only for a single native
        domain with the device Philox stream, whose rows are exchangeable, so no
        shuffle pass is needed (see ``_shuffle_mode``)."""
        import torch
        if not isinstance(N, int):
            raise TypeError(f"LintSampler.sample_to_host: Expected int N, got {type(N)}")
        if N <= 0:
            raise ValueError(f"LintSampler.sample_to_host: Expected positive N, got {N}")
        grid = self.grids[0]
        if self.ngrids != 1 or self.qmc or self.rng_backend != "philox" or not hasattr(grid, "_sample_philox"):
            raise ValueError("LintSampler.sample_to_host: needs one native domain and the Philox stream")
        dev = self._device()
        tdt = _device.torch_dtype(self._domain_dtype())
        if out is None:
            out = torch.empty((N, self.dim), dtype=tdt, pin_memory=True)
        if tuple(out.shape) != (N, self.dim) or out.dtype != tdt or out.is_cuda:
            raise ValueError("LintSampler.sample_to_host: `out` must be a host tensor of shape (N, k) "
                             "in the domain's dtype")
        chunk = min(int(chunk), N)
        if getattr(self, "_stage", None) is None or self._stage[0].shape[0] < chunk \
                or self._stage[0].dtype != tdt or self._stage[0].shape[1] != self.dim \
                or self._stage[0].device != dev:
            self._stage = [torch.empty((chunk, self.dim), dtype=tdt, device=dev) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream(dev)
        done = 0
        i = 0
        while done < N:
            n = min(chunk, N - done)
            buf = self._stage[i & 1]
            main.wait_event(self._stage_free[i & 1])            # previous copy out of this buffer
            if self.order == "cell":
                grid._sample_cellmajor(n, self._philox_key, self._philox_offset, out=buf[:n])
            else:
                grid._sample_philox(n, self._philox_key, self._philox_offset, out=buf[:n])
            self._philox_offset += n
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                out[done:done + n].copy_(buf[:n], non_blocking=True)
                self._stage_free[i & 1].record(self._copy_stream)
            done += n
            i += 1
        main.wait_stream(self._copy_stream)
        self._copy_stream.synchronize()         # the caller may read `out` as soon as this returns
        return out

    # ------------------------------------------------------------------ row order
    def _shuffle_mode(self):
        """'host', 'device' or None — see the ``shuffle`` option in the module docstring."""
        if self.rng_backend == "numpy":
            return "host" if self.shuffle in ("auto", True) else None
        if self.shuffle is True:
            return "device"
        if self.shuffle is False:
            return None
        return "device" if (self.qmc or self.ngrids > 1) else None

    def _permute_rows(self, X):
        import torch
        with torch.cuda.device(self._device()):
            return self._permute_rows_impl(X)

    def _permute_rows_impl(self, X):
        """Device row permutation keyed by the sampler's stream position."""
        import torch
        X = X.contiguous()
        out = torch.empty_like(X)
        n, k = X.shape
        # every call gets its own permutation: QMC draws do not advance the Philox offset, so
        # a per-sampler call counter is mixed in as well
        self._permute_calls = getattr(self, "_permute_calls", 0) + 1
        seed = (self._philox_key ^ (0x9E3779B97F4A7C15 * (self._philox_offset + n + 1))
                ^ (0xD1B54A32D192ED03 * self._permute_calls)) & (2 ** 64 - 1)
        _capi.call("lint_permute_rows", _device.ptr(X), _device.ptr(out), n, k,
                   _device.lint_dtype(np.float32 if X.dtype == torch.float32 else np.float64),
                   C.c_uint64(seed), _device.stream_ptr(X.device))
        return out
