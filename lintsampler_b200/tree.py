"""``DensityTree``: adaptive 2^k-tree domain.

Host-side mirror of the reference class (density_structures/tree.py:10-541):
same constructor, ``refine`` / ``get_leaf_at_pos`` / ``choose_cells`` contract,
exceptions, warnings and — crucially — the same *leaf order*, because the order
of ``tree.leaves`` defines the leaf CDF (tree.py:335-338).

Two builders produce the same tree.  For an arbitrary Python pdf, construction
and Romberg refinement are host logic that batches the pdf evaluations of every
cell opened in a sweep.  For the pdf families the library evaluates itself
(``GaussianMixture``, ``DoublePowerLaw``) the cells live in a device table and
``lint_tree_open`` evaluates corner densities, trapezoid masses and Romberg
chains of every batch of new cells (``_DeviceTree``); only the choice of the
cells to open — which fixes the leaf order — is made on the host, from three
numbers per new cell.  Sampling runs on the device over the flattened leaf
table (``lint_cells_sample_*``).
"""
import ctypes as C
from warnings import warn

import numpy as np

from . import _capi, _device
from .base import DensityStructure


class TreeCell:
    """One node of the tree (reference ``_TreeCell``, tree.py:544-703).

    Attributes kept API-compatible with the reference: ``parent``, ``idx``,
    ``level``, ``x`` (origin position), ``dx`` (extent), ``vol``,
    ``corner_densities`` (2^k, corner order), ``mass_raw``,
    ``romberg_estimates``, ``mass_romberg``, ``children`` (tuple or None).
    ``origin`` holds the integer lattice coordinates of the cell at its level.
    """
    __slots__ = ("parent", "idx", "level", "origin", "x", "dx", "vol", "corner_densities",
                 "mass_raw", "romberg_estimates", "mass_romberg", "children")

    def __init__(self, parent, idx, level, origin):
        self.parent = parent
        self.idx = idx
        self.level = level
        self.origin = origin
        self.children = None


def _unit_corners(k):
    """(2^k, k) corner offsets, dim 0 = most significant bit  [utils.py:37-74]."""
    c = np.arange(2 ** k)[:, None]
    return (c >> (k - 1 - np.arange(k))[None, :]) & 1


class _VertexEvaluator:
    """Evaluates the pdf on lattice vertices ``mins + c * (maxs-mins)/2^level``
    with an optional cache keyed by the vertex's coarsest (level, coords)
    representation (the reference's ``_GridCache``, tree.py:707-929)."""

    def __init__(self, mins, maxs, pdf, vectorized, args, kwargs, usecache):
        self.mins, self.maxs = mins, maxs
        self.pdf, self.vectorized, self.args, self.kwargs = pdf, vectorized, args, kwargs
        self.usecache = usecache
        self.k = len(mins)
        self.cache = {}

    def positions(self, coords, level):
        return self.mins + coords * ((self.maxs - self.mins) / 2 ** level)   # tree.py:866-867

    def _call(self, pos):
        """pdf on an (n,k) batch of positions -> (n,) fp64."""
        n = len(pos)
        if self.k == 1:
            pos = pos[..., 0]                                   # tree.py:813-814
        if self.vectorized:
            out = np.empty(n, dtype=np.float64)
            try:
                out[:] = self.pdf(pos, *self.args, **self.kwargs)   # NumPy assignment rules, as
            except TypeError:                                       # the reference (tree.py:816-822)
                raise ValueError("DensityTree: pdf function does not return appropriate shape.")
            return out
        out = np.empty(n, dtype=np.float64)
        for i in range(n):
            try:
                out[i] = self.pdf(pos[i], *self.args, **self.kwargs)
            except TypeError:
                raise ValueError("DensityTree: pdf function does not return appropriate shape.")
        return out

    def eval(self, coords, level):
        """Densities at integer lattice ``coords`` (n,k) of refinement ``level``."""
        coords = np.asarray(coords, dtype=np.int64)
        n = len(coords)
        if not self.usecache:
            return self._call(self.positions(coords, level))
        # canonical representation: halve while every coordinate is even
        cc = coords.copy()
        lev = np.full(n, level, dtype=np.int64)
        while True:
            m = (lev > 0) & np.all(cc % 2 == 0, axis=1)
            if not m.any():
                break
            cc[m] //= 2
            lev[m] -= 1
        keys = [(int(l),) + tuple(int(v) for v in row) for l, row in zip(lev, cc)]
        out = np.empty(n, dtype=np.float64)
        missing = {}
        for i, key in enumerate(keys):
            v = self.cache.get(key)
            if v is None:
                missing.setdefault(key, []).append(i)
            else:
                out[i] = v
        if missing:
            first = np.array([ix[0] for ix in missing.values()])
            vals = self._call(self.positions(coords[first], level))
            for (key, ix), v in zip(missing.items(), vals):
                self.cache[key] = v
                out[ix] = v
        return out


class _DeviceTree:
    """Cell table of a tree on the device (``lint_tree_t``) and the host mirror of the per-cell
    scalars the refinement logic reads (tree.py:413-488): level, parent, first child, raw mass,
    Romberg mass, last Romberg correction."""

    room_override = None        # (tests) openings per launch of the stage-2 loop

    def __init__(self, mins, maxs, pdf, device):
        import torch
        self.device = _device.require_cuda(device)
        self.k = len(mins)
        self.nc = 2 ** self.k
        self.mins = np.asarray(mins, dtype=np.float64)
        self.maxs = np.asarray(maxs, dtype=np.float64)
        self.pdf = pdf
        self.stride = int(_capi.load().lint_tree_est_stride())
        self.max_level = self.stride - 2
        self.n = 0
        self.cap = 0
        self._t = {}
        self.level = np.empty(0, dtype=np.int32)
        self.parent = np.empty(0, dtype=np.int64)
        self.first_child = np.empty(0, dtype=np.int64)
        self.mass_raw = np.empty(0)
        self.mass_rom = np.empty(0)
        self.err = np.empty(0)
        self._flags = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._grow(4096)
        self._launch(None)                                       # the root

    # per-cell device arrays: name -> (columns, dtype)
    def _layout(self):
        import torch
        return {"origin": (self.k, torch.int32), "level": (1, torch.int32), "parent": (1, torch.int32),
                "corners": (self.nc, torch.float64), "mass_raw": (1, torch.float64),
                "est": (self.stride, torch.float64)}

    def _grow(self, need):
        import torch
        if need <= self.cap:
            return
        cap = max(need, 2 * self.cap)
        with torch.cuda.device(self.device):
            for name, (cols, dt) in self._layout().items():
                new = torch.empty((cap, cols), dtype=dt, device=self.device)
                if self.n:
                    new[:self.n] = self._t[name][:self.n]
                self._t[name] = new
        for name in ("level", "parent", "first_child", "mass_raw", "mass_rom", "err"):
            old = getattr(self, name)
            new = np.empty(cap, dtype=old.dtype)
            new[:self.n] = old[:self.n]
            setattr(self, name, new)
        self.cap = cap

    def _struct(self):
        t = _capi.LintTree()
        t.dim = self.k
        for d in range(self.k):
            t.mins[d], t.maxs[d] = float(self.mins[d]), float(self.maxs[d])
        t.capacity = self.cap
        for name in self._layout():
            setattr(t, name + "_dev", self._t[name].data_ptr())
        return t

    def _pdf_structs(self):
        from .pdfs import GaussianMixture
        if isinstance(self.pdf, GaussianMixture):
            gs, keep = self.pdf._as_struct()
            return C.byref(gs), None, (gs, keep)
        hs, keep = self.pdf._as_struct()
        return None, C.byref(hs), (hs, keep)

    def _launch(self, ids):
        """Create the root (ids None) or the children of the cells ``ids``; returns the new ids."""
        import torch
        with torch.cuda.device(self.device):
            n_open = 0 if ids is None else len(ids)
            n_new = 1 if ids is None else n_open * self.nc
            first = self.n
            self._grow(first + n_new)
            ids_t = None if ids is None else torch.as_tensor(np.asarray(ids, dtype=np.int32), device=self.device)
            summary = torch.empty((n_new, 3), dtype=torch.float64, device=self.device)
            gmm, halos, keep = self._pdf_structs()
            t = self._struct()
            _capi.call("lint_tree_open", C.byref(t), gmm, halos, _device.ptr(ids_t), n_open, first,
                       _device.ptr(summary), _device.ptr(self._flags), _device.stream_ptr(self.device))
            del keep
            s = summary.cpu().numpy()
        new = np.arange(first, first + n_new, dtype=np.int64)
        if ids is None:
            self.level[0], self.parent[0] = 0, -1
        else:
            ids = np.asarray(ids, dtype=np.int64)
            self.level[new] = np.repeat(self.level[ids] + 1, self.nc)
            self.parent[new] = np.repeat(ids, self.nc)
            self.first_child[ids] = first + self.nc * np.arange(n_open, dtype=np.int64)
        self.first_child[new] = -1
        self.mass_raw[new], self.mass_rom[new], self.err[new] = s[:, 0], s[:, 1], s[:, 2]
        self.n = first + n_new
        return new

    def open(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        if len(ids) and int(self.level[ids].max()) >= self.max_level:
            raise ValueError(f"DensityTree: the device builder refines to level {self.max_level} at most")
        return self._launch(ids)

    def refine_worst(self, ids, tree_tol):
        """Stage 2 of ``refine`` (tree.py:448-488) on the device, one launch per batch of openings
        (``lint_tree_refine_worst``).  Returns ``(leaf ids, finished)``; not finished means the
        leaf list outgrew what the single-CTA loop handles and the caller carries on."""
        import torch
        lib = _capi.load()
        max_leaves = int(lib.lint_tree_max_refine_leaves())
        ids = np.asarray(ids, dtype=np.int64)
        with torch.cuda.device(self.device):
            while True:
                n = len(ids)
                room = self.room_override or max(512, n // 2)        # openings this launch has room for
                if n + room * (self.nc - 1) + self.nc > max_leaves:
                    return ids, False
                self._grow(self.n + room * self.nc)
                leaf_cap = n + room * (self.nc - 1) + self.nc
                dev = self.device
                ids_t = torch.empty(leaf_cap, dtype=torch.int32, device=dev)
                ids_t[:n] = torch.as_tensor(ids.astype(np.int32), device=dev)
                rom_t = torch.empty(leaf_cap, dtype=torch.float64, device=dev)
                raw_t = torch.empty(leaf_cap, dtype=torch.float64, device=dev)
                rom_t[:n] = torch.as_tensor(self.mass_rom[ids], device=dev)
                raw_t[:n] = torch.as_tensor(self.mass_raw[ids], device=dev)
                sq_t = torch.empty(leaf_cap, dtype=torch.float64, device=dev)
                state = torch.tensor([n, self.n, 0, 0], dtype=torch.int32, device=dev)
                gmm, halos, keep = self._pdf_structs()
                t = self._struct()
                _capi.call("lint_tree_refine_worst", C.byref(t), gmm, halos, _device.ptr(ids_t), _device.ptr(rom_t),
                           _device.ptr(raw_t), _device.ptr(sq_t), leaf_cap, float(tree_tol), room,
                           _device.ptr(state), _device.ptr(self._flags), _device.stream_ptr(dev))
                del keep
                n_leaves, n_cells, done, status = (int(v) for v in state.cpu().numpy())
                self._adopt(n_cells)
                ids = ids_t[:n_leaves].cpu().numpy().astype(np.int64)
                if status == 0:
                    return ids, True
                if status == 2:
                    raise ValueError(f"DensityTree: the device builder refines to level {self.max_level} at most")
                if done == 0:                                        # no room at all although we just made some
                    return ids, False

    def _adopt(self, n_cells):
        """Bring the host mirror up to date with cells [self.n, n_cells) made by a device-side loop."""
        if n_cells == self.n:
            return
        new = np.arange(self.n, n_cells, dtype=np.int64)
        n_before, self.n = self.n, n_cells                           # fetch() reads up to self.n
        lev = self.fetch("level", new)[:, 0]
        par = self.fetch("parent", new)[:, 0].astype(np.int64)
        est = self.fetch("est", new)
        rows = np.arange(len(new))
        self.level[new], self.parent[new] = lev, par
        self.first_child[new] = -1
        self.first_child[par[::self.nc]] = new[::self.nc]
        self.mass_raw[new] = self.fetch("mass_raw", new)[:, 0]
        self.mass_rom[new] = est[rows, lev]
        self.err[new] = np.abs(est[rows, lev] - est[rows, np.maximum(lev - 1, 0)])
        assert n_before + len(new) == n_cells

    def check(self):
        """Density checks of tree.py:598-605 for everything evaluated so far."""
        bad = int(self._flags.item())
        if bad & _capi.LINT_FLAG_NEGATIVE:
            raise ValueError("DensityTree: Densities can't be negative")
        if bad & _capi.LINT_FLAG_NONFINITE:
            raise ValueError("DensityTree: Detected non-finite density")

    def fetch(self, name, ids=None):
        """Host copy of a per-cell device array (rows ``ids``, default all cells)."""
        import torch
        t = self._t[name][:self.n]
        if ids is not None:
            t = t[torch.as_tensor(np.asarray(ids, dtype=np.int64), device=self.device)]
        return t.cpu().numpy()

    def geometry(self, ids):
        """(x, dx) of cells ``ids`` with the reference's arithmetic: dx = span / 2^level,
        x = mins + origin * dx  [tree.py:585-591]."""
        org = self.fetch("origin", ids).astype(np.int64)
        lev = self.level[np.asarray(ids, dtype=np.int64)]
        dx = (self.maxs - self.mins)[None, :] / (2.0 ** lev)[:, None]
        return self.mins[None, :] + org * dx, dx


class DensityTree(DensityStructure):
    """Tree-like domain over the box ``[mins, maxs]``.

    Parameters (as the reference, tree.py:228-232)
    ----------
    mins, maxs : scalars (1-D) or length-k sequences.
    pdf, vectorizedpdf, pdf_args, pdf_kwargs : as ``DensityGrid``.
    min_openings : number of full openings at construction.
    usecache : memoise pdf evaluations on shared vertices.
    batch : evaluate all vertices of a cell's children in one vectorised call.

    Keyword-only extensions: ``dtype``, ``device``, ``cdf_mode`` as ``DensityGrid``;
    ``build``: "device" builds the tree in a device cell table (``lint_tree_open``; needs one
    of the library's own pdf families), "host" keeps the host builder, None (default) picks
    the device builder whenever the pdf allows it and a CUDA device is present.
    """

    def __init__(self, mins, maxs, pdf, vectorizedpdf=False, pdf_args=(), pdf_kwargs={},
                 min_openings=1, usecache=True, batch=False,
                 *, dtype=np.float64, device=None, cdf_mode="parallel", build=None):
        if not hasattr(mins, "__len__"):
            mins, maxs = np.array([mins]), np.array([maxs])
        else:
            mins, maxs = np.array(mins), np.array(maxs)
        if len(maxs) != len(mins):
            raise ValueError("DensityTree.__init__: mins and maxs have different lengths.")
        for arr in (mins, maxs):                                # tree.py:248-262
            if arr.ndim != 1:
                raise ValueError("DensityTree.__init__: Need 1D array of minima/maxima.")
            if np.any((maxs - mins) <= 0):
                raise ValueError("DensityTree.__init__: Coordinates not monotically increasing.")
            if not np.all(np.isfinite(arr)):
                raise ValueError("DensityTree.__init__: Coordinates not finite-valued.")
        if len(mins) > _capi.LINT_MAX_DIM:
            raise ValueError(f"DensityTree: at most {_capi.LINT_MAX_DIM} dimensions are supported")
        if batch and usecache:
            warn("DensityTree.__init__: `usecache` set to True but cache not used if `batch=True`")
        self._mins, self._maxs, self._dim = mins, maxs, len(mins)
        self.usecache, self.batch = usecache, batch
        self.dtype = _device.check_dtype(dtype)
        self.cdf_mode = cdf_mode
        self._device_arg = device
        self._table = None          # device leaf table, rebuilt lazily
        self._corners01 = _unit_corners(self._dim)
        self._dt = None             # device cell table (device builder)
        self._leaf_ids = None       # its leaves, in the reference's leaf order
        self._cells = None          # TreeCell objects of a device-built tree, made on demand
        if self._wants_device_build(pdf, pdf_args, pdf_kwargs, build):
            self._dt = _DeviceTree(mins, maxs, pdf, device)
            ids = np.zeros(1, dtype=np.int64)
            for _ in range(min_openings):                       # tree.py:285-291
                ids = self._dt.open(ids)                        # children of leaf i follow those of leaf i-1
            self._leaf_ids = ids
            self._dt.check()
            return
        self._eval = _VertexEvaluator(mins, maxs, pdf, vectorizedpdf or batch, pdf_args, pdf_kwargs,
                                      usecache and not batch)
        self._root = TreeCell(None, 0, 0, np.zeros(self._dim, dtype=np.int64))
        self._finish_cells([self._root])
        leaves = [self._root]
        for _ in range(min_openings):                           # tree.py:285-291
            self._open(leaves)
            leaves = [child for leaf in leaves for child in leaf.children]
        self._leaves = leaves

    @staticmethod
    def _wants_device_build(pdf, pdf_args, pdf_kwargs, build):
        from .pdfs import FUSED_PDFS
        if build not in (None, "host", "device"):
            raise ValueError(f"DensityTree: build must be None, 'host' or 'device', got {build!r}")
        fused = isinstance(pdf, FUSED_PDFS) and not pdf_args and not pdf_kwargs
        if build == "device":
            if not fused:
                raise ValueError("DensityTree: build='device' needs a GaussianMixture or DoublePowerLaw pdf "
                                 "without extra arguments")
            return True
        if build == "host" or not fused:
            return False
        import torch
        return torch.cuda.is_available()

    # ------------------------------------------------------------------ cells as objects
    @property
    def leaves(self):
        """Leaf cells in the reference's order (tree.py:290-291, 440-446, 479-481)."""
        if self._dt is not None:
            cells = self._materialise()
            return [cells[i] for i in self._leaf_ids]
        return self._leaves

    @leaves.setter
    def leaves(self, value):
        if self._dt is not None:
            raise AttributeError("DensityTree: the leaves of a device-built tree cannot be reassigned")
        self._leaves = value

    @property
    def root(self):
        return self._materialise()[0] if self._dt is not None else self._root

    def _materialise(self):
        """TreeCell objects for every cell of the device table (API compatibility: ``leaves``,
        ``root``, ``get_leaf_at_pos``); built once per state of the tree."""
        if self._cells is not None and len(self._cells) == self._dt.n:
            return self._cells
        dt, k = self._dt, self._dim
        ids = np.arange(dt.n)
        org = dt.fetch("origin").astype(np.int64)
        x, dx = dt.geometry(ids)
        corners, est = dt.fetch("corners"), dt.fetch("est")
        cells = []
        for i in range(dt.n):
            par = cells[dt.parent[i]] if dt.parent[i] >= 0 else None
            j = 0 if par is None else i - dt.first_child[dt.parent[i]]
            c = TreeCell(par, 0 if par is None else par.idx * 2 ** k + int(j), int(dt.level[i]), org[i])
            c.x, c.dx, c.vol = x[i], dx[i], np.prod(dx[i])
            c.corner_densities = corners[i]
            c.mass_raw = dt.mass_raw[i]
            c.romberg_estimates = list(est[i, :dt.level[i] + 1])
            c.mass_romberg = dt.mass_rom[i]
            cells.append(c)
        for i in range(dt.n):
            fc = dt.first_child[i]
            if fc >= 0:
                cells[i].children = tuple(cells[fc:fc + 2 ** k])
        self._cells = cells
        return cells

    # ------------------------------------------------------------------ properties
    @property
    def dim(self):
        return self._dim

    @property
    def mins(self):
        return self._mins

    @property
    def maxs(self):
        return self._maxs

    @property
    def _leaf_masses(self):
        if self._dt is not None:
            return self._dt.mass_raw[self._leaf_ids].copy()
        return np.array([leaf.mass_raw for leaf in self.leaves])

    @property
    def total_mass(self):
        return np.sum(self._leaf_masses)

    # ------------------------------------------------------------------ construction
    def _open(self, cells):
        """Create the 2^k children of every cell in ``cells`` (tree.py:615-647),
        evaluating all their corner densities in one batch."""
        k = self._dim
        for cell in cells:
            kids = []
            for j in range(2 ** k):
                kids.append(TreeCell(cell, cell.idx * 2 ** k + j, cell.level + 1,
                                     2 * cell.origin + self._corners01[j]))
            cell.children = tuple(kids)
        self._finish_cells([kid for cell in cells for kid in cell.children])
        self._table = None

    def _finish_cells(self, cells):
        """Evaluate corner densities, trapezoid mass and Romberg estimates."""
        k = self._dim
        nc = 2 ** k
        span = self._maxs - self._mins
        # corner densities, one pdf batch per refinement level present
        by_level = {}
        for cell in cells:
            by_level.setdefault(cell.level, []).append(cell)
        for level, group in by_level.items():
            lattice = (np.stack([c.origin for c in group])[:, None, :]
                       + self._corners01[None, :, :]).reshape(-1, k)
            f = self._eval.eval(lattice, level).reshape(len(group), nc)
            if np.any(f < 0):                                   # tree.py:598-601
                raise ValueError("DensityTree: Densities can't be negative")
            if not np.all(np.isfinite(f)):                      # tree.py:602-605
                raise ValueError("DensityTree: Detected non-finite density")
            dx = span / 2 ** level
            vol = np.prod(dx)
            mass = vol * (np.sum(f, axis=1) / nc)               # vol * average  [tree.py:608]
            for cell, fc, m in zip(group, f, mass):
                cell.dx = dx
                cell.vol = vol
                cell.x = self._mins + cell.origin * dx
                cell.corner_densities = fc
                cell.mass_raw = m
        # Romberg estimates: siblings share their whole ancestor chain
        done = set()
        for cell in cells:
            if id(cell) in done:
                continue
            sibs = [cell] if cell.parent is None else [c for c in cell.parent.children]
            for s in sibs:
                done.add(id(s))
            self._romberg(sibs)

    def _romberg(self, sibs):
        """Richardson-extrapolated masses of sibling cells  [tree.py:649-703].

        ``raws[:, 0]`` is each cell's own trapezoid mass; ``raws[:, j]`` is the
        mass its j-th ancestor's multilinear interpolant assigns to it (value
        at the cell's midpoint times the cell volume)."""
        k = self._dim
        level = sibs[0].level
        n = len(sibs)
        raws = np.empty((n, level + 1))
        raws[:, 0] = [s.mass_raw for s in sibs]
        origins = np.stack([s.origin for s in sibs]).astype(np.float64)
        anc = sibs[0]
        for j in range(1, level + 1):
            anc = anc.parent
            # midpoint of each cell in the ancestor's unit cube (dyadic, exact)
            mid = (origins + 0.5) / 2 ** j - anc.origin
            w = np.stack((1 - mid[:, 0], mid[:, 0]), axis=-1)
            for d in range(1, k):
                rd = np.stack((1 - mid[:, d], mid[:, d]), axis=-1)
                w = (w[:, :, None] * rd[:, None, :]).reshape(n, -1)
            fmid = np.sum(anc.corner_densities[None, :] * w, axis=1)
            raws[:, j] = fmid * anc.vol / 2 ** (j * k)
        est = np.empty((n, level + 1))
        est[:, 0] = raws[:, 0]
        cur = raws.copy()
        for i in range(level):
            cur = cur[:, :-1] - np.diff(cur, axis=1) / (4 ** (i + 1) - 1)
            est[:, i + 1] = cur[:, 0]
        for s, e in zip(sibs, est):
            s.romberg_estimates = list(e)
            s.mass_romberg = e[-1]

    @staticmethod
    def _leaf_error(leaf):
        e = leaf.romberg_estimates
        return np.abs(e[-1] - e[-2])

    # ------------------------------------------------------------------ refinement
    def refine(self, tree_tol, leaf_tol=0.01, leaf_mass_contribution_tol=0.01, verbose=False):
        """Open leaves with large Romberg mass-error estimates  [tree.py:357-488].

        Stage 1 repeatedly opens every leaf whose last Romberg correction exceeds
        ``leaf_tol`` of its mass (ignoring leaves lighter than
        ``leaf_mass_contribution_tol`` times the mean leaf); stage 2 repeatedly
        opens the single worst leaf until the tree-wide fractional error drops
        below ``tree_tol``.
        """
        if (len(self._leaf_ids) if self._dt is not None else len(self.leaves)) == 1:
            raise RuntimeError(
                "DensityTree.refine: Romberg refinement needs at least two tree levels "
                "to estimate mass errors. Instantiate tree with min_openings>0.")
        if self._dt is not None:
            return self._refine_device(tree_tol, leaf_tol, leaf_mass_contribution_tol, verbose)
        if verbose:
            print(f"Pre-loop: {len(self.leaves)} leaves on tree. Total mass={self.total_mass}")

        m_rom = np.array([leaf.mass_romberg for leaf in self.leaves])
        errs = np.array([self._leaf_error(leaf) for leaf in self.leaves])
        sweep = 0
        while True:
            sweep += 1
            sel = np.where((errs > leaf_tol * m_rom)
                           & (m_rom > leaf_mass_contribution_tol * np.nanmean(m_rom)))[0]
            if len(sel):
                opened = [self.leaves[i] for i in sel[::-1]]       # reversed index order
                keep = np.ones(len(self.leaves), dtype=bool)
                keep[sel] = False
                self._open(opened)
                fresh = [kid for cell in opened for kid in cell.children]
                self.leaves = [leaf for leaf, kp in zip(self.leaves, keep) if kp] + fresh
                m_rom = np.concatenate((m_rom[keep], [c.mass_romberg for c in fresh]))
                errs = np.concatenate((errs[keep], [self._leaf_error(c) for c in fresh]))
            if verbose:
                print(f"End of leaf iteration {sweep}: {len(self.leaves)} leaves on tree. "
                      f"Total mass={np.sum(m_rom)}, with mean leaf mass={np.nanmean(m_rom)}")
            if not len(sel):
                break

        m_rom = np.array([leaf.mass_romberg for leaf in self.leaves])
        m_raw = np.array([leaf.mass_raw for leaf in self.leaves])
        it = 0
        while True:
            it += 1
            m_tot = np.sum(m_rom)
            sq = (m_rom - m_raw) ** 2
            err_tot = np.sqrt(np.sum(sq))
            if err_tot / m_tot < tree_tol:
                break
            worst = int(np.argmax(sq))
            cell = self.leaves.pop(worst)
            self._open([cell])
            self.leaves += list(cell.children)
            m_rom = np.concatenate((np.delete(m_rom, worst), [c.mass_romberg for c in cell.children]))
            m_raw = np.concatenate((np.delete(m_raw, worst), [c.mass_raw for c in cell.children]))
            if verbose:
                print(f"End of tree iteration {it}: {len(self.leaves)} leaves on tree. "
                      f"Total mass={m_tot}. Fractional error={err_tot / m_tot}.")
        self._table = None

    def _refine_device(self, tree_tol, leaf_tol, leaf_mass_contribution_tol, verbose):
        """``refine`` over the device cell table: the same two stages and the same open order on
        arrays of cell ids; every batch of new cells is one ``lint_tree_open`` launch."""
        dt = self._dt
        ids = self._leaf_ids
        if verbose:
            print(f"Pre-loop: {len(ids)} leaves on tree. Total mass={self.total_mass}")
        m_rom, errs = dt.mass_rom[ids].copy(), dt.err[ids].copy()
        sweep = 0
        while True:
            sweep += 1
            sel = np.where((errs > leaf_tol * m_rom)
                           & (m_rom > leaf_mass_contribution_tol * np.nanmean(m_rom)))[0]
            if len(sel):
                keep = np.ones(len(ids), dtype=bool)
                keep[sel] = False
                fresh = dt.open(ids[sel[::-1]])                    # reversed index order (tree.py:440-446)
                ids = np.concatenate((ids[keep], fresh))
                m_rom = np.concatenate((m_rom[keep], dt.mass_rom[fresh]))
                errs = np.concatenate((errs[keep], dt.err[fresh]))
            if verbose:
                print(f"End of leaf iteration {sweep}: {len(ids)} leaves on tree. "
                      f"Total mass={np.sum(m_rom)}, with mean leaf mass={np.nanmean(m_rom)}")
            if not len(sel):
                break
        if not verbose:                                            # stage 2 in one launch per batch of openings
            ids, finished = dt.refine_worst(ids, tree_tol)
            if finished:
                self._leaf_ids = ids
                self._table = None
                dt.check()
                return
            m_rom = dt.mass_rom[ids].copy()
        m_raw = dt.mass_raw[ids].copy()
        it = 0
        while True:
            it += 1
            m_tot = np.sum(m_rom)
            sq = (m_rom - m_raw) ** 2
            err_tot = np.sqrt(np.sum(sq))
            if err_tot / m_tot < tree_tol:
                break
            worst = int(np.argmax(sq))
            fresh = dt.open(ids[worst:worst + 1])
            ids = np.concatenate((np.delete(ids, worst), fresh))
            m_rom = np.concatenate((np.delete(m_rom, worst), dt.mass_rom[fresh]))
            m_raw = np.concatenate((np.delete(m_raw, worst), dt.mass_raw[fresh]))
            if verbose:
                print(f"End of tree iteration {it}: {len(ids)} leaves on tree. "
                      f"Total mass={m_tot}. Fractional error={err_tot / m_tot}.")
        self._leaf_ids = ids
        self._table = None
        dt.check()

    def get_leaf_at_pos(self, pos):
        """Leaf containing ``pos``  [tree.py:490-541]."""
        pos = np.array([pos]) if not hasattr(pos, "__len__") else np.array(pos)
        if pos.shape != (self._dim,):
            raise ValueError(
                f"DensityTree.get_leaf_at_pos: Shape of input pos: {pos.shape} does not make sense.")
        r = (pos - self._mins) / (self._maxs - self._mins)
        if not np.all((r >= 0) & (r <= 1)):
            raise ValueError("DensityTree.get_leaf_at_pos: Requested pos falls outside tree.")
        cell = self.root
        weights = 2 ** np.arange(self._dim)[::-1]
        while cell.children:
            upper = (r > 0.5).astype(int)
            cell = cell.children[int(upper.dot(weights))]
            r = 2 * r - upper
        return cell

    # ------------------------------------------------------------------ device leaf table
    def leaf_table(self):
        """Host arrays ``(x (L,k), dx (L,k), corners (L,2^k), mass (L,))`` in leaf order."""
        if self._dt is not None:
            ids = self._leaf_ids
            x, dx = self._dt.geometry(ids)
            return x, dx, self._dt.fetch("corners", ids), self._dt.mass_raw[ids].copy()
        L, k = len(self.leaves), self._dim
        x = np.array([leaf.x for leaf in self.leaves], dtype=np.float64).reshape(L, k)
        dx = np.array([leaf.dx for leaf in self.leaves], dtype=np.float64).reshape(L, k)
        cd = np.array([leaf.corner_densities for leaf in self.leaves], dtype=np.float64).reshape(L, 2 ** k)
        m = np.array([leaf.mass_raw for leaf in self.leaves], dtype=np.float64)
        return x, dx, cd, m

    def _ensure_device(self):
        if self._table is not None:
            return self._table
        self._table = _CellTable(*self.leaf_table(), dtype=self.dtype, device=self._device_arg,
                                 cdf_mode=self.cdf_mode)
        self.device = self._table.device
        return self._table

    def _sample_u(self, u_t, want_cells=False):
        return self._ensure_device().sample_u(u_t, want_cells)

    def _sample_philox(self, N, seed, offset, want_cells=False, out=None, stratum=(0.0, 1.0)):
        return self._ensure_device().sample_philox(N, seed, offset, want_cells, out, stratum)

    def _sample_sobol(self, N, sobol_struct, want_cells=False):
        return self._ensure_device().sample_sobol(N, sobol_struct, want_cells)

    def _sample_cellmajor(self, N, seed, offset, want_cells=False, out=None, stratum=(0.0, 1.0)):
        return self._ensure_device().sample_cellmajor(N, seed, offset, want_cells, out, stratum)

    def choose_cells(self, u):
        """Reference return contract  [tree.py:314-355]; the search runs on the device."""
        import torch
        table = self._ensure_device()
        u = np.asarray(u, dtype=np.float64).reshape(-1)
        ufull = np.zeros((len(u), self._dim + 1))
        ufull[:, -1] = u
        _, cells = table.sample_u(torch.as_tensor(ufull, device=table.device), want_cells=True)
        idx = cells.cpu().numpy()
        x, dx, cd, _ = self.leaf_table()
        return x[idx], x[idx] + dx[idx], tuple(cd[idx].T)


class _CellTable:
    """Device copy of a flat cell list + its CDF/guide table (``lint_cells_t``)."""

    def __init__(self, mins, widths, corners, masses, dtype, device, cdf_mode):
        self.device = _device.require_cuda(device)
        self._build(mins, widths, corners, masses, dtype, cdf_mode)

    @_device.on_device
    def _build(self, mins, widths, corners, masses, dtype, cdf_mode):
        import torch
        dev = self.device
        self.dtype = np.dtype(dtype)
        self.dim = mins.shape[1]
        self.ncells = mins.shape[0]
        tdt = _device.torch_dtype(dtype)
        self._mins = torch.as_tensor(np.ascontiguousarray(mins), device=dev)
        self._widths = torch.as_tensor(np.ascontiguousarray(widths), device=dev)
        corners = np.ascontiguousarray(corners, dtype=np.float64)
        if self.dtype == np.float32 and corners.size:
            # only ratios of corner densities matter to the sampler: scale into fp32 range
            peak = float(np.max(corners))
            if np.isfinite(peak) and peak > 0.0:
                corners = corners * 2.0 ** -np.floor(np.log2(peak))
        self._corners = torch.as_tensor(corners, device=dev).to(tdt).contiguous()
        self._masses = torch.as_tensor(np.ascontiguousarray(masses), device=dev)
        self._cdf = None
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        desc = self.descriptor(need_cdf=False)
        _capi.call("lint_cells_check", C.byref(desc), _device.ptr(flags), _device.stream_ptr(dev))
        bad = int(flags.item())
        if bad & _capi.LINT_FLAG_NEGATIVE:
            raise ValueError("Cell list: Densities can't be negative")
        if bad & _capi.LINT_FLAG_NONFINITE:
            raise ValueError("Cell list: Detected non-finite density")
        self._cdf, self._guide, self._guide_bits, self.total_mass = _device.build_cdf(
            self._masses, cdf_mode, dev)
        nbytes = int(_capi.load().lint_cells_record_bytes(self.dim, _device.lint_dtype(self.dtype)))
        self._rec = torch.empty(self.ncells * nbytes, dtype=torch.uint8, device=dev)
        _capi.call("lint_cells_records", C.byref(self.descriptor(need_cdf=False, records=False)),
                   _device.ptr(self._rec), _device.stream_ptr(dev))

    def descriptor(self, need_cdf=True, records=True):
        d = _capi.LintCells()
        d.dim = self.dim
        d.dtype = _device.lint_dtype(self.dtype)
        d.ncells = self.ncells
        d.mins_dev = self._mins.data_ptr()
        d.widths_dev = self._widths.data_ptr()
        d.corners_dev = self._corners.data_ptr()
        if need_cdf:
            d.cdf_dev = self._cdf.data_ptr()
            d.guide_dev = self._guide.data_ptr()
            d.guide_bits = self._guide_bits
        rec = getattr(self, "_rec", None)
        d.cell_records_dev = rec.data_ptr() if (records and rec is not None) else None
        return d

    def _alloc(self, N, want_cells, out=None):
        import torch
        if out is None:
            out = torch.empty((N, self.dim), dtype=_device.torch_dtype(self.dtype), device=self.device)
        cells = torch.empty(N, dtype=torch.int64, device=self.device) if want_cells else None
        return out, cells

    @_device.on_device
    def sample_u(self, u_t, want_cells=False):
        N = u_t.shape[0]
        out, cells = self._alloc(N, want_cells)
        desc = self.descriptor()
        _capi.call("lint_cells_sample_u", C.byref(desc), _device.ptr(u_t), N, _device.ptr(out),
                   _device.ptr(cells), _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    @_device.on_device
    def sample_philox(self, N, seed, offset, want_cells=False, out=None, stratum=(0.0, 1.0)):
        out, cells = self._alloc(N, want_cells, out)
        desc = self.descriptor()
        _capi.call("lint_cells_sample_philox", C.byref(desc), N, C.c_uint64(seed), C.c_uint64(offset),
                   C.c_double(stratum[0]), C.c_double(stratum[1]), _device.ptr(out), _device.ptr(cells), _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    @_device.on_device
    def sample_cellmajor(self, N, seed, offset, want_cells=False, out=None, stratum=(0.0, 1.0)):
        """Rows grouped by leaf (list order); ``lint_cells_sample_cellmajor``."""
        out, cells = self._alloc(N, want_cells, out)
        desc = self.descriptor()
        for start, rows in _device.cellmajor_chunks(N):
            scratch, nbytes = _device.cellmajor_scratch(self, self.ncells, rows)
            _capi.call("lint_cells_sample_cellmajor", C.byref(desc), rows, C.c_uint64(seed),
                       C.c_uint64(offset + start), C.c_double(stratum[0]), C.c_double(stratum[1]),
                       _device.ptr(scratch), nbytes, _device.ptr(out[start:start + rows]),
                       _device.ptr(cells[start:start + rows] if cells is not None else None),
                       _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    @_device.on_device
    def sample_sobol(self, N, sobol_struct, want_cells=False):
        out, cells = self._alloc(N, want_cells)
        desc = self.descriptor()
        _capi.call("lint_cells_sample_sobol", C.byref(desc), N, C.byref(sobol_struct), _device.ptr(out),
                   _device.ptr(cells), _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out
