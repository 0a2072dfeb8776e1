"""``DensityGrid``: rectilinear-grid domain, device resident.

Host-side mirror of the reference class (density_structures/grid.py:6-518):
same constructor, attributes, exceptions and ``choose_cells`` contract.  The
vertex densities, cell masses, cell CDF and guide table live in HBM as torch
tensors; every array step of the reference is one C-ABI kernel call.
"""
import ctypes as C

import numpy as np

from . import _capi, _device
from .base import DensityStructure
from .pdfs import FUSED_PDFS


def torch_range(ncells, device):
    """The whole-domain slab in lint_grid_build's layout (the cell-major sampler reads the cell range)."""
    import torch
    return torch.tensor([0, ncells, 0, 0], dtype=torch.int32, device=device)


def _is_flat_sequence(obj):
    """True for a sequence whose first element is not itself a sequence."""
    return hasattr(obj, "__len__") and not hasattr(obj[0], "__len__")


class _TensorDensities:
    """Marker wrapping pre-evaluated vertex densities held in a torch tensor."""

    def __init__(self, t):
        self.t = t

    def __call__(self, *a, **k):
        raise TypeError("pre-evaluated densities are not callable")


class DensityGrid(DensityStructure):
    """Rectilinear (not necessarily uniform) k-D grid with the pdf evaluated on
    its vertices.

    Parameters (as the reference, grid.py:253)
    ----------
    edges : 1-D sequence (k = 1) or tuple of k 1-D sequences of cell edges.
    pdf : callable returning the (unnormalised) density; a
        ``lintsampler_b200.GaussianMixture`` or ``DoublePowerLaw`` is evaluated by a fused kernel.
    vectorizedpdf, pdf_args, pdf_kwargs : as the reference.

    Keyword-only extensions (defaults reproduce the reference)
    ----------
    dtype : np.float64 (default) or np.float32 — storage of the vertex
        densities, arithmetic of the in-cell solve and dtype of the samples.
        The cell CDF is fp64 either way.
    device : torch CUDA device (default: current device).
    cdf_mode : "parallel" (default) or "sequential" (bit-identical to NumPy's
        ``cumsum``; slower, used for exact parity).
    pdf_on_device : if True the vectorised ``pdf`` is called with a torch CUDA
        tensor of shape (npts, k) and must return a torch tensor.
    guide_bits : log2 of the guide-table size (default: about one bin per cell).
        0: no guide table — enough for ``order="cell"``, whose only searches
        are its ~2.5e-4 N top-up rows.
    cell_records : keep a per-cell copy of the 2^k corner densities so a sample
        reads one contiguous record (default: on when it needs < 8 GiB).
    """

    def __init__(self, edges, pdf, vectorizedpdf=False, pdf_args=(), pdf_kwargs={},
                 *, dtype=np.float64, device=None, cdf_mode="parallel", pdf_on_device=False,
                 guide_bits=None, cell_records=None):
        # --- domain parsing: one flat sequence, or a tuple of them  [grid.py:256-285]
        if _is_flat_sequence(edges):
            arrays = [np.array(edges)]
        elif isinstance(edges, tuple) and _is_flat_sequence(edges[0]):
            arrays = [np.array(e) for e in edges]
        else:
            raise TypeError(
                "DensityGrid: the evaluation domain must be a single 1D iterable "
                "or a tuple of 1D iterables.")
        for arr in arrays:                                     # [grid.py:288-296]
            if np.any(np.diff(arr) <= 0):
                raise ValueError("DensityGrid: Edges not monotically increasing.")
            if not np.all(np.isfinite(arr)):
                raise ValueError("DensityGrid: Edges not finite-valued.")
        self._dim = len(arrays)
        if self._dim > _capi.LINT_MAX_DIM:
            raise ValueError(f"DensityGrid: at most {_capi.LINT_MAX_DIM} dimensions are supported")
        self.edgearrays = arrays
        self._edgeshape = tuple(len(a) for a in arrays)
        self._mins = np.array([a[0] for a in arrays])
        self._maxs = np.array([a[-1] for a in arrays])
        self.shape = tuple(n - 1 for n in self._edgeshape)
        self.ncells = np.prod(self.shape)
        self.dtype = _device.check_dtype(dtype)
        self.cdf_mode = cdf_mode
        self._device_arg = device
        self._guide_bits_arg = guide_bits
        self._cell_records_arg = cell_records
        self._host_densities = None
        self._masses_host = None
        self._evaluate_pdf(pdf, vectorizedpdf, pdf_args, pdf_kwargs, pdf_on_device)

    @classmethod
    def from_densities(cls, edges, densities, *, dtype=np.float64, device=None,
                       cdf_mode="parallel"):
        """Grid over vertex densities that were already evaluated: ``densities`` is
        an array (NumPy, or a torch tensor in host/pinned/device memory) of shape
        ``(n0+1, ..., nk-1+1)`` (or flat).  Same validation as the constructor;
        the density checks of grid.py:401-404 run on the device."""
        import torch
        if isinstance(densities, torch.Tensor):
            self = cls.__new__(cls)
            cls.__init__(self, edges, _TensorDensities(densities), vectorizedpdf=True,
                         dtype=dtype, device=device, cdf_mode=cdf_mode, pdf_on_device=True)
            return self
        table = np.asarray(densities, dtype=np.float64)
        return cls(edges, lambda pts: table.reshape(-1), vectorizedpdf=True, dtype=dtype,
                   device=device, cdf_mode=cdf_mode)

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
    def total_mass(self):
        if self._total_stale:                 # an unchecked update_densities left it behind
            self._check_build()
        return self._total_mass

    @property
    def vertex_densities(self):
        """(n0+1, ..., nk-1+1) NumPy array of vertex densities (fp64)."""
        if self._host_densities is None:
            self._host_densities = (self._V.cpu().numpy().astype(np.float64)
                                    / self._scale).reshape(self._edgeshape)
        return self._host_densities

    @property
    def masses(self):
        """Cell masses, shape ``self.shape`` (NumPy, fp64)  [grid.py:411]."""
        if self._masses_host is None:
            self._masses_host = (self._masses.cpu().numpy() / self._scale).reshape(self.shape)
        return self._masses_host

    # ------------------------------------------------------------------ build
    def _evaluate_pdf(self, pdf, vectorizedpdf, pdf_args, pdf_kwargs, pdf_on_device):
        """Vertex evaluation [grid.py:358-404] then the device build."""
        npts = int(np.prod(self._edgeshape))
        fused = isinstance(pdf, FUSED_PDFS) and not pdf_args and not pdf_kwargs
        if fused and pdf.dim != self._dim:
            raise ValueError(f"DensityGrid: {type(pdf).__name__} dimensionality does not match the grid")
        densities = None
        dev_densities = None
        if fused:
            pass                                    # evaluated by its fused kernel below
        elif isinstance(pdf, _TensorDensities):
            import torch
            dev = _device.require_cuda(self._device_arg)
            if pdf.t.numel() != npts:
                raise ValueError("DensityGrid: densities have the wrong size for these edges")
            dev_densities = pdf.t.reshape(npts).to(device=dev, non_blocking=True)
        elif vectorizedpdf and pdf_on_device:
            import torch
            dev = _device.require_cuda(self._device_arg)
            axes = [torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev) for a in self.edgearrays]
            if self._dim > 1:
                pts = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(npts, self._dim)
            else:
                pts = axes[0]
            f = pdf(pts, *pdf_args, **pdf_kwargs)
            if not isinstance(f, torch.Tensor):
                raise TypeError("DensityGrid: with pdf_on_device=True the pdf must return a torch tensor")
            if f.numel() != npts:
                raise ValueError("DensityGrid: pdf returned an array of the wrong size")
            dev_densities = f.to(device=dev, dtype=torch.float64).reshape(npts)
        else:
            if self._dim > 1:
                mesh = np.stack(np.meshgrid(*self.edgearrays, indexing='ij'), axis=-1)
                pts = mesh.reshape(npts, self._dim)
            else:
                pts = self.edgearrays[0]
            if vectorizedpdf:
                densities = np.array(pdf(pts, *pdf_args, **pdf_kwargs), dtype=np.float64)
            else:
                densities = np.zeros(npts, dtype=np.float64)
                for i in range(npts):
                    densities[i] = pdf(pts[i], *pdf_args, **pdf_kwargs)
            # validity, on the host because the values are already here  [grid.py:401-404]
            if np.any(densities < 0):
                raise ValueError("DensityGrid: Densities can't be negative")
            if not np.all(np.isfinite(densities)):
                raise ValueError("DensityGrid: Detected non-finite density")
            densities = densities.reshape(self._edgeshape)      # ValueError on a bad shape
            self._host_densities = densities
        self._build_device(pdf if fused else None, densities, dev_densities)

    def _build_device(self, gmm, densities, dev_densities):
        self.device = _device.require_cuda(self._device_arg)
        self._build_on_device(gmm, densities, dev_densities)

    @_device.on_device
    def _build_on_device(self, gmm, densities, dev_densities):
        import torch
        dev = self.device
        tdt = _device.torch_dtype(self.dtype)
        npts = int(np.prod(self._edgeshape))
        ncells = int(self.ncells)
        self._edges_t = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)
                         for a in self.edgearrays]
        # fp32 storage: densities are kept times a power of two that brings the largest
        # into [1, 2), so unnormalised pdfs in physical units (1e-40 .. 1e+40) neither
        # flush to zero nor overflow.  Sampling only sees ratios; masses are unscaled
        # again wherever they are reported.  (fp64: scale = 1.)
        self._scale = 1.0
        if self.dtype == np.float32:
            if densities is not None:
                peak = float(np.max(densities)) if densities.size else 0.0
            elif dev_densities is not None:
                peak = float(dev_densities.max().item())
            else:
                peak = gmm._peak_bound(self.edgearrays)            # upper bound of a fused pdf
            if np.isfinite(peak) and peak > 0.0:
                self._scale = float(2.0 ** -np.floor(np.log2(peak)))
        if densities is not None:
            self._V = torch.as_tensor(np.ascontiguousarray(densities).reshape(npts) * self._scale,
                                      device=dev).to(tdt)
        elif dev_densities is not None:
            self._V = (dev_densities * self._scale).to(tdt).contiguous()
        else:
            self._V = torch.empty(npts, dtype=tdt, device=dev)
        self._cdf = None
        self._guide = None
        self._guide_bits = 0
        st = _device.stream_ptr(dev)
        desc = self._descriptor()
        self._gmm = gmm
        if gmm is not None:
            gmm._eval_on_grid(desc, self._scale, st)
        self._masses_t = None                       # cell masses on the device: computed on demand
        self._flags = torch.zeros(1, dtype=torch.int32, device=dev)
        self._range = torch.zeros(4, dtype=torch.int32, device=dev)     # slab of the last build (lint_grid_build)
        self._stratum = (0.0, 1.0)
        self._cdf = torch.empty(ncells, dtype=torch.float64, device=dev)
        self._total = torch.empty(1, dtype=torch.float64, device=dev)
        self._guide_bits_planned = (int(self._guide_bits_arg) if self._guide_bits_arg is not None
                                    else _device.default_guide_bits(ncells))
        rec_bytes = ncells * (2 ** self._dim) * np.dtype(self.dtype).itemsize
        want_rec = self._cell_records_arg if self._cell_records_arg is not None else rec_bytes < (8 << 30)
        self._rec = torch.empty(ncells * 2 ** self._dim, dtype=tdt, device=dev) if want_rec else None
        self._guide = torch.empty((1 << self._guide_bits_planned) + 1, dtype=torch.int32, device=dev)
        nbytes = max(_capi.load().lint_cdf_scratch_bytes(ncells),
                     _capi.load().lint_grid_build_scratch_bytes(C.byref(self._descriptor())))
        self._scratch = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=dev)
        self._enqueue_build()
        self._check_build()

    @property
    def _masses(self):
        """Device tensor of the cell masses (scaled like the stored densities).  The parallel
        build never stores them; they are computed when somebody asks (grid.py:411)."""
        if self._masses_t is None:
            import torch
            with torch.cuda.device(self.device):
                self._masses_t = torch.empty(int(self.ncells), dtype=torch.float64, device=self.device)
                flags = torch.zeros(1, dtype=torch.int32, device=self.device)
                desc = self._descriptor()
                _capi.call("lint_grid_masses", C.byref(desc), _device.ptr(self._masses_t), _device.ptr(flags),
                           _device.stream_ptr(self.device))
        return self._masses_t

    @_device.on_device
    def _enqueue_build(self, stratum=(0.0, 1.0)):
        """Cell masses -> CDF -> guide table, enqueued on the current stream with no
        host synchronisation (grid.py:411-412, 431; utils.py:195-196).

        ``cdf_mode="parallel"``: one fused call, masses computed on the fly and never stored;
        with ``stratum=(u_lo, u_hi)`` only the slab of cells that stratum of the cell-selecting
        uniform can touch is scanned (a rank of a sharded run).  ``"sequential"``: the
        NumPy-identical path over stored masses, whole domain."""
        dev = self.device
        st = _device.stream_ptr(dev)
        ncells = int(self.ncells)
        mode = {"parallel": _capi.LINT_CDF_PARALLEL, "sequential": _capi.LINT_CDF_SEQUENTIAL}.get(self.cdf_mode)
        if mode is None:
            raise ValueError(f"cdf_mode must be 'parallel' or 'sequential', got {self.cdf_mode!r}")
        self._guide_bits = 0
        desc = self._descriptor()
        self._flags.zero_()
        self._masses_t = None
        self._stratum = (float(stratum[0]), float(stratum[1]))
        if self._rec is not None:
            _capi.call("lint_grid_records", C.byref(desc), _device.ptr(self._rec), st)
        if mode == _capi.LINT_CDF_PARALLEL:
            _capi.call("lint_grid_build", C.byref(desc), C.c_double(self._stratum[0]), C.c_double(self._stratum[1]),
                       _device.ptr(self._cdf), _device.ptr(self._total), _device.ptr(self._guide),
                       self._guide_bits_planned, None, _device.ptr(self._flags), _device.ptr(self._range),
                       _device.ptr(self._scratch), self._scratch.numel(), st)
        else:
            if self._stratum != (0.0, 1.0):
                raise ValueError("DensityGrid: a stratum-restricted build needs cdf_mode='parallel'")
            import torch
            self._masses_t = masses = torch.empty(ncells, dtype=torch.float64, device=dev)
            _capi.call("lint_grid_masses", C.byref(desc), _device.ptr(masses), _device.ptr(self._flags), st)
            _capi.call("lint_cdf_build", _device.ptr(masses), ncells, mode, _device.ptr(self._cdf),
                       _device.ptr(self._total), _device.ptr(self._scratch), self._scratch.numel(), st)
            if self._guide_bits_planned > 0:
                _capi.call("lint_guide_build", _device.ptr(self._cdf), ncells, self._guide_bits_planned,
                           _device.ptr(self._guide), st)
            self._range.copy_(torch_range(ncells, dev))
        self._guide_bits = self._guide_bits_planned
        self._masses_host = None
        self._total_stale = True

    @_device.on_device
    def update_densities(self, densities, check=True):
        """Replace the vertex densities (same edges) and rebuild masses, CDF and
        guide table on the current stream.  ``densities`` may be a NumPy array or
        a torch tensor in host (ideally pinned) or device memory.  With
        ``check=False`` nothing synchronises the host; the validity flags are
        then read by the next checked call."""
        import torch
        npts = self._V.numel()
        t = densities if isinstance(densities, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(densities))
        if t.numel() != npts:
            raise ValueError("DensityGrid.update_densities: wrong number of vertex densities")
        if self._scale != 1.0:
            t = t.to(self.device, non_blocking=True).to(torch.float64) * self._scale   # keeps the build's scale
        self._V.copy_(t.reshape(npts), non_blocking=True)       # H2D (+ cast) on the current stream
        self._host_densities = None
        self._enqueue_build()
        if check:
            self._check_build()

    def _check_build(self):
        bad = int(self._flags.item())                          # synchronises
        if bad & _capi.LINT_FLAG_NEGATIVE:                     # [grid.py:401-402]
            raise ValueError("DensityGrid: Densities can't be negative")
        if bad & _capi.LINT_FLAG_NONFINITE:                    # [grid.py:403-404]
            raise ValueError("DensityGrid: Detected non-finite density")
        self._total_mass = float(self._total.item()) / self._scale
        self._total_stale = False
        if not (self._total_mass > 0.0) or not np.isfinite(self._total_mass):
            raise ValueError(
                "DensityGrid: total mass of the grid is not positive and finite "
                f"({self._total_mass}); cannot build a cell CDF")

    def _descriptor(self):
        d = _capi.LintGrid()
        d.dim = self._dim
        d.dtype = _device.lint_dtype(self.dtype)
        for i, n in enumerate(self.shape):
            d.ncells[i] = n
            d.edges_dev[i] = self._edges_t[i].data_ptr()
        d.vertex_densities_dev = self._V.data_ptr()
        d.cdf_dev = self._cdf.data_ptr() if self._cdf is not None else None
        d.guide_dev = self._guide.data_ptr() if self._guide is not None else None
        d.guide_bits = self._guide_bits
        rec = getattr(self, "_rec", None)
        d.cell_records_dev = rec.data_ptr() if rec is not None else None
        return d

    # ------------------------------------------------------------------ sampling entry points
    @_device.on_device
    def _sample_u(self, u_t, want_cells=False):
        """(N,k+1) fp64 device uniforms -> (N,k) device samples  [sampling.py:97-127]."""
        import torch
        N = u_t.shape[0]
        out = torch.empty((N, self._dim), dtype=_device.torch_dtype(self.dtype), device=self.device)
        cells = torch.empty(N, dtype=torch.int64, device=self.device) if want_cells else None
        desc = self._descriptor()
        _capi.call("lint_grid_sample_u", C.byref(desc), _device.ptr(u_t), N, _device.ptr(out),
                   _device.ptr(cells), _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    @_device.on_device
    def _sample_philox(self, N, seed, offset, want_cells=False, out=None, stratum=(0.0, 1.0)):
        import torch
        if not (self._stratum[0] <= stratum[0] and stratum[1] <= self._stratum[1]):
            raise ValueError(f"DensityGrid: the tables were built for the stratum {self._stratum}; "
                             f"cannot draw from {tuple(stratum)}")
        if self._guide_bits == 0 and self._stratum != (0.0, 1.0):
            # without a guide table the u-driven search brackets the whole CDF, and only the
            # stratum's slab of it was built
            raise ValueError("DensityGrid: u-driven draws from a stratum-restricted build need guide_bits > 0")
        if out is None:
            out = torch.empty((N, self._dim), dtype=_device.torch_dtype(self.dtype), device=self.device)
        cells = torch.empty(N, dtype=torch.int64, device=self.device) if want_cells else None
        desc = self._descriptor()
        _capi.call("lint_grid_sample_philox", C.byref(desc), N, C.c_uint64(seed), C.c_uint64(offset),
                   C.c_double(stratum[0]), C.c_double(stratum[1]), _device.ptr(out), _device.ptr(cells), _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    @_device.on_device
    def _sample_cellmajor(self, N, seed, offset, want_cells=False, out=None, stratum=(0.0, 1.0)):
        """Throughput mode: rows grouped by cell (C order); exact multinomial cell
        counts, no per-sample search (``lint_grid_sample_cellmajor``)."""
        import torch
        if not (self._stratum[0] <= stratum[0] and stratum[1] <= self._stratum[1]):
            raise ValueError(f"DensityGrid: the tables were built for the stratum {self._stratum}; "
                             f"cannot draw from {tuple(stratum)}")
        if out is None:
            out = torch.empty((N, self._dim), dtype=_device.torch_dtype(self.dtype), device=self.device)
        cells = torch.empty(N, dtype=torch.int64, device=self.device) if want_cells else None
        desc = self._descriptor()
        for start, rows in _device.cellmajor_chunks(N):      # rows grouped by cell within each piece
            scratch, nbytes = _device.cellmajor_scratch(self, int(self.ncells), rows)
            _capi.call("lint_grid_sample_cellmajor", C.byref(desc), rows, C.c_uint64(seed),
                       C.c_uint64(offset + start), C.c_double(stratum[0]), C.c_double(stratum[1]),
                       _device.ptr(self._range), _device.ptr(scratch), nbytes,
                       _device.ptr(out[start:start + rows]),
                       _device.ptr(cells[start:start + rows] if cells is not None else None),
                       _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    @_device.on_device
    def _sample_sobol(self, N, sobol_struct, want_cells=False):
        import torch
        out = torch.empty((N, self._dim), dtype=_device.torch_dtype(self.dtype), device=self.device)
        cells = torch.empty(N, dtype=torch.int64, device=self.device) if want_cells else None
        desc = self._descriptor()
        _capi.call("lint_grid_sample_sobol", C.byref(desc), N, C.byref(sobol_struct), _device.ptr(out),
                   _device.ptr(cells), _device.stream_ptr(self.device))
        return (out, cells) if want_cells else out

    # ------------------------------------------------------------------ DensityStructure API
    def choose_cells(self, u):
        """Host-facing cell choice with the reference's return contract
        [grid.py:325-356]: ``(mins (N,k), maxs (N,k), corners 2^k-tuple of (N,))``.
        The search itself runs on the device."""
        import torch
        u = np.asarray(u, dtype=np.float64).reshape(-1)
        N = len(u)
        k = self._dim
        ufull = np.zeros((N, k + 1))
        ufull[:, -1] = u
        _, cells = self._sample_u(torch.as_tensor(ufull, device=self.device), want_cells=True)
        flat = cells.cpu().numpy()
        multi = np.unravel_index(flat, self.shape)
        mins = np.stack([self.edgearrays[d][multi[d]] for d in range(k)], axis=-1)
        maxs = np.stack([self.edgearrays[d][multi[d] + 1] for d in range(k)], axis=-1)
        V = self.vertex_densities
        corners = []
        for c in range(2 ** k):
            off = [(c >> (k - 1 - d)) & 1 for d in range(k)]
            corners.append(V[tuple(multi[d] + off[d] for d in range(k))])
        return mins, maxs, tuple(corners)
