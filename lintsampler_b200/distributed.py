"""Multi-GPU sampling: one process per GPU, the domain sharded by contiguous
ranges of the flat (dim-0-major) cell index — i.e. by slabs of the grid.

The reference has no distributed code; its nearest construct is the list of
domains (lintsampler.py:281-307): a domain is chosen with probability
proportional to its mass and the cell-selecting uniform is rescaled into the
domain's CDF interval.  A sharded run is exactly that with one domain per
rank: rank g owns the cells whose CDF interval intersects its stratum
``[b_g, b_{g+1})`` of the cell-selecting uniform and samples ``N_g`` rows with
``u_cell`` drawn from that stratum.

Strata are cut at equal mass (``b_g = g / world``) rather than at equal plane
counts, because per-rank work is proportional to the shard's *mass*, not its
cell count; a slab cut by plane index leaves the ranks that own the tails of
the density idle.  The per-shard masses are exchanged with one NCCL all-gather
of 8 bytes per rank per build, and ``N`` is split multinomially over them with a
seed shared by all ranks, so no sample data ever crosses NVLink.
"""
import numpy as np

from . import _device
from .grid import DensityGrid


def split_counts(n_total, shard_masses, seed):
    """``(N_0..N_{G-1}) ~ Multinomial(N, M / sum M)`` drawn identically on every
    rank from a shared seed (lintsampler.py:286-290 draws the same split one
    row at a time through ``_choice``)."""
    p = np.asarray(shard_masses, dtype=np.float64)
    p = p / p.sum()
    return np.random.default_rng(seed).multinomial(int(n_total), p)


def all_gather_masses(local_mass, world, group=None, out=None):
    """All-gather one fp64 scalar per rank (NCCL on device tensors, gloo on CPU
    tensors): the per-shard masses.  Returns a (world,) tensor on the same
    device; asynchronous with respect to the host for device tensors."""
    import torch
    local_mass = local_mass.reshape(1).to(torch.float64)
    if out is None:
        out = torch.empty(world, dtype=torch.float64, device=local_mass.device)
    if world > 1:
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, local_mass, group=group)
    else:
        out.copy_(local_mass)
    return out


def _collectives_available():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def route_rows(u_cell, bounds):
    """Parity-mode routing of caller-supplied uniforms: rank of each row and its
    cell uniform rescaled into that rank's stratum, exactly the list-of-domains
    arithmetic of lintsampler.py:290-306 (``_choice`` on the normalised
    cumulative masses, then ``(u - start) / (end - start)``)."""
    u_cell = np.asarray(u_cell, dtype=np.float64)
    p = np.diff(np.asarray(bounds, dtype=np.float64))
    c = p.cumsum()
    rank = (c / c[-1]).searchsorted(u_cell, side="right")
    edges = np.append(0, p.cumsum())
    return rank, (u_cell - edges[rank]) / (edges[rank + 1] - edges[rank])


def equal_mass_strata(world):
    """Stratum boundaries of the cell-selecting uniform, one stratum per rank."""
    return np.arange(world + 1, dtype=np.float64) / world


# Cost model of one rank's step (ns), measured on B200 at the 256^3 / N = 1e9 benchmark
# (profiles/README.md, r2g): a row of a split cell (94 % flat rows at ~1.9 ns, 6 % residual rows at
# ~6.4 ns), a row of a cell too small to split, and the per-cell work of the count / scan / build passes.
COST_NS_ROW_SPLIT, COST_NS_ROW_SMALL, COST_NS_CELL = 2.2, 6.4, 0.04
SPLIT_MIN_LAMBDA = 128.0          # csrc/lint_cellmajor_plan.h


def _group_end_values(grid):
    """CDF value at the end of every group of the grid's build (host array): the places where a
    stratum can be cut without a straddling cell."""
    import ctypes as C
    import torch
    from . import _capi
    ncells = int(grid.ncells)
    gc = int(_capi.load().lint_grid_group_cells(C.byref(grid._descriptor())))
    idx = torch.arange(gc - 1, ncells, gc, device=grid.device)
    if idx.numel() == 0 or int(idx[-1]) != ncells - 1:
        idx = torch.cat([idx, torch.tensor([ncells - 1], device=grid.device)])
    return grid._cdf[idx].cpu().numpy()


def cost_balanced_strata(grid, world, n_rows):
    """Stratum boundaries that equalise the modelled cost of the ranks instead of their mass.

    Equal-mass strata give every rank the same number of rows but very different numbers of
    cells: the ranks that own the tails of the density visit several times more cells (count,
    scan and build passes) and draw most of their rows from cells too small to split (the dearer
    emitter).  The cuts are placed at group boundaries of the grid's build (no cell straddles a
    cut) where the cumulative cost ``rows * c_row + cells * c_cell`` crosses g / world; every rank
    derives the same cuts from the same CDF entries, without communication.  ``n_rows``: the
    number of rows the whole domain will draw per call (sets rows per cell, hence which cells are
    split)."""
    import ctypes as C
    import torch
    from . import _capi
    ncells = int(grid.ncells)
    gc = int(_capi.load().lint_grid_group_cells(C.byref(grid._descriptor())))
    if world == 1 or gc <= 0 or ncells <= gc:
        return equal_mass_strata(world)
    ends_idx = torch.arange(gc - 1, ncells, gc, device=grid.device)
    if int(ends_idx[-1]) != ncells - 1:
        ends_idx = torch.cat([ends_idx, torch.tensor([ncells - 1], device=grid.device)])
    ends = grid._cdf[ends_idx].cpu().numpy()                       # CDF at the end of every group
    cells = np.diff(np.concatenate([[0], ends_idx.cpu().numpy() + 1])).astype(np.float64)
    rows = float(n_rows) * np.diff(np.concatenate([[0.0], ends]))
    lam = rows / cells
    cost = rows * np.where(lam >= SPLIT_MIN_LAMBDA, COST_NS_ROW_SPLIT, COST_NS_ROW_SMALL) + cells * COST_NS_CELL
    cum = np.cumsum(cost)
    cuts = np.searchsorted(cum, cum[-1] * np.arange(1, world) / world)
    bounds = np.concatenate([[0.0], ends[np.minimum(cuts, len(ends) - 1)], [1.0]])
    if np.any(np.diff(bounds) <= 0):
        return equal_mass_strata(world)
    return bounds


def slab_bounds(n0, world):
    """Cell-plane ranges along dim 0, one contiguous slab per rank:
    rank g owns cell planes [b[g], b[g+1]) and vertex planes [b[g], b[g+1]]."""
    return [(g * n0) // world for g in range(world + 1)]


class SlabShardedGrid:
    """The literal slab decomposition: rank g evaluates, reduces and scans ONLY its
    own slab of dim-0 planes (one halo vertex plane, re-evaluated locally), the slab
    masses are all-gathered (NCCL, 8 B per rank), N is split multinomially over
    them with a shared seed and every rank samples its share from its own slab.

    Work per rank is proportional to the slab's mass, so a peaked density leaves the
    ranks that own its tails idle; ``ShardedGrid`` (equal-mass strata) is the
    balanced alternative and what ``bench.py`` uses.  ``pdf`` must be evaluable on a
    sub-grid (any callable or ``GaussianMixture``).
    """

    def __init__(self, edges, pdf, *, dtype=np.float64, device=None, rank=0, world=1, group=None,
                 vectorizedpdf=True, cdf_mode="parallel", **grid_kw):
        self.rank, self.world, self.group = int(rank), int(world), group
        edges = [np.asarray(e) for e in (edges if isinstance(edges, tuple) else (edges,))]
        n0 = len(edges[0]) - 1
        if n0 < world:
            raise ValueError("SlabShardedGrid: fewer cell planes along dim 0 than ranks")
        b = slab_bounds(n0, world)
        self.planes = (b[self.rank], b[self.rank + 1])
        sub = [edges[0][b[self.rank]:b[self.rank + 1] + 1]] + edges[1:]
        dom = tuple(sub) if len(sub) > 1 else sub[0]
        self.grid = DensityGrid(dom, pdf, vectorizedpdf=vectorizedpdf, dtype=dtype, device=device,
                                cdf_mode=cdf_mode, **grid_kw)
        self.device = self.grid.device
        self._gather_buf = None
        self.shard_masses = self.exchange_masses()

    def exchange_masses(self):
        """NCCL all-gather of the slab masses (unscaled), as a NumPy array."""
        mine = self.grid._total / self.grid._scale
        self._gather_buf = all_gather_masses(mine, self.world, self.group, self._gather_buf)
        return self._gather_buf.cpu().numpy().copy()

    def sample(self, n_total, seed, offset=0, order="draw"):
        """This rank's rows of an ``n_total``-row draw: (device tensor (N_g, k), counts)."""
        import torch
        counts = split_counts(n_total, self.shard_masses, seed)
        n = int(counts[self.rank])
        start = int(counts[:self.rank].sum())
        out = torch.empty((n, self.grid.dim), dtype=_device.torch_dtype(self.grid.dtype), device=self.device)
        if n:
            if order == "cell":
                self.grid._sample_cellmajor(n, seed, offset + start, out=out)
            else:
                self.grid._sample_philox(n, seed, offset + start, out=out)
        return out, counts


class ShardedGrid:
    """A ``DensityGrid`` whose sampling work is sharded over ``world`` ranks.

    Every rank holds the (small) vertex-density table and builds the cell CDF;
    rank ``g`` samples only the cells of its equal-mass stratum.  With
    ``world == 1`` this is a plain ``DensityGrid``.
    """

    def __init__(self, edges, pdf, *, dtype=np.float64, device=None, rank=0, world=1,
                 group=None, vectorizedpdf=True, cdf_mode="parallel", balance_rows=None, **grid_kw):
        self.rank, self.world, self.group = int(rank), int(world), group
        self.grid = DensityGrid(edges, pdf, vectorizedpdf=vectorizedpdf, dtype=dtype, device=device,
                                cdf_mode=cdf_mode, **grid_kw)
        self.device = self.grid.device
        # strata of the cell-selecting uniform, one per rank: equal mass, or (balance_rows = rows
        # per call over the whole domain) equal modelled cost — see cost_balanced_strata
        self.balanced = bool(balance_rows) and cdf_mode == "parallel" and self.world > 1
        self.bounds = (cost_balanced_strata(self.grid, self.world, balance_rows) if self.balanced
                       else equal_mass_strata(self.world))
        self.shard_masses = None
        self._gather_buf = None
        self.enqueue_build(gather=False)            # restrict the tables to this rank's stratum
        self.exchange_masses()

    # -- build -------------------------------------------------------------
    def enqueue_build(self, gather=True):
        """This rank's share of the build on the current stream: every rank reduces the vertex
        table to tile totals (total mass and tile offsets, identical on all ranks, no
        communication) and scans only the slab of cells of its own stratum (the parallel CDF;
        ``cdf_mode="sequential"`` builds the whole table).  ``gather``: also all-gather the
        per-shard masses over NCCL (8 bytes per rank, asynchronous; equal-mass strata make it
        redundant for the split of N, so the benchmark leaves it out)."""
        lo, hi = float(self.bounds[self.rank]), float(self.bounds[self.rank + 1])
        if self.grid.cdf_mode == "parallel":
            self.grid._enqueue_build(stratum=(lo, hi))
        else:
            self.grid._enqueue_build()
        if gather:
            self._enqueue_gather()

    def _stratum_mass_tensor(self):
        """This rank's shard mass as a device scalar: (b_{g+1} - b_g) * total."""
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        return self.grid._total * ((hi - lo) / self.grid._scale)

    def _enqueue_gather(self):
        # NCCL all-gather, 8 bytes per rank.  Without a process group (ranks emulated one after the
        # other in one process, as the single-GPU tests do) the masses are filled in locally: with
        # equal-mass strata every rank can compute all of them from the total.
        if self.world > 1 and not _collectives_available():
            import torch
            widths = torch.as_tensor(np.diff(self.bounds), dtype=torch.float64, device=self.device)
            self._gather_buf = self.grid._total * (widths / self.grid._scale)
            return
        self._gather_buf = all_gather_masses(self._stratum_mass_tensor(), self.world, self.group,
                                             self._gather_buf)

    def exchange_masses(self):
        """Synchronous version: returns the per-shard masses as a NumPy array."""
        self._enqueue_gather()
        self.shard_masses = self._gather_buf.cpu().numpy().copy()
        return self.shard_masses

    # -- sampling ------------------------------------------------------------
    def sample_cellmajor_into(self, out, seed, offset):
        """Throughput mode: this rank's rows grouped by cell (see DensityGrid._sample_cellmajor).
        ``offset``: this rank's position in the Philox stream — ranks must pass disjoint ranges
        ``[offset, offset + len(out))`` (e.g. the prefix sums of the split of N)."""
        n = out.shape[0]
        lo, hi = float(self.bounds[self.rank]), float(self.bounds[self.rank + 1])
        self.grid._sample_cellmajor(n, seed, offset, out=out, stratum=(lo, hi))
        return out

    def sample_philox_into(self, out, seed, offset):
        """Fill the device tensor ``out`` (n, k) with this rank's rows.  Row i uses Philox
        counter ``offset + i``: ranks must pass disjoint ranges ``[offset, offset + n)``."""
        n = out.shape[0]
        lo, hi = float(self.bounds[self.rank]), float(self.bounds[self.rank + 1])
        self.grid._sample_philox(n, seed, offset, out=out, stratum=(lo, hi))
        return out

    def sample(self, n_total, seed, offset=0):
        """Draw this rank's share of ``n_total`` samples (multinomial split with a
        shared seed); returns a device tensor (N_g, k)."""
        import torch
        counts = split_counts(n_total, self.exchange_masses(), seed)
        n = int(counts[self.rank])
        start = int(counts[:self.rank].sum())
        out = torch.empty((n, self.grid.dim), dtype=_device.torch_dtype(self.grid.dtype),
                          device=self.device)
        if n:
            lo, hi = float(self.bounds[self.rank]), float(self.bounds[self.rank + 1])
            self.grid._sample_philox(n, seed, offset + start, out=out, stratum=(lo, hi))
        return out, counts

    # -- reporting -------------------------------------------------------------
    # -- calibration -----------------------------------------------------------
    def calibrate(self, n_rows, seed=0, iters=4, fixed_ms=0.07):
        """Re-cut the strata from MEASURED step times (multi-GPU, collective).  The analytic cost
        model of ``cost_balanced_strata`` misjudges the ranks that own the tails of the density
        (their rows are dearer than the model's average).  Every iteration each rank times its
        own step (build of its slab + its share of ``n_rows`` grouped rows) with CUDA events, the
        times are all-gathered, the cost per unit of CDF mass is taken as constant inside each
        current stratum and the cuts are moved to equalise it, snapped to group boundaries of the
        build.  ``fixed_ms``: the part of a step that does not shrink with the stratum.  All ranks
        compute the same cuts from the same gathered numbers."""
        import torch
        import torch.distributed as dist
        if self.world == 1 or not _collectives_available() or self.grid.cdf_mode != "parallel":
            return self.bounds
        ends = _group_end_values(self.grid)
        for it in range(iters):
            widths = np.diff(self.bounds)
            counts = split_counts(n_rows, widths, seed + it)
            n = int(counts[self.rank])
            out = torch.empty((max(n, 1), self.grid.dim), dtype=_device.torch_dtype(self.grid.dtype),
                              device=self.device)
            first = int(counts[:self.rank].sum())

            def step():
                self.enqueue_build(gather=False)
                self.sample_cellmajor_into(out[:n], seed=seed, offset=first)

            step()                                                 # scratch allocated, kernels loaded
            torch.cuda.synchronize(self.device)
            # timed as the benchmark runs it: one CUDA graph replay per step (eager launches add
            # host-side gaps that differ from rank to rank and hide the differences that matter)
            graph = None
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step()
            except Exception:
                graph = None
            run = graph.replay if graph is not None else step
            for _ in range(2):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 8
            a.record()
            for _ in range(reps):
                run()
            b.record()
            torch.cuda.synchronize(self.device)
            best = a.elapsed_time(b) / reps
            del graph
            times = torch.tensor([best], dtype=torch.float64, device=self.device)
            gathered = torch.empty(self.world, dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(gathered, times, group=self.group)
            t = np.maximum(gathered.cpu().numpy() - fixed_ms, 1e-3)
            # piecewise-constant cost density t_r / width_r: cumulative cost at the current cuts
            cum = np.concatenate([[0.0], np.cumsum(t)])
            targets = cum[-1] * np.arange(1, self.world) / self.world
            cuts = np.interp(targets, cum, self.bounds)                    # new cuts in u
            snapped = ends[np.clip(np.searchsorted(ends, cuts), 0, len(ends) - 1)]
            new = np.concatenate([[0.0], snapped, [1.0]])
            if np.any(np.diff(new) <= 0):
                break
            self.bounds = new
            self.balanced = True
            self.enqueue_build(gather=False)
        torch.cuda.synchronize(self.device)
        return self.bounds

    def launches_per_step(self, order="draw"):
        """Kernels of this library launched by enqueue_build + one sampling call: build 3 (vertex
        row sums, group totals, group scan) + 1 for the guide table when there is one, cell records
        0/1; then the u-driven sampler
        (1) or the grouped path (Poisson counts, 2 x scan+compact, row budget + chunk starts, flat
        boxes, flat emitter, general emitter, top-up = 8)."""
        build = 3 + (1 if self.grid._guide_bits_planned > 0 else 0) + (1 if self.grid._rec is not None else 0)
        return build + (8 if order == "cell" else 1)

    def describe(self):
        if self.world == 1:
            return "single GPU"
        kind = "cost-balanced strata (cut at group boundaries of the build)" if self.balanced else "equal-mass strata"
        return (f"{self.world} ranks, {kind} of the flat cell CDF (contiguous dim-0-major "
                "cell ranges): every rank reduces the vertex table to tile totals, scans only its own slab; "
                "N split multinomially with a shared seed; no sample data crosses NVLink")
