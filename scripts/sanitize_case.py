#!/usr/bin/env python
"""Small end-to-end case for a checked run: hierarchical build (whole domain and a stratum),
one order="cell" draw and one order="draw" draw of a 3-D grid in fp32 and fp64, a cell list
(tree leaves), the row permutation.

    LINTSAMPLER_B200_LIB=$(bash scripts/checked_build.sh) python scripts/sanitize_case.py   # device asserts
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py                       # where the tool is open
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb  # noqa: E402

rng = np.random.default_rng(0)
edges = tuple(np.linspace(-4, 4, n + 1) for n in (24, 16, 40))
gmm = lb.GaussianMixture(rng.dirichlet(np.ones(3)), rng.uniform(-2, 2, (3, 3)), sigmas=rng.uniform(0.3, 0.8, 3))
for dtype in (np.float32, np.float64):
    grid = lb.DensityGrid(edges, gmm, vectorizedpdf=True, dtype=dtype)
    N = 600_000                                     # > 32 rows per cell: the flat / residual split is on
    x, cells = grid._sample_cellmajor(N, 3, 0, want_cells=True)
    y = grid._sample_philox(50_000, 3, 0)
    grid._enqueue_build(stratum=(0.2, 0.7))
    z = grid._sample_cellmajor(100_000, 3, 7, stratum=(0.25, 0.6))
    grid._enqueue_build()                           # back to the whole domain
    s = lb.LintSampler(grid, seed=1, order="cell", shuffle=True).sample(20_000)
    torch.cuda.synchronize()
    assert x.shape == (N, 3) and bool(torch.isfinite(x).all()) and bool(torch.isfinite(z).all())
    assert int(cells.min()) >= 0 and int(cells.max()) < grid.ncells
tree = lb.DensityTree(np.full(2, -4.0), np.full(2, 4.0), pdf=lambda p: np.exp(-0.5 * (p ** 2).sum(-1)),
                      vectorizedpdf=True, min_openings=3)
t = lb.LintSampler(tree, seed=2, order="cell").sample(50_000)
u = lb.LintSampler(tree, seed=2).sample(5_000)
torch.cuda.synchronize()
print("sanitize_case ok", float(x.mean()), float(np.mean(t)), float(np.mean(u)))
