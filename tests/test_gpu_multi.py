"""Sharded sampling: the union of the ranks' draws must be one draw from the whole domain.

`test_emulated_ranks_*` plays the ranks one after the other on a single GPU (no process group:
equal-mass strata need no exchange), so it runs wherever the other GPU tests run;
`test_two_gpu_nccl_run` launches the real thing with torchrun when two GPUs are visible."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import lint_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import lintsampler_b200
    return lintsampler_b200


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("order", ["cell", "draw"])
@pytest.mark.parametrize("balanced", [False, True])
def test_emulated_ranks_reproduce_the_cell_distribution(lb, world, order, balanced):
    """Strong-scaling decomposition of bench.py: N split multinomially (shared seed) over equal-mass
    strata, rank g builds only its slab (lint_grid_build) and draws its rows; the combined per-cell
    histogram must match the exact cell probabilities, the slabs must tile the cells the strata
    touch, and every rank must find the same total mass."""
    import torch
    from scipy import stats
    from lintsampler_b200 import distributed as lbd
    rng = np.random.default_rng(0)
    n = 64
    edges = tuple(np.linspace(-4, 4, n + 1) for _ in range(3))
    gmm = lb.GaussianMixture(rng.dirichlet(np.ones(4)), rng.uniform(-2, 2, (4, 3)), sigmas=rng.uniform(0.3, 0.8, 4))
    N = 20_000_000
    counts = None
    hist = torch.zeros(n ** 3, dtype=torch.float64, device="cuda")
    totals, slabs = [], []
    for rank in range(world):
        # the grouped order needs no guide table (bench.py runs it without one)
        sh = lbd.ShardedGrid(edges, gmm, dtype=np.float32, rank=rank, world=world, cell_records=False,
                             guide_bits=0 if order == "cell" and balanced else 14,
                             balance_rows=N if balanced else None)
        if counts is None:                                         # shard probabilities = stratum widths
            widths = np.diff(sh.bounds)
            assert abs(widths.sum() - 1) < 1e-15 and (balanced or np.allclose(widths, 1.0 / world))
            counts = lbd.split_counts(N, widths, 5)
        else:
            assert np.array_equal(np.diff(sh.bounds), widths)      # every rank derives the same strata
        x = torch.empty((int(counts[rank]), 3), dtype=torch.float32, device="cuda")
        first = int(counts[:rank].sum())
        if order == "cell":
            sh.sample_cellmajor_into(x, seed=5, offset=first)
        else:
            sh.sample_philox_into(x, seed=5, offset=first)
        idx = torch.clamp(((x.double() + 4) / 8 * n).long(), 0, n - 1)
        hist += torch.bincount((idx[:, 0] * n + idx[:, 1]) * n + idx[:, 2], minlength=n ** 3).double()
        totals.append(float(sh.grid._total.item()))
        slabs.append(sh.grid._range.tolist()[:2])
        np.testing.assert_allclose(sh.shard_masses, sh.grid.total_mass * widths, rtol=1e-12)
    assert len(set(totals)) == 1                                   # same tile totals on every rank, bit for bit
    assert slabs[0][0] == 0 and slabs[-1][1] == n ** 3
    assert all(a[1] > b[0] for a, b in zip(slabs[:-1], slabs[1:]))  # consecutive slabs overlap (margins): no gap
    mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1)
    V = O.gmm_pdf(mesh, gmm.weights, gmm.means, gmm.covs).astype(np.float32).astype(np.float64)
    p = O.cell_masses(edges, V).ravel()
    p /= p.sum()
    h = hist.cpu().numpy()
    assert h.sum() == N
    big = p * N >= 20
    obs = np.append(h[big], h[~big].sum())
    exp = np.append(p[big], p[~big].sum()) * N
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, len(obs) - 1) > 1e-4, chi2


def test_two_gpu_nccl_run():
    """scripts/check_sharding_gpu.py under torchrun on two GPUs: NCCL all-gather of shard masses,
    both decompositions (equal-mass strata, dim-0 slabs), both row orders, chi^2 of the combined
    histogram.  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "check_sharding_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
