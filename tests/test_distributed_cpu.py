"""World-size-2 tests of the sharding logic on CPU (gloo): the collective that
exchanges per-shard masses, the shared-seed multinomial split of N, and the
parity-mode routing of uniforms to ranks.  The kernels themselves are covered by
the gpu-marked tests; here the per-rank sampling is played by the oracle so the
sharded result can be compared with the single-domain result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, golden_edges
from lintsampler_b200 import distributed as lbd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lint_oracle as O
        g = load_golden("g3d_iso")
        edges, V = golden_edges(g), g["vertex_densities"]
        cdf = O.cdf_from_masses(O.cell_masses(edges, V))
        total = float(np.sum(O.cell_masses(edges, V)))
        bounds = lbd.equal_mass_strata(world)
        mine = torch.tensor(total * (bounds[rank + 1] - bounds[rank]), dtype=torch.float64)
        masses = lbd.all_gather_masses(mine, world).numpy()
        counts = lbd.split_counts(100_000, masses, seed=7)
        # parity routing: every rank sees the same reference uniforms, keeps its rows,
        # samples them in its stratum; the union must equal the single-domain result
        u = g["u"].copy()
        owner, u_local = lbd.route_rows(u[:, -1], bounds)
        keep = owner == rank
        u_mine = u[keep]
        u_mine[:, -1] = bounds[rank] + u_local[keep] * (bounds[rank + 1] - bounds[rank])
        x = O.grid_sample(edges, V, u_mine, cdf=cdf)
        q.put((rank, masses, counts, np.flatnonzero(keep), x))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, m0, c0, rows0, x0), (_, m1, c1, rows1, x1) = res
    assert np.array_equal(m0, m1) and np.allclose(m0[0], m0[1])       # same gathered masses
    assert np.array_equal(c0, c1) and c0.sum() == 100_000              # same split, exact total
    assert abs(c0[0] - 50_000) < 1000                                  # ~ Multinomial(N, 1/2)
    g = load_golden("g3d_iso")
    assert len(rows0) + len(rows1) == len(g["u"]) and not set(rows0) & set(rows1)
    x = np.empty_like(g["x_pre"])
    x[rows0], x[rows1] = x0, x1
    # rescaling u into the stratum and back costs one rounding of the cell uniform:
    # a row can change cell only if u sits within 1 ulp of a CDF boundary
    same = np.all(x == g["x_pre"], axis=1)
    assert same.mean() > 0.999
    np.testing.assert_allclose(x[same], g["x_pre"][same], rtol=0, atol=0)


def test_split_and_routing_helpers():
    b = lbd.equal_mass_strata(4)
    assert np.array_equal(b, [0, 0.25, 0.5, 0.75, 1.0])
    r, ul = lbd.route_rows(np.array([0.0, 0.24, 0.25, 0.5, 0.99]), b)
    assert list(r) == [0, 0, 1, 2, 3]
    assert np.allclose(ul, [0.0, 0.96, 0.0, 0.0, 0.96])
    assert lbd.slab_bounds(256, 8) == [0, 32, 64, 96, 128, 160, 192, 224, 256]
    assert lbd.slab_bounds(10, 3) == [0, 3, 6, 10]
    c = lbd.split_counts(10**9, [1.0, 3.0], seed=1)
    assert c.sum() == 10**9 and abs(c[1] / 10**9 - 0.75) < 1e-4
