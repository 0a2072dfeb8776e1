#!/usr/bin/env python
"""Multi-GPU check (torchrun, NCCL): both sharding modes must reproduce the
single-domain distribution.  Every rank samples its share; rank 0 gathers per-cell
histograms and compares them with the exact cell probabilities (chi^2), and reports
the work balance of the two decompositions."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
from scipy import stats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb
from lintsampler_b200 import distributed as lbd
from oracle import lint_oracle as O      # checker only

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(0)
n = 64
edges = tuple(np.linspace(-4, 4, n + 1) for _ in range(3))
gmm = lb.GaussianMixture(rng.dirichlet(np.ones(4)), rng.uniform(-2, 2, (4, 3)), sigmas=rng.uniform(0.3, 0.8, 4))
N = 40_000_000
res = {}
for mode in ("mass", "slab"):
    for order in ("draw", "cell"):
        if mode == "mass":
            sh = lbd.ShardedGrid(edges, gmm, dtype=np.float32, rank=rank, world=world, group=dist.group.WORLD)
            x, counts = sh.sample(N, seed=5) if order == "draw" else (None, None)
            if x is None:
                counts = lbd.split_counts(N, sh.exchange_masses(), 5)
                x = torch.empty((int(counts[rank]), 3), dtype=torch.float32, device="cuda")
                lo, hi = float(sh.bounds[rank]), float(sh.bounds[rank + 1])
                sh.grid._sample_cellmajor(x.shape[0], 5, int(counts[:rank].sum()), out=x, stratum=(lo, hi))
        else:
            sh = lbd.SlabShardedGrid(edges, gmm, dtype=np.float32, rank=rank, world=world, group=dist.group.WORLD)
            x, counts = sh.sample(N, seed=5, order=order)
        idx = torch.clamp(((x.double() + 4) / 8 * n).long(), 0, n - 1)
        flat = (idx[:, 0] * n + idx[:, 1]) * n + idx[:, 2]
        hist = torch.bincount(flat, minlength=n ** 3).double()
        dist.all_reduce(hist)
        if rank == 0:
            mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1)
            V = O.gmm_pdf(mesh, gmm.weights, gmm.means, gmm.covs).astype(np.float32).astype(np.float64)
            p = O.cell_masses(edges, V).ravel()
            p /= p.sum()
            h = hist.cpu().numpy()
            big = p * N >= 20
            obs = np.append(h[big], h[~big].sum())
            exp = np.append(p[big], p[~big].sum()) * N
            chi2 = ((obs - exp) ** 2 / exp).sum()
            res[f"{mode}/{order}"] = dict(rows=int(h.sum()), chi2_p=float(stats.chi2.sf(chi2, len(obs) - 1)),
                                          counts=[int(c) for c in counts],
                                          balance=float(max(counts) / (N / world)))
if rank == 0:
    print(json.dumps(res))
    assert all(v["rows"] == N and v["chi2_p"] > 1e-4 for v in res.values())
dist.destroy_process_group()
