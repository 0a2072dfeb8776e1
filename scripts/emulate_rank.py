#!/usr/bin/env python
"""One rank of the strong-scaling benchmark, emulated on a single GPU: rank r of `world` builds
its slab and draws its multinomial share of N.  For per-kernel launch lists of the sharded step.

    python scripts/emulate_rank.py RANK WORLD [N_TOTAL] [REPS]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb  # noqa: E402
from lintsampler_b200 import distributed as lbd  # noqa: E402
from bench import synthetic_c3, SEED  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
n_total = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1_000_000_000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
edges, w, means, sig = synthetic_c3()
shard = lbd.ShardedGrid(edges, lb.GaussianMixture(w, means, sigmas=sig), dtype=np.float32, rank=rank, world=world,
                        cell_records=False, guide_bits=16,
                        balance_rows=None if os.environ.get("EQUAL_MASS") else n_total)
counts = lbd.split_counts(n_total, np.diff(shard.bounds), SEED * 1_000_003)
n, first = int(counts[rank]), int(counts[:rank].sum())
out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
res = []
for i in range(reps + 2):
    ev[0].record()
    shard.enqueue_build(gather=False)
    ev[1].record()
    shard.sample_cellmajor_into(out, seed=SEED, offset=i * n_total + first)
    ev[2].record()
    torch.cuda.synchronize()
    if i >= 2:
        res.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
print(json.dumps({"rank": rank, "world": world, "rows": n, "slab": shard.grid._range.tolist(),
                  "build_ms": float(np.mean([r[0] for r in res])), "sample_ms": float(np.mean([r[1] for r in res]))}))
