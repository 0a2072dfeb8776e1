#!/usr/bin/env python
"""One rank of the strong-scaling benchmark, emulated on a single GPU: rank r of `world` builds
its slab and draws its multinomial share of N, as one CUDA graph replay per step like bench.py.
For per-kernel launch lists and step times of the sharded step without an 8-GPU box.

    python scripts/emulate_rank.py RANK WORLD [N_TOTAL] [REPS] [BENCH_JSON]

BENCH_JSON: a bench.py line of a real WORLD-GPU run; its calibrated strata are used.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb  # noqa: E402
from lintsampler_b200 import _capi  # noqa: E402
from lintsampler_b200 import distributed as lbd  # noqa: E402
from bench import synthetic_c3, SEED  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
n_total = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1_000_000_000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
edges, w, means, sig = synthetic_c3()
shard = lbd.ShardedGrid(edges, lb.GaussianMixture(w, means, sigmas=sig), dtype=np.float32, rank=rank, world=world,
                        cell_records=False, guide_bits=int(os.environ.get("GUIDE_BITS", "0")),
                        balance_rows=None if os.environ.get("EQUAL_MASS") else n_total)
if len(sys.argv) > 5:
    for line in open(sys.argv[5]):
        if line.startswith("{"):
            strata = json.loads(line)["config"]["strata"]
            assert len(strata) == world + 1
            shard.bounds = np.asarray(strata, dtype=np.float64)
counts = lbd.split_counts(n_total, np.diff(shard.bounds), SEED * 1_000_003)
n, first = int(counts[rank]), int(counts[:rank].sum())
out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
counter = torch.zeros(1, dtype=torch.int64, device="cuda")
_capi.call("lint_set_offset_counter", C.c_void_p(counter.data_ptr()))
lane = torch.cuda.Stream()


def body(ev=None):
    st = torch.cuda.current_stream()
    if ev:
        ev[0].record(st)
    lane.wait_stream(st)
    _capi.call("lint_advance_counter", C.c_void_p(counter.data_ptr()), C.c_uint64(n_total), C.c_void_p(lane.cuda_stream))
    shard.enqueue_build(gather=False)
    st.wait_stream(lane)
    if ev:
        ev[1].record(st)
    shard.sample_cellmajor_into(out, seed=SEED, offset=first)
    if ev:
        ev[2].record(st)


res = []
for i in range(5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    body(ev)
    torch.cuda.synchronize()
    if i >= 2:
        res.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
line = {"rank": rank, "world": world, "rows": n, "slab": shard.grid._range.tolist(),
        "eager_build_ms": float(np.mean([r[0] for r in res])), "eager_sample_ms": float(np.mean([r[1] for r in res]))}
if not os.environ.get("NO_GRAPH"):
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        graph.replay()
    t1.record()
    torch.cuda.synchronize()
    line["graph_ms_per_step"] = t0.elapsed_time(t1) / reps
print(json.dumps(line))
