#!/usr/bin/env python
"""Time one cell-major draw of the headline workload (256^3 grid, fp32) with CUDA events,
optionally under the tuning overrides LINT_FLAT_BPS / LINT_REST_BPS / LINT_SIDE.

    python scripts/time_emit.py [N] [reps] [KEY=VALUE ...]      # one JSON line per configuration
"""
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb  # noqa: E402
from bench import synthetic_c3  # noqa: E402


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    sweeps = {}
    for a in sys.argv[3:]:
        k, v = a.split("=")
        sweeps[k] = v.split(",")
    edges, w, means, sig = synthetic_c3()
    dtype = np.float64 if os.environ.get("LINT_F64") else np.float32
    grid = lb.DensityGrid(edges, lb.GaussianMixture(w, means, sigmas=sig), vectorizedpdf=True,
                          dtype=dtype, cell_records=False, guide_bits=0)
    out = torch.empty((N, 3), dtype=torch.float32 if dtype == np.float32 else torch.float64, device="cuda")
    keys = sorted(sweeps)
    for combo in itertools.product(*(sweeps[k] for k in keys)) if keys else [()]:
        for k, v in zip(keys, combo):
            os.environ[k] = v
        for i in range(2):
            grid._sample_cellmajor(N, 1, i * N, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            grid._sample_cellmajor(N, 1, (2 + i) * N, out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        print(json.dumps({"N": N, **dict(zip(keys, combo)), "ms_min": min(ms), "ms_med": float(np.median(ms)),
                          "GBps": N * 3 * out.element_size() / (np.median(ms) * 1e-3) / 1e9}), flush=True)
    # the output-bandwidth ceiling: a plain fill of the same buffer
    for _ in range(2):
        out.fill_(1.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out.fill_(1.0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"fill_ms": ms, "fill_GBps": out.numel() * out.element_size() / (ms * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
