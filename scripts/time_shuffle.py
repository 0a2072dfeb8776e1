"""Cost of the device row permutation (random 12-byte gathers) against table size: 60 MB
(L2 resident) to 12 GB (the headline size)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from lintsampler_b200 import _capi, _device
k = 3
for N in (5_000_000, 20_000_000, 80_000_000, 320_000_000, 1_000_000_000):
    x = torch.rand((N, k), dtype=torch.float32, device="cuda")
    out = torch.empty_like(x)
    def run():
        _capi.call("lint_permute_rows", _device.ptr(x), _device.ptr(out), N, k, 0, C.c_uint64(7), _device.stream_ptr(x.device))
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [run() for _ in range(3)]; b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(json.dumps({"rows": N, "table_MB": N * k * 4 / 1e6, "permute_rows_ms": ms, "rows_per_s": N / ms * 1e3}))
    del x, out
