#!/usr/bin/env python
"""C5 tree: build + refine time of the device builder and of the host builder."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb  # noqa: E402
from bench import _gmm  # noqa: E402

pdf = _gmm(lb, 3, 8, 0.02, 0.2)
res = {}
for build in ("device", "device", "host"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tree = lb.DensityTree([-4.0] * 3, [4.0] * 3, pdf, vectorizedpdf=True, min_openings=4, dtype=np.float32, build=build)
    t1 = time.perf_counter()
    tree.refine(1e-3, leaf_tol=1e-3)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = len(tree._leaf_ids) if tree._dt is not None else len(tree.leaves)
    res.setdefault(build, []).append({"construct_s": t1 - t0, "refine_s": t2 - t1, "leaves": n,
                                      "cells": int(tree._dt.n) if tree._dt is not None else None})
print(json.dumps(res))
