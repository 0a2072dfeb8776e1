#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (C1, C2, C4, C5 of SURVEY.md §8d)
on one B200.  Prints one JSON object per config; `python scripts/bench_configs.py > out`.
C3 (the headline) is bench.py."""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lintsampler_b200 as lb  # noqa: E402

PEAK = 6455.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, N, k, itemsize, ms, **extra):
    rate = N / (ms * 1e-3)
    gbs = N * k * itemsize / (ms * 1e-3) / 1e9
    print(json.dumps(dict(config=name, N=N, ms=ms, samples_per_s=rate, out_GBps=gbs,
                          frac_of_hbm_peak=gbs / PEAK, **extra)), flush=True)


def gmm(k, K, smin, smax, full=False):
    rng = np.random.default_rng(0)
    means = rng.uniform(-2, 2, size=(K, k))
    if full:
        A = rng.uniform(-0.5, 0.5, size=(K, k, k))
        covs = A @ np.swapaxes(A, 1, 2) + 0.05 * np.eye(k)
        w = rng.dirichlet(np.ones(K))
        return lb.GaussianMixture(w, means, covs=covs)
    sig = rng.uniform(smin, smax, size=K)
    w = rng.dirichlet(np.ones(K))
    return lb.GaussianMixture(w, means, sigmas=sig)


def c1():
    mu, sig, w = np.array([-3.0, 0.5, 2.5]), np.array([1.0, 0.25, 0.75]), np.array([0.4, 0.25, 0.35])

    def pdf(x):
        return np.sum(w * np.exp(-0.5 * ((x - mu) / sig) ** 2) / (sig * np.sqrt(2 * np.pi)))
    lb.LintSampler(np.linspace(-7, 7, 100), pdf, seed=42).sample(10000)     # CUDA context, library load
    t0 = time.perf_counter()
    for _ in range(20):
        x = lb.LintSampler(np.linspace(-7, 7, 100), pdf, seed=42).sample(10000)
    dt = (time.perf_counter() - t0) / 20
    print(json.dumps(dict(config="C1 README 1-D quickstart (99 cells, scalar pdf, N=1e4), construct+sample wall",
                          N=10000, ms=dt * 1e3, samples_per_s=10000 / dt, mean=float(x.mean()))), flush=True)


def grid_case(name, edges, pdf, N, qmc=False, dtype=np.float32):
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    isz = 4 if dtype == np.float32 else 8
    for order in (("draw",) if qmc else ("cell", "draw")):
        kw = dict(cell_records=False, guide_bits=0) if order == "cell" else {}
        t0 = time.perf_counter()
        grid = lb.DensityGrid(edges, pdf, vectorizedpdf=True, dtype=dtype, **kw)
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        k = grid.dim
        out = torch.empty((N, k), dtype=tdt, device="cuda")
        ms_build = timed(grid._enqueue_build)
        if qmc:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                s = lb.LintSampler(grid, seed=1, qmc=True, return_tensor=True, shuffle=False)
                st = s._sobol_struct()
                st_sorted = s._sobol_struct(sorted_n=N)
            ms = timed(lambda: grid._sample_sobol(N, st), reps=3)
            report(f"{name} [sobol, direct enumeration, no shuffle]", N, k, 4, ms, cells=int(grid.ncells))
            ms = timed(lambda: s._permute_rows(grid._sample_sobol(N, st)), reps=3)
            report(f"{name} [sobol, direct enumeration + device shuffle]", N, k, 4, ms, cells=int(grid.ncells))
            ms = timed(lambda: s._permute_rows(grid._sample_sobol(N, st_sorted)), reps=3)
        elif order == "cell":
            ms = timed(lambda: grid._sample_cellmajor(N, 1, 0, out=out))
        else:
            ms = timed(lambda: grid._sample_philox(N, 1, 0, out=out))
        report(f"{name} [{'sobol, sorted enumeration + device shuffle' if qmc else 'order=' + order}]", N, k, isz, ms,
               ms_build_kernels=ms_build, first_construct_s=t_build, cells=int(grid.ncells))
        del grid, out
        torch.cuda.empty_cache()


def c5():
    pdf = gmm(3, 8, 0.02, 0.2)
    t0 = time.perf_counter()
    tree = lb.DensityTree([-4.0] * 3, [4.0] * 3, pdf, vectorizedpdf=True, min_openings=4, dtype=np.float32)
    t1 = time.perf_counter()
    tree.refine(1e-3, leaf_tol=1e-3)
    t2 = time.perf_counter()
    levels = np.bincount([leaf.level for leaf in tree.leaves])
    N = 100_000_000
    out = torch.empty((N, 3), dtype=torch.float32, device="cuda")
    tree._ensure_device()
    for order in ("cell", "draw"):
        fn = (lambda: tree._sample_cellmajor(N, 1, 0, out=out)) if order == "cell" else \
             (lambda: tree._sample_philox(N, 1, 0, out=out))
        ms = timed(fn)
        report(f"C5 3-D DensityTree, K=8 sharp GMM [order={order}]", N, 3, 4, ms, leaves=len(tree.leaves),
               levels={int(i): int(c) for i, c in enumerate(levels) if c}, total_mass=float(tree.total_mass),
               host_build_s=t1 - t0, host_refine_s=t2 - t1)


if __name__ == "__main__":
    only = set(sys.argv[1:])            # e.g. `bench_configs.py c4p c5`; default: all

    def want(tag):
        return not only or tag in only

    if want("c1"):
        c1()
    if want("c2"):
        grid_case("C2 2-D 1024^2 full-cov GMM, N=1e8 fp32", tuple(np.linspace(-4, 4, 1025) for _ in range(2)),
                  gmm(2, 4, 0, 0, full=True), 100_000_000)
    if want("c3d"):
        # the headline grid (bench.py's generator) with fp64 densities and samples
        rng = np.random.default_rng(0)
        means = rng.uniform(-2, 2, size=(4, 3))
        sig = rng.uniform(0.3, 0.8, size=4)
        w = rng.dirichlet(np.ones(4))
        grid_case("C3 3-D 256^3 isotropic GMM, N=1e9 fp64", tuple(np.linspace(-4, 4, 257) for _ in range(3)),
                  lb.GaussianMixture(w, means, sigmas=sig), 1_000_000_000, dtype=np.float64)
    if want("c4"):
        grid_case("C4 5-D 32^5 GMM, N=2^27 fp32", tuple(np.linspace(-4, 4, 33) for _ in range(5)),
                  gmm(5, 4, 0.5, 0.9), 1 << 27, qmc=True)
    if want("c4p"):
        grid_case("C4' 5-D 32^5 GMM (pseudo-random), N=1e8 fp32", tuple(np.linspace(-4, 4, 33) for _ in range(5)),
                  gmm(5, 4, 0.5, 0.9), 100_000_000)
    if want("c5"):
        c5()
