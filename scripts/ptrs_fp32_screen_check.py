#!/usr/bin/env python
"""Check the fp32 screen of the PTRS acceptance test (csrc/lint_cellmajor.cuh, `poisson`).

The Poisson count sampler accepts a PTRS proposal k when

    log(lhs_arg) <= -lam + k log(lam) - log(k!)                      (exact test, fp64)

and decides most proposals from an fp32 evaluation of t = log(lhs_arg) - rhs with
rhs = k log1p(-x) + d - log(k1)/2 - log(2 pi)/2 - s(k1), k1 = k + 1, d = k1 - lam, x = d / k1:
accept if t < -tol, reject if t > tol, fall through to the fp64 test otherwise, with
tol = 2e-3 + 4e-6 |d|.  A wrong decision is possible only if the fp32 value of t differs from the
exact one by more than tol.  This script replays the screen in NumPy float32 arithmetic over
random proposals of every regime the kernel uses it in (k >= 4, lam < 1e5), adds the documented
worst-case error of the device intrinsics that NumPy does not reproduce bit for bit
(`__logf`: 2^-21.41 absolute on [0.5, 2], 3 ulp elsewhere; `log1pf`: 1 ulp), and reports

    max |t_fp32 - t_exact| / tol       (must stay below 1; it stays below ~0.3)

and the number of wrong decisions (must be 0).  `check(n_per_lambda)` is called by
tests/test_host_logic.py with a reduced sample.
"""
import sys

import numpy as np
from scipy.special import gammaln

F = np.float32


def _logf_bound(x):
    """Documented error bound of CUDA's __logf at x (absolute)."""
    y = np.abs(np.log(x.astype(np.float64)))
    ulp = np.spacing(np.maximum(y, 1e-30).astype(F)).astype(np.float64)
    return np.where((x >= 0.5) & (x <= 2.0), 2.0 ** -21.41, 3.0 * ulp)


def proposals(lam, n, rng):
    """PTRS proposals (k, lhs_arg) that reach the acceptance test, as the kernel forms them."""
    b = 0.931 + 2.53 * np.sqrt(lam)
    a = -0.059 + 0.02483 * b
    vr = 0.9277 - 3.6224 / (b - 2.0)
    u = rng.random(n) - 0.5
    v = rng.random(n)
    us = 0.5 - np.abs(u)
    k = np.floor((2.0 * a / us + b) * u + lam + 0.43)
    squeeze = (us >= 0.07) & (v <= vr)
    drop = (k < 0.0) | ((us < 0.013) & (v > us))
    keep = ~squeeze & ~drop & (k >= 4.0)
    invalpha = 1.1239 + 1.1328 / (b - 3.4)
    lhs = v * invalpha / (a / (us * us) + b)
    return k[keep], lhs[keep]


def screen(lam, k, lhs):
    """fp32 replay of the kernel's screen: (t, tol, error bound of the intrinsics)."""
    d = F(1.0) * (k + 1.0 - lam).astype(F)            # (float)(k + 1.0 - lam): one rounding of an fp64 value
    k1 = k.astype(F) + F(1.0)
    x = d / k1
    series = (F(1.0 / 12.0) - F(1.0 / 360.0) / (k1 * k1)) / k1
    l1p = np.log1p(-x).astype(F)
    logk1 = np.log(k1).astype(F)
    rhs = ((k.astype(F) * l1p + d) - F(0.5) * logk1) - F(0.9189385) - series
    loglhs = np.log(lhs.astype(F)).astype(F)
    t = loglhs - rhs
    tol = F(2e-3) + F(4e-6) * np.abs(d)
    # intrinsics: k * (1 ulp of log1pf) + 0.5 * bound(__logf(k1)) + bound(__logf(lhs))
    intr = (k * np.spacing(np.abs(l1p)).astype(np.float64) + 0.5 * _logf_bound(k1)
            + _logf_bound(lhs.astype(F)))
    return t.astype(np.float64), tol.astype(np.float64), intr


def check(n_per_lambda=200_000, seed=0, verbose=False):
    rng = np.random.default_rng(seed)
    worst, wrong, total = 0.0, 0, 0
    for lam in [10.0, 12.0, 17.3, 30.0, 64.0, 100.0, 130.0, 333.0, 1e3, 3.3e3, 1e4, 3e4, 9e4, 99999.0]:
        k, lhs = proposals(lam, n_per_lambda, rng)
        exact = np.log(lhs) - (-lam + k * np.log(lam) - gammaln(k + 1.0))
        t, tol, intr = screen(lam, k, lhs)
        err = np.abs(t - exact) + intr
        ratio = float(np.max(err / tol)) if len(k) else 0.0
        # a decision is wrong if the screen decides and the exact test says the opposite
        bad = int(np.sum(((t + intr < -tol) & (exact > 0)) | ((t - intr > tol) & (exact <= 0))))
        decided = float(np.mean(np.abs(t) > tol)) if len(k) else 0.0
        worst, wrong, total = max(worst, ratio), wrong + bad, total + len(k)
        if verbose:
            print(f"lam {lam:9.1f}: {len(k):7d} proposals, screen decides {100 * decided:5.1f} %, "
                  f"max err/tol {ratio:.3f}, wrong {bad}")
    return {"proposals": total, "max_err_over_tol": worst, "wrong_decisions": wrong}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    res = check(n, verbose=True)
    print(res)
    sys.exit(0 if res["wrong_decisions"] == 0 and res["max_err_over_tol"] < 1.0 else 1)
