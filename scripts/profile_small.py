"""Host-overhead profile of a tiny problem (README quickstart): where do the
milliseconds of LintSampler(...).sample(N) go when the kernels take microseconds?"""
import cProfile, pstats, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lintsampler_b200 as lb
mu, sig, w = np.array([-3.0, 0.5, 2.5]), np.array([1.0, 0.25, 0.75]), np.array([0.4, 0.25, 0.35])
def pdf(x):
    return np.sum(w * np.exp(-0.5 * ((x - mu) / sig) ** 2) / (sig * np.sqrt(2 * np.pi)))
g = np.linspace(-7, 7, 100)
lb.LintSampler(g, pdf, seed=42).sample(10000)
t0 = time.perf_counter()
for _ in range(20):
    s = lb.LintSampler(g, pdf, seed=42)
t1 = time.perf_counter()
for _ in range(20):
    s.sample(10000)
t2 = time.perf_counter()
print("construct ms", (t1 - t0) / 20 * 1e3, "sample ms", (t2 - t1) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    lb.LintSampler(g, pdf, seed=42).sample(10000)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
