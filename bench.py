#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): samples/s for a 256^3 3-D DensityGrid,
N = 1e9 samples per step in total (strong scaling over the GPUs: each rank owns an equal-mass
stratum of the cell CDF), fp32 samples / fp64 CDF, synthetic K=4 Gaussian mixture.

    python bench.py --gpus N --steps K --warmup W            # this build
    python bench.py --impl reference ...                      # CPU reference arm

One "step" = the whole hot path on one batch: cell masses -> cell CDF -> guide
table -> N fused Philox/search/gather/inverse-CDF samples written to HBM, with
the vertex densities already resident in HBM.  `value` counts the samples of
all ranks over the max-over-ranks device time.  `e2e` is the same step driven
through the public API from host buffers: pinned vertex densities copied in,
samples copied out to pinned host memory, both inside the timed region.

Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_N = 256          # cells per dimension (C3 in SURVEY.md §8)
DIM = 3
N_SAMPLES = 1_000_000_000
SEED = 1


def synthetic_c3():
    """SURVEY.md §8(d) C3 generator: all draws from default_rng(0), in this order."""
    rng = np.random.default_rng(0)
    means = rng.uniform(-2, 2, size=(4, DIM))
    sig = rng.uniform(0.3, 0.8, size=4)
    w = rng.dirichlet(np.ones(4))
    edges = tuple(np.linspace(-4, 4, GRID_N + 1) for _ in range(DIM))
    return edges, w, means, sig


def algorithmic_bytes(n_samples, dens_bytes=4, out_bytes=4):
    """B_mass + B_scan + B_sample."""
    V = (GRID_N + 1) ** DIM
    Cn = GRID_N ** DIM
    b_mass = V * dens_bytes + Cn * 8
    b_scan = 2 * Cn * 8
    b_sample = n_samples * DIM * out_bytes
    return b_mass, b_scan, b_sample


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.index = index
        self.enabled = enabled
        self.rows = []
        self._nv = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self._nv is not None:
            return self._run_nvml()
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                     "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def _setup_nvml(self):
        """The same fields through NVML (what nvidia-smi reads), so that they can be read every
        millisecond: a timed region of a few milliseconds still gets a median over several samples."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            reasons(h)
            return nv, h, reasons, mx
        except Exception:
            return None

    def _run_nvml(self):
        nv, h, reasons, mx = self._nv
        bits = [0x8, 0x40, 0x20, 0x4]       # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self._stop.is_set():
            try:
                r = int(reasons(h))
                self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)]
                                 + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._stop.wait(0.001)

    def __enter__(self):
        if self.enabled:
            self._nv = self._setup_nvml()       # before the timed region starts
            self.source = "nvml" if self._nv is not None else "nvidia-smi"
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.enabled:
            self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names)
                   if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "source": getattr(self, "source", "nvidia-smi")}


# --------------------------------------------------------------------------- CPU legs (oracle)
_CPU_TABLES = None      # (edges, V, cdf): set before the pool forks, inherited copy-on-write


def _cpu_worker(args):
    seed, n = args
    from oracle import lint_oracle as O
    edges, V, cdf = _CPU_TABLES
    u = np.random.default_rng(seed).random((n, DIM + 1))
    x = O.grid_sample(edges, V, u, cdf=cdf)
    return float(x[0, 0])


def cpu_baseline_port(n_cpu=3_000_000):
    """The oracle (NumPy port of the reference's algorithm) on ONE host core, on a
    bounded sample of the same workload: CDF rebuild + uniforms + _grid_sample."""
    from oracle import lint_oracle as O
    edges, w, means, sig = synthetic_c3()
    mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1)
    V = O.gmm_pdf(mesh, w, means, sig[:, None, None] ** 2 * np.eye(DIM)[None])
    del mesh
    masses = O.cell_masses(edges, V)
    t0 = time.perf_counter()
    cdf = O.cdf_from_masses(masses)                 # the reference redoes this on every call
    u = np.random.default_rng(SEED).random((n_cpu, DIM + 1))
    x = O.grid_sample(edges, V, u, cdf=cdf)
    t1 = time.perf_counter()
    np.random.default_rng(SEED).shuffle(x, axis=0)  # the reference's final rng.shuffle
    t2 = time.perf_counter()
    return {"value": n_cpu / (t1 - t0), "unit": "samples/s", "cores": 1, "kind": "port",
            "sample": f"{n_cpu} samples of the same 256^3 grid: CDF rebuild + PCG64 uniforms + "
                      f"_grid_sample (oracle/lint_oracle.py) in {t1 - t0:.2f} s on 1 core; with the "
                      f"reference's final rng.shuffle {n_cpu / (t2 - t0):.3g} samples/s"}


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU algorithm (the oracle port — the
    reference is pure NumPy and cannot travel to the GPU box) on all host cores:
    one process per core, each sampling its share of a bounded batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import lint_oracle as O
    edges, w, means, sig = synthetic_c3()
    mesh = np.stack(np.meshgrid(*edges, indexing="ij"), axis=-1)
    V = O.gmm_pdf(mesh, w, means, sig[:, None, None] ** 2 * np.eye(DIM)[None])
    del mesh
    cdf = O.cdf_from_masses(O.cell_masses(edges, V))
    global _CPU_TABLES
    _CPU_TABLES = (edges, V, cdf)
    cores = min(len(os.sched_getaffinity(0)), 64)
    per = 1_000_000
    n_step = cores * per
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(s):
            pool.map(_cpu_worker, [(1000 * s + c, per) for c in range(cores)])
        for s in range(args.warmup):
            step(s)
        t0 = time.perf_counter()
        for s in range(args.steps):
            step(100 + s)
        dt = (time.perf_counter() - t0) / args.steps
    val = n_step / dt
    line = {
        "impl": "reference", "metric": "samples/s", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C3: 3-D DensityGrid 256^3, K=4 isotropic GMM on [-4,4]^3",
                   "samples_per_step": n_step, "note": "bounded sample of the N=1e9 workload"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{n_step} samples/step ({per} per process), uniforms + "
                                   "searchsorted + corner gather + in-cell inverse CDF; CDF prebuilt"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def emit_traffic(order, rows_flat, rows_general, n_rows):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, scaled from the ncu
    --set full capture recorded in profiles/traffic.json (bytes per row, with the commit and the
    profile file it came from)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, None
    if order == "cell":
        k = t.get("emit_flat_kernel<3,float>")
        return (k["dram_bytes_per_row"] * rows_flat, k["source"]) if k else (None, None)
    k = t.get("sample_kernel<3,float,GridSource,UniformsPhilox>")
    return (k["dram_bytes_per_row"] * n_rows, k["source"]) if k else (None, None)


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import lintsampler_b200 as lb
    from lintsampler_b200 import _capi
    from lintsampler_b200 import distributed as lbd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    edges, w, means, sig = synthetic_c3()
    gmm = lb.GaussianMixture(w, means, sigmas=sig)
    # strong scaling: N samples in total, split multinomially over the ranks'
    # equal-mass strata with a seed every rank shares (no communication); --weak: N per rank
    n_total = args.samples * (world if args.weak else 1)
    grid_kw = {}
    if args.order == "cell":
        # the grouped emitter loads each cell once, so it needs neither per-cell records
        # nor a fine guide table (only the ~2.5e-4 N top-up rows search the CDF)
        grid_kw.update(cell_records=False, guide_bits=0)
    if args.guide_bits is not None:
        grid_kw["guide_bits"] = args.guide_bits
    if args.no_records:
        grid_kw["cell_records"] = False
    shard = lbd.ShardedGrid(edges, gmm, dtype=np.float32, device=dev, rank=rank, world=world,
                            group=dist.group.WORLD if world > 1 else None,
                            balance_rows=None if args.equal_mass else n_total, **grid_kw)
    if world > 1 and not args.equal_mass and not args.no_calibrate and args.order == "cell":
        shard.calibrate(n_total, seed=SEED)          # strata re-cut from measured step times (setup, collective)
    widths = np.diff(shard.bounds)                   # shard probabilities (identical on every rank)
    cap = n_total if world == 1 else int(n_total * widths.max() + 8 * np.sqrt(n_total) + 64)
    out = torch.empty((cap, DIM), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def share(i):
        """(rows, first row) of this rank in step i: Multinomial(N, 1/G ... 1/G), shared seed."""
        if world == 1:
            return n_total, 0
        counts = lbd.split_counts(n_total, widths, SEED * 1_000_003 + i)
        return int(counts[rank]), int(counts[:rank].sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The split of N over the ranks is drawn once (shared seed) and kept for the run, so that a step
    # is one CUDA graph replay; the graph starts by advancing a device-resident Philox offset, so
    # every replay draws fresh samples (lint_set_offset_counter / lint_advance_counter).
    n_mine, first_mine = share(0)
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    _capi.call("lint_set_offset_counter", C.c_void_p(counter.data_ptr()))
    lane = torch.cuda.Stream(dev)

    def body(ev=None, order=None):
        if ev:
            ev[0].record(stream_now())
        # the Philox offset advances beside the build (which draws nothing)
        main = stream_now()
        lane.wait_stream(main)
        _capi.call("lint_advance_counter", C.c_void_p(counter.data_ptr()), C.c_uint64(n_total),
                   C.c_void_p(lane.cuda_stream))
        shard.enqueue_build(gather=False)           # vertex row sums + group offsets of the whole grid, CDF (+ guide) of the own slab
        main.wait_stream(lane)
        if ev:
            ev[1].record(stream_now())
        if (order or args.order) == "cell":
            shard.sample_cellmajor_into(out[:n_mine], seed=SEED, offset=first_mine)
        else:
            shard.sample_philox_into(out[:n_mine], seed=SEED, offset=first_mine)
        if ev:
            ev[2].record(stream_now())

    def stream_now():
        return torch.cuda.current_stream(dev)

    for i in range(args.warmup):
        body()
    barrier()
    graph = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        for _ in range(2):
            graph.replay()
        barrier()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, enabled=rank == 0) as clocks:
        barrier()
        h0 = time.perf_counter()
        t_beg.record(stream)
        for i in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                body()
        t_end.record(stream)
        host_ms = (time.perf_counter() - h0) * 1e3 / args.steps      # host time to enqueue a step
        barrier()
    ms_total = t_beg.elapsed_time(t_end)
    clock_line = clocks.summary()
    if rank == 0 and clock_line["samples"] < 5:
        # a timed region of a few milliseconds is shorter than a handful of NVML reads: sample the same
        # steps again, untimed, right after it (every rank runs them, so the load is the same)
        clock_line["note"] = ("timed region %.1f ms: clocks sampled over extra untimed steps right after it"
                              % ms_total)
    extra = clock_line["samples"] < 5 if rank == 0 else False
    if world > 1:
        flag = torch.tensor([int(extra)], device=dev)
        dist.broadcast(flag, 0)
        extra = bool(flag.item())
    if extra:
        with ClockSampler(local, enabled=rank == 0) as clocks2:
            barrier()
            for i in range(max(args.steps, int(60.0 / max(ms_total / args.steps, 1e-3)))):
                graph.replay() if graph is not None else body()
            barrier()
        if rank == 0:
            note = clock_line["note"]
            clock_line = clocks2.summary()
            clock_line["note"] = note
    # stage breakdown: a few eager steps with events between the stages
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(3)]
    for e in evs:
        body(e)
    barrier()
    ms_build = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    ms_sample = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    t = torch.tensor([ms_total, ms_build, ms_sample, host_ms], dtype=torch.float64, device=dev)
    rank_ms = [ms_total / args.steps]
    if world > 1:
        every = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(every, t[:1].clone() / args.steps)
        rank_ms = [float(v) for v in every.tolist()]             # each rank's own step time (the value uses the max)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_build, ms_sample, host_ms = t.tolist()
    ms_step = ms_total / args.steps
    value = n_total / (ms_step * 1e-3)

    # ---- the dominant kernels on their own: CUDA events recorded by the library around the flat
    # emitter and the general emitter (3 extra steps, after the timed region)
    kern = None
    if args.order == "cell":
        _capi.call("lint_cellmajor_timing", 1)
        acc = np.zeros(5)
        reps = 3
        for i in range(reps):
            body()
            buf = (C.c_double * 5)()
            _capi.call("lint_cellmajor_last_times", buf)
            acc += np.array(list(buf))
        _capi.call("lint_cellmajor_timing", 0)
        kern = dict(zip(["ms_flat", "ms_general", "rows_flat", "rows_general", "rows_grouped"], acc / reps))
    barrier()

    # the other row order, for the record (2 steps, same workload)
    other = "draw" if args.order == "cell" else "cell"
    other_line = None
    if not args.no_other and world == 1:
        # its own tables: the u-driven kernel wants per-cell records and a fine guide table
        shard.grid.__dict__.pop("_cm_scratch", None)
        okw = dict(cell_records=False, guide_bits=0) if other == "cell" else dict(guide_bits=22)
        shard_main, shard = shard, lbd.ShardedGrid(edges, gmm, dtype=np.float32, device=dev, rank=rank,
                                                   world=world, group=None, **okw)
        body(order=other)
        barrier()
        o_beg, o_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o_beg.record(stream)
        for i in range(2):
            body(order=other)
        o_end.record(stream)
        barrier()
        ms_other = o_beg.elapsed_time(o_end) / 2
        other_line = {"order": other, "value": n_total / (ms_other * 1e-3), "ms_per_step": ms_other, "steps": 2}
        shard = shard_main

    _capi.call("lint_set_offset_counter", None)
    # ---- the same workload in the reference's own precision (fp64 densities and samples), 3 steps
    f64_line = None
    if world == 1 and not args.no_other:
        out = None
        torch.cuda.empty_cache()
        g64 = lb.DensityGrid(edges, gmm, vectorizedpdf=True, dtype=np.float64, device=dev, cell_records=False,
                             guide_bits=0)
        out64 = torch.empty((n_total, DIM), dtype=torch.float64, device=dev)
        for i in range(2):
            g64._enqueue_build()
            g64._sample_cellmajor(n_total, SEED, i * n_total, out=out64)
        barrier()
        o_beg, o_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o_beg.record(stream)
        for i in range(3):
            g64._enqueue_build()
            g64._sample_cellmajor(n_total, SEED, (2 + i) * n_total, out=out64)
        o_end.record(stream)
        barrier()
        ms64 = o_beg.elapsed_time(o_end) / 3
        bm, bs, bo = algorithmic_bytes(n_total, 8, 8)
        f64_line = {"dtype": "f64", "value": n_total / (ms64 * 1e-3), "ms_per_step": ms64, "steps": 3,
                    "step_frac_of_hbm_roofline": (bm + bs + bo) / (ms64 * 1e-3) / 1e9 /
                    float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
                    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None}
        del g64, out64
        torch.cuda.empty_cache()
        out = None
    # ---- e2e: public API, host buffers in and out
    e2e = None
    if not args.no_e2e:
        out = None
        torch.cuda.empty_cache()
        e2e = run_e2e(args, lb, edges, gmm, dev, world, n_total)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        b_mass, b_scan, b_sample = algorithmic_bytes(n_total)
        step_bytes_per_gpu = (b_mass + b_scan + b_sample) / world
        if args.order == "cell":
            rows_dom, ms_dom = kern["rows_flat"], kern["ms_flat"]
            name = "emit_flat_kernel<3,float> (rows of the constant parts of the split cells)"
        else:
            rows_dom, ms_dom = n_total / world, ms_sample
            name = "sample_kernel<3,float,GridSource<3,float>,UniformsPhilox<3>>"
        achieved = rows_dom * DIM * 4 / (max(ms_dom, 1e-9) * 1e-3) / 1e9
        traffic, traffic_src = emit_traffic(args.order, rows_dom, kern["rows_general"] if kern else 0, n_total / world)
        line = {
            "metric": "samples/s", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "C3: 3-D DensityGrid 256^3 (257^3 fp32 vertex densities, fp64 CDF), "
                            "K=4 isotropic GMM on [-4,4]^3, N=%d samples per step in total" % n_total,
                "samples_total": n_total, "samples_per_gpu": n_total / world,
                "l2": "outputs (12 GB/step in total) and tables (>200 MB) exceed L2",
                "sharding": shard.describe(),
                "strata": [float(b) for b in shard.bounds],
                "launch": ("one CUDA graph replay per step: %d kernels captured once; the graph advances a "
                           "device-resident Philox offset, so every step draws fresh samples; the multinomial "
                           "split of N over the ranks is drawn once per run from the shared seed"
                           % (shard.launches_per_step(args.order) + 1)) if graph is not None else "eager launches",
                "host_ms_per_step": host_ms,
                "rank_ms_per_step": rank_ms,
                "order": ("rows grouped by cell: exact multinomial cell counts (Poisson counts + u-driven "
                          "top-up), i.i.d. as a multiset, NOT exchangeable in row order"
                          if args.order == "cell" else
                          "rows in draw order (i.i.d. Philox rows, exchangeable like the reference's shuffle)"),
            },
            "roofline": {
                "bound": "hbm", "kernel": name,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured; a copy: a pure write stream "
                               "reaches ~7470 GB/s on this part, scripts/time_emit.py)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_launch": rows_dom * DIM * 4, "ms_per_launch": ms_dom,
                "timing": "CUDA events recorded by the library on the launching stream around this kernel",
                "traffic": traffic, "traffic_source": traffic_src,
                "stages_ms": {"build": ms_build, "sampling": ms_sample,
                              **({"emit_flat": kern["ms_flat"], "emit_general": kern["ms_general"],
                                  "planning_and_topup": ms_sample - kern["ms_flat"] - kern["ms_general"]}
                                 if kern else {})},
                "rows": ({k: kern[k] for k in ("rows_flat", "rows_general", "rows_grouped")} if kern else None),
                "step_frac_of_hbm_roofline": step_bytes_per_gpu / (ms_step * 1e-3) / 1e9 / peak,
            },
            "clocks": clock_line,
            "gpu_launches": (shard.launches_per_step(args.order) + 1) * args.steps,   # + the offset-counter kernel
        }
        if other_line:
            line["other_order"] = other_line
        if f64_line:
            line["c3_f64"] = f64_line
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_port()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _numa_summary():
    """Where the pinned staging memory can live: NUMA nodes of the host and of this process."""
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        return {"host_numa_nodes": len(nodes), "cpus_allowed": len(os.sched_getaffinity(0))}
    except Exception:
        return None


def run_e2e(args, lb, edges, gmm, dev, world, n_total):
    import torch
    import torch.distributed as dist
    n = n_total // world                                    # rows per rank
    steps = max(1, min(args.steps, 3))
    # host inputs: pinned fp32 vertex densities (evaluated once, outside the timed region)
    kw = dict(cell_records=False, guide_bits=0) if args.order == "cell" else dict(guide_bits=22)
    g0 = lb.DensityGrid(edges, gmm, vectorizedpdf=True, dtype=np.float32, device=dev, **kw)
    v_host = torch.empty(g0._V.numel(), dtype=torch.float32, pin_memory=True)
    v_host.copy_(g0._V)
    out_host = torch.empty((n, DIM), dtype=torch.float32, pin_memory=True)
    sampler = lb.LintSampler(g0, seed=SEED, dtype=np.float32, order=args.order)
    torch.cuda.synchronize(dev)

    def one():
        g0.update_densities(v_host, check=True)              # H2D + masses + CDF + guide + validity read-back
        sampler.sample_to_host(n, out=out_host)              # sample + D2H, pipelined; returns when the rows are in host memory
    one()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    assert np.isfinite(out_host[-1].numpy()).all()
    d2h = int(n * DIM * 4)
    return {"value": n * world / dt, "unit": "samples/s", "h2d_bytes_per_step": int(v_host.numel() * 4),
            "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": steps,
            "d2h_gb_per_s_per_gpu": d2h / dt / 1e9, "numa": _numa_summary(),
            "api": "DensityGrid.update_densities(pinned, check=True) + LintSampler.sample_to_host(N, out=pinned)"}


# --------------------------------------------------------------------------- the other named configs
def _gmm(lb, k, K, smin, smax, full=False):
    """SURVEY.md 8(d) generators: all draws from default_rng(0) in this order."""
    rng = np.random.default_rng(0)
    means = rng.uniform(-2, 2, size=(K, k))
    if full:
        A = rng.uniform(-0.5, 0.5, size=(K, k, k))
        covs = A @ np.swapaxes(A, 1, 2) + 0.05 * np.eye(k)
        return lb.GaussianMixture(rng.dirichlet(np.ones(K)), means, covs=covs)
    sig = rng.uniform(smin, smax, size=K)
    return lb.GaussianMixture(rng.dirichlet(np.ones(K)), means, sigmas=sig)


def run_config(args):
    """`--config C2|C3f64|C4|C5`: the other configurations BASELINE.json names, one
    GPU, the same JSON shape.  A step = build of the tables (grids) + the N-row draw through the
    same entry points the headline uses; `roofline` is the step against its algorithmic bytes."""
    import warnings
    import torch
    import lintsampler_b200 as lb
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    name = args.config.upper()
    dtype, qmc, tree = np.float32, False, None
    if name == "C2":
        edges = tuple(np.linspace(-4, 4, 1025) for _ in range(2))
        pdf, N = _gmm(lb, 2, 4, 0, 0, full=True), 100_000_000
        what = "C2: 2-D DensityGrid 1024^2, K=4 full-covariance GMM, N=1e8, fp32"
    elif name == "C3F64":
        edges, w, means, sig = synthetic_c3()
        pdf, N, dtype = lb.GaussianMixture(w, means, sigmas=sig), 1_000_000_000, np.float64
        what = "C3 in the reference's precision: 3-D DensityGrid 256^3, N=1e9, fp64 densities and samples"
    elif name == "C4":
        edges = tuple(np.linspace(-4, 4, 33) for _ in range(5))
        pdf, N, qmc = _gmm(lb, 5, 4, 0.5, 0.9), 100_000_000, True
        what = ("C4: 5-D DensityGrid 32^5 (32 corners per cell), N=1e8, qmc=True: device Sobol points "
                "bit-identical to scipy's Sobol(d=6, bits=32, seed=1), rows permuted on the device")
    elif name == "C5":
        pdf, N = _gmm(lb, 3, 8, 0.02, 0.2), 100_000_000
        t0 = time.perf_counter()
        tree = lb.DensityTree([-4.0] * 3, [4.0] * 3, pdf, vectorizedpdf=True, min_openings=4, dtype=np.float32)
        tree.refine(1e-3, leaf_tol=1e-3)
        tree_s = time.perf_counter() - t0
        what = ("C5: 3-D DensityTree, K=8 sharp GMM, min_openings=4, refine(1e-3, leaf_tol=1e-3): %d leaves "
                "(built in the device cell table, lint_tree_open, in %.3f s incl. the first-call warm-up, outside "
                "the timed region), N=1e8, fp32" % (len(tree._leaf_ids if tree._dt is not None else tree.leaves), tree_s))
    else:
        raise SystemExit(f"unknown --config {args.config}")
    isz = np.dtype(dtype).itemsize
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    if tree is None:
        kw = {} if qmc else dict(cell_records=False, guide_bits=0)
        grid = lb.DensityGrid(edges, pdf, vectorizedpdf=True, dtype=dtype, device=dev, **kw)
        k, C_, V_ = grid.dim, int(grid.ncells), int(np.prod([len(e) for e in edges]))
        alg = V_ * isz + C_ * 8 + 2 * C_ * 8 + N * k * isz
    else:
        tree._ensure_device()
        k = 3
        alg = N * k * isz + len(tree.leaves) * (8 + 2 * 3) * isz
    out = None if qmc else torch.empty((N, k), dtype=tdt, device=dev)
    sampler = None
    if qmc:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sampler = lb.LintSampler(grid, seed=SEED, qmc=True, return_tensor=True)

    def step(i):
        if tree is not None:
            tree._sample_cellmajor(N, SEED, i * N, out=out)
        elif qmc:
            grid._enqueue_build()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")                 # scipy's non-power-of-two warning
                sampler.sample(N)
        else:
            grid._enqueue_build()
            grid._sample_cellmajor(N, SEED, i * N, out=out)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clocks:
        a.record()
        for i in range(args.steps):
            step(args.warmup + i)
        b.record()
        torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "samples/s", "value": N / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
        "config": {"workload": what, "samples_total": N,
                   "order": "scipy-identical Sobol point set, device row permutation" if qmc else
                            "rows grouped by cell (order=\"cell\")"},
        "roofline": {"bound": "hbm", "kernel": "whole step (tables + draw)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "algorithmic_bytes_per_step": alg, "traffic": None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650"},
        "clocks": clocks.summary(),
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=N_SAMPLES,
                    help="samples per step in total (strong scaling); with --weak: per GPU")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --samples rows per GPU per step")
    ap.add_argument("--config", default=None, help="one of the other named configs: C2, C3f64, C4, C5 (one GPU)")
    ap.add_argument("--order", default="cell", choices=["cell", "draw"],
                    help="cell: rows grouped by cell (throughput mode); draw: u-driven per-sample search")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--equal-mass", action="store_true",
                    help="equal-mass strata instead of cost-balanced ones (multi-GPU)")
    ap.add_argument("--no-calibrate", action="store_true",
                    help="keep the analytic cost-balanced strata (no timed trial steps at setup)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the 2-step run of the other row order")
    ap.add_argument("--guide-bits", type=int, default=None, help="tuning: log2 guide-table size")
    ap.add_argument("--no-records", action="store_true", help="tuning: gather corners from the vertex array")
    ap.add_argument("--grid-n", type=int, default=None, help="tuning: cells per dim (default 256)")
    args = ap.parse_args()
    if args.grid_n:
        global GRID_N
        GRID_N = args.grid_n
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config:
        run_config(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
