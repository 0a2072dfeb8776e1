"""In-tree build of the CUDA library: nvcc, sm_100a only.

    python -m lintsampler_b200._build          # build if stale
    python -m lintsampler_b200._build --force

The resulting ``_lib/liblintsampler_b200.so`` is git-ignored but travels to the
GPU box with the working tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "liblintsampler_b200.so")
# (source, object stem, extra defines): the cell-major kernels are compiled once per
# dimension so that the eight translation units build in parallel
SOURCES = [("lint_build.cu", "lint_build", []), ("lint_sample.cu", "lint_sample", []),
           ("lint_cellmajor.cu", "lint_cellmajor", []), ("lint_tree.cu", "lint_tree", [])]
# LINT_DIMS=3 (development only): build the cell-major kernels of the listed dimensions, stubs for the rest
_DIMS = [int(d) for d in os.environ.get("LINT_DIMS", "1,2,3,4,5,6").split(",")]
SOURCES += [("lint_cellmajor_inst.cu", f"lint_cellmajor_k{k}",
             [f"-DLINT_K={k}"] + ([] if k in _DIMS else ["-DLINT_STUB"])) for k in range(1, 7)]
HEADERS = ["lint_device.cuh", "lint_pdf.cuh", "lint_pairwise.cuh", "lint_cellmajor.cuh", "lint_cellmajor_plan.h",
           os.path.join("..", "..", "include", "lintsampler_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _plan(verbose=False):
    """[(stem, obj, stamp, cmd, up_to_date)] for every translation unit."""
    nvcc = _nvcc()
    plan = []
    for src, stem, defines in SOURCES:
        obj = os.path.join(OUT_DIR, stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *defines, *os.environ.get("LINT_EXTRA_NVCC", "").split(), "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        # an object is reused when it is newer than its sources and was built by the same command
        heads = HEADERS if "cellmajor" in src else [h for h in HEADERS if "cellmajor" not in h]
        deps = [os.path.join(CSRC, d) for d in [src] + heads]
        stamp = obj + ".cmd"
        fresh = (os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == " ".join(cmd)
                 and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps))
        plan.append((stem, obj, stamp, cmd, fresh))
    return plan


def is_stale():
    if not os.path.exists(OUT):
        return True
    if not os.path.isdir(OUT_DIR) or not any(f.endswith(".o") for f in os.listdir(OUT_DIR)):
        # only the library travelled (no objects): compare it with the sources
        t = os.path.getmtime(OUT)
        deps = [os.path.join(CSRC, s) for s in sorted({src for src, _, _ in SOURCES}) + HEADERS]
        return any(os.path.getmtime(d) > t for d in deps)
    plan = _plan()
    return any(not fresh or os.path.getmtime(obj) > os.path.getmtime(OUT) for _, obj, _, _, fresh in plan)


def build(force=False, verbose=False):
    """Compile every stale .cu (in parallel) and link one shared library; returns its path."""
    if not force and not verbose and not is_stale():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    procs = []
    plan = _plan(verbose)
    for stem, obj, stamp, cmd, fresh in plan:
        if fresh and not force and not verbose:
            continue
        if verbose:
            print(" ".join(cmd))
        if os.path.exists(stamp):
            os.remove(stamp)
        procs.append((stem, stamp, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = None
    for stem, stamp, cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stdout.write(out.decode())
        if p.returncode:
            failed = failed or stem
        else:
            with open(stamp, "w") as f:
                f.write(" ".join(cmd))
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    cmd = [_nvcc(), "-Wno-deprecated-gpu-targets", "-shared", "-o", OUT, *[obj for _, obj, _, _, _ in plan]]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
