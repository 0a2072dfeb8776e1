"""Device plumbing: PyTorch owns device memory and streams, nothing else."""
import ctypes as C

import numpy as np

from . import _capi


def on_device(method):
    """Run a method with the owner's CUDA device current: kernels launch on the current
    device, which must be the one that holds the buffers (``self.device``)."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        dev = getattr(self, "device", None)
        if dev is None:
            return method(self, *args, **kwargs)
        import torch
        with torch.cuda.device(dev):
            return method(self, *args, **kwargs)
    return wrapper


def require_cuda(device=None):
    """Return a torch CUDA device or raise: the product has no CPU path."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError(
            "lintsampler_b200 needs a CUDA device (B200, sm_100a); none is visible. "
            "There is no CPU fallback.")
    _capi.load()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"lintsampler_b200: device must be a CUDA device, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Raw device pointer of a (contiguous) torch tensor, or NULL for None."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def torch_dtype(np_dtype):
    import torch
    return torch.float32 if np.dtype(np_dtype) == np.float32 else torch.float64


def lint_dtype(np_dtype):
    return _capi.LINT_F32 if np.dtype(np_dtype) == np.float32 else _capi.LINT_F64


def check_dtype(dtype):
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError(f"lintsampler_b200: dtype must be float32 or float64, got {dt}")
    return dt


def default_guide_bits(ncells):
    """Guide-table size.  One bin per cell while the table stays small (<= 2^20 entries,
    4 MB): bins that lie inside one cell answer a query without touching the CDF, so a
    finer table means fewer probes.  Beyond that the table — addressed uniformly at
    random — competes with the hot cells for L2, and one bin per four cells measured
    best (2^22 beat 2^24 and 2^20 at 256^3; profiles/README.md)."""
    bits = max(4, int(np.ceil(np.log2(max(int(ncells), 2)))))
    if bits > 20:
        bits = max(20, bits - 2)
    return min(bits, 26)


def build_cdf(masses_t, mode, device):
    """masses (n,) fp64 device tensor -> (cdf, guide, guide_bits, total_mass float).

    Wraps lint_cdf_build + lint_guide_build; raises the ValueError the drop-in
    documents for an all-zero density (the reference silently divides 0/0 there,
    grid.py:431 — SURVEY.md §8b quirks).
    """
    import torch
    n = masses_t.numel()
    mode_id = {"parallel": _capi.LINT_CDF_PARALLEL, "sequential": _capi.LINT_CDF_SEQUENTIAL}.get(mode)
    if mode_id is None:
        raise ValueError(f"cdf_mode must be 'parallel' or 'sequential', got {mode!r}")
    lib = _capi.load()
    nbytes = lib.lint_cdf_scratch_bytes(n)
    scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=device)
    cdf = torch.empty(n, dtype=torch.float64, device=device)
    total = torch.empty(1, dtype=torch.float64, device=device)
    st = stream_ptr(device)
    _capi.call("lint_cdf_build", ptr(masses_t), n, mode_id, ptr(cdf), ptr(total), ptr(scratch),
               nbytes, st)
    total_mass = float(total.item())
    if not (total_mass > 0.0) or not np.isfinite(total_mass):
        raise ValueError(
            "lintsampler_b200: total mass of the domain is not positive and finite "
            f"({total_mass}); cannot build a cell CDF")
    bits = default_guide_bits(n)
    guide = torch.empty((1 << bits) + 1, dtype=torch.int32, device=device)
    _capi.call("lint_guide_build", ptr(cdf), n, bits, ptr(guide), st)
    return cdf, guide, bits, total_mass


def cellmajor_scratch(owner, ncells, N):
    """Scratch buffer for the cell-major sampler, cached on ``owner`` and grown on demand."""
    import torch
    nbytes = int(_capi.load().lint_cellmajor_scratch_bytes(ncells, N))
    buf = getattr(owner, "_cm_scratch", None)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=owner.device)
        owner._cm_scratch = buf
    return buf, nbytes


CELLMAJOR_MAX_ROWS = 1 << 31     # lint_*_sample_cellmajor takes N <= 2^32 - 2^16; larger draws are chunked


def cellmajor_chunks(N):
    """(start, rows) pieces of an N-row draw, each within the C ABI's per-call limit."""
    start = 0
    while start < N:
        rows = min(N - start, CELLMAJOR_MAX_ROWS)
        yield start, rows
        start += rows
