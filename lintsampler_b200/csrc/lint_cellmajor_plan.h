// Host-side plan of a cell-major draw: scratch layout and row-budget helpers shared by the
// C-ABI file (lint_cellmajor.cu) and the per-dimension kernel translation units
// (lint_cellmajor_inst.cu, compiled once per LINT_K so that the build runs in parallel).
#pragma once
#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lintsampler_b200.h"

namespace lint {

constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
// Rows per warp work item ("chunk"): 4096 amortises the per-chunk set-up best (4.64 against
// 4.73 ms per 1e9 rows with 2048), but a draw needs enough chunks to balance ~9500 resident warps.
constexpr uint32_t EMIT_CHUNK_MIN = 2048, EMIT_CHUNK_MAX = 4096;      // powers of two (scan_compact_kernel shifts by them)
static_assert((EMIT_CHUNK_MIN & (EMIT_CHUNK_MIN - 1)) == 0 && (EMIT_CHUNK_MAX & (EMIT_CHUNK_MAX - 1)) == 0, "chunk sizes");
inline uint32_t emit_chunk_rows(uint64_t N) { return N >= (1ull << 28) ? EMIT_CHUNK_MAX : EMIT_CHUNK_MIN; }

// One segment list per run: the flat parts (cid = cell) and the rest (cid = 2 cell + 1 with
// the split, the cell itself without).
struct SegList {
    uint32_t *cid, *coff, *chunk_start, *chunk_start1, *nnz, *rows, *tile_counter;
    unsigned long long *tile_state;
};
struct Scratch {
    uint32_t *counts_flat, *counts_rest;
    SegList flat, rest;
    void *boxes;                 // (lo, w) of every flat segment, 2 * LINT_MAX_DIM doubles at most
    uint32_t *plan;              // [0] T_f, [1] T_r, [2] T = first top-up row
    uint32_t *chunk_counter;     // the flat emitter's work counter
    uint32_t chunk_slots;        // entries of every chunk_start table
    void *zero_begin;            // [zero_begin, zero_begin + zero_bytes): cleared before every draw
    size_t zero_bytes;
    size_t bytes;
};

// A cell is split into its flat and residual parts only where that pays: below this
// many expected rows the residual part would be a handful of rows with a segment of
// their own, dearer than drawing all of the cell's rows from the full interpolant.
#ifndef LINT_SPLIT_MIN_LAMBDA
#define LINT_SPLIT_MIN_LAMBDA 128.0
#endif
constexpr double SPLIT_MIN_LAMBDA = LINT_SPLIT_MIN_LAMBDA;

// `box_bytes`: bytes of one flat segment's box (2 * dim values of the output type; the default is
// the largest).  Only split cells have a flat segment: those expecting >= SPLIT_MIN_LAMBDA rows, of
// which there are at most N / SPLIT_MIN_LAMBDA.
inline Scratch carve(void *base, uint64_t ncells, uint64_t N, size_t box_bytes = 2 * LINT_MAX_DIM * 8) {
    const uint64_t nscan = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    const uint64_t nchunks = (N + 1 + EMIT_CHUNK_MIN - 1) / EMIT_CHUNK_MIN;  // the most any draw uses
    uintptr_t p = (uintptr_t)base;
    auto take = [&](size_t n) { uintptr_t q = p; p += (n + 255) & ~(size_t)255; return (void *)q; };
    Scratch s;
    s.chunk_slots = (uint32_t)(nchunks + 2);
    s.counts_flat = (uint32_t *)take(ncells * 4);
    s.counts_rest = (uint32_t *)take(ncells * 4);
    for (SegList *l : {&s.flat, &s.rest}) {
        l->cid = (uint32_t *)take((ncells + 1) * 4);
        l->coff = (uint32_t *)take((ncells + 1) * 4);
        // first segment of every output chunk, written by the scan: for a run that starts on an even
        // global row, and (chunk_start1) on an odd one — the general run is laid over even rows
        l->chunk_start = (uint32_t *)take((nchunks + 2) * 4);
        l->chunk_start1 = (uint32_t *)take((nchunks + 2) * 4);
    }
    const uint64_t nbox = std::min<uint64_t>(ncells, (uint64_t)((double)N / SPLIT_MIN_LAMBDA) + 1);
    s.boxes = take(nbox * box_bytes);
    s.plan = (uint32_t *)take(16);
    s.zero_begin = (void *)p;
    s.chunk_counter = (uint32_t *)take(4);
    for (SegList *l : {&s.flat, &s.rest}) {
        l->tile_state = (unsigned long long *)take(nscan * 8);
        l->nnz = (uint32_t *)take(4);
        l->rows = (uint32_t *)take(4);
        l->tile_counter = (uint32_t *)take(4);
    }
    s.zero_bytes = (size_t)(p - (uintptr_t)s.zero_begin);
    s.bytes = (size_t)(p - (uintptr_t)base);
    return s;
}

// Expected number of grouped rows and the bound on the top-up tail: the Poisson
// total T has mean L = N - 8 sqrt(N) and standard deviation sqrt(L) < sqrt(N), so
// N - T lies in (0, 16 sqrt(N)) except for 8-sigma events on either side.
inline double grouped_mean(uint64_t N) { return std::max(0.0, (double)N - 8.0 * std::sqrt((double)N)); }
inline uint64_t topup_bound(uint64_t N) {
    return std::min<uint64_t>(N, (uint64_t)(16.0 * std::sqrt((double)N)) + 64);
}

struct Stratum { double u_lo, u_hi; };

constexpr uint64_t SPLIT_MIN_ROWS_PER_CELL = 32;
// (a draw from a stratum of width w of the cell CDF is judged as the N / w rows of the whole
// domain it is a part of: a rank of a sharded run splits exactly the cells a single GPU would)
inline bool use_split(uint64_t ncells, uint64_t N, double stratum_width = 1.0) {
    return (double)N >= (double)(SPLIT_MIN_ROWS_PER_CELL * ncells) * stratum_width;
}


// Optional timing of the two emitters (lint_cellmajor_timing): CUDA events recorded on the
// caller's stream around each of them.  Defined in lint_cellmajor.cu.
struct EmitTimers {
    bool enabled = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // flat begin/end, general begin/end
    bool flat_ran = false;
    const uint32_t *plan_dev = nullptr;                          // row budget of the timed call
};
EmitTimers &emit_timers();

// Side streams of a draw (one set per device, created on first use; lint_cellmajor.cu).  The
// pieces of a draw that do not depend on one another are enqueued side by side — the two
// scans, the flat emitter next to the general one, the top-up next to both — joined by events,
// so the call is still complete, in stream order, on the caller's stream (and a CUDA graph
// captured from it holds the same fork / join structure).  NULL: everything on the caller's
// stream (LINT_SIDE=0, or while lint_cellmajor_timing is on).
struct Lanes {
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
Lanes *side_lanes();

// per-dimension entry points, instantiated in lint_cellmajor_inst.cu
template <int K>
int grid_cellmajor_k(const lint_grid_t *g, uint64_t nc, uint64_t N, uint2 key, uint64_t offset, Stratum u,
                     const uint32_t *range, const Scratch &s, void *out, int64_t *cell_idx, cudaStream_t st);
template <int K>
int cells_cellmajor_k(const lint_cells_t *c, uint64_t N, uint2 key, uint64_t offset, Stratum u,
                      const Scratch &s, void *out, int64_t *cell_idx, cudaStream_t st);

}  // namespace lint
