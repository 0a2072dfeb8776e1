// Cell-major ("grouped") sampling: the throughput path.
//
// The u-driven sampler of lint_sample.cu spends its time on per-sample random
// access (guide, CDF, cell record) — work the reference also does per sample
// (utils.py:197, grid.py:440-467).  N i.i.d. draws of a cell are however fully
// described by the per-cell counts (n_c) ~ Multinomial(N, p_c), so this path
//   1. draws independent counts n_c ~ Poisson(L p_c) with L = N - 8 sqrt(N), one
//      thread per cell (probabilities are differences of the same fp64 CDF the
//      search path uses, grid.py:431 + utils.py:195-196).  Given their total T the
//      counts are exactly Multinomial(T, p), and T < N except with probability
//      ~1e-15 (an 8-sigma event),
//   2. turns them into row offsets with a single-pass decoupled look-back scan
//      (integer sums, hence deterministic) that also compacts away empty cells,
//   3. emits each cell's rows contiguously: the cell's box and corners are
//      loaded once per cell, every row costs only its k in-cell uniforms and the
//      k-D inverse CDF of sampling.py:41-94,
//   4. tops up the remaining N - T rows (~2.5e-4 N) with the u-driven kernel;
//      Multinomial(T, p) + Multinomial(N - T, p) = Multinomial(N, p), so the N
//      rows are an exact i.i.d. sample as a multiset.
// Flat/residual split.  Inside a cell the interpolant is f = f_min + (f - f_min): a
// constant plus a multilinear density that vanishes at one corner.  Rows of the
// constant part are plain uniforms (z = u, no inverse CDF at all); only rows of the
// residual need sampling.py:41-94.  A split cell c becomes two segments, 2c (flat)
// and 2c + 1 (residual), with independent Poisson counts of means lambda q and
// lambda (1 - q), q = f_min / mean(corners) — Poisson thinning, exact.  On a fine
// grid of a smooth pdf q is close to 1 (0.94 mass-weighted on the 256^3 benchmark
// grid), so most rows cost their uniforms, an fma and a store.  The split doubles
// the segment table and adds a corner gather to the count pass, so it is used only
// when N >= SPLIT_MIN_ROWS_PER_CELL * cells, and within such a draw only for cells
// expecting >= SPLIT_MIN_LAMBDA rows (the others keep all rows in segment 2c + 1,
// drawn from the full interpolant; the emitter re-derives the decision from the
// same CDF entries).
// Row ORDER: three runs.  Rows [0, T_f) are the flat parts of the split cells, grouped by
// cell in C order; rows [T_f, T) are everything else drawn by count — the residual parts of
// the split cells and all rows of the cells that are not split — again grouped by cell in C
// order; rows [T, N) are the top-up tail in draw order.  Keeping the flat rows apart lets a
// lean kernel of their own write them (no quantile, no per-cell coefficients, 16-byte
// coalesced stores) while the general emitter works on the rest.  Callers that need the
// reference's exchangeable row order (lintsampler.py:309-311) apply the row permutation
// afterwards.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <cuda_pipeline.h>
#pragma once
#include "lint_device.cuh"
#include "lint_cellmajor_plan.h"

// Checked build (-DLINT_CHECKED; scripts/checked_build.sh): every index the emitters and the scan
// compute is asserted against its bounds on the device.  compute-sanitizer is closed on the GPU
// pool this was developed on, so the cell-major tests are run once against this build instead.
#ifdef LINT_CHECKED
#include <cassert>
#define LINT_CHECK(cond) assert(cond)
#else
#define LINT_CHECK(cond) ((void)0)
#endif

namespace lint {

int validate_grid(const lint_grid_t *g, bool need_cdf, uint64_t *ncells_out, uint64_t *nverts_out);
template <int K> GridView<K> make_grid_view(const lint_grid_t *g, uint64_t ncells, bool stage_edges);

// Philox counter word 1 tags (word 0 = cell / first row of a group, words 2-3 = call offset)
constexpr uint32_t TAG_COUNT = 0x53000000u;   // | attempt   (per-cell count draws)
constexpr uint32_t TAG_COORD = 0x43000000u;   // | block index

// ---------------------------------------------------------------------------
// Poisson counts
// ---------------------------------------------------------------------------
struct CellRng {
    uint2 key;
    uint32_t cell, w2, w3;
    uint32_t attempt = 0;
    __device__ __forceinline__ void next(double &a, double &b) {
        const uint4 r = philox4x32_10(make_uint4(cell, TAG_COUNT | attempt, w2, w3), key);
        ++attempt;
        a = u53(r.x, r.y);
        b = u53(r.z, r.w);
    }
};

// log(k!) for integer-valued k >= 0: table below 16, Stirling series above
// (truncation error < 2e-12 there)
static __constant__ double LOG_FACTORIAL_TAB[16] = {
    0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458, 4.787491742782046,
    6.579251212010101, 8.525161361065415, 10.604602902745251, 12.801827480081469,
    15.104412573075516, 17.502307845873887, 19.987214495661885, 22.552163853123425,
    25.19122118273868, 27.89927138384089};
__device__ __forceinline__ double log_factorial(double k) {
    if (k < 16.0) return LOG_FACTORIAL_TAB[(int)k];
    const double x = k + 1.0, r = 1.0 / x, r2 = r * r;
    return (x - 0.5) * log(x) - x + 0.9189385332046727 +
           r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
}

// Poisson(lam): sequential inversion with one uniform below 10, Hörmann's
// transformed rejection with squeeze (PTRS, 1993) above.  Split in two so that the count pass
// can settle the common cases at once (zero mean; lam < 10 and u <= 1 - lam <= exp(-lam): no
// rows, most cells of a fine grid) and queue the others for a second phase in which every lane
// of a warp has real work of the same kind.
constexpr double POISSON_PTRS_FROM = 10.0;

// lam < 10, first uniform u already drawn (rng.attempt == 1) and u > 1 - lam
__device__ __forceinline__ uint32_t poisson_inversion(double lam, double u, CellRng &rng) {
    const double p0 = exp(-lam);
    for (;;) {
        double p = p0;
        uint32_t x = 0;
        bool ok = true;
        while (u > p) {
            u -= p;
            ++x;
            if (x > 200) { ok = false; break; }     // fp dust past the far tail: redraw
            p *= lam * __drcp_rn((double)x);
        }
        if (ok) return x;
        double unused;
        rng.next(u, unused);
    }
}

// PTRS in two stages, so that the count pass can run each stage with full warps.
// Stage A, one attempt: proposal k and the squeeze.  Returns 1 accepted (k final), 0 rejected
// outright (try again), 2 undecided: stage B must test (k, lhs_arg).
__device__ __forceinline__ int ptrs_propose(double lam, double b, CellRng &rng, double &k, double &lhs_arg) {
    const double a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 * __drcp_rn(b - 2.0);
    double u, v;
    rng.next(u, v);
    u -= 0.5;
    const double us = 0.5 - fabs(u);
    const double ius = __drcp_rn(us);
    k = floor((2.0 * a * ius + b) * u + lam + 0.43);
    if (us >= 0.07 && v <= vr) return 1;                          // the squeeze accepts ~86 %
    if (k < 0.0 || (us < 0.013 && v > us)) return 0;
    const double invalpha = 1.1239 + 1.1328 * __drcp_rn(b - 3.4);
    lhs_arg = v * invalpha * __drcp_rn(a * ius * ius + b);
    return 2;
}

// Stage B: the acceptance test log(lhs_arg) <= -lam + k log(lam) - log(k!).  It is first decided
// in fp32 wherever the margin exceeds a generous bound on the fp32 error; the fp64 test settles
// the rest (a few per thousand).  With k1 = k + 1, d = k1 - lam, x = d / k1 and Stirling's series
// the right-hand side is  k log1p(-x) + d - log(k1)/2 - log(2 pi)/2 - s(k1).
// (scripts/ptrs_fp32_screen_check.py replays this screen and bounds its error.)
__device__ __forceinline__ bool ptrs_accept(double lam, double k, double lhs_arg) {
    if (k >= 4.0 && lam < 1.0e5) {
        const float d = (float)(k + 1.0 - lam), k1 = (float)k + 1.0f, x = d / k1;
        const float series = (1.0f / 12.0f - (1.0f / 360.0f) / (k1 * k1)) / k1;
        const float rhs = ((float)k * log1pf(-x) + d) - 0.5f * __logf(k1) - 0.9189385f - series;
        const float t = __logf((float)lhs_arg) - rhs;
        const float tol = 2e-3f + 4e-6f * fabsf(d);
        if (t < -tol) return true;
        if (t > tol) return false;
    }
    return log(lhs_arg) <= -lam + k * log(lam) - log_factorial(k);
}

// The table the counts are read from: the normalised CDF restricted to the stratum
// [u_lo, u_hi) of the cell-selecting uniform ((0, 1) = the whole domain; a rank of
// a sharded run passes its own slice).
struct CdfSlice {
    const double *cdf;
    uint32_t ncells;
    double u_lo, u_hi;
};

// Expected number of grouped rows of cell c — evaluated identically by the count pass and
// by the emitter (same operations on the same table).
__device__ __forceinline__ double cell_lambda(const CdfSlice &t, uint32_t c, double lam_per_unit) {
    // mass of the cell inside the stratum: clamp both CDF values, then difference
    const double f0 = c == 0 ? 0.0 : __ldg(t.cdf + c - 1), f1 = __ldg(t.cdf + c);
    const double m = fmin(fmax(f1, t.u_lo), t.u_hi) - fmin(fmax(f0, t.u_lo), t.u_hi);
    return __dmul_rn(m, lam_per_unit);
}


// One Poisson draw from the stream of (cell, part)
__device__ __forceinline__ uint32_t poisson(double lam, CellRng &rng) {
    if (!(lam > 0.0)) return 0;
    if (lam < POISSON_PTRS_FROM) {
        double u, unused;
        rng.next(u, unused);
        // exp(-lam) >= 1 - lam: most cells of a fine grid hold lam << 1 and end here,
        // without the exponential
        if (u <= 1.0 - lam) return 0;
        return poisson_inversion(lam, u, rng);
    }
    const double b = 0.931 + 2.53 * sqrt(lam);
    for (;;) {
        double k, arg;
        const int verdict = ptrs_propose(lam, b, rng, k, arg);
        if (verdict == 1 || (verdict == 2 && ptrs_accept(lam, k, arg))) return (uint32_t)k;
    }
}

// counts_rest[c] = rows of cell c drawn from its full interpolant (no split), or with the
// split: counts_flat[c] = flat rows, counts_rest[c] = residual rows — or all rows of a cell
// that is not split.  (Settling the cheap cases first and working the expensive ones off from
// shared-memory queues with full warps was tried twice and lost to this plain form by 30-45 %:
// profiles/README.md, r2e.)
constexpr int POISSON_THREADS = 256;
#ifndef LINT_POISSON_MINB
#define LINT_POISSON_MINB 6      // 40 registers: 1536 threads per SM (measured best of 4 / 6 / 8)
#endif
template <typename Cells>
__global__ void __launch_bounds__(POISSON_THREADS, LINT_POISSON_MINB)
poisson_counts_kernel(Cells src, CdfSlice t, double lam_per_unit, uint2 key, StreamPos pos, bool split,
                      const uint32_t *__restrict__ range, uint32_t *__restrict__ counts_flat,
                      uint32_t *__restrict__ counts_rest, uint4 *__restrict__ zero, uint32_t nzero) {
    pdl_release();
    pdl_wait();                                            // the CDF and the slab: the build, if it is the kernel before
    // the scans' tile states and totals (Scratch::zero_begin), cleared here instead of by a memset
    // node of its own: the scans are the next kernels in the stream
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nzero; i += gridDim.x * blockDim.x)
        zero[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint64_t offset = pos.get();
    const uint32_t w2 = (uint32_t)offset, w3 = (uint32_t)(offset >> 32);
    // cells [c0, c1): the slab of the stratum (lint_grid_build), or every cell
    const uint32_t c0 = range ? range[0] : 0u, c1 = range ? range[1] : t.ncells;
    for (uint32_t c = c0 + blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += gridDim.x * blockDim.x) {
        const double lam = cell_lambda(t, c, lam_per_unit);
        if (!split) {
            CellRng rng{key, c, w2, w3};
            counts_rest[c] = poisson(lam, rng);
            continue;
        }
        double mean[2] = {0.0, lam};
        if (lam >= SPLIT_MIN_LAMBDA) {
            mean[0] = lam * src.flat_fraction(c);
            mean[1] = lam - mean[0];
        }
        uint32_t n[2];
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {           // one copy of the sampler's code
            CellRng rng{key, 2 * c + part, w2, w3};
            n[part] = poisson(mean[part], rng);
        }
        counts_flat[c] = n[0];
        counts_rest[c] = n[1];
    }
}

// ---------------------------------------------------------------------------
// decoupled look-back scan of (count, non-empty) pairs + compaction
// ---------------------------------------------------------------------------
// A tile publishes its aggregate, then its inclusive prefix, in one 64-bit word:
// bits 63-62 status, bits 61-32 non-empty cells, bits 31-0 rows.  Integer sums
// are associative, so the look-back's variable association is harmless here.
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long pack(uint32_t rows, uint32_t cells) {
    return ((unsigned long long)cells << 32) | rows;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_compact_kernel(const uint32_t *counts, uint32_t ncells_all, const uint32_t *__restrict__ range,
                    uint32_t id_mul, uint32_t id_add, unsigned long long *tile_state, uint32_t *tile_counter,
                    uint32_t *cid, uint32_t *coff, uint32_t *nnz_out, uint32_t *rows_out, uint32_t chunk_rows,
                    uint32_t chunk_slots, uint32_t *__restrict__ chunk0, uint32_t *__restrict__ chunk1) {
    pdl_wait();                                            // the counts: poisson_counts_kernel
    const int chunk_shift = 31 - __clz(chunk_rows);
    // cells [first, first + ncells) of the count array; ids are global cell indices
    const uint32_t first = range ? range[0] : 0u;
    const uint32_t ncells = range ? range[1] - range[0] : ncells_all;
    counts += first;
    id_add += first * id_mul;
    if ((uint64_t)blockIdx.x * SCAN_TILE >= ncells) return;       // (the grid is sized for every cell)
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);   // tiles start in index order
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS];
    unsigned long long run = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        c[i] = base + i < ncells ? __ldg(counts + base + i) : 0;
        run += pack(c[i], c[i] != 0);
    }
    // block-wide inclusive scan of the per-thread totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long wex = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) wex += s_warp[w];
        tile_total += s_warp[w];
    }
    // look-back (warp 0)
    if (warp == 0) {
        if (lane == 0) {
            const unsigned long long st = tile == 0 ? (ST_INC | tile_total) : (ST_AGG | tile_total);
            atomicExch(tile_state + tile, st);
        }
        unsigned long long excl = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            for (;;) {
                const int idx = look - lane;
                unsigned long long v = ST_INC;                      // virtual tile -1: empty inclusive
                if (idx >= 0) {
                    do { v = atomicAdd(tile_state + idx, 0ull); } while ((v & ST_MASK) == 0);
                }
                const uint32_t inc_mask = __ballot_sync(0xffffffffu, (v & ST_MASK) == ST_INC);
                const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32;
                unsigned long long contrib = lane <= first_inc ? (v & ~ST_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                excl += contrib;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(tile_state + tile, ST_INC | (excl + tile_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const unsigned long long pre0 = s_excl + wex + (inc - run);    // exclusive prefix of this thread
    unsigned long long pre = pre0;
    // Chunk t of the run starts in the segment that holds its first row: t * chunk_rows when the run
    // begins on an even global row, max(t * chunk_rows, 1) - 1 when on an odd one (the general run
    // is laid over even rows, see emit_kernel).  A segment of rows [lo, hi) owns the chunks
    // [ceil(lo / R), ceil(hi / R)) of the first numbering and [(lo + R) / R, (hi + R) / R) — and
    // chunk 0 if lo = 0 — of the second: none or one for almost every segment, written by its lane;
    // a segment of more than eight chunks' worth of rows (a heavy tree leaf) is left to the whole warp.
    // (R is a power of two: shifts.  t < chunk_slots: the tables are sized for the N rows asked for;
    // an 8-sigma Poisson total beyond that is clamped away by run_rows and never emitted.)
    auto chunk_span = [&](uint64_t lo, uint64_t hi, uint32_t &t0, uint32_t &t1, uint32_t &u0, uint32_t &u1) {
        t0 = (uint32_t)min((lo + chunk_rows - 1) >> chunk_shift, (uint64_t)chunk_slots);
        t1 = (uint32_t)min((hi + chunk_rows - 1) >> chunk_shift, (uint64_t)chunk_slots);
        u0 = (uint32_t)min((lo + chunk_rows) >> chunk_shift, (uint64_t)chunk_slots);
        u1 = (uint32_t)min((hi + chunk_rows) >> chunk_shift, (uint64_t)chunk_slots);
    };
    constexpr uint32_t WIDE_CHUNKS = 8;                            // up to nine chunk starts per table by the lane itself
    bool any_wide = false;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (c[i] != 0) {
            const uint32_t slot = (uint32_t)(pre >> 32);
            LINT_CHECK(slot < ncells && base + i < ncells);
            cid[slot] = (base + i) * id_mul + id_add;
            coff[slot] = (uint32_t)pre;
            const uint64_t lo = (uint32_t)pre, hi = lo + c[i];
            if (chunk1 && lo == 0) chunk1[0] = slot;
            if (c[i] > WIDE_CHUNKS * chunk_rows) {
                any_wide = true;
            } else {
                uint32_t t0, t1, u0, u1;
                chunk_span(lo, hi, t0, t1, u0, u1);
                for (uint32_t t = t0; t < t1; ++t) chunk0[t] = slot;
                if (chunk1)
                    for (uint32_t t = u0; t < u1; ++t) chunk1[t] = slot;
            }
        }
        pre += pack(c[i], c[i] != 0);
    }
    if (__ballot_sync(0xffffffffu, any_wide)) {                   // rare: second pass, warp per heavy segment
        unsigned long long q = pre0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const bool wide = c[i] > WIDE_CHUNKS * chunk_rows;
            uint32_t t0 = 0, t1 = 0, u0 = 0, u1 = 0;
            if (wide) chunk_span((uint32_t)q, (uint64_t)(uint32_t)q + c[i], t0, t1, u0, u1);
            const uint32_t slot = (uint32_t)(q >> 32);
            uint32_t todo = __ballot_sync(0xffffffffu, wide);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t sl = __shfl_sync(0xffffffffu, slot, src);
                const uint32_t a0 = __shfl_sync(0xffffffffu, t0, src), a1 = __shfl_sync(0xffffffffu, t1, src);
                const uint32_t b0 = __shfl_sync(0xffffffffu, u0, src), b1 = __shfl_sync(0xffffffffu, u1, src);
                for (uint32_t t = a0 + lane; t < a1; t += 32) chunk0[t] = sl;
                if (chunk1)
                    for (uint32_t t = b0 + lane; t < b1; t += 32) chunk1[t] = sl;
            }
            q += pack(c[i], c[i] != 0);
        }
    }
    if (base <= ncells - 1 && ncells - 1 < base + SCAN_ITEMS) {    // owner of the last cell
        const uint32_t nnz = (uint32_t)(pre >> 32);
        *nnz_out = nnz;
        *rows_out = (uint32_t)pre;                                 // T = total grouped rows
        coff[nnz] = (uint32_t)pre;                                 // sentinel
    }
}

// Row budget of the two grouped runs (for the top-up, which starts at row T, and for callers):
// plan[0] = T_f, plan[1] = T_r, plan[2] = T = T_f + T_r.  T <= N but for an 8-sigma event of the
// Poisson total; clamping keeps every writer inside the output whatever happened.  The emitters
// derive the same numbers themselves (run_rows) so that they need not wait for this kernel.
__device__ __forceinline__ void run_rows(const uint32_t *flat_rows, const uint32_t *rest_rows, uint32_t n_requested,
                                         uint32_t &tf, uint32_t &tr) {
    tf = flat_rows ? min(*flat_rows, n_requested) : 0u;
    tr = min(*rest_rows, n_requested - tf);
}
static __global__ void plan_totals_kernel(const uint32_t *flat_rows, const uint32_t *rest_rows, uint32_t n_requested,
                                          uint32_t *plan) {
    uint32_t tf, tr;
    run_rows(flat_rows, rest_rows, n_requested, tf, tr);
    plan[0] = tf;
    plan[1] = tr;
    plan[2] = tf + tr;
}

// ---------------------------------------------------------------------------
// emission
// ---------------------------------------------------------------------------
template <int K, typename T>
__device__ __forceinline__ double flat_fraction_of(const T (&c)[1 << K]) {
    T lo = c[0], sum = c[0];
#pragma unroll
    for (int j = 1; j < (1 << K); ++j) { lo = c[j] < lo ? c[j] : lo; sum += c[j]; }
    return sum > (T)0 ? fmin(fmax((double)lo * (double)(1 << K) / (double)sum, 0.0), 1.0) : 0.0;
}

// c <- c - min(c): the residual density of the flat/residual split
template <int NC, typename T>
__device__ __forceinline__ void subtract_min(T (&c)[NC]) {
    T lo = c[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) lo = c[j] < lo ? c[j] : lo;
#pragma unroll
    for (int j = 0; j < NC; ++j) c[j] -= lo;
}

template <int K, typename DensT>
struct GridCells {
    GridView<K> g;
    __host__ size_t smem_bytes(size_t elem) const { return (size_t)g.edge_total * elem; }
    template <typename T>
    __device__ __forceinline__ void prepare(T *smem) const { grid_stage_edges<K, T>(g, smem); }
    template <typename T>
    __device__ __forceinline__ void fetch(const T *smem, uint32_t idx, T (&c)[1 << K], T (&lo)[K],
                                          T (&hi)[K]) const {
        grid_fetch<K, DensT, T>(g, smem, idx, c, lo, hi);
    }
    __device__ __forceinline__ bool prenormalised() const { return g.rec != nullptr; }
    // q = min(corners) / mean(corners): the share of the cell's mass under the constant f_min
    __device__ __forceinline__ double flat_fraction(uint32_t idx) const {
        uint32_t i[K];
        grid_unravel<K>(g, idx, i);
        DensT c[1 << K];
        grid_corners<K, DensT, DensT>(g, idx, i, c);
        return flat_fraction_of<K>(c);
    }
    // the cell's box only: lower corner and widths (maxs - mins, sampling.py:126) in T
    template <typename T>
    __device__ __forceinline__ void box(const T *smem, uint32_t idx, T (&lo)[K], T (&w)[K]) const {
        uint32_t i[K];
        grid_unravel<K>(g, idx, i);
#pragma unroll
        for (int d = 0; d < K; ++d) {
            T hi;
            if (g.edge_total && smem) {
                lo[d] = smem[g.edge_off[d] + i[d]];
                hi = smem[g.edge_off[d] + i[d] + 1];
            } else {
                lo[d] = (T)__ldg(g.edges[d] + i[d]);
                hi = (T)__ldg(g.edges[d] + i[d] + 1);
            }
            w[d] = hi - lo[d];
        }
    }
};

template <int K, typename DensT>
struct ListCells {
    CellsView s;
    __host__ size_t smem_bytes(size_t) const { return 0; }
    template <typename T>
    __device__ __forceinline__ void prepare(T *) const {}
    template <typename T>
    __device__ __forceinline__ void fetch(const T *, uint32_t idx, T (&c)[1 << K], T (&lo)[K],
                                          T (&hi)[K]) const {
        cells_fetch<K, DensT, T>(s, idx, c, lo, hi);
    }
    __device__ __forceinline__ bool prenormalised() const { return s.rec != nullptr; }
    __device__ __forceinline__ double flat_fraction(uint32_t idx) const {
        DensT c[1 << K];
        cells_corners<K, DensT, DensT>(s, idx, c);
        return flat_fraction_of<K>(c);
    }
    template <typename T>
    __device__ __forceinline__ void box(const T *, uint32_t idx, T (&lo)[K], T (&w)[K]) const {
#pragma unroll
        for (int d = 0; d < K; ++d) {
            const double x = __ldg(s.mins + (size_t)idx * K + d);
            const double dx = __ldg(s.widths + (size_t)idx * K + d);
            lo[d] = (T)x;
            w[d] = (T)__dadd_rn(x, dx) - lo[d];       // maxs = leaf.x + leaf.dx [tree.py:348]
        }
    }
};

// ---------------------------------------------------------------------------
// flat emitter: rows [0, T_f), every segment the constant part of a split cell
// ---------------------------------------------------------------------------
// A flat row is K independent uniforms mapped into the cell's box, x_d = lo_d + w_d u
// (sampling.py:126 with z = u), so the run is a stream of independent ELEMENTS: element e
// of the row-major output belongs to row e / K, dimension e % K, and needs nothing but the
// box of that row's segment and a fresh uniform.  The kernel therefore works on 16-byte
// output vectors, not rows, in STEPS laid over the output itself: step s of a chunk covers
// vectors [v, v + 32 J) — a whole number of rows — and lane l writes vectors v + l,
// v + 32 + l, ...: every store instruction covers 512 contiguous bytes.  Because steps
// start at multiples of K elements, the dimension of a lane's i-th element is the same in
// every step: a lane keeps the box constants rotated once and for all (dsel below).  No row
// is ever assembled, nothing is shuffled.
//   * A step that lies inside one segment (the common case: flat segments are hundreds of
//     rows long) runs straight-line code with that segment's constants in registers.
//   * A step that holds a segment boundary (or the end of the run) stages the boxes of the
//     segments it touches in shared memory, one segment per lane, and every element picks
//     its own segment's constants; vectors are still written whole.
// Uniforms: fp32 elements take 21-bit fields, six to a Philox4x32-10 block (the in-cell
// position is resolved to 2^-21 of the cell width, finer than an fp32 ulp of any
// coordinate with |x| >= 4 cell widths), built as the mantissa of a float in [1, 2);
// fp64 elements take 53 bits, two to a block.  Counters name the lane's first vector of
// the step, so the value of every element depends on (seed, offset, N, position) only —
// not on which path or which warp produced it.
constexpr uint32_t TAG_FLAT = 0x46000000u;   // | bits 32.. of the vector index << 8 | block
#ifndef LINT_FLAT_THREADS
#define LINT_FLAT_THREADS 256
#endif
#ifndef LINT_FLAT_MINB
#define LINT_FLAT_MINB 3
#endif
constexpr int FLAT_THREADS = LINT_FLAT_THREADS;

template <typename T> struct FlatTraits;
template <> struct FlatTraits<float> {
    static constexpr int VEC = 4, FIELDS = 6;
    // field f of a block: bits [21 f, 21 f + 21) of the 128-bit word, centred, in [2^-22, 1 - 2^-22].
    // (m, o) = (0x007ffffc, 0x3f800002) held in registers by the caller, so that mask and
    // exponent are one three-input logic instruction.
    template <int F>
    static __device__ __forceinline__ float uniform(const uint4 &r, uint32_t m, uint32_t o) {
        uint32_t s;
        if (F == 0) s = r.x << 2;
        else if (F == 1) s = __funnelshift_r(r.x, r.y, 19);
        else if (F == 2) s = __funnelshift_r(r.y, r.z, 8);
        else if (F == 3) s = __funnelshift_r(r.y, r.z, 29);
        else if (F == 4) s = __funnelshift_r(r.z, r.w, 18);
        else s = r.w >> 7;
        return __uint_as_float((s & m) | o) - 1.0f;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
#if defined(LINT_FLAT_NOSTORE)
        if (v[0] == 12345.678f && v[1] == v[2] && v[3] == 1.0f)          // experiment: never true
#endif
#if defined(LINT_FLAT_PLAINSTORE)
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
#else
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
#endif
    }
};
template <> struct FlatTraits<double> {
    static constexpr int VEC = 2, FIELDS = 2;
    template <int F>
    static __device__ __forceinline__ double uniform(const uint4 &r, uint32_t, uint32_t) {
        return F == 0 ? u53(r.x, r.y) : u53(r.z, r.w);
    }
    static __device__ __forceinline__ void store(double *p, const double (&v)[2]) {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
    }
};

// Vectors per lane per step: a step (32 J vectors) must span a whole number of rows and
// divide a chunk; J = 3 or 5 where K has that factor, 4 otherwise.
template <int K, int VEC>
__host__ __device__ constexpr int flat_vectors_per_lane() {
    int base = 1;
    while ((32 * VEC * base) % K) ++base;
    return base == 1 ? 4 : base;
}

template <int F, int NF, typename T>
struct FlatFieldOf {                       // uniform #F of a lane's step: block F / NF, field F % NF
    static __device__ __forceinline__ T get(const uint4 *r, uint32_t m, uint32_t o) {
        return FlatTraits<T>::template uniform<F % NF>(r[F / NF], m, o);
    }
};

// f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>)
template <int I, int N, typename Fn>
__device__ __forceinline__ void static_for(Fn &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// Box of every flat segment, once per draw: boxes[j] = (lo[K], w[K]) of segment j's cell in the
// output type (w = maxs - mins, sampling.py:126).  The emitter reads a segment's constants from
// this table (a few coalesced loads) instead of unravelling the cell index and gathering its
// edges; the table is a fraction of a percent of the output and stays in L2.
template <int K, typename T, typename Cells>
__global__ void __launch_bounds__(256)
flat_boxes_kernel(Cells src, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ nnz_p,
                  T *__restrict__ boxes) {
    const uint32_t nnz = *nnz_p;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += gridDim.x * blockDim.x) {
        T lo[K], w[K];
        src.box((const T *)nullptr, __ldg(cid + j), lo, w);
#pragma unroll
        for (int d = 0; d < K; ++d) {
            boxes[(size_t)j * (2 * K) + d] = lo[d];
            boxes[(size_t)j * (2 * K) + K + d] = w[d];
        }
    }
}

template <int K, typename T, bool CELLS>
__global__ void __launch_bounds__(FLAT_THREADS, LINT_FLAT_MINB)
emit_flat_kernel(const T *__restrict__ boxes, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ coff,
                 const uint32_t *__restrict__ nnz_p, const uint32_t *__restrict__ flat_rows, uint32_t n_requested,
                 const uint32_t *__restrict__ chunk_start, uint32_t chunk_rows, uint32_t *chunk_counter,
                 const __grid_constant__ PhiloxKeys key, StreamPos pos, T *__restrict__ out,
                 int64_t *__restrict__ cell_idx) {
    using FT = FlatTraits<T>;
    const uint64_t offset = pos.get();
    const uint32_t fm = key.field_mask, fo = key.field_one;      // see FlatTraits<float>::uniform
    constexpr int VEC = FT::VEC, NF = FT::FIELDS;
    constexpr int J = flat_vectors_per_lane<K, VEC>();
    constexpr int NB = (VEC * J + NF - 1) / NF;        // Philox blocks per lane per step
    constexpr uint32_t STEP_ELEMS = 32u * VEC * J, STEP_ROWS = STEP_ELEMS / K;
    static_assert(STEP_ELEMS % K == 0 && EMIT_CHUNK_MIN % STEP_ROWS == 0, "steps tile rows and chunks");
    constexpr int WARPS = FLAT_THREADS / 32;
    constexpr int SLOT = 2 * K + 1;                    // staged box: lo[K], w[K]; odd stride
    __shared__ T s_box_all[WARPS][32 * SLOT];
    __shared__ uint32_t s_end_all[WARPS][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *s_box = s_box_all[warp];
    uint32_t *s_end = s_end_all[warp];
    const uint32_t nnz = *nnz_p;
    const uint32_t N = min(*flat_rows, n_requested);   // rows of the flat run (T_f)
    const uint32_t nchunks = (N + chunk_rows - 1) / chunk_rows;
    const uint32_t olo = (uint32_t)offset, ohi = (uint32_t)(offset >> 32);
    constexpr uint32_t FULL = 0xffffffffu;
    // dimension of the lane's elements: element i of vector jj has dimension
    // (VEC * lane + 32 * VEC * jj + i) % K = dsel[(32 * VEC * jj + i) % K]
    uint32_t dsel[K];
#pragma unroll
    for (int t = 0; t < K; ++t) dsel[t] = (VEC * lane + t) % K;

    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(chunk_counter, 1u);
        chunk = __shfl_sync(FULL, chunk, 0);
        if (chunk >= nchunks) break;
        const uint32_t r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, N);
        uint32_t j = __ldg(chunk_start + chunk);                           // segment that holds row r0
        LINT_CHECK(j < nnz && coff[j] <= r0 && r0 < coff[j + 1]);          // (the scan wrote the right owner)
        // 32-entry register window of the segment table: lane i holds segment jbase + i
        uint32_t jbase = j, w_off = 0xffffffffu;
        auto reload = [&](uint32_t jb) {
            jbase = jb;
            w_off = jb + lane <= nnz ? __ldg(coff + jb + lane) : 0xffffffffu;
        };
        reload(j);
        // Stage the boxes of the segments from jb on that start at or before row `upto`, as far as
        // the window reaches: window entry b (= segment jbase + b, one per lane) goes to slot b;
        // s_end[b] = the segment's end as an element index relative to row `rel` (saturated),
        // s_box[b] = (lo, w).  Returns the number staged (>= 1: segment jb) and the first slot.
        auto stage = [&](uint32_t jb, uint32_t upto, uint32_t rel, uint32_t &first) {
            if (jb - jbase >= 24u) reload(jb);
            first = jb - jbase;
            const uint32_t end = __shfl_down_sync(FULL, w_off, 1);         // lane 31: unused
            const bool valid = lane >= first && lane < 31 && jbase + lane < nnz && w_off <= upto;
            __syncwarp();
            if (valid) {
                const T *b = boxes + (size_t)(jbase + lane) * (2 * K);
#pragma unroll
                for (int d = 0; d < 2 * K; ++d) s_box[lane * SLOT + d] = __ldg(b + d);
                const uint64_t e = (uint64_t)(min(end, N) - rel) * K;      // end > rel: the segment reaches row rel
                s_end[lane] = (uint32_t)min(e, (uint64_t)0xfffffff0u);
            }
            __syncwarp();
            return (uint32_t)__popc(__ballot_sync(FULL, valid));
        };
        T rlo[K], rw[K];                               // constants of segment `held`, rotated for this lane
        uint32_t held = 0xffffffffu;
        auto take = [&](uint32_t slot) {               // registers <- staged slot
#pragma unroll
            for (int t = 0; t < K; ++t) {
                rlo[t] = s_box[slot * SLOT + dsel[t]];
                rw[t] = s_box[slot * SLOT + K + dsel[t]];
            }
        };

        for (uint32_t R = r0; R < r1; R += STEP_ROWS) {
            // invariant: segment j holds row R
            const uint64_t vl = (uint64_t)R * K / VEC + lane;              // the lane's first vector
            uint4 r[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q)
#if defined(LINT_FLAT_NOPHILOX)
                r[q] = make_uint4((uint32_t)vl * 0x9E3779B9u, q ^ olo, ohi, (uint32_t)vl + key.k0[3]);   // experiment
#else
                r[q] = philox4x32_10(make_uint4((uint32_t)vl, TAG_FLAT | ((uint32_t)(vl >> 32) << 8) | q, olo, ohi),
                                     key);
#endif
            T u[J][VEC];                                                   // the lane's uniforms of this step
            static_for<0, J * VEC>([&](auto f_tag) {
                constexpr int f = decltype(f_tag)::value;
                u[f / VEC][f % VEC] = FlatFieldOf<f, NF, T>::get(r, fm, fo);
            });
            T *p = out + vl * VEC;
            if (j + 3 - jbase >= 32u) reload(j);
            const uint32_t seg_end = __shfl_sync(FULL, w_off, (int)(j + 1 - jbase));
            const bool whole = R + STEP_ROWS <= N;                         // the step ends inside the run
            // constants of segment sg, rotated for this lane, from the box table
            auto load = [&](uint32_t sg, T (&lo)[K], T (&w)[K]) {
                const T *b = boxes + (size_t)sg * (2 * K);
#pragma unroll
                for (int t = 0; t < K; ++t) { lo[t] = __ldg(b + dsel[t]); w[t] = __ldg(b + K + dsel[t]); }
            };
            if (whole && seg_end >= R + STEP_ROWS) {
                // ---- the whole step lies in segment j
                if (held != j) {
                    load(j, rlo, rw);
                    held = j;
                }
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    T x[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const int t = (32 * VEC * jj + i) % K;
                        x[i] = affine_w(rlo[t], rw[t], u[jj][i]);
                    }
                    LINT_CHECK((vl + 32u * jj + 1) * VEC <= (uint64_t)N * K && held == j && j < nnz);
                    FT::store(p + (size_t)32 * VEC * jj, x);
                }
                if (seg_end == R + STEP_ROWS) ++j;
            } else if (whole && __shfl_sync(FULL, w_off, (int)(j + 2 - jbase)) >= R + STEP_ROWS) {
                // ---- one boundary inside the step: elements below it belong to segment j, the
                // others to segment j + 1; both boxes in registers, every element picks its own
                if (held != j) load(j, rlo, rw);
                T blo[K], bw[K];
                load(j + 1, blo, bw);
                const uint32_t eb = (seg_end - R) * K;                     // first element of segment j + 1
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    T x[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const int t = (32 * VEC * jj + i) % K;
                        const T xa = affine_w(rlo[t], rw[t], u[jj][i]), xb = affine_w(blo[t], bw[t], u[jj][i]);
                        x[i] = VEC * (lane + 32u * jj) + i >= eb ? xb : xa;
                    }
                    LINT_CHECK((vl + 32u * jj + 1) * VEC <= (uint64_t)N * K && j + 1 < nnz && eb > 0 && eb < STEP_ELEMS);
                    FT::store(p + (size_t)32 * VEC * jj, x);
                }
#pragma unroll
                for (int t = 0; t < K; ++t) { rlo[t] = blo[t]; rw[t] = bw[t]; }
                held = j + 1;
                j += __shfl_sync(FULL, w_off, (int)(j + 2 - jbase)) == R + STEP_ROWS ? 2 : 1;
            } else {
                // ---- segment boundaries (or the end of the run) inside the step: every element
                // takes the box of its own segment.  At most 31 segments are staged at a time;
                // elements past the last staged one wait for the next pass.
                const uint32_t step_end = min(R + STEP_ROWS, N);           // rows
                const uint32_t lim = (step_end - R) * K;                   // elements of the step that exist
                uint32_t done = 0;                                         // elements [0, done) are settled
                uint32_t jp = j;
                for (;;) {
                    uint32_t first;
                    const uint32_t ns = stage(jp, step_end, R, first);
                    // elements below `upto` belong to the staged segments
                    const uint32_t last_end = s_end[first + ns - 1];
                    const uint32_t upto = min(last_end, lim);
                    // slot of every element: first + the number of staged segments that end at or
                    // before it; the element's uniform is replaced by its coordinate in place
#pragma unroll
                    for (int jj = 0; jj < J; ++jj) {
                        uint32_t sidx[VEC];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) sidx[i] = first;
                        for (uint32_t b = first; b + 1 < first + ns; ++b) {
                            const uint32_t eb = s_end[b];
#pragma unroll
                            for (int i = 0; i < VEC; ++i) sidx[i] += VEC * (lane + 32u * jj) + i >= eb ? 1u : 0u;
                        }
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const int t = (32 * VEC * jj + i) % K;
                            const uint32_t e = VEC * (lane + 32u * jj) + i;
                            LINT_CHECK(!(e >= done && e < upto) || (sidx[i] >= first && sidx[i] < first + ns && sidx[i] < 31));
                            if (e >= done && e < upto)
                                u[jj][i] = affine_w(s_box[sidx[i] * SLOT + dsel[t]], s_box[sidx[i] * SLOT + K + dsel[t]],
                                                    u[jj][i]);
                        }
                    }
                    done = upto;
                    if (done >= lim) {
                        // the segment that holds row step_end: the first staged one that ends after it
                        uint32_t adv = 0;
                        for (uint32_t b = first; b < first + ns; ++b) adv += s_end[b] <= lim ? 1u : 0u;
                        j = jp + adv;
                        held = 0xffffffffu;
                        if (adv < ns) {                                    // ... and it is staged: keep its constants
                            take(first + adv);
                            held = j;
                        }
                        break;
                    }
                    jp += ns;                                              // all staged segments ended before `lim`
                }
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    const uint32_t e0 = VEC * (lane + 32u * jj);
                    LINT_CHECK(lim <= STEP_ELEMS && (uint64_t)R * K + lim <= (uint64_t)N * K && done >= lim);
                    if (e0 + VEC <= lim) {
                        FT::store(p + (size_t)32 * VEC * jj, u[jj]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i)
                            if (e0 + i < lim) store_stream(p + (size_t)32 * VEC * jj + i, u[jj][i]);
                    }
                }
            }
        }
        if (CELLS) {
            // cell index of the chunk's rows: a second walk over its segments
            uint32_t jj = __ldg(chunk_start + chunk);
            uint32_t row = r0;
            while (row < r1) {
                const uint32_t end = min(__ldg(coff + jj + 1), r1);
                const long long c = (long long)__ldg(cid + jj);
                for (uint32_t rr = row + lane; rr < end; rr += 32)
                    __stcs(reinterpret_cast<long long *>(cell_idx) + rr, c);
                row = end;
                ++jj;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// general emitter: rows [T_f, T_f + T_r), every segment drawn from an interpolant
// ---------------------------------------------------------------------------
// Segments here are short (a few dozen rows: the residual parts of the split cells and the
// cells too small to split), so the kernel never specialises on segment length: a warp step
// covers 64 consecutive rows, lane l owns the two rows 2 l and 2 l + 1 of the step, and each
// row finds its own segment from one warp-wide OR of the segment starts inside the step.
//   * One Philox block yields the in-cell uniforms of a lane's two rows (fp32, k <= 3: six
//     21-bit fields, see FlatTraits).
//   * fp32, k <= 3: the per-cell constants (box, hoisted quantile coefficients) of 64
//     consecutive segments are computed one segment per lane and staged in shared memory;
//     a row then costs a few 128-bit loads instead of a fetch + normalisation.
//   * Steps are laid over EVEN global rows (the run starts at the arbitrary row T_f), so a lane's
//     two rows are 2 k contiguous, 8- or 16-byte aligned values and leave as k vector stores.
// Every warp works on its own chunks of `chunk_rows` rows and walks the chunk's segments
// through a private shared-memory window of the segment table (the next chunk's first window
// is prefetched with cp.async).  No CTA-wide barrier.
#ifndef LINT_EMIT_MINB
#define LINT_EMIT_MINB 4       // resident CTAs per SM the fp32 k<=3 emitter is compiled for
#endif
#ifndef LINT_EMIT_THREADS
#define LINT_EMIT_THREADS 256
#endif
constexpr int EMIT_THREADS = LINT_EMIT_THREADS;
constexpr int EMIT_WIN = 96;                 // window entries (>= 66: a step looks at 65 segment starts)
constexpr uint32_t EMIT_STEP = 64;           // rows per warp step
constexpr uint32_t EMIT_PRE = 32;            // staged segments (= lanes); a step that draws from more
                                             // (segments of less than two rows on average) fetches per row
constexpr uint32_t EMIT_PRE_NONE = 0x80000000u;

template <int K, typename T, typename Cells, bool CELLS>
__global__ void __launch_bounds__(EMIT_THREADS, sizeof(T) == 4 ? (K <= 3 ? LINT_EMIT_MINB : 2) : 1)
emit_kernel(Cells src, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ coff,
            const uint32_t *__restrict__ nnz_p, const uint32_t *__restrict__ flat_rows,
            const uint32_t *__restrict__ rest_rows, uint32_t n_requested, const uint32_t *__restrict__ chunk_start0,
            const uint32_t *__restrict__ chunk_start1, uint32_t chunk_rows, const __grid_constant__ PhiloxKeys key,
            StreamPos pos, uint32_t split, CdfSlice table, double lam_per_unit, T *__restrict__ out,
            int64_t *__restrict__ cell_idx) {
    using FT = FlatTraits<T>;
    const uint64_t offset = pos.get();
    constexpr int NF = FT::FIELDS, NB = (2 * K + NF - 1) / NF;       // Philox blocks per pair of rows
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sedges = reinterpret_cast<T *>(smem_raw);
    __shared__ uint32_t s_off_all[EMIT_THREADS / 32][2][EMIT_WIN];
    __shared__ uint32_t s_cid_all[EMIT_THREADS / 32][2][EMIT_WIN];
    // staged constants: CellCoefs words, then the box (lo[K], w[K]); the slot stride is an
    // odd number of 16-byte units so the per-lane 128-bit stores do not conflict
    constexpr bool PRE = sizeof(T) == 4 && K <= 3;
    constexpr int PRE_WORDS = CellCoefs<K, T>::WORDS + 2 * K;
    constexpr int PRE_Q = ((PRE_WORDS + 3) / 4) | 1;
    __shared__ float4 s_pre_all[PRE ? EMIT_THREADS / 32 : 1][PRE ? EMIT_PRE * PRE_Q : 1];
    src.prepare(sedges);
    const uint32_t fm = key.field_mask, fo = key.field_one;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nnz = *nnz_p;
    // Rows are counted from the even global row at or before the start of the run: row r of this
    // frame is global row (T_f - shift) + r and belongs to the run when shift <= r < end.
    uint32_t t_flat, t_rest;                           // rows of the flat run before this one, rows of this run
    run_rows(flat_rows, rest_rows, n_requested, t_flat, t_rest);
    const uint32_t shift = t_flat & 1u, end = t_rest + shift;
    const uint32_t *__restrict__ chunk_start = shift ? chunk_start1 : chunk_start0;
    out += (size_t)(t_flat - shift) * K;
    if (CELLS) cell_idx += t_flat - shift;
    const uint32_t pair0 = (t_flat - shift) >> 1;      // global index of the frame's first pair of rows
    const uint32_t nchunks = t_rest ? (end + chunk_rows - 1) / chunk_rows : 0u;       // (an empty run has no chunk)
    const uint32_t nwarps = gridDim.x * (EMIT_THREADS / 32);
    const uint32_t olo = (uint32_t)offset, ohi = (uint32_t)(offset >> 32);
    constexpr uint32_t FULL = 0xffffffffu;
    // window entry i of buffer b <- segment jb + i (offsets exist for 0..nnz, cells for 0..nnz-1)
    auto prefetch = [&](uint32_t jb, int b) {
#pragma unroll
        for (uint32_t i = lane; i < EMIT_WIN; i += 32) {
            if (jb + i <= nnz) __pipeline_memcpy_async(&s_off_all[warp][b][i], coff + jb + i, 4);
            if (jb + i < nnz) __pipeline_memcpy_async(&s_cid_all[warp][b][i], cid + jb + i, 4);
        }
        __pipeline_commit();
    };
    uint32_t chunk = blockIdx.x * (EMIT_THREADS / 32) + warp;
    int buf = 0;
    uint32_t j0 = chunk < nchunks ? __ldg(chunk_start + chunk) : 0;
    uint32_t j0_next = chunk + nwarps < nchunks ? __ldg(chunk_start + chunk + nwarps) : 0;
    if (chunk < nchunks) prefetch(j0, 0);
    for (; chunk < nchunks; chunk += nwarps, buf ^= 1) {
        const uint32_t r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, end);
        const bool more = chunk + nwarps < nchunks;
        if (more) prefetch(j0_next, buf ^ 1);
        uint32_t jbase = j0;                             // segment held in window slot 0
        LINT_CHECK(j0 < nnz && coff[j0] <= max(r0, shift) - shift && max(r0, shift) - shift < coff[j0 + 1]);
        j0 = j0_next;
        if (chunk + 2 * (uint64_t)nwarps < nchunks) j0_next = __ldg(chunk_start + chunk + 2 * nwarps);
        if (more) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);
        __syncwarp();
        uint32_t *s_off = s_off_all[warp][buf], *s_cid = s_cid_all[warp][buf];
        // first row (in the frame) of segment sg (sg > nnz: +inf); callers keep sg - jbase < EMIT_WIN
        auto off = [&](uint32_t sg) { return sg <= nnz ? s_off[sg - jbase] + shift : 0xffffffffu; };
        auto refill = [&](uint32_t jb) {                 // slide the window so that slot 0 = segment jb
            __syncwarp();
#pragma unroll
            for (uint32_t i = lane; i < EMIT_WIN; i += 32) {
                if (jb + i <= nnz) s_off[i] = __ldg(coff + jb + i);
                if (jb + i < nnz) s_cid[i] = __ldg(cid + jb + i);
            }
            jbase = jb;
            __syncwarp();
        };
        uint32_t cbase = EMIT_PRE_NONE;                  // staged block = segments [cbase, cbase + EMIT_PRE)
        float4 *s_pre = s_pre_all[PRE ? warp : 0];
        // segment ids: the cell itself, or with the flat/residual split 2 cell + 1
        // constants of one segment into the caller's registers
        auto constants = [&](uint32_t v, CellCoefs<K, T> &cf, T (&lo)[K], T (&w)[K]) {
            T c0[1 << K], hi[K];
            src.fetch(sedges, v >> split, c0, lo, hi);
#pragma unroll
            for (int d = 0; d < K; ++d) w[d] = hi[d] - lo[d];              // maxs - mins, sampling.py:126
            // the residual of a split cell, or the whole of a cell that is not split
            const bool residual = split && cell_lambda(table, v >> 1, lam_per_unit) >= SPLIT_MIN_LAMBDA;
            if (residual) subtract_min(c0);
            cf.set(c0, !residual && src.prenormalised());
        };
        // stage the constants of segments [jb, jb + EMIT_PRE) that start inside this chunk;
        // the window must hold entries jb .. jb + EMIT_PRE - 1
        auto stage = [&](uint32_t jb) {
            __syncwarp();
            {
                constexpr uint32_t h = 0;
                const uint32_t js = jb + lane;
                if (js < nnz && off(js) < r1) {
                    T lo[K], wd[K];
                    CellCoefs<K, T> cf;
                    float w[4 * PRE_Q] = {};
                    constants(s_cid[js - jbase], cf, lo, wd);
                    cf.pack(w);
#pragma unroll
                    for (int d = 0; d < K; ++d) {
                        w[CellCoefs<K, T>::WORDS + d] = (float)lo[d];
                        w[CellCoefs<K, T>::WORDS + K + d] = (float)wd[d];
                    }
#pragma unroll
                    for (int q = 0; q < (PRE_WORDS + 3) / 4; ++q)
                        s_pre[(32 * h + lane) * PRE_Q + q] =
                            make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                }
            }
            cbase = jb;
            __syncwarp();
        };
        // one row of segment sg: constants (staged, or fetched), in-cell draw, affine map
        auto draw_row = [&](uint32_t sg, bool staged, const T (&u)[K], T (&x)[K], uint32_t &cell) {
            LINT_CHECK(sg >= jbase && sg - jbase < (uint32_t)EMIT_WIN && sg < nnz);
            LINT_CHECK(!(PRE && staged) || (sg >= cbase && sg - cbase < EMIT_PRE));
            const uint32_t v = s_cid[sg - jbase];
            cell = v >> split;
            T xlo[K], xw[K], z[K];
            CellCoefs<K, T> coef;
            if (PRE && staged) {
                float w[4 * PRE_Q];
#pragma unroll
                for (int q = 0; q < (PRE_WORDS + 3) / 4; ++q) {
                    const float4 t4 = s_pre[(sg - cbase) * PRE_Q + q];
                    w[4 * q] = t4.x; w[4 * q + 1] = t4.y; w[4 * q + 2] = t4.z; w[4 * q + 3] = t4.w;
                }
                coef.unpack(w);
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    xlo[d] = (T)w[CellCoefs<K, T>::WORDS + d];
                    xw[d] = (T)w[CellCoefs<K, T>::WORDS + K + d];
                }
            } else {
                constants(v, coef, xlo, xw);
            }
            coef.draw(u, z);                                               // sampling.py:41-94
#pragma unroll
            for (int d = 0; d < K; ++d) x[d] = affine_w(xlo[d], xw[d], z[d]);   // sampling.py:126
        };

        uint32_t j = jbase;        // warp-uniform: the segment that holds row R (or ends exactly there)
        for (uint32_t R = r0; R < r1; R += EMIT_STEP) {
            if (j + EMIT_STEP + 2 - jbase >= (uint32_t)EMIT_WIN) refill(j);
            // bit b of (lo_m, hi_m): a segment starts at row R + b, 0 <= b < 64 (lane i looks at
            // segments j + 1 + i and j + 33 + i; segments hold >= 1 row, so no other can)
            const uint32_t o1 = off(j + 1 + lane) - R, o2 = off(j + 33 + lane) - R;
            uint32_t lo_m = (o1 < 32u ? 1u << o1 : 0u) | (o2 < 32u ? 1u << o2 : 0u);
            uint32_t hi_m = (o1 - 32u < 32u ? 1u << (o1 - 32u) : 0u) | (o2 - 32u < 32u ? 1u << (o2 - 32u) : 0u);
            lo_m = __reduce_or_sync(FULL, lo_m);
            hi_m = __reduce_or_sync(FULL, hi_m);
            const uint32_t nstart = __popc(lo_m) + __popc(hi_m);
            // segments this step draws from: j (unless it ends at R: bit 0) .. j + nstart, at most 64
            const uint32_t first = j + (lo_m & 1u);
            const bool staged = PRE && j + nstart - first < EMIT_PRE;
            if (staged && (first - cbase >= EMIT_PRE || j + nstart - cbase >= EMIT_PRE)) stage(first);
            // the lane's rows: R + 2 lane and the next one; a row's segment is j + the number of
            // starts in (R, row]
            const uint32_t x0 = 2u * lane, ra = R + x0;
            uint32_t sa;
            if (x0 < 32u) sa = j + __popc(lo_m & ((2u << x0) - 1u));
            else sa = j + __popc(lo_m) + __popc(hi_m & ((2u << (x0 - 32u)) - 1u));
            const uint32_t bit_b = x0 + 1u < 32u ? (lo_m >> (x0 + 1u)) & 1u : (hi_m >> (x0 - 31u)) & 1u;
            const uint32_t sb = sa + bit_b;
            const bool va = ra >= shift && ra < end, vb = ra + 1u >= shift && ra + 1u < end;
            if (va || vb) {
                const uint32_t pair = pair0 + (ra >> 1);
                uint4 r[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    r[q] = philox4x32_10(make_uint4(pair, TAG_COORD | q, olo, ohi), key);
                T ua[K], ub[K], xa[K], xb[K];
                static_for<0, K>([&](auto d_tag) {
                    constexpr int d = decltype(d_tag)::value;
                    ua[d] = FlatFieldOf<d, NF, T>::get(r, fm, fo);
                    ub[d] = FlatFieldOf<K + d, NF, T>::get(r, fm, fo);
                });
                uint32_t ca = 0, cb = 0;
                draw_row(va ? sa : sb, staged, ua, xa, ca);
                draw_row(vb ? sb : sa, staged, ub, xb, cb);
                T *p = out + (size_t)ra * K;
                LINT_CHECK(sa >= j && sb <= j + nstart && (!va || (off(sa) <= ra && ra < off(sa + 1))) &&
                           (!vb || (off(sb) <= ra + 1 && ra + 1 < off(sb + 1))));
                if (va && vb) {
                    T e[2 * K];
#pragma unroll
                    for (int d = 0; d < K; ++d) { e[d] = xa[d]; e[K + d] = xb[d]; }
#pragma unroll
                    for (int m = 0; m < K; ++m) store_stream2(p + 2 * m, e[2 * m], e[2 * m + 1]);
                } else {
#pragma unroll
                    for (int d = 0; d < K; ++d) {
                        if (va) store_stream(p + d, xa[d]);
                        if (vb) store_stream(p + K + d, xb[d]);
                    }
                }
                if (CELLS) {
                    if (va) __stcs(reinterpret_cast<long long *>(cell_idx) + ra, (long long)ca);
                    if (vb) __stcs(reinterpret_cast<long long *>(cell_idx) + ra + 1, (long long)cb);
                }
            }
            j += nstart;                                 // the segment that holds row R + 64
        }
        __syncwarp();
    }
}

// tuning overrides read from the environment at every call (profiling sweeps only)
inline int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
static void scan_list(const uint32_t *counts, uint32_t n, const uint32_t *range, uint32_t id_mul, uint32_t id_add,
                      const SegList &l, uint32_t chunk_rows, uint32_t chunk_slots, bool odd_frame_too, bool chained,
                      cudaStream_t st) {
    const uint32_t nscan = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *chunk1 = odd_frame_too ? l.chunk_start1 : nullptr;
    if (chained)                                               // behind the Poisson kernel on its own stream
        launch_chained(scan_compact_kernel, nscan, SCAN_THREADS, 0, st, counts, n, range, id_mul, id_add, l.tile_state,
                       l.tile_counter, l.cid, l.coff, l.nnz, l.rows, chunk_rows, chunk_slots, l.chunk_start, chunk1);
    else
        scan_compact_kernel<<<nscan, SCAN_THREADS, 0, st>>>(counts, n, range, id_mul, id_add, l.tile_state,
                                                           l.tile_counter, l.cid, l.coff, l.nnz, l.rows, chunk_rows,
                                                           chunk_slots, l.chunk_start, chunk1);
}

// counts -> offsets, compaction and chunk starts, for both runs; the row budget (plan[]) for the
// top-up on its lane.  Lanes: the flat scan runs on side[0] beside the other scan on the caller's
// stream; ev[1] = flat scan done, ev[4] = other scan done.
template <typename Cells>
static int plan_rows(const Cells &src, const double *cdf, uint64_t ncells, uint64_t N, uint2 key,
                     uint64_t offset, Stratum u, const uint32_t *range, const Scratch &s, cudaStream_t st) {
    const bool split = use_split(ncells, N, u.u_hi - u.u_lo);
    const CdfSlice t{cdf, (uint32_t)ncells, u.u_lo, u.u_hi};
    const double lam_per_unit = grouped_mean(N) / (u.u_hi - u.u_lo);
    const int cblocks = (int)std::min<uint64_t>((ncells + POISSON_THREADS - 1) / POISSON_THREADS, 148 * 32);
    const uint32_t chunk_rows = emit_chunk_rows(N);
    launch_chained(poisson_counts_kernel<Cells>, cblocks, POISSON_THREADS, 0, st, src, t, lam_per_unit, key,
                   StreamPos{offset, offset_counter()}, split, range, s.counts_flat, s.counts_rest,
                   (uint4 *)s.zero_begin, (uint32_t)(s.zero_bytes / 16));
    Lanes *L = side_lanes();
    if (L && split) {                                          // the two scans side by side
        cudaEventRecord(L->ev[0], st);
        cudaStreamWaitEvent(L->side[0], L->ev[0], 0);
        scan_list(s.counts_flat, (uint32_t)ncells, range, 1, 0, s.flat, chunk_rows, s.chunk_slots, false, false, L->side[0]);
        cudaEventRecord(L->ev[1], L->side[0]);
    } else if (split) {
        scan_list(s.counts_flat, (uint32_t)ncells, range, 1, 0, s.flat, chunk_rows, s.chunk_slots, false, true, st);
    }
    scan_list(s.counts_rest, (uint32_t)ncells, range, split ? 2 : 1, split ? 1 : 0, s.rest, chunk_rows, s.chunk_slots, split, true, st);
    // the row budget, for the top-up (rows [T, N)) on its lane: it needs both scans, the emitters
    // need neither it nor each other's scan (emit_flat: the flat scan; emit_kernel: both)
    cudaStream_t st_plan = st;
    if (L) {
        cudaEventRecord(L->ev[4], st);
        cudaStreamWaitEvent(L->side[1], L->ev[4], 0);
        if (split) cudaStreamWaitEvent(L->side[1], L->ev[1], 0);
        st_plan = L->side[1];
    }
    plan_totals_kernel<<<1, 1, 0, st_plan>>>(split ? s.flat.rows : nullptr, s.rest.rows, (uint32_t)N, s.plan);
    if (L && split) cudaStreamWaitEvent(st, L->ev[1], 0);      // the general emitter reads T_f
    return check_launch("cell-major row planning");
}

// The flat emitter runs on a side lane beside the general one: the first fills the store path,
// the second the issue slots (profiles/README.md r2g).
template <int K, typename T, typename Cells>
int launch_emit(const Cells &src, const double *cdf, const Scratch &s, uint64_t ncells, uint64_t N, uint2 key,
                uint64_t offset, Stratum u, void *out, int64_t *cell_idx, cudaStream_t st) {
    const bool split = use_split(ncells, N, u.u_hi - u.u_lo);
    const CdfSlice t{cdf, (uint32_t)ncells, u.u_lo, u.u_hi};
    const double lam_per_unit = grouped_mean(N) / (u.u_hi - u.u_lo);
    const uint32_t chunk_rows = emit_chunk_rows(N);
    const uint32_t nchunks = (uint32_t)((N + 1 + chunk_rows - 1) / chunk_rows);     // (+1: the general emitter's frame)
    const size_t smem = src.smem_bytes(sizeof(T));
    EmitTimers &tm = emit_timers();
    if (tm.enabled) {
        tm.flat_ran = split;
        tm.plan_dev = s.plan;
        cudaEventRecord(tm.ev[0], st);
    }
    // The flat emitter beside the general one.  It needs the flat scan only, but it is held back
    // until the general emitter is about to be launched (and then builds its box table first): both
    // kernels are sized to fill the machine, and the step is shortest when the general one, bound by
    // instruction issue, takes its share of every SM first and the flat one, bound by the store
    // path, fills what is left — starting them together or the other way round costs 4 % of the step.
    Lanes *L = split ? side_lanes() : nullptr;
    cudaStream_t st_main = st;
    if (L) {
        cudaEventRecord(L->ev[2], st);
        cudaStreamWaitEvent(L->side[0], L->ev[2], 0);
        st = L->side[0];
    }
    if (split) {
        const uint32_t fper = FLAT_THREADS / 32;
        const int fblocks = (int)std::min<uint32_t>((nchunks + fper - 1) / fper,
                                                     148u * env_int("LINT_FLAT_BPS", LINT_FLAT_MINB));
        T *boxes = (T *)s.boxes;
        flat_boxes_kernel<K, T, Cells><<<(int)std::min<uint64_t>((ncells + 255) / 256, 148 * 8), 256, 0, st>>>(
            src, s.flat.cid, s.flat.nnz, boxes);
        if (cell_idx)
            emit_flat_kernel<K, T, true><<<fblocks, FLAT_THREADS, 0, st>>>(
                boxes, s.flat.cid, s.flat.coff, s.flat.nnz, s.flat.rows, (uint32_t)N, s.flat.chunk_start, chunk_rows, s.chunk_counter,
                philox_round_keys(key), StreamPos{offset, offset_counter()}, (T *)out, cell_idx);
        else
            emit_flat_kernel<K, T, false><<<fblocks, FLAT_THREADS, 0, st>>>(
                boxes, s.flat.cid, s.flat.coff, s.flat.nnz, s.flat.rows, (uint32_t)N, s.flat.chunk_start, chunk_rows, s.chunk_counter,
                philox_round_keys(key), StreamPos{offset, offset_counter()}, (T *)out, cell_idx);
        if (int rc = check_launch("cell-major flat emission")) return rc;
    }
    if (L) {
        cudaEventRecord(L->ev[3], st);
        st = st_main;
    }
    if (tm.enabled) {
        cudaEventRecord(tm.ev[1], st);
        cudaEventRecord(tm.ev[2], st);
    }
    const uint32_t per_cta = EMIT_THREADS / 32;
    const int blocks = (int)std::min<uint32_t>((nchunks + per_cta - 1) / per_cta,
                                                148u * env_int("LINT_REST_BPS", 8));
    if (cell_idx)
        emit_kernel<K, T, Cells, true><<<blocks, EMIT_THREADS, smem, st>>>(
            src, s.rest.cid, s.rest.coff, s.rest.nnz, split ? s.flat.rows : nullptr, s.rest.rows, (uint32_t)N,
            s.rest.chunk_start, s.rest.chunk_start1, chunk_rows,
            philox_round_keys(key), StreamPos{offset, offset_counter()}, split ? 1u : 0u, t, lam_per_unit, (T *)out, cell_idx);
    else
        emit_kernel<K, T, Cells, false><<<blocks, EMIT_THREADS, smem, st>>>(
            src, s.rest.cid, s.rest.coff, s.rest.nnz, split ? s.flat.rows : nullptr, s.rest.rows, (uint32_t)N,
            s.rest.chunk_start, s.rest.chunk_start1, chunk_rows,
            philox_round_keys(key), StreamPos{offset, offset_counter()}, split ? 1u : 0u, t, lam_per_unit, (T *)out, cell_idx);
    if (tm.enabled) cudaEventRecord(tm.ev[3], st);
    if (L) cudaStreamWaitEvent(st, L->ev[3], 0);
    return check_launch("cell-major emission");
}

template <int K>
int grid_cellmajor_k(const lint_grid_t *g, uint64_t nc, uint64_t N, uint2 key, uint64_t offset, Stratum u,
                     const uint32_t *range, const Scratch &s, void *out, int64_t *cell_idx, cudaStream_t st) {
    if (g->dtype == LINT_F32) {
        GridCells<K, float> src{make_grid_view<K>(g, nc, true)};
        if (int rc = plan_rows(src, g->cdf_dev, nc, N, key, offset, u, range, s, st)) return rc;
        return launch_emit<K, float>(src, g->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
    }
    GridCells<K, double> src{make_grid_view<K>(g, nc, true)};
    if (int rc = plan_rows(src, g->cdf_dev, nc, N, key, offset, u, range, s, st)) return rc;
    return launch_emit<K, double>(src, g->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
}

// defined in lint_sample.cu: u-driven rows [*first_row_dev, N), at most max_rows of them
int topup_grid(const lint_grid_t *g, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
               uint64_t seed, uint64_t offset, double u_lo, double u_hi, const uint32_t *cell_range_dev, void *out,
               int64_t *cell_idx, cudaStream_t st);
int topup_cells(const lint_cells_t *c, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
                uint64_t seed, uint64_t offset, double u_lo, double u_hi, void *out, int64_t *cell_idx,
                cudaStream_t st);

template <int K>
int cells_cellmajor_k(const lint_cells_t *c, uint64_t N, uint2 key, uint64_t offset, Stratum u,
                      const Scratch &s, void *out, int64_t *cell_idx, cudaStream_t st) {
    CellsView v;
    v.mins = c->mins_dev;
    v.widths = c->widths_dev;
    v.corners = c->corners_dev;
    v.rec = c->cell_records_dev;
    v.table.cdf = c->cdf_dev;
    v.table.guide = nullptr;
    v.table.guide_bits = 0;
    v.table.n = (uint32_t)c->ncells;
    const uint64_t nc = (uint64_t)c->ncells;
    if (c->dtype == LINT_F32) {
        ListCells<K, float> src{v};
        if (int rc = plan_rows(src, c->cdf_dev, nc, N, key, offset, u, nullptr, s, st)) return rc;
        return launch_emit<K, float>(src, c->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
    }
    ListCells<K, double> src{v};
    if (int rc = plan_rows(src, c->cdf_dev, nc, N, key, offset, u, nullptr, s, st)) return rc;
    return launch_emit<K, double>(src, c->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
}

}  // namespace lint
