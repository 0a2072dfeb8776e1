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
// Row ORDER is by cell (C order; a split cell's flat rows, then its residual rows)
// for the first T rows, draw order for the top-up tail.  Callers that need the reference's exchangeable row order
// (lintsampler.py:309-311) apply the row permutation afterwards.
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <cuda_pipeline.h>
#include "lint_device.cuh"

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
__constant__ double LOG_FACTORIAL_TAB[16] = {
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
// transformed rejection with squeeze (PTRS, 1993) above.
__device__ __forceinline__ uint32_t poisson(double lam, CellRng &rng) {
    if (!(lam > 0.0)) return 0;
    if (lam < 10.0) {
        double u, unused;
        rng.next(u, unused);
        // exp(-lam) >= 1 - lam: most cells of a fine grid hold lam << 1 and end here,
        // without the exponential
        if (u <= 1.0 - lam) return 0;
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
            rng.next(u, unused);
        }
    }
    const double b = 0.931 + 2.53 * sqrt(lam);
    const double a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        double u, v;
        rng.next(u, v);
        u -= 0.5;
        const double us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= vr) return (uint32_t)k;            // the squeeze accepts ~86 %
        if (k < 0.0 || (us < 0.013 && v > us)) continue;
        const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
        const double lhs_arg = v * invalpha / (a / (us * us) + b);
        // Acceptance test log(lhs_arg) <= -lam + k log(lam) - log(k!).  ~14 % of the draws get
        // here, i.e. nearly every warp, so it is first decided in fp32 wherever the margin
        // exceeds a generous bound on the fp32 error; the fp64 test below settles the rest
        // (a few per thousand).  With k1 = k + 1, d = k1 - lam, x = d / k1 and Stirling's
        // series the right-hand side is  k log1p(-x) + d - log(k1)/2 - log(2 pi)/2 - s(k1).
        if (k >= 4.0 && lam < 1.0e5) {
            const float d = (float)(k + 1.0 - lam), k1 = (float)k + 1.0f, x = d / k1;
            const float series = (1.0f / 12.0f - (1.0f / 360.0f) / (k1 * k1)) / k1;
            const float rhs = ((float)k * log1pf(-x) + d) - 0.5f * __logf(k1) - 0.9189385f - series;
            const float t = __logf((float)lhs_arg) - rhs;
            const float tol = 2e-3f + 4e-6f * fabsf(d);
            if (t < -tol) return (uint32_t)k;
            if (t > tol) continue;
        }
        if (log(lhs_arg) <= -lam + k * log(lam) - log_factorial(k)) return (uint32_t)k;
    }
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

// A cell is split into its flat and residual parts only where that pays: below this
// many expected rows the residual part would be a handful of rows with a segment of
// their own, dearer than drawing all of the cell's rows from the full interpolant.
#ifndef LINT_SPLIT_MIN_LAMBDA
#define LINT_SPLIT_MIN_LAMBDA 128.0
#endif
constexpr double SPLIT_MIN_LAMBDA = LINT_SPLIT_MIN_LAMBDA;

// counts[c] (no split), or counts[2c] = flat rows, counts[2c + 1] = residual rows — or all
// rows of a cell that is not split
template <typename Cells>
__global__ void __launch_bounds__(256)
poisson_counts_kernel(Cells src, CdfSlice t, double lam_per_unit, uint2 key, uint64_t offset, bool split,
                      uint32_t *counts) {
    const uint32_t w2 = (uint32_t)offset, w3 = (uint32_t)(offset >> 32);
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < t.ncells; c += gridDim.x * blockDim.x) {
        const double lam = cell_lambda(t, c, lam_per_unit);
        if (!split) {
            CellRng rng{key, c, w2, w3};
            counts[c] = poisson(lam, rng);
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
        reinterpret_cast<uint2 *>(counts)[c] = make_uint2(n[0], n[1]);
    }
}

// ---------------------------------------------------------------------------
// decoupled look-back scan of (count, non-empty) pairs + compaction
// ---------------------------------------------------------------------------
// A tile publishes its aggregate, then its inclusive prefix, in one 64-bit word:
// bits 63-62 status, bits 61-32 non-empty cells, bits 31-0 rows.  Integer sums
// are associative, so the look-back's variable association is harmless here.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long pack(uint32_t rows, uint32_t cells) {
    return ((unsigned long long)cells << 32) | rows;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_compact_kernel(const uint32_t *counts, uint32_t ncells, unsigned long long *tile_state,
                    uint32_t *tile_counter, uint32_t *cid, uint32_t *coff, uint32_t *nnz_out,
                    uint32_t *rows_out) {
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
    unsigned long long pre = s_excl + wex + (inc - run);           // exclusive prefix of this thread
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (c[i] != 0) {
            const uint32_t slot = (uint32_t)(pre >> 32);
            cid[slot] = base + i;
            coff[slot] = (uint32_t)pre;
        }
        pre += pack(c[i], c[i] != 0);
    }
    if (base <= ncells - 1 && ncells - 1 < base + SCAN_ITEMS) {    // owner of the last cell
        const uint32_t nnz = (uint32_t)(pre >> 32);
        *nnz_out = nnz;
        *rows_out = (uint32_t)pre;                                 // T = total grouped rows
        coff[nnz] = (uint32_t)pre;                                 // sentinel
    }
}

// first segment of every output chunk: last j with coff[j] <= chunk * chunk_rows
__global__ void chunk_start_kernel(const uint32_t *coff, const uint32_t *nnz_p, const uint32_t *rows_p,
                                  uint32_t chunk_rows, uint32_t *chunk_start) {
    const uint32_t nnz = *nnz_p;
    const uint32_t nchunks = (*rows_p + chunk_rows - 1) / chunk_rows;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nchunks; t += gridDim.x * blockDim.x) {
        const uint32_t r0 = t * chunk_rows;
        uint32_t lo = 0, hi = nnz;                                 // coff[nnz] = N > r0
        while (lo < hi) {                                          // first j with coff[j] > r0
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(coff + mid) > r0) hi = mid; else lo = mid + 1;
        }
        chunk_start[t] = lo - 1;
    }
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
};

#ifndef LINT_EMIT_U
#define LINT_EMIT_U 4          // rows per lane in the straight-line fast path
#endif
#ifndef LINT_EMIT_MINB
#define LINT_EMIT_MINB 4       // resident CTAs per SM the fp32 k<=3 emitter is compiled for
#endif
#ifndef LINT_EMIT_THREADS
#define LINT_EMIT_THREADS 256
#endif
constexpr int EMIT_THREADS = LINT_EMIT_THREADS, EMIT_U = LINT_EMIT_U;
// Rows per warp work item ("chunk"): 4096 amortises the per-chunk set-up best (4.64 against
// 4.73 ms per 1e9 rows with 2048), but a draw needs enough chunks to balance ~9500 resident warps.
constexpr uint32_t EMIT_CHUNK_MIN = 2048, EMIT_CHUNK_MAX = 4096;
static uint32_t emit_chunk_rows(uint64_t N) { return N >= (1ull << 28) ? EMIT_CHUNK_MAX : EMIT_CHUNK_MIN; }

// In-cell uniforms of a group of R rows of one lane: R*K coordinates from
// ceil(R*K*words/4) Philox blocks.  fp32 rows use one 32-bit word per coordinate
// (top 24 bits); fp64 rows two words (53 bits).  The counter names the group's
// first row, `sub` separates the multi-row groups (0) from single rows (1).
template <int K, typename T, int R>
struct GroupUniforms {
    static constexpr int WORDS = K * R * (sizeof(T) == 8 ? 2 : 1);
    static constexpr int CALLS = (WORDS + 3) / 4;
    uint32_t w[4 * CALLS];
    __device__ __forceinline__ void fill(uint2 key, uint32_t first_row, uint32_t sub, uint64_t offset) {
#pragma unroll
        for (int q = 0; q < CALLS; ++q) {
            const uint4 r = philox4x32_10(
                make_uint4(first_row, TAG_COORD | (sub << 16) | q, (uint32_t)offset, (uint32_t)(offset >> 32)),
                key);
            w[4 * q] = r.x; w[4 * q + 1] = r.y; w[4 * q + 2] = r.z; w[4 * q + 3] = r.w;
        }
    }
    __device__ __forceinline__ void get(int i, float (&u)[K]) const {
#pragma unroll
        // round-toward-zero conversion keeps the top 24 bits: u in [0, 1 - 2^-24]
        for (int d = 0; d < K; ++d) u[d] = __uint2float_rz(w[i * K + d]) * (1.0f / 4294967296.0f);
    }
    __device__ __forceinline__ void get(int i, double (&u)[K]) const {
#pragma unroll
        for (int d = 0; d < K; ++d) u[d] = u53(w[2 * (i * K + d)], w[2 * (i * K + d) + 1]);
    }
    // RAW: u / raw_scale — fp32 callers that only multiply u by a per-cell constant fold
    // the 2^-32 into that constant
    template <bool RAW>
    __device__ __forceinline__ void get_as(int i, float (&u)[K]) const {
        if constexpr (RAW) {
#pragma unroll
            for (int d = 0; d < K; ++d) u[d] = __uint2float_rz(w[i * K + d]);
        } else {
            get(i, u);
        }
    }
    template <bool RAW>
    __device__ __forceinline__ void get_as(int i, double (&u)[K]) const { get(i, u); }
};
template <typename T> __device__ __forceinline__ T raw_scale();
template <> __device__ __forceinline__ float raw_scale<float>() { return 1.0f / 4294967296.0f; }
template <> __device__ __forceinline__ double raw_scale<double>() { return 1.0; }

// FLAT: a row of the constant part of the density, z = u (the caller passes the raw
// uniforms and a width that carries raw_scale)
template <int K, typename T, bool CELLS, bool FLAT>
__device__ __forceinline__ void emit_row(const CellCoefs<K, T> &coef, const T (&xlo)[K], const T (&xw)[K],
                                         const T (&u)[K], uint32_t row, uint32_t cell,
                                         T *__restrict__ out, int64_t *__restrict__ cell_idx) {
    T z[K];
    if constexpr (FLAT) {
#pragma unroll
        for (int d = 0; d < K; ++d) z[d] = u[d];
    } else {
        coef.draw(u, z);                                           // sampling.py:41-94
    }
#pragma unroll
    for (int d = 0; d < K; ++d)
        store_stream(out + (size_t)row * K + d, affine_w(xlo[d], xw[d], z[d]));   // sampling.py:126
    if (CELLS) __stcs(reinterpret_cast<long long *>(cell_idx) + row, (long long)cell);
}

// Every warp works on its own chunks of `chunk_rows` consecutive rows and walks the
// segments (a non-empty cell, or the non-empty flat / residual part of a split cell)
// that intersect the chunk in order, through a small private shared-memory window of
// the segment table (EMIT_WIN entries, refilled as the walk advances; the next chunk's
// first window is prefetched with cp.async while the current chunk is computed).  No
// CTA-wide barrier.  Segments tile the rows, so the walk never searches.  Lanes map to
// consecutive rows:
//   * blocks of 32*EMIT_U rows inside one segment: the segment's constants are in
//     registers and every lane draws EMIT_U rows from straight-line code (a last block
//     of more than half that size runs in the same form with the rows past the end
//     masked);
//   * the rest of the segment 32 rows at a time, one row per lane;
//   * where the next 32 rows cover two or more whole segments: one row per lane, each
//     lane with the constants of its own segment (found from one warp-wide OR of the
//     segment starts), so the cost stays bounded however small the cells are.
// fp32, k <= 3: the per-cell constants (box, hoisted quantile coefficients) of 32
// consecutive segments are computed at once, one segment per lane, and staged in
// shared memory; entering a cell then costs a few broadcast loads instead of a
// warp-redundant fetch + set.
constexpr int EMIT_WIN = 64;
constexpr uint32_t EMIT_PRE = 32;            // segments per staged block (= lanes)
constexpr uint32_t EMIT_PRE_NONE = 0x80000000u;

template <int K, typename T, typename Cells, bool CELLS>
__global__ void __launch_bounds__(EMIT_THREADS, sizeof(T) == 4 ? (K <= 3 ? LINT_EMIT_MINB : 2) : 1)
emit_kernel(Cells src, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ coff,
            const uint32_t *__restrict__ nnz_p, const uint32_t *__restrict__ rows_p,
            const uint32_t *__restrict__ chunk_start, uint32_t n_requested, uint32_t chunk_rows, uint2 key,
            uint64_t offset, uint32_t split, CdfSlice table, double lam_per_unit, T *__restrict__ out,
            int64_t *__restrict__ cell_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sedges = reinterpret_cast<T *>(smem_raw);
    __shared__ uint32_t s_off_all[EMIT_THREADS / 32][2][EMIT_WIN];
    __shared__ uint32_t s_cid_all[EMIT_THREADS / 32][2][EMIT_WIN];
    // staged constants: CellCoefs words, then the box (lo[K], hi[K]); the slot stride is an
    // odd number of 16-byte units so the per-lane 128-bit stores do not conflict
    constexpr bool PRE = sizeof(T) == 4 && K <= 3;
    constexpr int PRE_WORDS = CellCoefs<K, T>::WORDS + 2 * K;
    constexpr int PRE_Q = ((PRE_WORDS + 3) / 4) | 1;
    __shared__ float4 s_pre_all[PRE ? EMIT_THREADS / 32 : 1][PRE ? EMIT_PRE * PRE_Q : 1];
    src.prepare(sedges);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nnz = *nnz_p;
    const uint32_t N = min(*rows_p, n_requested);      // grouped rows T (T <= N but for an 8-sigma event)
    const uint32_t nchunks = (N + chunk_rows - 1) / chunk_rows;
    const uint32_t nwarps = gridDim.x * (EMIT_THREADS / 32);
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
        const uint32_t r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, N);
        const bool more = chunk + nwarps < nchunks;
        if (more) prefetch(j0_next, buf ^ 1);
        uint32_t jbase = j0;                             // segment held in window slot 0
        j0 = j0_next;
        if (chunk + 2 * (uint64_t)nwarps < nchunks) j0_next = __ldg(chunk_start + chunk + 2 * nwarps);
        if (more) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);
        __syncwarp();
        uint32_t *s_off = s_off_all[warp][buf], *s_cid = s_cid_all[warp][buf];
        // offset of segment j (j > nnz: +inf); callers keep j - jbase < EMIT_WIN
        auto off = [&](uint32_t j) { return j <= nnz ? s_off[j - jbase] : 0xffffffffu; };
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

        uint32_t j = jbase;                              // warp-uniform segment cursor
        uint32_t row = r0;
        uint32_t loaded = 0xffffffffu;                   // segment whose constants the registers hold
        T xlo[K], xw[K];                                 // box: lower corner, width (x raw_scale if flat)
        CellCoefs<K, T> coef;
        uint32_t cell = 0;
        bool flat = false;                               // the segment is the flat part of its cell
        uint32_t cbase = EMIT_PRE_NONE;                  // staged block = segments [cbase, cbase + 32)
        float4 *s_pre = s_pre_all[PRE ? warp : 0];
        // segment ids: the cell itself, or with the flat/residual split 2 cell + (residual ? 1 : 0)
        auto is_flat = [&](uint32_t v) { return split && !(v & 1); };
        // constants of one segment into the caller's registers
        auto constants = [&](uint32_t v, CellCoefs<K, T> &cf, T (&lo)[K], T (&w)[K]) {
            T c0[1 << K], hi[K];
            src.fetch(sedges, v >> split, c0, lo, hi);
#pragma unroll
            for (int d = 0; d < K; ++d) w[d] = hi[d] - lo[d];              // maxs - mins, sampling.py:126
            if (!is_flat(v)) {
                // the residual of a split cell, or the whole of a cell that is not split
                const bool residual = split && cell_lambda(table, v >> 1, lam_per_unit) >= SPLIT_MIN_LAMBDA;
                if (residual) subtract_min(c0);
                cf.set(c0, !residual && src.prenormalised());
            } else {
#pragma unroll
                for (int d = 0; d < K; ++d) w[d] *= raw_scale<T>();
            }
        };
        // stage the constants of segments [jb, jb + 32) that start inside this chunk;
        // the window must hold entries jb .. jb + 32
        auto stage = [&](uint32_t jb) {
            __syncwarp();
            const uint32_t js = jb + lane;
            if (js < nnz && s_off[js - jbase] < r1) {
                T lo[K], wd[K];
                CellCoefs<K, T> cf;
                float w[4 * PRE_Q] = {};
                constants(s_cid[js - jbase], cf, lo, wd);
                if (!is_flat(s_cid[js - jbase])) cf.pack(w);
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    w[CellCoefs<K, T>::WORDS + d] = (float)lo[d];
                    w[CellCoefs<K, T>::WORDS + K + d] = (float)wd[d];
                }
#pragma unroll
                for (int q = 0; q < (PRE_WORDS + 3) / 4; ++q)
                    s_pre[lane * PRE_Q + q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
            cbase = jb;
            __syncwarp();
        };
        // registers <- constants of segment seg (in the window; PRE: also in the staged block)
        auto enter = [&](uint32_t seg) {
            const uint32_t v = s_cid[seg - jbase];
            cell = v >> split;
            flat = is_flat(v);
            if constexpr (PRE) {
                float w[4 * PRE_Q];
#pragma unroll
                for (int q = 0; q < (PRE_WORDS + 3) / 4; ++q) {
                    const float4 t4 = s_pre[(seg - cbase) * PRE_Q + q];
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
        };

        // invariant: segment j holds `row` (segments are non-empty, so they tile the rows)
        while (row < r1) {
            if (j + 2 - jbase >= EMIT_WIN) refill(j);
            const uint32_t seg_end = min(off(j + 1), r1);
            // the per-lane path pays where the next 32 rows cover at least two whole segments
            if (off(j + 2) > min(row + 32u, r1)) {
                // ---- rows [row, seg_end) of one segment, lanes = consecutive rows
                if (loaded != j) {
                    if constexpr (PRE) {
                        if (j - cbase >= EMIT_PRE) {
                            if (j + EMIT_PRE + 1 - jbase >= EMIT_WIN) refill(j);
                            stage(j);
                        }
                    }
                    enter(j);
                    loaded = j;
                }
                auto rows = [&](auto tag) {
                    constexpr bool FLAT = decltype(tag)::value;
                    // blocks of 32 * EMIT_U rows
                    while (seg_end - row >= 32u * EMIT_U) {
                        const uint32_t first = row + lane;
                        GroupUniforms<K, T, EMIT_U> rnd;
                        rnd.fill(key, first, 0, offset);
#pragma unroll
                        for (int t = 0; t < EMIT_U; ++t) {
                            T u[K];
                            rnd.template get_as<FLAT>(t, u);
                            emit_row<K, T, CELLS, FLAT>(coef, xlo, xw, u, first + 32u * t, cell, out, cell_idx);
                        }
                        row += 32u * EMIT_U;
                    }
                    // a last partial block of more than half that is still cheaper in this form
                    // (rows past the end masked) than one warp-row at a time
                    if (seg_end - row > 16u * EMIT_U) {
                        const uint32_t first = row + lane;
                        GroupUniforms<K, T, EMIT_U> rnd;
                        rnd.fill(key, first, 0, offset);
#pragma unroll
                        for (int t = 0; t < EMIT_U; ++t) {
                            T u[K];
                            rnd.template get_as<FLAT>(t, u);
                            if (first + 32u * t < seg_end)
                                emit_row<K, T, CELLS, FLAT>(coef, xlo, xw, u, first + 32u * t, cell, out, cell_idx);
                        }
                        row = seg_end;
                    }
                    // the rest of the segment, however short: its constants are in registers
                    // already, which makes idle lanes cheaper than the per-lane path
                    while (row < seg_end) {
                        const uint32_t r = row + lane;
                        if (r < seg_end) {
                            GroupUniforms<K, T, 1> rnd;
                            rnd.fill(key, r, 1, offset);
                            T u[K];
                            rnd.template get_as<FLAT>(0, u);
                            emit_row<K, T, CELLS, FLAT>(coef, xlo, xw, u, r, cell, out, cell_idx);
                        }
                        row = min(row + 32u, seg_end);
                    }
                };
                if (flat) rows(std::true_type{}); else rows(std::false_type{});
                ++j;                                     // row == seg_end: the next segment, if row < r1
            } else {
                // ---- the next 32 rows span several segments: one row per lane, own segment each.
                // Segments hold >= 1 row, so lane l's segment is within [j, j + l].
                if (j + 33 - jbase >= EMIT_WIN) refill(j);
                const uint32_t r = row + lane;
                const uint32_t lim = min(row + 32u, r1);
                // bit b of `starts`: a segment begins at row + b (lane i looks at segment j + 1 + i);
                // a row's segment is j + the number of starts in (row, r]
                const uint32_t o = off(j + 1 + lane) - row;                   // >= 1
                const uint32_t starts = __reduce_or_sync(0xffffffffu, o < 32u ? 1u << o : 0u);
                const uint32_t lo = j + __popc(starts & ((2u << lane) - 1u));
                const uint32_t last = __shfl_sync(0xffffffffu, lo, (int)(lim - row) - 1);   // the last row's segment
                if constexpr (PRE) {
                    if (j - cbase >= EMIT_PRE || last - cbase >= EMIT_PRE) stage(j);       // last <= j + 31
                }
                if (r < lim) {
                    enter(lo);
                    GroupUniforms<K, T, 1> rnd;
                    rnd.fill(key, r, 1, offset);
                    T u[K];
                    if (flat) {
                        rnd.template get_as<true>(0, u);
                        emit_row<K, T, CELLS, true>(coef, xlo, xw, u, r, cell, out, cell_idx);
                    } else {
                        rnd.template get_as<false>(0, u);
                        emit_row<K, T, CELLS, false>(coef, xlo, xw, u, r, cell, out, cell_idx);
                    }
                }
                loaded = 0xffffffffu;
                row = lim;
                j = off(last + 1) <= row ? last + 1 : last;     // (window holds up to j + 32)
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
struct Scratch {
    uint32_t *counts, *cid, *coff, *chunk_start, *nnz, *rows, *tile_counter;
    unsigned long long *tile_state;
    size_t bytes;
};

// `ncells` real cells = 2 ncells virtual cells (flat, residual) in the count / segment tables
static Scratch carve(void *base, uint64_t ncells_real, uint64_t N) {
    const uint64_t ncells = 2 * ncells_real;
    const uint64_t nscan = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    const uint64_t nchunks = (N + EMIT_CHUNK_MIN - 1) / EMIT_CHUNK_MIN;      // the most any draw uses
    uintptr_t p = (uintptr_t)base;
    auto take = [&](size_t n) { uintptr_t q = p; p += (n + 255) & ~(size_t)255; return (void *)q; };
    Scratch s;
    s.counts = (uint32_t *)take(ncells * 4);
    s.cid = (uint32_t *)take((ncells + 1) * 4);
    s.coff = (uint32_t *)take((ncells + 1) * 4);
    s.chunk_start = (uint32_t *)take(std::max<uint64_t>(nchunks, 1) * 4);
    s.tile_state = (unsigned long long *)take(nscan * 8);
    s.nnz = (uint32_t *)take(4);
    s.rows = (uint32_t *)take(4);
    s.tile_counter = (uint32_t *)take(4);
    s.bytes = (size_t)(p - (uintptr_t)base);
    return s;
}

// Expected number of grouped rows and the bound on the top-up tail: the Poisson
// total T has mean L = N - 8 sqrt(N) and standard deviation sqrt(L) < sqrt(N), so
// N - T lies in (0, 16 sqrt(N)) except for 8-sigma events on either side.
static double grouped_mean(uint64_t N) { return std::max(0.0, (double)N - 8.0 * std::sqrt((double)N)); }
static uint64_t topup_bound(uint64_t N) {
    return std::min<uint64_t>(N, (uint64_t)(16.0 * std::sqrt((double)N)) + 64);
}

// counts -> offsets (+ compaction) -> chunk starts
struct Stratum { double u_lo, u_hi; };

constexpr uint64_t SPLIT_MIN_ROWS_PER_CELL = 32;
static bool use_split(uint64_t ncells, uint64_t N) { return N >= SPLIT_MIN_ROWS_PER_CELL * ncells; }

template <typename Cells>
static int plan_rows(const Cells &src, const double *cdf, uint64_t ncells, uint64_t N, uint2 key,
                     uint64_t offset, Stratum u, const Scratch &s, cudaStream_t st) {
    const bool split = use_split(ncells, N);
    const CdfSlice t{cdf, (uint32_t)ncells, u.u_lo, u.u_hi};
    const double lam_per_unit = grouped_mean(N) / (u.u_hi - u.u_lo);
    const int cblocks = (int)std::min<uint64_t>((ncells + 255) / 256, 148 * 16);
    poisson_counts_kernel<<<cblocks, 256, 0, st>>>(src, t, lam_per_unit, key, offset, split, s.counts);
    const uint32_t nvirt = (uint32_t)(split ? 2 * ncells : ncells);
    const uint32_t nscan = (nvirt + SCAN_TILE - 1) / SCAN_TILE;
    cudaMemsetAsync(s.tile_state, 0, (size_t)nscan * 8, st);
    cudaMemsetAsync(s.tile_counter, 0, 4, st);
    scan_compact_kernel<<<nscan, SCAN_THREADS, 0, st>>>(s.counts, nvirt, s.tile_state,
                                                       s.tile_counter, s.cid, s.coff, s.nnz, s.rows);
    const uint32_t chunk_rows = emit_chunk_rows(N);
    const uint32_t nchunks = (uint32_t)((N + chunk_rows - 1) / chunk_rows);
    chunk_start_kernel<<<std::max(1u, std::min((nchunks + 255) / 256, 148u * 8)), 256, 0, st>>>(
        s.coff, s.nnz, s.rows, chunk_rows, s.chunk_start);
    return check_launch("cell-major row planning");
}

template <int K, typename T, typename Cells>
int launch_emit(const Cells &src, const double *cdf, const Scratch &s, uint64_t ncells, uint64_t N, uint2 key,
                uint64_t offset, Stratum u, void *out, int64_t *cell_idx, cudaStream_t st) {
    const uint32_t split = use_split(ncells, N) ? 1u : 0u;
    const CdfSlice t{cdf, (uint32_t)ncells, u.u_lo, u.u_hi};
    const double lam_per_unit = grouped_mean(N) / (u.u_hi - u.u_lo);
    const uint32_t chunk_rows = emit_chunk_rows(N);
    const uint32_t nchunks = (uint32_t)((N + chunk_rows - 1) / chunk_rows);
    const int blocks = (int)std::min<uint32_t>((nchunks + EMIT_THREADS / 32 - 1) / (EMIT_THREADS / 32), 148 * 8);
    if (cell_idx)
        emit_kernel<K, T, Cells, true><<<blocks, EMIT_THREADS, src.smem_bytes(sizeof(T)), st>>>(
            src, s.cid, s.coff, s.nnz, s.rows, s.chunk_start, (uint32_t)N, chunk_rows, key, offset, split, t, lam_per_unit, (T *)out, cell_idx);
    else
        emit_kernel<K, T, Cells, false><<<blocks, EMIT_THREADS, src.smem_bytes(sizeof(T)), st>>>(
            src, s.cid, s.coff, s.nnz, s.rows, s.chunk_start, (uint32_t)N, chunk_rows, key, offset, split, t, lam_per_unit, (T *)out, cell_idx);
    return check_launch("cell-major emission");
}

static int check_common(int64_t N, uint64_t ncells, double u_lo, double u_hi, void *out, void *scratch,
                        size_t scratch_bytes, const char *what) {
    if (!(u_lo >= 0.0) || !(u_hi <= 1.0) || !(u_lo < u_hi))
        return fail(LINT_ERR_BAD_ARG, "%s: stratum [%g, %g) is not inside [0, 1]", what, u_lo, u_hi);
    // (row + 32 must not wrap in the emitter's 32-bit row arithmetic)
    if (N < 0 || N > (1ll << 32) - (1ll << 16))
        return fail(LINT_ERR_BAD_ARG, "%s: N must be in [0, 2^32 - 2^16]", what);
    if (ncells >= (1ull << 29)) return fail(LINT_ERR_UNSUPPORTED, "%s: needs < 2^29 cells", what);
    if (N > 0 && !out) return fail(LINT_ERR_BAD_ARG, "%s: out_dev is NULL", what);
    if (!scratch || scratch_bytes < carve(nullptr, ncells, (uint64_t)N).bytes)
        return fail(LINT_ERR_BAD_ARG, "%s: scratch too small", what);
    return LINT_OK;
}

template <int K>
int grid_cellmajor_k(const lint_grid_t *g, uint64_t nc, uint64_t N, uint2 key, uint64_t offset, Stratum u,
                     const Scratch &s, void *out, int64_t *cell_idx, cudaStream_t st) {
    if (g->dtype == LINT_F32) {
        GridCells<K, float> src{make_grid_view<K>(g, nc, true)};
        if (int rc = plan_rows(src, g->cdf_dev, nc, N, key, offset, u, s, st)) return rc;
        return launch_emit<K, float>(src, g->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
    }
    GridCells<K, double> src{make_grid_view<K>(g, nc, true)};
    if (int rc = plan_rows(src, g->cdf_dev, nc, N, key, offset, u, s, st)) return rc;
    return launch_emit<K, double>(src, g->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
}

// defined in lint_sample.cu: u-driven rows [*first_row_dev, N), at most max_rows of them
int topup_grid(const lint_grid_t *g, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
               uint64_t seed, uint64_t offset, double u_lo, double u_hi, void *out, int64_t *cell_idx,
               cudaStream_t st);
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
        if (int rc = plan_rows(src, c->cdf_dev, nc, N, key, offset, u, s, st)) return rc;
        return launch_emit<K, float>(src, c->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
    }
    ListCells<K, double> src{v};
    if (int rc = plan_rows(src, c->cdf_dev, nc, N, key, offset, u, s, st)) return rc;
    return launch_emit<K, double>(src, c->cdf_dev, s, nc, N, key, offset, u, out, cell_idx, st);
}

}  // namespace lint

using namespace lint;

extern "C" {

size_t lint_cellmajor_scratch_bytes(int64_t ncells, int64_t N) {
    if (ncells < 1 || N < 0) return 0;
    return carve(nullptr, (uint64_t)ncells, (uint64_t)N).bytes;
}

int lint_grid_sample_cellmajor(const lint_grid_t *g, int64_t N, uint64_t seed, uint64_t offset,
                               double u_lo, double u_hi, void *scratch_dev, size_t scratch_bytes, void *out_dev,
                               int64_t *cell_idx_dev, lint_stream_t stream) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, true, &nc, &nv)) return rc;
    if (int rc = check_common(N, nc, u_lo, u_hi, out_dev, scratch_dev, scratch_bytes, "lint_grid_sample_cellmajor")) return rc;
    if (N == 0) return LINT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const Scratch s = carve(scratch_dev, nc, (uint64_t)N);
    const Stratum u{u_lo, u_hi};
    int rc;
    switch (g->dim) {
        case 1: rc = grid_cellmajor_k<1>(g, nc, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 2: rc = grid_cellmajor_k<2>(g, nc, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 3: rc = grid_cellmajor_k<3>(g, nc, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 4: rc = grid_cellmajor_k<4>(g, nc, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 5: rc = grid_cellmajor_k<5>(g, nc, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        default: rc = grid_cellmajor_k<6>(g, nc, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
    }
    if (rc) return rc;
    // rows [T, N): the u-driven sampler on the same stratum
    return topup_grid(g, N, s.rows, (int64_t)topup_bound((uint64_t)N), seed, offset, u_lo, u_hi, out_dev,
                      cell_idx_dev, st);
}

int lint_cells_sample_cellmajor(const lint_cells_t *c, int64_t N, uint64_t seed, uint64_t offset,
                                double u_lo, double u_hi, void *scratch_dev, size_t scratch_bytes, void *out_dev,
                                int64_t *cell_idx_dev, lint_stream_t stream) {
    if (!c || c->dim < 1 || c->dim > LINT_MAX_DIM || c->ncells < 1 || !c->mins_dev || !c->widths_dev ||
        !c->corners_dev || !c->cdf_dev || (c->dtype != LINT_F32 && c->dtype != LINT_F64))
        return fail(LINT_ERR_BAD_ARG, "lint_cells_sample_cellmajor: bad cell-list descriptor");
    const uint64_t nc = (uint64_t)c->ncells;
    if (int rc = check_common(N, nc, u_lo, u_hi, out_dev, scratch_dev, scratch_bytes, "lint_cells_sample_cellmajor")) return rc;
    if (N == 0) return LINT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const Scratch s = carve(scratch_dev, nc, (uint64_t)N);
    const Stratum u{u_lo, u_hi};
    int rc;
    switch (c->dim) {
        case 1: rc = cells_cellmajor_k<1>(c, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 2: rc = cells_cellmajor_k<2>(c, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 3: rc = cells_cellmajor_k<3>(c, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 4: rc = cells_cellmajor_k<4>(c, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        case 5: rc = cells_cellmajor_k<5>(c, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
        default: rc = cells_cellmajor_k<6>(c, N, key, offset, u, s, out_dev, cell_idx_dev, st); break;
    }
    if (rc) return rc;
    return topup_cells(c, N, s.rows, (int64_t)topup_bound((uint64_t)N), seed, offset, u_lo, u_hi, out_dev,
                       cell_idx_dev, st);
}

}  // extern "C"
