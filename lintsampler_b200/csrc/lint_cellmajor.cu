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
// Row ORDER is by cell (C order) for the first T rows, draw order for the
// top-up tail.  Callers that need the reference's exchangeable row order
// (lintsampler.py:309-311) apply the row permutation afterwards.
#include <algorithm>
#include <cmath>
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
__device__ __forceinline__ double log_factorial(double k) {
    const double tab[16] = {0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458,
                            4.787491742782046, 6.579251212010101, 8.525161361065415,
                            10.604602902745251, 12.801827480081469, 15.104412573075516,
                            17.502307845873887, 19.987214495661885, 22.552163853123425,
                            25.19122118273868, 27.89927138384089};
    if (k < 16.0) return tab[(int)k];
    const double x = k + 1.0, r = 1.0 / x, r2 = r * r;
    return (x - 0.5) * log(x) - x + 0.9189385332046727 +
           r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
}

// Poisson(lam): sequential inversion with one uniform below 10, Hörmann's
// transformed rejection with squeeze (PTRS, 1993) above.
__device__ uint32_t poisson(double lam, CellRng &rng) {
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
                p *= lam / (double)x;
            }
            if (ok) return x;
            rng.next(u, unused);
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        double u, v;
        rng.next(u, v);
        u -= 0.5;
        const double us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= vr) return (uint32_t)k;
        if (k < 0.0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - log_factorial(k))
            return (uint32_t)k;
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

__global__ void __launch_bounds__(256)
poisson_counts_kernel(CdfSlice t, double lam_per_unit, uint2 key, uint64_t offset, uint32_t *counts) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < t.ncells; c += gridDim.x * blockDim.x) {
        // mass of the cell inside the stratum: clamp both CDF values, then difference
        const double f0 = c == 0 ? 0.0 : __ldg(t.cdf + c - 1), f1 = __ldg(t.cdf + c);
        const double m = fmin(fmax(f1, t.u_lo), t.u_hi) - fmin(fmax(f0, t.u_lo), t.u_hi);
        CellRng rng{key, c, (uint32_t)offset, (uint32_t)(offset >> 32)};
        counts[c] = poisson(m * lam_per_unit, rng);
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
constexpr int EMIT_THREADS = LINT_EMIT_THREADS, EMIT_CHUNK = 2048, EMIT_U = LINT_EMIT_U;

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
};

template <int K, typename T, bool CELLS>
__device__ __forceinline__ void emit_row(const CellCoefs<K, T> &coef, const T (&xlo)[K], const T (&xhi)[K],
                                         const T (&u)[K], uint32_t row, uint32_t cell,
                                         T *__restrict__ out, int64_t *__restrict__ cell_idx) {
    T z[K];
    coef.draw(u, z);                                               // sampling.py:41-94
#pragma unroll
    for (int d = 0; d < K; ++d)
        store_stream(out + (size_t)row * K + d, affine(xlo[d], xhi[d], z[d]));   // sampling.py:126
    if (CELLS) __stcs(reinterpret_cast<long long *>(cell_idx) + row, (long long)cell);
}

// Every warp works on its own chunks of EMIT_CHUNK consecutive rows and walks the
// segments (= non-empty cells) that intersect the chunk in order, through a small
// private shared-memory window of the segment table (EMIT_WIN entries, refilled
// as the walk advances; the next chunk's first window is prefetched with cp.async
// while the current chunk is computed).  No CTA-wide barrier.  Lanes map to
// consecutive rows:
//   * 32*EMIT_U rows inside one cell: the cell's box and corners are loaded once
//     (all lanes the same address) and every lane draws EMIT_U rows from
//     straight-line code — no per-row bookkeeping, coalesced stores;
//   * 32 rows inside one cell: the same with one row per lane;
//   * 32 rows spanning several cells: each lane finds its own cell in the window
//     (the cost of this path is bounded however small the cells are).
// fp32, k <= 3: the per-cell constants (box, hoisted quantile coefficients) of 32
// consecutive segments are computed at once, one segment per lane, and staged in
// shared memory; entering a cell then costs a few broadcast loads instead of a
// warp-redundant fetch + set.
constexpr int EMIT_WIN = 64;
constexpr uint32_t EMIT_PRE = 32;            // segments per staged block (= lanes)
constexpr uint32_t EMIT_PRE_NONE = 0x80000000u;
// A run of at least this many rows of one cell is emitted by the uniform path even if it
// does not fill the warp (idle lanes are cheaper than the per-lane path); shorter runs go
// through the per-lane path together with the cells that follow.
constexpr uint32_t EMIT_MIN_UNIFORM = 16;

template <int K, typename T, typename Cells, bool CELLS>
__global__ void __launch_bounds__(EMIT_THREADS, sizeof(T) == 4 && K <= 3 ? LINT_EMIT_MINB : 1)
emit_kernel(Cells src, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ coff,
            const uint32_t *__restrict__ nnz_p, const uint32_t *__restrict__ rows_p,
            const uint32_t *__restrict__ chunk_start, uint32_t n_requested, uint2 key, uint64_t offset,
            T *__restrict__ out, int64_t *__restrict__ cell_idx) {
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
    const bool pren = src.prenormalised();
    const uint32_t nnz = *nnz_p;
    const uint32_t N = min(*rows_p, n_requested);      // grouped rows T (T <= N but for an 8-sigma event)
    const uint32_t nchunks = (N + EMIT_CHUNK - 1) / EMIT_CHUNK;
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
        const uint32_t r0 = chunk * EMIT_CHUNK, r1 = min(r0 + EMIT_CHUNK, N);
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
        uint32_t loaded = 0xffffffffu;                   // segment whose cell the registers hold
        T xlo[K], xhi[K];
        CellCoefs<K, T> coef;
        uint32_t cell = 0;
        uint32_t cbase = EMIT_PRE_NONE;                  // staged block = segments [cbase, cbase + 32)
        float4 *s_pre = s_pre_all[PRE ? warp : 0];
        // stage the constants of segments [jb, jb + 32) that start inside this chunk;
        // the window must hold entries jb .. jb + 32
        auto stage = [&](uint32_t jb) {
            __syncwarp();
            const uint32_t js = jb + lane;
            if (js < nnz && s_off[js - jbase] < r1) {
                T c0[1 << K], lo[K], hi[K];
                src.fetch(sedges, s_cid[js - jbase], c0, lo, hi);
                CellCoefs<K, T> cf;
                cf.set(c0, pren);
                float w[4 * PRE_Q];
                cf.pack(w);
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    w[CellCoefs<K, T>::WORDS + d] = (float)lo[d];
                    w[CellCoefs<K, T>::WORDS + K + d] = (float)hi[d];
                }
#pragma unroll
                for (int q = 0; q < (PRE_WORDS + 3) / 4; ++q)
                    s_pre[lane * PRE_Q + q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
            cbase = jb;
            __syncwarp();
        };
        auto unstage = [&](uint32_t seg) {               // seg in [cbase, cbase + 32) and in the window
            float w[4 * PRE_Q];
#pragma unroll
            for (int q = 0; q < (PRE_WORDS + 3) / 4; ++q) {
                const float4 v = s_pre[(seg - cbase) * PRE_Q + q];
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
            coef.unpack(w);
#pragma unroll
            for (int d = 0; d < K; ++d) {
                xlo[d] = (T)w[CellCoefs<K, T>::WORDS + d];
                xhi[d] = (T)w[CellCoefs<K, T>::WORDS + K + d];
            }
            cell = s_cid[seg - jbase];
        };
        while (row < r1) {
            for (;;) {                                   // segment containing `row`
                if (j + 1 - jbase >= EMIT_WIN) refill(j);
                if (off(j + 1) > row) break;
                ++j;
            }
            const uint32_t seg_end = min(off(j + 1), r1);
            if (seg_end - row >= EMIT_MIN_UNIFORM || seg_end == r1) {
                // ---- rows [row, ...) of one cell, lanes = consecutive rows
                if (loaded != j) {
                    if constexpr (PRE) {
                        if (j - cbase >= EMIT_PRE) {
                            if (j + EMIT_PRE + 1 - jbase >= EMIT_WIN) refill(j);
                            stage(j);
                        }
                        unstage(j);
                    } else {
                        T c0[1 << K];
                        cell = s_cid[j - jbase];
                        src.fetch(sedges, cell, c0, xlo, xhi);
                        coef.set(c0, pren);
                    }
                    loaded = j;
                }
                while (seg_end - row >= 32u * EMIT_U) {
                    const uint32_t first = row + lane;
                    GroupUniforms<K, T, EMIT_U> rnd;
                    rnd.fill(key, first, 0, offset);
#pragma unroll
                    for (int t = 0; t < EMIT_U; ++t) {
                        T u[K];
                        rnd.get(t, u);
                        emit_row<K, T, CELLS>(coef, xlo, xhi, u, first + 32u * t, cell, out, cell_idx);
                    }
                    row += 32u * EMIT_U;
                }
                // the rest of the cell, however short: its constants are in registers already,
                // which makes idle lanes cheaper than handing the tail to the per-lane path
                while (row < seg_end) {
                    const uint32_t r = row + lane;
                    if (r < seg_end) {
                        GroupUniforms<K, T, 1> rnd;
                        rnd.fill(key, r, 1, offset);
                        T u[K];
                        rnd.get(0, u);
                        emit_row<K, T, CELLS>(coef, xlo, xhi, u, r, cell, out, cell_idx);
                    }
                    row = min(row + 32u, seg_end);
                }
            } else {
                // ---- the next 32 rows span several cells: one row per lane, own cell each.
                // Segments hold >= 1 row, so lane l's segment is within [j, j + l].
                if (j + 33 - jbase >= EMIT_WIN) refill(j);
                const uint32_t r = row + lane;
                const uint32_t lim = min(row + 32u, r1);
                uint32_t lo = j;
                if (r < lim) {
                    uint32_t hi = min(j + lane, nnz - 1);
                    while (lo < hi) {
                        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
                        if (s_off[mid - jbase] <= r) lo = mid; else hi = mid - 1;
                    }
                    if constexpr (!PRE) {
                        T c0[1 << K];
                        cell = s_cid[lo - jbase];
                        src.fetch(sedges, cell, c0, xlo, xhi);
                        coef.set(c0, pren);
                    }
                }
                const uint32_t last = __shfl_sync(0xffffffffu, lo, (int)(lim - row) - 1);   // the last row's segment
                if constexpr (PRE) {
                    if (j - cbase >= EMIT_PRE || last - cbase >= EMIT_PRE) stage(j);       // last <= j + 31
                }
                if (r < lim) {
                    if constexpr (PRE) unstage(lo);
                    GroupUniforms<K, T, 1> rnd;
                    rnd.fill(key, r, 1, offset);
                    T u[K];
                    rnd.get(0, u);
                    emit_row<K, T, CELLS>(coef, xlo, xhi, u, r, cell, out, cell_idx);
                }
                loaded = 0xffffffffu;
                j = last;                                // resume at the last row's segment
                row = lim;
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

static Scratch carve(void *base, uint64_t ncells, uint64_t N) {
    const uint64_t nscan = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    const uint64_t nchunks = (N + EMIT_CHUNK - 1) / EMIT_CHUNK;
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
static int plan_rows(const double *cdf, uint64_t ncells, uint64_t N, uint2 key, uint64_t offset,
                     double u_lo, double u_hi, const Scratch &s, cudaStream_t st) {
    const CdfSlice t{cdf, (uint32_t)ncells, u_lo, u_hi};
    const double lam_per_unit = grouped_mean(N) / (u_hi - u_lo);
    const int cblocks = (int)std::min<uint64_t>((ncells + 255) / 256, 148 * 16);
    poisson_counts_kernel<<<cblocks, 256, 0, st>>>(t, lam_per_unit, key, offset, s.counts);
    const uint32_t nscan = (uint32_t)((ncells + SCAN_TILE - 1) / SCAN_TILE);
    cudaMemsetAsync(s.tile_state, 0, (size_t)nscan * 8, st);
    cudaMemsetAsync(s.tile_counter, 0, 4, st);
    scan_compact_kernel<<<nscan, SCAN_THREADS, 0, st>>>(s.counts, (uint32_t)ncells, s.tile_state,
                                                       s.tile_counter, s.cid, s.coff, s.nnz, s.rows);
    const uint32_t nchunks = (uint32_t)((N + EMIT_CHUNK - 1) / EMIT_CHUNK);
    chunk_start_kernel<<<std::max(1u, std::min((nchunks + 255) / 256, 148u * 8)), 256, 0, st>>>(
        s.coff, s.nnz, s.rows, EMIT_CHUNK, s.chunk_start);
    return check_launch("cell-major row planning");
}

template <int K, typename T, typename Cells>
int launch_emit(const Cells &src, const Scratch &s, uint64_t N, uint2 key, uint64_t offset, void *out,
                int64_t *cell_idx, cudaStream_t st) {
    const uint32_t nchunks = (uint32_t)((N + EMIT_CHUNK - 1) / EMIT_CHUNK);
    const int blocks = (int)std::min<uint32_t>((nchunks + EMIT_THREADS / 32 - 1) / (EMIT_THREADS / 32), 148 * 8);
    if (cell_idx)
        emit_kernel<K, T, Cells, true><<<blocks, EMIT_THREADS, src.smem_bytes(sizeof(T)), st>>>(
            src, s.cid, s.coff, s.nnz, s.rows, s.chunk_start, (uint32_t)N, key, offset, (T *)out, cell_idx);
    else
        emit_kernel<K, T, Cells, false><<<blocks, EMIT_THREADS, src.smem_bytes(sizeof(T)), st>>>(
            src, s.cid, s.coff, s.nnz, s.rows, s.chunk_start, (uint32_t)N, key, offset, (T *)out, cell_idx);
    return check_launch("cell-major emission");
}

static int check_common(int64_t N, uint64_t ncells, double u_lo, double u_hi, void *out, void *scratch,
                        size_t scratch_bytes, const char *what) {
    if (!(u_lo >= 0.0) || !(u_hi <= 1.0) || !(u_lo < u_hi))
        return fail(LINT_ERR_BAD_ARG, "%s: stratum [%g, %g) is not inside [0, 1]", what, u_lo, u_hi);
    if (N < 0 || N >= (1ll << 32)) return fail(LINT_ERR_BAD_ARG, "%s: N must be in [0, 2^32)", what);
    if (ncells >= (1ull << 30)) return fail(LINT_ERR_UNSUPPORTED, "%s: needs < 2^30 cells", what);
    if (N > 0 && !out) return fail(LINT_ERR_BAD_ARG, "%s: out_dev is NULL", what);
    if (!scratch || scratch_bytes < carve(nullptr, ncells, (uint64_t)N).bytes)
        return fail(LINT_ERR_BAD_ARG, "%s: scratch too small", what);
    return LINT_OK;
}

template <int K>
int grid_cellmajor_k(const lint_grid_t *g, uint64_t nc, uint64_t N, uint2 key, uint64_t offset,
                     const Scratch &s, void *out, int64_t *cell_idx, cudaStream_t st) {
    if (g->dtype == LINT_F32) {
        GridCells<K, float> src{make_grid_view<K>(g, nc, true)};
        return launch_emit<K, float>(src, s, N, key, offset, out, cell_idx, st);
    }
    GridCells<K, double> src{make_grid_view<K>(g, nc, true)};
    return launch_emit<K, double>(src, s, N, key, offset, out, cell_idx, st);
}

// defined in lint_sample.cu: u-driven rows [*first_row_dev, N), at most max_rows of them
int topup_grid(const lint_grid_t *g, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
               uint64_t seed, uint64_t offset, double u_lo, double u_hi, void *out, int64_t *cell_idx,
               cudaStream_t st);
int topup_cells(const lint_cells_t *c, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
                uint64_t seed, uint64_t offset, double u_lo, double u_hi, void *out, int64_t *cell_idx,
                cudaStream_t st);

template <int K>
int cells_cellmajor_k(const lint_cells_t *c, uint64_t N, uint2 key, uint64_t offset, const Scratch &s,
                      void *out, int64_t *cell_idx, cudaStream_t st) {
    CellsView v;
    v.mins = c->mins_dev;
    v.widths = c->widths_dev;
    v.corners = c->corners_dev;
    v.rec = c->cell_records_dev;
    v.table.cdf = c->cdf_dev;
    v.table.guide = nullptr;
    v.table.guide_bits = 0;
    v.table.n = (uint32_t)c->ncells;
    if (c->dtype == LINT_F32) {
        ListCells<K, float> src{v};
        return launch_emit<K, float>(src, s, N, key, offset, out, cell_idx, st);
    }
    ListCells<K, double> src{v};
    return launch_emit<K, double>(src, s, N, key, offset, out, cell_idx, st);
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
    if (int rc = plan_rows(g->cdf_dev, nc, (uint64_t)N, key, offset, u_lo, u_hi, s, st)) return rc;
    int rc;
    switch (g->dim) {
        case 1: rc = grid_cellmajor_k<1>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 2: rc = grid_cellmajor_k<2>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 3: rc = grid_cellmajor_k<3>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 4: rc = grid_cellmajor_k<4>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 5: rc = grid_cellmajor_k<5>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        default: rc = grid_cellmajor_k<6>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st); break;
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
    if (int rc = plan_rows(c->cdf_dev, nc, (uint64_t)N, key, offset, u_lo, u_hi, s, st)) return rc;
    int rc;
    switch (c->dim) {
        case 1: rc = cells_cellmajor_k<1>(c, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 2: rc = cells_cellmajor_k<2>(c, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 3: rc = cells_cellmajor_k<3>(c, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 4: rc = cells_cellmajor_k<4>(c, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        case 5: rc = cells_cellmajor_k<5>(c, N, key, offset, s, out_dev, cell_idx_dev, st); break;
        default: rc = cells_cellmajor_k<6>(c, N, key, offset, s, out_dev, cell_idx_dev, st); break;
    }
    if (rc) return rc;
    return topup_cells(c, N, s.rows, (int64_t)topup_bound((uint64_t)N), seed, offset, u_lo, u_hi, out_dev,
                       cell_idx_dev, st);
}

}  // extern "C"
