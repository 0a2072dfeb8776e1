// Cell-major ("grouped") sampling: the throughput path.
//
// The u-driven sampler of lint_sample.cu spends its time on per-sample random
// access (guide, CDF, cell record) — work the reference also does per sample
// (utils.py:197, grid.py:440-467).  N i.i.d. draws of a cell are however fully
// described by the per-cell counts (n_c) ~ Multinomial(N, p_c), so this path
//   1. draws the counts exactly, by binomial splitting down a binary tree over
//      the flat cell range (probabilities are differences of the same fp64 CDF
//      the search path uses, grid.py:431 + utils.py:195-196),
//   2. turns them into row offsets with a single-pass decoupled look-back scan
//      (integer sums, hence deterministic) that also compacts away empty cells,
//   3. emits each cell's rows contiguously: the cell's box and corners are
//      loaded once per cell, every row costs only its k in-cell uniforms and the
//      k-D inverse CDF of sampling.py:41-94.
// The rows of one call are an exact i.i.d. sample as a multiset; their ORDER is
// by cell (C order).  Callers that need the reference's exchangeable row order
// (lintsampler.py:309-311) apply the row permutation afterwards.
#include <algorithm>
#include "lint_device.cuh"

namespace lint {

int validate_grid(const lint_grid_t *g, bool need_cdf, uint64_t *ncells_out, uint64_t *nverts_out);
template <int K> GridView<K> make_grid_view(const lint_grid_t *g, uint64_t ncells, bool stage_edges);

// Philox counter word 1 tags (word 0 = node / row group, words 2-3 = call offset)
constexpr uint32_t TAG_SPLIT = 0x53000000u;   // | level | attempt << 8
constexpr uint32_t TAG_COORD = 0x43000000u;   // | block index

// ---------------------------------------------------------------------------
// exact binomial sampler
// ---------------------------------------------------------------------------
struct SplitRng {
    uint2 key;
    uint32_t node, level, w2, w3;
    uint32_t attempt = 0;
    __device__ __forceinline__ void next(double &a, double &b) {
        const uint4 r = philox4x32_10(make_uint4(node, TAG_SPLIT | level | (attempt << 8), w2, w3), key);
        ++attempt;
        a = u53(r.x, r.y);
        b = u53(r.z, r.w);
    }
};

// Binomial(n, p) for 0 < p <= 1/2.  Small means: sequential inversion (BINV);
// otherwise Hörmann's transformed rejection with squeeze (BTRS, 1993).
__device__ uint32_t binomial_half(uint32_t n, double p, SplitRng &rng) {
    const double dn = (double)n;
    if (dn * p < 10.0) {
        const double q = 1.0 - p, s = p / q, a = (dn + 1.0) * s;
        for (;;) {
            double u, unused;
            rng.next(u, unused);
            double r = exp(dn * log1p(-p));
            uint32_t x = 0;
            bool ok = true;
            while (u > r) {
                u -= r;
                ++x;
                if (x > n || x > 200) { ok = false; break; }   // fp dust in the far tail: redraw
                r *= a / (double)x - s;
            }
            if (ok) return x;
        }
    }
    const double spq = sqrt(dn * p * (1.0 - p));
    const double b = 1.15 + 2.53 * spq;
    const double a = -0.0873 + 0.0248 * b + 0.01 * p;
    const double c = dn * p + 0.5;
    const double vr = 0.92 - 4.2 / b;
    const double alpha = (2.83 + 5.1 / b) * spq;
    const double lpq = log(p / (1.0 - p));
    const double m = floor((dn + 1.0) * p);
    const double h = lgamma(m + 1.0) + lgamma(dn - m + 1.0);
    for (;;) {
        double u, v;
        rng.next(u, v);
        u -= 0.5;
        const double us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + c);
        if (k < 0.0 || k > dn) continue;
        if (us >= 0.07 && v <= vr) return (uint32_t)k;
        v = log(v * alpha / (a / (us * us) + b));
        if (v <= h - lgamma(k + 1.0) - lgamma(dn - k + 1.0) + (k - m) * lpq) return (uint32_t)k;
    }
}

__device__ __forceinline__ uint32_t binomial(uint32_t n, double p, SplitRng &rng) {
    if (n == 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    return p <= 0.5 ? binomial_half(n, p, rng) : n - binomial_half(n, 1.0 - p, rng);
}

// The table the splitting tree reads: the normalised CDF restricted to the stratum
// [u_lo, u_hi) of the cell-selecting uniform ((0, 1) = the whole domain; a rank of
// a sharded run passes its own slice).
struct CdfSlice {
    const double *cdf;
    uint32_t ncells;
    double u_lo, u_hi;
};

// CDF just below cell i (F(0) = 0, F(i >= C) = 1), clamped to the stratum: the mass
// of cells [a, b) inside the stratum is F(b) - F(a)
__device__ __forceinline__ double cdf_below(const CdfSlice &t, uint64_t i) {
    const double f = i == 0 ? 0.0 : (i >= t.ncells ? 1.0 : __ldg(t.cdf + i - 1));
    return fmin(fmax(f, t.u_lo), t.u_hi);
}

// One level of the splitting tree: node j of `nnodes` covers cells
// [j*span, (j+1)*span); its count is split between its two halves.
__device__ __forceinline__ void split_node(const CdfSlice &t, uint64_t span, uint32_t level,
                                           uint32_t j, uint32_t n, uint2 key, uint64_t offset,
                                           uint32_t &left, uint32_t &right) {
    left = right = 0;
    if (n == 0) return;
    const uint64_t lo = (uint64_t)j * span, mid = lo + span / 2, hi = lo + span;
    if (mid >= t.ncells) { left = n; return; }            // right half lies beyond the grid
    const double f0 = cdf_below(t, lo), f1 = cdf_below(t, mid), f2 = cdf_below(t, hi);
    const double tot = f2 - f0;
    const double q = tot > 0.0 ? (f1 - f0) / tot : 0.5;
    SplitRng rng{key, j, level, (uint32_t)offset, (uint32_t)(offset >> 32)};
    left = binomial(n, q, rng);
    right = n - left;
}

// levels [0, top_levels): one CTA walks down the small top of the tree in shared memory
constexpr int TOP_LEVELS = 11;                // 2^11 = 2048 nodes after the top kernel
__global__ void __launch_bounds__(1024)
split_top_kernel(CdfSlice t, int depth, uint32_t n_total, uint2 key,
                 uint64_t offset, int levels, uint32_t *out) {
    __shared__ uint32_t a[2][1 << TOP_LEVELS];
    if (threadIdx.x == 0) a[0][0] = n_total;
    __syncthreads();
    for (int lv = 0; lv < levels; ++lv) {
        const uint32_t nn = 1u << lv;
        const uint64_t span = 1ull << (depth - lv);
        for (uint32_t j = threadIdx.x; j < nn; j += blockDim.x) {
            uint32_t l, r;
            split_node(t, span, lv, j, a[lv & 1][j], key, offset, l, r);
            a[(lv + 1) & 1][2 * j] = l;
            a[(lv + 1) & 1][2 * j + 1] = r;
        }
        __syncthreads();
    }
    for (uint32_t j = threadIdx.x; j < (1u << levels); j += blockDim.x) out[j] = a[levels & 1][j];
}

__global__ void __launch_bounds__(256)
split_level_kernel(CdfSlice t, int depth, int level, uint2 key, uint64_t offset,
                   const uint32_t *in, uint32_t *out) {
    const uint32_t nn = 1u << level;
    const uint64_t span = 1ull << (depth - level);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nn; j += gridDim.x * blockDim.x) {
        uint32_t l, r;
        split_node(t, span, level, j, __ldg(in + j), key, offset, l, r);
        reinterpret_cast<uint2 *>(out)[j] = make_uint2(l, r);
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
                    uint32_t n_total) {
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
        coff[nnz] = n_total;                                       // sentinel
    }
}

// first segment of every output tile: last j with coff[j] <= tile * tile_rows
__global__ void tile_start_kernel(const uint32_t *coff, const uint32_t *nnz_p, uint32_t ntiles,
                                  uint32_t tile_rows, uint32_t *tile_start) {
    const uint32_t nnz = *nnz_p;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        const uint32_t r0 = t * tile_rows;
        uint32_t lo = 0, hi = nnz;                                 // coff[nnz] = N > r0
        while (lo < hi) {                                          // first j with coff[j] > r0
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(coff + mid) > r0) hi = mid; else lo = mid + 1;
        }
        tile_start[t] = lo - 1;
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
};

constexpr int EMIT_THREADS = 256, EMIT_RPT = 8, EMIT_TILE = EMIT_THREADS * EMIT_RPT;

// K in-cell uniforms for each of the EMIT_RPT rows of a thread.  fp32 rows use one
// 32-bit word per coordinate (top 24 bits); fp64 rows two words (53 bits).
template <int K, typename T>
struct RowUniforms {
    static constexpr int WORDS = K * EMIT_RPT * (sizeof(T) == 8 ? 2 : 1);
    static constexpr int CALLS = (WORDS + 3) / 4;
    uint32_t w[4 * CALLS];
    __device__ __forceinline__ void fill(uint2 key, uint32_t rowgroup, uint64_t offset) {
#pragma unroll
        for (int q = 0; q < CALLS; ++q) {
            const uint4 r = philox4x32_10(
                make_uint4(rowgroup, TAG_COORD | q, (uint32_t)offset, (uint32_t)(offset >> 32)), key);
            w[4 * q] = r.x; w[4 * q + 1] = r.y; w[4 * q + 2] = r.z; w[4 * q + 3] = r.w;
        }
    }
    __device__ __forceinline__ void get(int i, float (&u)[K]) const {
#pragma unroll
        for (int d = 0; d < K; ++d) u[d] = (float)(w[i * K + d] >> 8) * (1.0f / 16777216.0f);
    }
    __device__ __forceinline__ void get(int i, double (&u)[K]) const {
#pragma unroll
        for (int d = 0; d < K; ++d) u[d] = u53(w[2 * (i * K + d)], w[2 * (i * K + d) + 1]);
    }
};

template <int K, typename T, typename Cells>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(Cells src, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ coff,
            const uint32_t *__restrict__ nnz_p, const uint32_t *__restrict__ tile_start, uint32_t N,
            uint2 key, uint64_t offset, T *__restrict__ out, int64_t *__restrict__ cell_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sedges = reinterpret_cast<T *>(smem_raw);
    __shared__ uint32_t s_off[EMIT_TILE + 1];
    __shared__ uint32_t s_cid[EMIT_TILE + 1];
    src.prepare(sedges);
    const uint32_t nnz = *nnz_p;
    const uint32_t ntiles = (N + EMIT_TILE - 1) / EMIT_TILE;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t r0 = tile * EMIT_TILE;
        const uint32_t j0 = __ldg(tile_start + tile);
        // every staged segment but the first starts inside the tile and holds >= 1 row,
        // so EMIT_TILE + 1 consecutive offsets always cover the tile
        const uint32_t cnt = min((uint32_t)EMIT_TILE + 1, nnz + 1 - j0);
        for (uint32_t i = threadIdx.x; i < cnt; i += EMIT_THREADS) {
            s_off[i] = __ldg(coff + j0 + i);
            s_cid[i] = j0 + i < nnz ? __ldg(cid + j0 + i) : 0;
        }
        __syncthreads();
        const uint32_t row0 = r0 + threadIdx.x * EMIT_RPT;
        if (row0 < N) {
            // last staged segment starting at or before row0
            uint32_t lo = 0, hi = cnt - 1;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo + 1) >> 1);
                if (s_off[mid] <= row0) lo = mid; else hi = mid - 1;
            }
            uint32_t seg = lo;
            uint32_t seg_end = s_off[seg + 1];
            T c0[1 << K], xlo[K], xhi[K];
            uint32_t cell = s_cid[seg];
            src.fetch(sedges, cell, c0, xlo, xhi);
            RowUniforms<K, T> rnd;
            rnd.fill(key, row0 / EMIT_RPT, offset);
            T x[EMIT_RPT][K];
#pragma unroll
            for (int i = 0; i < EMIT_RPT; ++i) {
                const uint32_t row = row0 + i;
                if (row >= seg_end && row < N) {                   // next non-empty cell
                    ++seg;
                    seg_end = s_off[seg + 1];
                    cell = s_cid[seg];
                    src.fetch(sedges, cell, c0, xlo, xhi);
                }
                T c[1 << K], u[K], z[K];
#pragma unroll
                for (int j = 0; j < (1 << K); ++j) c[j] = c0[j];
                rnd.get(i, u);
                incell<K>(c, u, z);                                // sampling.py:41-94
#pragma unroll
                for (int d = 0; d < K; ++d) x[i][d] = affine(xlo[d], xhi[d], z[d]);   // sampling.py:126
                if (cell_idx && row < N) __stcs(reinterpret_cast<long long *>(cell_idx) + row, (long long)cell);
            }
            // EMIT_RPT consecutive rows per thread: K * EMIT_RPT contiguous values
            T *dst = out + (size_t)row0 * K;
            if (row0 + EMIT_RPT <= N) {
                constexpr int VEC = 16 / (int)sizeof(T);
                static_assert((K * EMIT_RPT) % VEC == 0, "row group must be a whole number of 16-byte stores");
                const T *flat = &x[0][0];
#pragma unroll
                for (int v = 0; v < K * EMIT_RPT / VEC; ++v)
                    __stcs(reinterpret_cast<float4 *>(dst) + v, reinterpret_cast<const float4 *>(flat)[v]);
            } else {
                for (int i = 0; i < EMIT_RPT && row0 + i < N; ++i)
                    for (int d = 0; d < K; ++d) dst[i * K + d] = x[i][d];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
struct Scratch {
    uint32_t *counts_a, *counts_b, *cid, *coff, *tile_start, *nnz, *tile_counter;
    unsigned long long *tile_state;
    size_t bytes;
};

static int tree_depth(uint64_t ncells) {
    int d = 0;
    while ((1ull << d) < ncells) ++d;
    return d;
}

static Scratch carve(void *base, uint64_t ncells, uint64_t N) {
    const uint64_t cpad = 1ull << tree_depth(ncells);
    const uint64_t nscan = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    const uint64_t ntiles = (N + EMIT_TILE - 1) / EMIT_TILE;
    uintptr_t p = (uintptr_t)base;
    auto take = [&](size_t n) { uintptr_t q = p; p += (n + 255) & ~(size_t)255; return (void *)q; };
    Scratch s;
    s.counts_a = (uint32_t *)take(std::max<uint64_t>(cpad, 2) * 4);
    s.counts_b = (uint32_t *)take(std::max<uint64_t>(cpad, 2) * 4);
    s.cid = (uint32_t *)take((ncells + 1) * 4);
    s.coff = (uint32_t *)take((ncells + 1) * 4);
    s.tile_start = (uint32_t *)take(std::max<uint64_t>(ntiles, 1) * 4);
    s.tile_state = (unsigned long long *)take(nscan * 8);
    s.nnz = (uint32_t *)take(4);
    s.tile_counter = (uint32_t *)take(4);
    s.bytes = (size_t)(p - (uintptr_t)base);
    return s;
}

// counts -> offsets -> tile starts; returns the buffer holding the per-cell counts
static int plan_rows(const double *cdf, uint64_t ncells, uint64_t N, uint2 key, uint64_t offset,
                     double u_lo, double u_hi, const Scratch &s, cudaStream_t st) {
    const int depth = tree_depth(ncells);
    const CdfSlice t{cdf, (uint32_t)ncells, u_lo, u_hi};
    uint32_t *cur = s.counts_a, *nxt = s.counts_b;
    const int top = std::min(depth, TOP_LEVELS);
    split_top_kernel<<<1, 1024, 0, st>>>(t, depth, (uint32_t)N, key, offset, top, cur);
    for (int lv = top; lv < depth; ++lv) {
        const uint32_t nn = 1u << lv;
        const int blocks = (int)std::min<uint32_t>((nn + 255) / 256, 148 * 16);
        split_level_kernel<<<blocks, 256, 0, st>>>(t, depth, lv, key, offset, cur, nxt);
        std::swap(cur, nxt);
    }
    const uint32_t nscan = (uint32_t)((ncells + SCAN_TILE - 1) / SCAN_TILE);
    cudaMemsetAsync(s.tile_state, 0, (size_t)nscan * 8, st);
    cudaMemsetAsync(s.tile_counter, 0, 4, st);
    scan_compact_kernel<<<nscan, SCAN_THREADS, 0, st>>>(cur, (uint32_t)ncells, s.tile_state, s.tile_counter,
                                                       s.cid, s.coff, s.nnz, (uint32_t)N);
    const uint32_t ntiles = (uint32_t)((N + EMIT_TILE - 1) / EMIT_TILE);
    tile_start_kernel<<<std::max(1u, std::min((ntiles + 255) / 256, 148u * 8)), 256, 0, st>>>(
        s.coff, s.nnz, ntiles, EMIT_TILE, s.tile_start);
    return check_launch("cell-major row planning");
}

template <int K, typename T, typename Cells>
int launch_emit(const Cells &src, const Scratch &s, uint64_t N, uint2 key, uint64_t offset, void *out,
                int64_t *cell_idx, cudaStream_t st) {
    const uint32_t ntiles = (uint32_t)((N + EMIT_TILE - 1) / EMIT_TILE);
    const int blocks = (int)std::min<uint32_t>(ntiles, 148 * 8);
    emit_kernel<K, T, Cells><<<blocks, EMIT_THREADS, src.smem_bytes(sizeof(T)), st>>>(
        src, s.cid, s.coff, s.nnz, s.tile_start, (uint32_t)N, key, offset, (T *)out, cell_idx);
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

template <int K>
int cells_cellmajor_k(const lint_cells_t *c, uint64_t N, uint2 key, uint64_t offset, const Scratch &s,
                      void *out, int64_t *cell_idx, cudaStream_t st) {
    CellsView v;
    v.mins = c->mins_dev;
    v.widths = c->widths_dev;
    v.corners = c->corners_dev;
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
    switch (g->dim) {
        case 1: return grid_cellmajor_k<1>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 2: return grid_cellmajor_k<2>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 3: return grid_cellmajor_k<3>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 4: return grid_cellmajor_k<4>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 5: return grid_cellmajor_k<5>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st);
        default: return grid_cellmajor_k<6>(g, nc, N, key, offset, s, out_dev, cell_idx_dev, st);
    }
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
    switch (c->dim) {
        case 1: return cells_cellmajor_k<1>(c, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 2: return cells_cellmajor_k<2>(c, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 3: return cells_cellmajor_k<3>(c, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 4: return cells_cellmajor_k<4>(c, N, key, offset, s, out_dev, cell_idx_dev, st);
        case 5: return cells_cellmajor_k<5>(c, N, key, offset, s, out_dev, cell_idx_dev, st);
        default: return cells_cellmajor_k<6>(c, N, key, offset, s, out_dev, cell_idx_dev, st);
    }
}

}  // extern "C"
