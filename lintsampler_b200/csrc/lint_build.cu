// Grid-build kernels: analytic vertex evaluation, trapezoid cell masses,
// NumPy-identical pairwise sum, CDF construction (parallel / NumPy-sequential),
// guide table.  C ABI declared in include/lintsampler_b200.h.
//
// Reference citations are relative to /root/reference/src/lintsampler.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include "lint_device.cuh"
#include "lint_pdf.cuh"
#include "lint_pairwise.cuh"

namespace lint {

std::string &last_error_slot() {
    static thread_local std::string s;
    return s;
}

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_slot() = buf;
    return code;
}

static const uint64_t *g_offset_counter = nullptr;
const uint64_t *offset_counter() { return g_offset_counter; }

__global__ void advance_counter_kernel(uint64_t *counter, uint64_t by) { *counter += by; }

bool pdl_enabled() {
    static const bool on = [] {
        const char *v = getenv("LINT_PDL");
        return v && v[0] == '1';
    }();
    return on;
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LINT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return LINT_OK;
}

// ---------------------------------------------------------------------------
// descriptor validation shared with lint_sample.cu
// ---------------------------------------------------------------------------
int validate_grid(const lint_grid_t *g, bool need_cdf, uint64_t *ncells_out, uint64_t *nverts_out) {
    if (!g) return fail(LINT_ERR_BAD_ARG, "grid descriptor is NULL");
    if (g->dim < 1 || g->dim > LINT_MAX_DIM)
        return fail(LINT_ERR_BAD_ARG, "grid dim %d outside 1..%d", g->dim, LINT_MAX_DIM);
    if (g->dtype != LINT_F32 && g->dtype != LINT_F64)
        return fail(LINT_ERR_BAD_ARG, "grid dtype %d is not LINT_F32/LINT_F64", g->dtype);
    uint64_t nc = 1, nv = 1;
    for (int d = 0; d < g->dim; ++d) {
        if (g->ncells[d] < 1) return fail(LINT_ERR_BAD_ARG, "grid ncells[%d] < 1", d);
        if (!g->edges_dev[d]) return fail(LINT_ERR_BAD_ARG, "grid edges_dev[%d] is NULL", d);
        nc *= (uint64_t)g->ncells[d];
        nv *= (uint64_t)g->ncells[d] + 1;
        if (nv >= (1ull << 31))
            return fail(LINT_ERR_UNSUPPORTED, "grids with >= 2^31 vertices are not supported");
    }
    if (!g->vertex_densities_dev) return fail(LINT_ERR_BAD_ARG, "grid vertex_densities_dev is NULL");
    if (need_cdf && !g->cdf_dev) return fail(LINT_ERR_BAD_ARG, "grid cdf_dev is NULL");
    if (g->guide_bits < 0 || g->guide_bits > 30 || (g->guide_bits > 0 && !g->guide_dev))
        return fail(LINT_ERR_BAD_ARG, "grid guide table inconsistent (bits=%d)", g->guide_bits);
    if (ncells_out) *ncells_out = nc;
    if (nverts_out) *nverts_out = nv;
    return LINT_OK;
}

// Shared-memory budget for the staged edge arrays (elements of the arithmetic type)
constexpr uint32_t MAX_STAGED_EDGES = 4096;

template <int K>
GridView<K> make_grid_view(const lint_grid_t *g, uint64_t ncells, bool stage_edges) {
    GridView<K> v;
    v.V = g->vertex_densities_dev;
    v.rec = g->cell_records_dev;
    uint32_t stride = 1, eoff = 0;
    for (int d = K - 1; d >= 0; --d) {
        v.edges[d] = g->edges_dev[d];
        v.n[d] = (uint32_t)g->ncells[d];
        v.vstride[d] = stride;
        stride *= (uint32_t)g->ncells[d] + 1;
        // magic division by n[d] for dividends < 2^31 (Granlund & Montgomery):
        // L = ceil(log2 n), m = floor(2^(31+L) / n) + 1, q = (r * m) >> (31 + L)
        const uint32_t n = v.n[d];
        if (n >= 2) {
            uint32_t L = 0;
            while ((1ull << L) < n) ++L;
            v.div_m[d] = (uint32_t)(((1ull << (31 + L)) / n) + 1);
            v.div_s[d] = L - 1;
        } else {
            v.div_m[d] = 0;
            v.div_s[d] = 0;
        }
    }
    for (int d = 0; d < K; ++d) {
        v.edge_off[d] = eoff;
        eoff += v.n[d] + 1;
    }
    v.edge_total = (stage_edges && eoff <= MAX_STAGED_EDGES) ? eoff : 0;
    for (int c = 0; c < (1 << K); ++c) {
        uint32_t off = 0;
        for (int d = 0; d < K; ++d)
            if ((c >> (K - 1 - d)) & 1) off += v.vstride[d];
        v.corner_off[c] = off;
    }
    v.table.cdf = g->cdf_dev;
    v.table.guide = g->guide_dev;
    v.table.guide_bits = g->guide_bits;
    v.table.n = (uint32_t)ncells;
    return v;
}
template GridView<1> make_grid_view<1>(const lint_grid_t *, uint64_t, bool);
template GridView<2> make_grid_view<2>(const lint_grid_t *, uint64_t, bool);
template GridView<3> make_grid_view<3>(const lint_grid_t *, uint64_t, bool);
template GridView<4> make_grid_view<4>(const lint_grid_t *, uint64_t, bool);
template GridView<5> make_grid_view<5>(const lint_grid_t *, uint64_t, bool);
template GridView<6> make_grid_view<6>(const lint_grid_t *, uint64_t, bool);

// ---------------------------------------------------------------------------
// fused Gaussian-mixture vertex evaluation            [replaces grid.py:384-398]
// ---------------------------------------------------------------------------
template <int K, typename DensT>
__global__ void __launch_bounds__(256)
gmm_eval_kernel(GridView<K> g, GmmParams<K> p, double scale, uint32_t nverts, DensT *out) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nverts; v += gridDim.x * blockDim.x) {
        double x[K];
        uint32_t r = v;
#pragma unroll
        for (int d = K - 1; d >= 0; --d) {
            const uint32_t nd = g.n[d] + 1;
            uint32_t i;
            if (d == 0) { i = r; } else { const uint32_t q = r / nd; i = r - q * nd; r = q; }
            x[d] = __ldg(g.edges[d] + i);
        }
        const double f = p(x);
        out[v] = (DensT)(f * scale);
    }
}

template <int K>
int eval_gmm_impl(const lint_grid_t *g, const lint_gmm_t *gmm, double scale, uint64_t nverts,
                  cudaStream_t st) {
    const GmmParams<K> p = fill_gmm<K>(gmm);
    GridView<K> v = make_grid_view<K>(g, 1, false);
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((nverts + threads - 1) / threads, 148 * 32);
    if (g->dtype == LINT_F32)
        gmm_eval_kernel<K, float><<<blocks, threads, 0, st>>>(v, p, scale, (uint32_t)nverts,
                                                             (float *)g->vertex_densities_dev);
    else
        gmm_eval_kernel<K, double><<<blocks, threads, 0, st>>>(v, p, scale, (uint32_t)nverts,
                                                              (double *)g->vertex_densities_dev);
    return check_launch("lint_grid_eval_gmm");
}

// ---------------------------------------------------------------------------
// fused double-power-law halo evaluation   [3_dark_matter.ipynb cells 11, 16 through grid.py:384-398]
// ---------------------------------------------------------------------------
template <int K, typename DensT>
__global__ void __launch_bounds__(256)
halo_eval_kernel(GridView<K> g, HaloParams<K> p, double scale, uint32_t nverts, DensT *out) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nverts; v += gridDim.x * blockDim.x) {
        double x[K];
        uint32_t r = v;
#pragma unroll
        for (int d = K - 1; d >= 0; --d) {
            const uint32_t nd = g.n[d] + 1;
            uint32_t i;
            if (d == 0) { i = r; } else { const uint32_t q = r / nd; i = r - q * nd; r = q; }
            x[d] = __ldg(g.edges[d] + i);
        }
        const double f = p(x);
        out[v] = (DensT)(f * scale);
    }
}

template <int K>
int eval_halos_impl(const lint_grid_t *g, const lint_halos_t *h, double scale, uint64_t nverts,
                    cudaStream_t st) {
    const HaloParams<K> p = fill_halos<K>(h);
    GridView<K> v = make_grid_view<K>(g, 1, false);
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((nverts + threads - 1) / threads, 148 * 32);
    if (g->dtype == LINT_F32)
        halo_eval_kernel<K, float><<<blocks, threads, 0, st>>>(v, p, scale, (uint32_t)nverts,
                                                              (float *)g->vertex_densities_dev);
    else
        halo_eval_kernel<K, double><<<blocks, threads, 0, st>>>(v, p, scale, (uint32_t)nverts,
                                                               (double *)g->vertex_densities_dev);
    return check_launch("lint_grid_eval_halos");
}

// ---------------------------------------------------------------------------
// trapezoid cell masses + density validity check   [grid.py:469-518, 411, 401-404]
// ---------------------------------------------------------------------------
// mass of one cell and the validity bits of its corners: the sum starts from 0 and adds the
// corners in corner order (`sum = np.zeros(shape)` then += per corner, grid.py:489-497), one
// multiplication by 1/2^k, volume ((w0*w1)*w2)... (np.outer chain, grid.py:516-517)
template <int K, typename DensT, bool CHECK = true>
__device__ __forceinline__ double cell_mass_one(const GridView<K> &g, uint32_t cell, uint32_t &bad) {
    double c[1 << K], lo[K], hi[K];
    grid_fetch<K, DensT, double>(g, (const double *)nullptr, cell, c, lo, hi);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < (1 << K); ++j) {
        acc = __dadd_rn(acc, c[j]);
        if (CHECK) {
            if (c[j] < 0.0) bad |= LINT_FLAG_NEGATIVE;
            if (!isfinite(c[j])) bad |= LINT_FLAG_NONFINITE;
        }
    }
    const double mean = __dmul_rn(acc, 1.0 / (1 << K));     // exact /2^K
    double vol = __dsub_rn(hi[0], lo[0]);                   // np.diff, then outer products
#pragma unroll
    for (int d = 1; d < K; ++d) vol = __dmul_rn(vol, __dsub_rn(hi[d], lo[d]));
    return __dmul_rn(mean, vol);
}

template <int K, typename DensT>
__global__ void __launch_bounds__(256)
cell_mass_kernel(GridView<K> g, uint32_t ncells, double *masses, uint32_t *flags) {
    uint32_t bad = 0;
    for (uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncells;
         cell += gridDim.x * blockDim.x)
        masses[cell] = cell_mass_one<K, DensT>(g, cell, bad);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad && (threadIdx.x & 31) == 0) atomicOr(flags, bad);
}

template <int K>
int masses_impl(const lint_grid_t *g, uint64_t ncells, double *masses, uint32_t *flags,
                cudaStream_t st) {
    GridView<K> v = make_grid_view<K>(g, ncells, false);
    v.rec = nullptr;                       // masses always come from the vertex array
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((ncells + threads - 1) / threads, 148 * 32);
    if (g->dtype == LINT_F32)
        cell_mass_kernel<K, float><<<blocks, threads, 0, st>>>(v, (uint32_t)ncells, masses, flags);
    else
        cell_mass_kernel<K, double><<<blocks, threads, 0, st>>>(v, (uint32_t)ncells, masses, flags);
    return check_launch("lint_grid_masses");
}

// ---------------------------------------------------------------------------
// per-cell corner records: the 2^K corner densities of every cell (fp32: divided by
// the cell's largest, since a draw depends on their ratios only) laid out contiguously, so the sampler reads one record instead of 2^(K-1) scattered
// vertex pairs  [layout for grid.py:440-467]
// ---------------------------------------------------------------------------
template <int K, typename DensT>
__global__ void __launch_bounds__(256)
cell_records_kernel(GridView<K> g, uint32_t ncells, DensT *rec) {
    for (uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncells;
         cell += gridDim.x * blockDim.x) {
        uint32_t i[K];
        grid_unravel<K>(g, cell, i);
        uint32_t base = 0;
#pragma unroll
        for (int d = 0; d < K; ++d) base += i[d] * g.vstride[d];
        const DensT *V = reinterpret_cast<const DensT *>(g.V);
        DensT v[1 << K];
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) v[j] = __ldg(V + base + g.corner_off[j]);
        if (sizeof(DensT) == 4) {                 // fp32 records: corners / largest corner
            float m = 0.f;
#pragma unroll
            for (int j = 0; j < (1 << K); ++j) m = fmaxf(m, (float)v[j]);
            const float inv = m > 0.f ? 1.0f / m : 0.f;
#pragma unroll
            for (int j = 0; j < (1 << K); ++j) v[j] = (DensT)((float)v[j] * inv);
        }
        DensT *dst = rec + (size_t)cell * (1 << K);
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) dst[j] = v[j];
    }
}

template <int K>
int records_impl(const lint_grid_t *g, uint64_t ncells, void *rec, cudaStream_t st) {
    GridView<K> v = make_grid_view<K>(g, ncells, false);
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((ncells + threads - 1) / threads, 148 * 32);
    if (g->dtype == LINT_F32)
        cell_records_kernel<K, float><<<blocks, threads, 0, st>>>(v, (uint32_t)ncells, (float *)rec);
    else
        cell_records_kernel<K, double><<<blocks, threads, 0, st>>>(v, (uint32_t)ncells, (double *)rec);
    return check_launch("lint_grid_records");
}

template <typename DensT>
__global__ void cells_check_kernel(const DensT *c, uint64_t n, uint32_t *flags) {
    uint32_t bad = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const double v = (double)c[i];
        if (v < 0.0) bad |= LINT_FLAG_NEGATIVE;
        if (!isfinite(v)) bad |= LINT_FLAG_NONFINITE;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad && (threadIdx.x & 31) == 0) atomicOr(flags, bad);
}

// ---------------------------------------------------------------------------
// NumPy-identical pairwise sum                      [np.sum at grid.py:412,431]
// ---------------------------------------------------------------------------
// NumPy sums a contiguous fp64 array with a recursion that halves the range
// (left half rounded down to a multiple of 8) until a block has <= 128
// elements, which is summed with 8 interleaved accumulators.  The recursion
// tree depends only on n.  Every node at depth D (the deepest level at which
// all nodes still have > 128 elements) is evaluated by one thread with the
// scalar recursion; the complete binary tree above is then folded level by
// level, left + right, exactly as the recursion returns.
__global__ void pw_nodes_kernel(const double *x, uint64_t n, int depth, double *node_out) {
    const uint32_t nnodes = 1u << depth;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < nnodes; p += gridDim.x * blockDim.x) {
        uint64_t start = 0, len = n;
        for (int b = depth - 1; b >= 0; --b) {
            const uint64_t n2 = pw_left(len);
            if ((p >> b) & 1) { start += n2; len -= n2; } else { len = n2; }
        }
        node_out[p] = pw_node(x + start, len);
    }
}

__global__ void pw_fold_kernel(const double *in, uint32_t nout, double *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += gridDim.x * blockDim.x)
        out[i] = __dadd_rn(in[2 * i], in[2 * i + 1]);
}

static int sum_numpy_launch(const double *x, uint64_t n, double *out, double *scratch,
                            cudaStream_t st) {
    const int depth = pw_depth(n);
    uint32_t nnodes = 1u << depth;
    double *a = scratch, *b = scratch + nnodes;
    double *dst = depth == 0 ? out : a;
    pw_nodes_kernel<<<std::max(1u, std::min(nnodes / 128, 148u * 16)), 128, 0, st>>>(x, n, depth, dst);
    for (int lv = depth; lv > 0; --lv) {
        nnodes >>= 1;
        double *o = lv == 1 ? out : b;
        pw_fold_kernel<<<std::max(1u, std::min(nnodes / 256, 148u * 8)), 256, 0, st>>>(a, nnodes, o);
        std::swap(a, b);
    }
    return check_launch("lint_sum_numpy_f64");
}

// ---------------------------------------------------------------------------
// CDF, NumPy-sequential mode                [grid.py:431 + utils.py:195-196]
// ---------------------------------------------------------------------------
constexpr int SEQ_CHUNK = 1024;
constexpr int SEQ_THREADS = 256;

// One CTA.  Thread 0 carries the strictly sequential fp64 accumulation (that is
// what np.cumsum does) over chunk c-1 while warps 1.. stage p = m / total for
// chunk c into shared memory and stream chunk c-2 back out.
__global__ void __launch_bounds__(SEQ_THREADS)
cdf_sequential_kernel(const double *masses, uint64_t n, const double *total, double *cdf) {
    __shared__ double sin[2][SEQ_CHUNK];
    __shared__ double sout[2][SEQ_CHUNK];
    const double tot = *total;
    const uint64_t nchunks = (n + SEQ_CHUNK - 1) / SEQ_CHUNK;
    double acc = 0.0;
    for (uint64_t c = 0; c < nchunks + 2; ++c) {
        if (threadIdx.x >= 32) {
            const int t = threadIdx.x - 32;
            if (c < nchunks) {                           // stage chunk c
                const uint64_t base = c * SEQ_CHUNK;
                for (int i = t; i < SEQ_CHUNK; i += SEQ_THREADS - 32)
                    if (base + i < n) sin[c & 1][i] = __ddiv_rn(masses[base + i], tot);
            }
            if (c >= 2) {                                // drain chunk c-2
                const uint64_t base = (c - 2) * SEQ_CHUNK;
                for (int i = t; i < SEQ_CHUNK; i += SEQ_THREADS - 32)
                    if (base + i < n) cdf[base + i] = sout[c & 1][i];
            }
        } else if (threadIdx.x == 0 && c >= 1 && c - 1 < nchunks) {   // scan chunk c-1
            const uint64_t base = (c - 1) * SEQ_CHUNK;
            const int m = (int)min((uint64_t)SEQ_CHUNK, n - base);
            const double *in = sin[(c - 1) & 1];
            double *o = sout[(c - 1) & 1];
            int i = 0;
            if (c == 1) { acc = in[0]; o[0] = acc; i = 1; }   // cumsum starts from p[0]
#pragma unroll 8
            for (; i < m; ++i) { acc = __dadd_rn(acc, in[i]); o[i] = acc; }
        }
        __syncthreads();
    }
}

__global__ void cdf_normalise_kernel(double *cdf, uint64_t n) {
    const double last = cdf[n - 1];   // element n-1 is rewritten by a later launch
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i + 1 < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        cdf[i] = __ddiv_rn(cdf[i], last);
}
__global__ void cdf_set_last_kernel(double *cdf, uint64_t n) { cdf[n - 1] = 1.0; }

// ---------------------------------------------------------------------------
// CDF, parallel mode: deterministic, monotone, zero-preserving tiled scan
// ---------------------------------------------------------------------------
// S_i = E_tile (+) (E_warp (+) (E_lane (+) P_i)) where P_i is the sequential
// prefix inside a lane's run of PAR_ITEMS values and every exclusive offset is
// *exactly* the inclusive value of the element before it.  Because fp addition
// of non-negative numbers is monotone in each argument this gives a
// non-decreasing sequence with S_i == S_{i-1} whenever m_i == 0 (zero-mass
// cells stay unselectable, utils.py:197) and it is independent of scheduling,
// unlike a decoupled look-back whose partial-sum association varies run to run
// (the reference's determinism tests, tests/test_lintsampler.py:520-541, need
// bit-reproducible CDFs).
constexpr int PAR_THREADS = 256;
constexpr int PAR_ITEMS = 8;
constexpr int PAR_TILE = PAR_THREADS * PAR_ITEMS;

// Inclusive prefix of one tile in registers: x[] is replaced by
// E_warp (+) (E_oct (+) (E_lane (+) P_i)); returns the tile's inclusive total (all threads).
// The 32 lanes of a warp are nested in two levels — octets of lanes, then the four octets — so the
// sequential chains are 7 + 3 steps long instead of 31; at every level an exclusive offset is
// exactly the inclusive value (at that level) of the element before it, which is what keeps the
// sequence monotone and zero-mass cells equal to their predecessor.
__device__ __forceinline__ double tile_scan(double (&x)[PAR_ITEMS], double *warp_tot /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 1; j < PAR_ITEMS; ++j) x[j] = __dadd_rn(x[j - 1], x[j]);
    // chain across the lanes of an octet: lane l's exclusive offset is lane l-1's inclusive total
    double run = x[PAR_ITEMS - 1];      // becomes the inclusive total through this lane (within the octet)
    double excl = 0.0;
#pragma unroll
    for (int l = 1; l < 8; ++l) {
        const double prev = __shfl_sync(0xffffffffu, run, (lane & ~7) + l - 1);
        if ((lane & 7) == l) { excl = prev; run = __dadd_rn(prev, run); }
    }
    if (lane & 7) {
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) x[j] = __dadd_rn(excl, x[j]);
    }
    // chain across the four octets: an octet's total is its last lane's inclusive value
    double orun = __shfl_sync(0xffffffffu, run, lane | 7);
    double oexcl = 0.0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
        const double prev = __shfl_sync(0xffffffffu, orun, 8 * (q - 1));
        if ((lane >> 3) == q) { oexcl = prev; orun = __dadd_rn(prev, orun); }
    }
    if (lane >> 3) {
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) x[j] = __dadd_rn(oexcl, x[j]);
    }
    if (lane == 31) warp_tot[warp] = x[PAR_ITEMS - 1];
    __syncthreads();
    // chain across warps: warp_tot[w] is warp w's own inclusive total
    double wex = 0.0;
    for (int w = 0; w < warp; ++w) wex = (w == 0) ? warp_tot[0] : __dadd_rn(wex, warp_tot[w]);
    double tile_total = 0.0;
    for (int w = 0; w < PAR_THREADS / 32; ++w)
        tile_total = (w == 0) ? warp_tot[0] : __dadd_rn(tile_total, warp_tot[w]);
    if (warp > 0) {
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) x[j] = __dadd_rn(wex, x[j]);
    }
    __syncthreads();
    return tile_total;
}

__global__ void __launch_bounds__(PAR_THREADS)
scan_tile_totals_kernel(const double *m, uint64_t n, double *tile_tot) {
    __shared__ double warp_tot[PAR_THREADS / 32];
    for (uint64_t t = blockIdx.x; t * PAR_TILE < n; t += gridDim.x) {
        double x[PAR_ITEMS];
        const uint64_t base = t * PAR_TILE + (uint64_t)threadIdx.x * PAR_ITEMS;
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) x[j] = base + j < n ? m[base + j] : 0.0;
        const double tot = tile_scan(x, warp_tot);
        if (threadIdx.x == 0) tile_tot[t] = tot;
    }
}

// Offsets above the tile level keep the same nesting: tiles are chained inside
// groups of PAR_GROUP tiles, groups are chained by one thread.
//   S_i = G_group (+) (T_tile (+) x_i)
// tile_tot[] is replaced in place by T_tile; group_off[] receives G_group.
constexpr int PAR_GROUP = 64;

__global__ void __launch_bounds__(256)
scan_group_kernel(double *tile_tot, uint32_t ntiles, double *group_tot) {
    const uint32_t ngroups = (ntiles + PAR_GROUP - 1) / PAR_GROUP;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
        const uint32_t lo = g * PAR_GROUP, hi = min(lo + PAR_GROUP, ntiles);
        double run = 0.0;
        for (uint32_t t = lo; t < hi; ++t) {
            const double mine = tile_tot[t];
            tile_tot[t] = run;
            run = __dadd_rn(run, mine);          // 0.0 + x is exact for the first tile
        }
        group_tot[g] = run;
    }
}

__global__ void scan_group_chain_kernel(double *group_tot, uint32_t ngroups, double *total) {
    double run = 0.0;
    for (uint32_t g = 0; g < ngroups; ++g) {
        const double mine = group_tot[g];
        group_tot[g] = run;
        run = __dadd_rn(run, mine);
    }
    *total = run;
}

__global__ void __launch_bounds__(PAR_THREADS)
scan_apply_kernel(const double *m, uint64_t n, const double *tile_off, const double *group_off,
                  const double *total, double *cdf) {
    __shared__ double warp_tot[PAR_THREADS / 32];
    const double tot = *total;
    for (uint64_t t = blockIdx.x; t * PAR_TILE < n; t += gridDim.x) {
        double x[PAR_ITEMS];
        const uint64_t base = t * PAR_TILE + (uint64_t)threadIdx.x * PAR_ITEMS;
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) x[j] = base + j < n ? m[base + j] : 0.0;
        tile_scan(x, warp_tot);
        const double toff = tile_off[t], goff = group_off[t / PAR_GROUP];
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) {
            if (base + j < n) {
                const double s = __dadd_rn(goff, __dadd_rn(toff, x[j]));
                cdf[base + j] = __ddiv_rn(s, tot);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// guide table
// ---------------------------------------------------------------------------
// guide[j] = first i with cdf[i] > j/G (bit 31: the bin lies wholly inside that
// cell, i.e. guide[j+1] is the same cell).  Cell i owns the bins j with
// cdf[i-1] <= j/G < cdf[i], i.e. j in [ceil(cdf[i-1]*G), ceil(cdf[i]*G)).
// One warp walks 32 consecutive cells; short runs are written by the owning
// lane, long runs cooperatively.
__global__ void __launch_bounds__(256)
guide_build_kernel(const double *cdf, uint32_t n, int bits, uint32_t *guide, const uint32_t *range) {
    const double G = (double)(1u << bits);
    const uint32_t Gi = 1u << bits;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    // cells [c0, c1) only; the others own bins nobody asks for
    const uint32_t c0 = range ? range[0] : 0u, c1 = range ? range[1] : n;
    for (uint32_t w = (c0 >> 5) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); (uint64_t)w * 32 < c1; w += warps) {
        const uint32_t i = w * 32 + lane;
        uint32_t j0 = 0, j1 = 0;
        if (i >= c0 && i < c1) {
            const double lo = i == 0 ? 0.0 : cdf[i - 1];
            j0 = (uint32_t)min(ceil(lo * G), G);
            j1 = i == n - 1 ? Gi + 1 : (uint32_t)min(ceil(cdf[i] * G), G);   // last cell also owns j = G
        }
        const uint32_t len = j1 > j0 ? j1 - j0 : 0;
        // every bin but the cell's last lies wholly inside the cell: flag it
        if (len > 0 && len <= 32)
            for (uint32_t j = j0; j < j1; ++j) guide[j] = j + 1 < j1 ? (i | GUIDE_SAME) : i;
        uint32_t longmask = __ballot_sync(0xffffffffu, len > 32);
        while (longmask) {
            const int src = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            const uint32_t a = __shfl_sync(0xffffffffu, j0, src), b = __shfl_sync(0xffffffffu, j1, src);
            const uint32_t ci = w * 32 + src;
            for (uint32_t j = a + lane; j < b; j += 32) guide[j] = j + 1 < b ? (ci | GUIDE_SAME) : ci;
        }
    }
}

// ---------------------------------------------------------------------------
// hierarchical parallel build for grids: vertices -> group totals -> slab CDF
// ---------------------------------------------------------------------------
// The grid's cells are cut into GROUPS of consecutive cell rows (a row = the n_last cells along
// the last dimension), about 2 K cells each (whole rows: more where a row is longer).
//   1. Group totals the cheap way.  The trapezoid mass of a set of whole rows is a weighted sum
//      of vertex densities with separable weights: for every vertex row, Q = sum_i V[i] omega[i]
//      (omega = half the widths of the adjacent last-dim cells); the mass of a cell row is
//      (prod of its leading widths) / 2^(k-1) times the sum of the Q of its 2^(k-1) vertex rows.
//      One streaming pass over the vertex table (which also performs the validity check of
//      grid.py:401-404 on every vertex exactly once), then a few thousand adds.  Group offsets
//      Off[g] are the chained sums, Off[ngroups] is the total mass.
//   2. Inside a group the CDF is the exact chained scan of the cell masses (the arithmetic of
//      cell_mass_one; tiles of 2048 cells chained by the block that owns the group), laid on the
//      group's offset and kept below the next one:
//          S_i = min(Off[g] (+) L_i, Off[g + 1]),   cdf_i = S_i / Off[ngroups],   cdf[n - 1] = 1.
//      Monotone by construction; a zero-mass cell repeats the value before it (utils.py:197 never
//      selects it); bit-reproducible and independent of which rank or block computes it.
// Step 1 costs one read of the vertex table and is done by every rank of a sharded run — that
// is all a rank needs to know the total mass and where its stratum lies.  Step 2, the expensive
// one, runs only over the groups the rank's stratum [u_lo, u_hi) can touch (plus one group of
// margin on either side).  Cell masses are computed once and never stored.
constexpr uint32_t GROUP_TARGET_CELLS = PAR_TILE;      // one scan tile: a group is one block-step of the scan

struct GroupPlan {
    uint32_t n_last, nrow, nvrow, rows_per_group, ngroups, group_cells, sub;
    double u_lo, u_hi;
};

// Q[vr] = sum_i V[vr, i] * omega_last[i], one warp per vertex row, fixed summation order
constexpr int ROWSUM_BATCH = 8;
template <int K, typename DensT>
__global__ void __launch_bounds__(256)
vertex_rowsum_kernel(GridView<K> g, GroupPlan plan, double *__restrict__ Q, uint32_t *flags,
                     uint32_t *done_counter) {
    extern __shared__ double s_omega[];                            // n_last + 1 weights
    pdl_release();
    if (blockIdx.x == 0 && threadIdx.x == 0) *done_counter = 0;    // for group_totals_kernel, next in the stream
    const double *el = g.edges[K - 1];
    for (uint32_t i = threadIdx.x; i <= plan.n_last; i += blockDim.x) {
        const double wl = i > 0 ? __dsub_rn(__ldg(el + i), __ldg(el + i - 1)) : 0.0;
        const double wr = i < plan.n_last ? __dsub_rn(__ldg(el + i + 1), __ldg(el + i)) : 0.0;
        s_omega[i] = 0.5 * (wl + wr);
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5, nv = plan.n_last + 1;
    const DensT *V = reinterpret_cast<const DensT *>(g.V);
    uint32_t bad = 0;
    // Issue-bound like everything integer on this part (profiles/README.md r2a): five instructions
    // per element — load, min, convert, weight, fma.  No per-element predicates (full 32-wide slots
    // first, the leftover lanes once), and the validity check rides on the arithmetic: a negative
    // element drags the running minimum below zero, a NaN or an infinity makes the row sum
    // non-finite (finite fp32/fp64 densities times finite weights cannot overflow an fp64 sum).
    const uint32_t nfull = nv >> 5, rem = nv & 31u;
    for (uint32_t vr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; vr < plan.nvrow; vr += warps) {
        const DensT *row = V + (size_t)vr * nv + lane;
        const double *om = s_omega + lane;
        const DensT tail = lane < rem ? __ldg(row + nfull * 32) : DensT(0);    // in flight with the first batch
        double acc = 0.0;
        DensT vmin = DensT(0);
        uint32_t s = 0;
        for (; s + ROWSUM_BATCH <= nfull; s += ROWSUM_BATCH) {
            DensT v[ROWSUM_BATCH];
#pragma unroll
            for (int j = 0; j < ROWSUM_BATCH; ++j) v[j] = __ldg(row + (s + j) * 32);
#pragma unroll
            for (int j = 0; j < ROWSUM_BATCH; ++j) {
                vmin = v[j] < vmin ? v[j] : vmin;
                acc = fma((double)v[j], om[(s + j) * 32], acc);
            }
        }
        for (; s < nfull; ++s) {
            const DensT v = __ldg(row + s * 32);
            vmin = v < vmin ? v : vmin;
            acc = fma((double)v, om[s * 32], acc);
        }
        if (lane < rem) {
            vmin = tail < vmin ? tail : vmin;
            acc = fma((double)tail, om[nfull * 32], acc);
        }
        if (vmin < DensT(0)) bad |= LINT_FLAG_NEGATIVE;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (!isfinite(acc)) bad |= LINT_FLAG_NONFINITE;
        if (lane == 0) Q[vr] = acc;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad && lane == 0) atomicOr(flags, bad);
}

// Gt[g] = sum of the row totals of group g (one warp per group: lane-strided partial sums, then a
// fixed shuffle tree); the last block to finish chains them into Off[] and finds the slab of the
// stratum.  range_out: [0] first cell, [1] end cell, [2] first group, [3] end group.
template <int K>
__global__ void __launch_bounds__(256)
group_totals_kernel(GridView<K> g, GroupPlan plan, uint32_t ncells, const double *__restrict__ Q, double *Off,
                    double *total, uint32_t *done_counter, uint32_t *range_out) {
    __shared__ bool s_last;
    const uint32_t lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    pdl_release();
    pdl_wait();                                                    // Q, done_counter: vertex_rowsum_kernel
    // plan.sub lanes per group (a power of two >= the rows of a group, at most 32): 32 / sub groups
    // per warp step.  Lanes past the rows hold zero, so the sum is that of a full-warp tree.
    const uint32_t sub = plan.sub, per_warp = 32 / sub, sl = lane & (sub - 1);
    for (uint32_t g0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * per_warp; g0 < plan.ngroups;
         g0 += warps * per_warp) {
        const uint32_t gi = g0 + lane / sub;
        const uint32_t r0 = gi * plan.rows_per_group;
        const uint32_t r1 = gi < plan.ngroups ? min(r0 + plan.rows_per_group, plan.nrow) : r0;
        double sum = 0.0;
        for (uint32_t rho = r0 + sl; rho < r1; rho += sub) {
            uint32_t i[K];
            grid_unravel<K>(g, rho * plan.n_last, i);
            uint32_t vbase = 0;
            double pw = 1.0 / (double)(1u << (K - 1));
#pragma unroll
            for (int d = 0; d + 1 < K; ++d) {
                vbase += i[d] * g.vstride[d];
                pw *= __ldg(g.edges[d] + i[d] + 1) - __ldg(g.edges[d] + i[d]);
            }
            double q = 0.0;
#pragma unroll
            for (int h = 0; h < (1 << (K - 1)); ++h) q += Q[(vbase + g.corner_off[2 * h]) / (plan.n_last + 1)];
            sum += pw * q;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            if ((uint32_t)d < sub) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (sl == 0 && gi < plan.ngroups) Off[gi + 1] = sum;       // chained below
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // Off[g + 1] <- Gt[0] (+) ... (+) Gt[g]: chunks of 2048 totals scanned with the nested exact
    // chains of tile_scan (an offset is always the inclusive value before it: non-decreasing),
    // chunk after chunk
    volatile double *vo = Off;
    __shared__ double warp_tot[PAR_THREADS / 32];
    double carry = 0.0;
    double x[PAR_ITEMS], nxt[PAR_ITEMS];
#pragma unroll
    for (int j = 0; j < PAR_ITEMS; ++j) {
        const uint32_t i = threadIdx.x * PAR_ITEMS + j;
        nxt[j] = i < plan.ngroups ? vo[i + 1] : 0.0;
    }
    for (uint32_t c0 = 0; c0 < plan.ngroups; c0 += PAR_TILE) {
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) {
            x[j] = nxt[j];
            const uint32_t i = c0 + PAR_TILE + threadIdx.x * PAR_ITEMS + j;        // the next chunk, in flight
            nxt[j] = i < plan.ngroups ? vo[i + 1] : 0.0;                           // during this one's scan
        }
        const double chunk_total = tile_scan(x, warp_tot);
#pragma unroll
        for (int j = 0; j < PAR_ITEMS; ++j) {
            const uint32_t i = c0 + threadIdx.x * PAR_ITEMS + j;
            if (i < plan.ngroups) vo[i + 1] = c0 == 0 ? x[j] : __dadd_rn(carry, x[j]);
        }
        carry = c0 == 0 ? chunk_total : __dadd_rn(carry, chunk_total);
    }
    __threadfence();
    __syncthreads();
    const double run = carry;
    // groups the stratum can touch: Off[g + 1] / total > u_lo and Off[g] / total < u_hi.  Off is
    // non-decreasing, so the two bracketing groups are counts (every thread tests a few entries):
    //   lo = last g in [0, ngroups) with Off[g] / total <= u_lo   (Off[0] = 0 always passes)
    //   a  = first g in [0, ngroups] with Off[g] / total >= u_hi
    __shared__ uint32_t s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t n_le = 0, n_lt = 0;
    // (Off[g] <= u total instead of Off[g] / total <= u: one product instead of a division per entry;
    // the two can differ for an entry within an ulp of a stratum boundary, which the group of margin
    // on either side of the slab absorbs)
    const double t_lo = plan.u_lo * run, t_hi = plan.u_hi * run;
    for (uint32_t base = threadIdx.x; base < plan.ngroups; base += blockDim.x * 8) {
        double v[8];                                               // eight L2 loads in flight, not one
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t gi = base + j * blockDim.x;
            v[j] = gi > 0 && gi < plan.ngroups ? __ldcg(Off + gi) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + j * blockDim.x < plan.ngroups) {
                n_le += v[j] <= t_lo;
                n_lt += v[j] < t_hi;
            }
        }
    }
    n_le = __reduce_add_sync(0xffffffffu, n_le);
    n_lt = __reduce_add_sync(0xffffffffu, n_lt);
    if (lane == 0) {
        atomicAdd(&s_cnt[0], n_le);
        atomicAdd(&s_cnt[1], n_lt);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    vo[0] = 0.0;
    *total = run;
    *done_counter = 0;
    const uint32_t lo = s_cnt[0] > 0 ? s_cnt[0] - 1 : 0, a = s_cnt[1];
    uint32_t g_lo = lo > 0 ? lo - 1 : 0, g_hi = min(a + 1, plan.ngroups);      // one group of margin
    if (plan.u_lo <= 0.0) g_lo = 0;                                // a stratum that starts / ends with the domain
    if (plan.u_hi >= 1.0) g_hi = plan.ngroups;                     // covers its zero-mass head / tail too
    range_out[0] = g_lo * plan.group_cells;
    range_out[1] = (uint32_t)min((uint64_t)g_hi * plan.group_cells, (uint64_t)ncells);
    range_out[2] = g_lo;
    range_out[3] = g_hi;
}

// One block per group of the slab: exact chained scan of the group's cell masses -> cdf.
// Thread t computes the masses of cells t, t + 256, ... of a tile (consecutive lanes, consecutive
// cells: every corner load is coalesced); the tile is transposed through shared memory so that a
// thread scans eight consecutive cells.
constexpr int SCAN_PAD = PAR_ITEMS + 1;         // 9 doubles per thread's run: conflict-free 64-bit reads
#ifndef LINT_GSCAN_MINB
#define LINT_GSCAN_MINB 6        // 40 registers: 1536 threads per SM (measured best of 4 / 6 / 8)
#endif
template <int K, typename DensT>
__global__ void __launch_bounds__(PAR_THREADS, LINT_GSCAN_MINB)
group_scan_kernel(GridView<K> g, GroupPlan plan, uint32_t ncells, const double *__restrict__ Off,
                  const double *__restrict__ total, const uint32_t *__restrict__ range, double *cdf,
                  double *masses) {
    __shared__ double warp_tot[PAR_THREADS / 32];
    __shared__ double s_m[PAR_THREADS * SCAN_PAD];
    pdl_release();
    pdl_wait();                                                    // Off, total, range: group_totals_kernel
    const double inv_tot = 1.0 / *total;
    const uint32_t g_lo = range[2], g_hi = range[3];
    uint32_t unused = 0;
    for (uint32_t gi = g_lo + blockIdx.x; gi < g_hi; gi += gridDim.x) {
        const uint32_t c_begin = gi * plan.group_cells;
        const uint32_t c_end = (uint32_t)min((uint64_t)c_begin + plan.group_cells, (uint64_t)ncells);
        const double off = Off[gi], off_next = Off[gi + 1];
        double run = 0.0;                                          // chained total of the tiles before this one
        for (uint32_t tb = c_begin; tb < c_end; tb += PAR_TILE) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PAR_ITEMS; ++j) {
                const uint32_t k = threadIdx.x + j * PAR_THREADS;              // cell tb + k of the tile
                const double mk = tb + k < c_end ? cell_mass_one<K, DensT, false>(g, tb + k, unused) : 0.0;
                s_m[(k / PAR_ITEMS) * SCAN_PAD + (k % PAR_ITEMS)] = mk;
                if (masses && tb + k < c_end) masses[tb + k] = mk;
            }
            __syncthreads();
            double x[PAR_ITEMS];
#pragma unroll
            for (int j = 0; j < PAR_ITEMS; ++j) x[j] = s_m[threadIdx.x * SCAN_PAD + j];
            const uint32_t base = tb + threadIdx.x * PAR_ITEMS;
            const double tile_total = tile_scan(x, warp_tot);
            // back through shared memory: a thread holds eight consecutive entries, and eight-byte stores
            // at a 64-byte stride reach L2 as partial sectors; transposed, a warp stores 256 contiguous bytes
#pragma unroll
            for (int j = 0; j < PAR_ITEMS; ++j) {
                const double s = fmin(__dadd_rn(off, tb == c_begin ? x[j] : __dadd_rn(run, x[j])), off_next);
                // (a product, not a division: twenty-odd instructions per cell less; monotone all the same)
                s_m[threadIdx.x * SCAN_PAD + j] = base + j == ncells - 1 ? 1.0 : fmin(__dmul_rn(s, inv_tot), 1.0);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PAR_ITEMS; ++j) {
                const uint32_t k = threadIdx.x + j * PAR_THREADS;
                if (tb + k < c_end) cdf[tb + k] = s_m[(k / PAR_ITEMS) * SCAN_PAD + (k % PAR_ITEMS)];
            }
            run = tb == c_begin ? tile_total : __dadd_rn(run, tile_total);
        }
        // the entry just before the slab: readers difference cdf[c] - cdf[c - 1]
        if (gi == g_lo && gi > 0 && threadIdx.x == 0) cdf[c_begin - 1] = fmin(__dmul_rn(off, inv_tot), 1.0);
    }
}

static GroupPlan make_group_plan(const lint_grid_t *g, uint64_t ncells, double u_lo, double u_hi) {
    GroupPlan p;
    const int K = g->dim;
    p.n_last = (uint32_t)g->ncells[K - 1];
    p.nrow = (uint32_t)(ncells / p.n_last);
    uint64_t nv = 1;
    for (int d = 0; d + 1 < K; ++d) nv *= (uint64_t)g->ncells[d] + 1;
    p.nvrow = (uint32_t)nv;
    p.rows_per_group = std::max<uint32_t>(1, GROUP_TARGET_CELLS / p.n_last);
    p.ngroups = (p.nrow + p.rows_per_group - 1) / p.rows_per_group;
    p.group_cells = p.rows_per_group * p.n_last;
    p.sub = 1;
    while (p.sub < 32 && p.sub < p.rows_per_group) p.sub <<= 1;
    p.u_lo = u_lo;
    p.u_hi = u_hi;
    return p;
}

// Small guide tables (far fewer bins than cells of the slab): one thread per bin searches the
// slab for the first cell with cdf > j / G instead of every cell looking for the bins it owns.
// Same table, same flags; only bins whose answer lies in the slab [c0, c1) are written.
__global__ void __launch_bounds__(256)
guide_search_kernel(const double *__restrict__ cdf, uint32_t n, int bits, uint32_t *guide,
                    const uint32_t *__restrict__ range) {
    const uint32_t Gi = 1u << bits;
    const double G = (double)Gi;
    const uint32_t c0 = range ? range[0] : 0u, c1 = range ? range[1] : n;
    // bins a slab cell can own: ceil(cdf[c0 - 1] G) <= j < ceil(cdf[c1 - 1] G)  (the last cell of the
    // table also owns j = G)
    const double before = c0 > 0 ? __ldg(cdf + c0 - 1) : 0.0, last = __ldg(cdf + c1 - 1);
    const uint32_t j_lo = (uint32_t)fmin(ceil(before * G), G);
    const uint32_t j_hi = c1 == n ? Gi : (uint32_t)fmin(ceil(last * G), G);           // inclusive, generous
    for (uint32_t j = j_lo + blockIdx.x * blockDim.x + threadIdx.x; j <= j_hi; j += gridDim.x * blockDim.x) {
        const double u = (double)j / G;
        uint32_t lo = c0, hi = c1;                               // first i in [c0, c1) with cdf[i] > j / G, or c1
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
        }
        uint32_t i = lo;
        if (i >= c1) {
            if (c1 != n) continue;                               // the owner lies beyond the slab: not ours
            i = n - 1;                                           // j = G: the last cell
        }
        // flagged when the next bin has the same owner: the whole bin lies inside cell i
        const bool same = i == n - 1 ? j < Gi : __ldg(cdf + i) > (double)(j + 1) / G;
        guide[j] = same ? (i | GUIDE_SAME) : i;
    }
}

template <int K>
int group_build_impl(const lint_grid_t *g, uint64_t ncells, double u_lo, double u_hi, double *cdf, double *total,
                     uint32_t *guide, int guide_bits, double *masses, uint32_t *flags, uint32_t *range,
                     void *scratch, cudaStream_t st) {
    GridView<K> v = make_grid_view<K>(g, ncells, false);
    v.rec = nullptr;
    const GroupPlan plan = make_group_plan(g, ncells, u_lo, u_hi);
    double *Q = (double *)scratch;
    double *Off = Q + plan.nvrow;
    uint32_t *done = (uint32_t *)(Off + plan.ngroups + 1);
    const size_t smem = ((size_t)plan.n_last + 1) * sizeof(double);
    if (smem > 96 * 1024) return fail(LINT_ERR_UNSUPPORTED, "lint_grid_build: more than 12287 cells along the last dimension");
    const int qblocks = (int)std::min<uint64_t>(((uint64_t)plan.nvrow + 7) / 8, 148 * 8);
    const int sblocks = (int)std::min<uint32_t>(plan.ngroups, 148 * 8);
    const uint32_t per_block = 8 * (32 / plan.sub);               // groups per block step
    const int tblocks = (int)std::min<uint32_t>((plan.ngroups + per_block - 1) / per_block, 148 * 4);
    auto run = [&](auto tag) {
        using D = decltype(tag);
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(vertex_rowsum_kernel<K, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        vertex_rowsum_kernel<K, D><<<qblocks, 256, smem, st>>>(v, plan, Q, flags, done);
        launch_chained(group_totals_kernel<K>, tblocks, 256, 0, st, v, plan, (uint32_t)ncells, (const double *)Q, Off,
                       total, done, range);
        launch_chained(group_scan_kernel<K, D>, sblocks, PAR_THREADS, 0, st, v, plan, (uint32_t)ncells,
                       (const double *)Off, (const double *)total, (const uint32_t *)range, cdf, masses);
    };
    if (g->dtype == LINT_F32) run(float{}); else run(double{});
    if (guide_bits > 0 && (1ull << guide_bits) * 16 <= ncells) {
        const int gblocks = (int)std::min<uint64_t>(((1ull << guide_bits) + 256) / 256, 148 * 8);
        guide_search_kernel<<<gblocks, 256, 0, st>>>(cdf, (uint32_t)ncells, guide_bits, guide, range);
    } else if (guide_bits > 0) {
        const uint64_t nwarps = (ncells + 31) / 32;
        const int gblocks = (int)std::min<uint64_t>((nwarps + 7) / 8, 148 * 16);
        guide_build_kernel<<<gblocks, 256, 0, st>>>(cdf, (uint32_t)ncells, guide_bits, guide, range);
    }
    return check_launch("lint_grid_build");
}

}  // namespace lint

using namespace lint;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int lint_abi_version(void) { return LINT_ABI_VERSION; }

int lint_set_offset_counter(const uint64_t *counter_dev) {
    g_offset_counter = counter_dev;
    return LINT_OK;
}

int lint_advance_counter(uint64_t *counter_dev, uint64_t by, lint_stream_t stream) {
    if (!counter_dev) return fail(LINT_ERR_BAD_ARG, "lint_advance_counter: counter_dev is NULL");
    advance_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, by);
    return check_launch("lint_advance_counter");
}
const char *lint_last_error(void) { return last_error_slot().c_str(); }

int lint_grid_eval_gmm(const lint_grid_t *g, const lint_gmm_t *gmm, double scale,
                       lint_stream_t stream) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, false, &nc, &nv)) return rc;
    if (!gmm || gmm->dim != g->dim) return fail(LINT_ERR_BAD_ARG, "gmm dim does not match grid dim");
    if (gmm->ncomp < 1 || gmm->ncomp > LINT_GMM_MAX_COMP)
        return fail(LINT_ERR_BAD_ARG, "gmm ncomp %d outside 1..%d", gmm->ncomp, LINT_GMM_MAX_COMP);
    if (!gmm->log_norm || !gmm->mean || !gmm->prec_chol)
        return fail(LINT_ERR_BAD_ARG, "gmm parameter arrays must not be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    switch (g->dim) {
        case 1: return eval_gmm_impl<1>(g, gmm, scale, nv, st);
        case 2: return eval_gmm_impl<2>(g, gmm, scale, nv, st);
        case 3: return eval_gmm_impl<3>(g, gmm, scale, nv, st);
        case 4: return eval_gmm_impl<4>(g, gmm, scale, nv, st);
        case 5: return eval_gmm_impl<5>(g, gmm, scale, nv, st);
        default: return eval_gmm_impl<6>(g, gmm, scale, nv, st);
    }
}

int lint_grid_eval_halos(const lint_grid_t *g, const lint_halos_t *h, double scale,
                         lint_stream_t stream) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, false, &nc, &nv)) return rc;
    if (!h || h->dim != g->dim) return fail(LINT_ERR_BAD_ARG, "halo dim does not match grid dim");
    if (h->ncomp < 1 || h->ncomp > LINT_GMM_MAX_COMP)
        return fail(LINT_ERR_BAD_ARG, "halo ncomp %d outside 1..%d", h->ncomp, LINT_GMM_MAX_COMP);
    if (!h->rho0 || !h->centre || !h->r_s)
        return fail(LINT_ERR_BAD_ARG, "halo parameter arrays must not be NULL");
    if (!(h->alpha > 0.0) || !(h->eps >= 0.0))
        return fail(LINT_ERR_BAD_ARG, "halo profile needs alpha > 0 and eps >= 0");
    for (int j = 0; j < h->ncomp; ++j)
        if (!(h->r_s[j] > 0.0)) return fail(LINT_ERR_BAD_ARG, "halo scale radius %d is not positive", j);
    cudaStream_t st = (cudaStream_t)stream;
    switch (g->dim) {
        case 1: return eval_halos_impl<1>(g, h, scale, nv, st);
        case 2: return eval_halos_impl<2>(g, h, scale, nv, st);
        case 3: return eval_halos_impl<3>(g, h, scale, nv, st);
        case 4: return eval_halos_impl<4>(g, h, scale, nv, st);
        case 5: return eval_halos_impl<5>(g, h, scale, nv, st);
        default: return eval_halos_impl<6>(g, h, scale, nv, st);
    }
}

int lint_grid_masses(const lint_grid_t *g, double *masses_dev, uint32_t *flags_dev,
                     lint_stream_t stream) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, false, &nc, &nv)) return rc;
    if (!masses_dev || !flags_dev) return fail(LINT_ERR_BAD_ARG, "masses_dev / flags_dev is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    switch (g->dim) {
        case 1: return masses_impl<1>(g, nc, masses_dev, flags_dev, st);
        case 2: return masses_impl<2>(g, nc, masses_dev, flags_dev, st);
        case 3: return masses_impl<3>(g, nc, masses_dev, flags_dev, st);
        case 4: return masses_impl<4>(g, nc, masses_dev, flags_dev, st);
        case 5: return masses_impl<5>(g, nc, masses_dev, flags_dev, st);
        default: return masses_impl<6>(g, nc, masses_dev, flags_dev, st);
    }
}

int lint_grid_records(const lint_grid_t *g, void *records_dev, lint_stream_t stream) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, false, &nc, &nv)) return rc;
    if (!records_dev) return fail(LINT_ERR_BAD_ARG, "lint_grid_records: records_dev is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    switch (g->dim) {
        case 1: return records_impl<1>(g, nc, records_dev, st);
        case 2: return records_impl<2>(g, nc, records_dev, st);
        case 3: return records_impl<3>(g, nc, records_dev, st);
        case 4: return records_impl<4>(g, nc, records_dev, st);
        case 5: return records_impl<5>(g, nc, records_dev, st);
        default: return records_impl<6>(g, nc, records_dev, st);
    }
}

int lint_cells_check(const lint_cells_t *c, uint32_t *flags_dev, lint_stream_t stream) {
    if (!c || !c->corners_dev || !flags_dev) return fail(LINT_ERR_BAD_ARG, "NULL argument");
    if (c->dim < 1 || c->dim > LINT_MAX_DIM || c->ncells < 1)
        return fail(LINT_ERR_BAD_ARG, "bad cell-list shape");
    const uint64_t n = (uint64_t)c->ncells << c->dim;
    const int blocks = (int)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->dtype == LINT_F32)
        cells_check_kernel<float><<<blocks, 256, 0, st>>>((const float *)c->corners_dev, n, flags_dev);
    else
        cells_check_kernel<double><<<blocks, 256, 0, st>>>((const double *)c->corners_dev, n, flags_dev);
    return check_launch("lint_cells_check");
}

size_t lint_cdf_scratch_bytes(int64_t n) {
    if (n < 1) return 0;
    // pairwise-sum node buffers (2 * 2^depth <= 2 * n/129 + 2) and parallel tile totals
    const size_t nodes = 2 * ((size_t)1 << pw_depth((uint64_t)n));
    const size_t tiles = ((size_t)n + PAR_TILE - 1) / PAR_TILE;
    const size_t groups = (tiles + PAR_GROUP - 1) / PAR_GROUP;
    return (nodes + tiles + groups + 8) * sizeof(double);
}

int lint_sum_numpy_f64(const double *x_dev, int64_t n, double *out_dev, void *scratch_dev,
                       size_t scratch_bytes, lint_stream_t stream) {
    if (!x_dev || !out_dev || n < 1) return fail(LINT_ERR_BAD_ARG, "lint_sum_numpy_f64: bad argument");
    if (scratch_bytes < lint_cdf_scratch_bytes(n) || !scratch_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_sum_numpy_f64: scratch too small");
    return sum_numpy_launch(x_dev, (uint64_t)n, out_dev, (double *)scratch_dev, (cudaStream_t)stream);
}

int lint_cdf_build(const double *masses_dev, int64_t n, int mode, double *cdf_dev,
                   double *total_dev, void *scratch_dev, size_t scratch_bytes,
                   lint_stream_t stream) {
    if (!masses_dev || !cdf_dev || !total_dev || n < 1)
        return fail(LINT_ERR_BAD_ARG, "lint_cdf_build: bad argument");
    if (n >= (1ll << 31)) return fail(LINT_ERR_UNSUPPORTED, "lint_cdf_build: n >= 2^31");
    if (scratch_bytes < lint_cdf_scratch_bytes(n) || !scratch_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_cdf_build: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t un = (uint64_t)n;
    if (mode == LINT_CDF_SEQUENTIAL) {
        if (int rc = sum_numpy_launch(masses_dev, un, total_dev, (double *)scratch_dev, st)) return rc;
        cdf_sequential_kernel<<<1, SEQ_THREADS, 0, st>>>(masses_dev, un, total_dev, cdf_dev);
        // `cdf /= cdf[-1]`: the divisor must be read before any element changes, so the
        // last element is rewritten by a separate launch after the others
        const int blocks = (int)std::min<uint64_t>((un + 255) / 256, 148 * 16);
        cdf_normalise_kernel<<<blocks, 256, 0, st>>>(cdf_dev, un);
        cdf_set_last_kernel<<<1, 1, 0, st>>>(cdf_dev, un);
        return check_launch("lint_cdf_build(sequential)");
    }
    if (mode != LINT_CDF_PARALLEL) return fail(LINT_ERR_BAD_ARG, "lint_cdf_build: unknown mode %d", mode);
    double *tile_tot = (double *)scratch_dev;
    const uint32_t ntiles = (uint32_t)((un + PAR_TILE - 1) / PAR_TILE);
    const int blocks = (int)std::min<uint32_t>(ntiles, 148 * 8);
    scan_tile_totals_kernel<<<blocks, PAR_THREADS, 0, st>>>(masses_dev, un, tile_tot);
    const uint32_t ngroups = (ntiles + PAR_GROUP - 1) / PAR_GROUP;
    double *group_off = tile_tot + ntiles;
    scan_group_kernel<<<(ngroups + 255) / 256, 256, 0, st>>>(tile_tot, ntiles, group_off);
    scan_group_chain_kernel<<<1, 1, 0, st>>>(group_off, ngroups, total_dev);
    scan_apply_kernel<<<blocks, PAR_THREADS, 0, st>>>(masses_dev, un, tile_tot, group_off, total_dev,
                                                      cdf_dev);
    return check_launch("lint_cdf_build(parallel)");
}

int lint_guide_build(const double *cdf_dev, int64_t n, int guide_bits, uint32_t *guide_dev,
                     lint_stream_t stream) {
    if (!cdf_dev || !guide_dev || n < 1 || n >= (1ll << 31))
        return fail(LINT_ERR_BAD_ARG, "lint_guide_build: bad argument");
    if (guide_bits < 1 || guide_bits > 30)
        return fail(LINT_ERR_BAD_ARG, "lint_guide_build: guide_bits %d outside 1..30", guide_bits);
    const uint64_t nwarps = ((uint64_t)n + 31) / 32;
    const int blocks = (int)std::min<uint64_t>((nwarps + 7) / 8, 148 * 16);
    guide_build_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cdf_dev, (uint32_t)n, guide_bits, guide_dev,
                                                                 nullptr);
    return check_launch("lint_guide_build");
}

int64_t lint_grid_group_cells(const lint_grid_t *g) {
    uint64_t nc, nv;
    if (validate_grid(g, false, &nc, &nv)) return 0;
    return (int64_t)make_group_plan(g, nc, 0.0, 1.0).group_cells;
}

size_t lint_grid_build_scratch_bytes(const lint_grid_t *g) {
    uint64_t nc, nv;
    if (validate_grid(g, false, &nc, &nv)) return 0;
    const GroupPlan p = make_group_plan(g, nc, 0.0, 1.0);
    return ((size_t)p.nvrow + p.ngroups + 3) * sizeof(double);
}

int lint_grid_build(const lint_grid_t *g, double u_lo, double u_hi, double *cdf_dev, double *total_dev,
                    uint32_t *guide_dev, int guide_bits, double *masses_dev, uint32_t *flags_dev,
                    uint32_t *range_dev, void *scratch_dev, size_t scratch_bytes, lint_stream_t stream) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, false, &nc, &nv)) return rc;
    if (!cdf_dev || !total_dev || !flags_dev || !range_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_grid_build: cdf_dev / total_dev / flags_dev / range_dev is NULL");
    if (!(u_lo >= 0.0) || !(u_hi <= 1.0) || !(u_lo < u_hi))
        return fail(LINT_ERR_BAD_ARG, "lint_grid_build: stratum [%g, %g) is not inside [0, 1]", u_lo, u_hi);
    if (guide_bits < 0 || guide_bits > 30 || (guide_bits > 0 && !guide_dev))
        return fail(LINT_ERR_BAD_ARG, "lint_grid_build: guide table inconsistent (bits=%d)", guide_bits);
    if (!scratch_dev || scratch_bytes < lint_grid_build_scratch_bytes(g))
        return fail(LINT_ERR_BAD_ARG, "lint_grid_build: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    switch (g->dim) {
        case 1: return group_build_impl<1>(g, nc, u_lo, u_hi, cdf_dev, total_dev, guide_dev, guide_bits, masses_dev, flags_dev, range_dev, scratch_dev, st);
        case 2: return group_build_impl<2>(g, nc, u_lo, u_hi, cdf_dev, total_dev, guide_dev, guide_bits, masses_dev, flags_dev, range_dev, scratch_dev, st);
        case 3: return group_build_impl<3>(g, nc, u_lo, u_hi, cdf_dev, total_dev, guide_dev, guide_bits, masses_dev, flags_dev, range_dev, scratch_dev, st);
        case 4: return group_build_impl<4>(g, nc, u_lo, u_hi, cdf_dev, total_dev, guide_dev, guide_bits, masses_dev, flags_dev, range_dev, scratch_dev, st);
        case 5: return group_build_impl<5>(g, nc, u_lo, u_hi, cdf_dev, total_dev, guide_dev, guide_bits, masses_dev, flags_dev, range_dev, scratch_dev, st);
        default: return group_build_impl<6>(g, nc, u_lo, u_hi, cdf_dev, total_dev, guide_dev, guide_bits, masses_dev, flags_dev, range_dev, scratch_dev, st);
    }
}

}  // extern "C"
