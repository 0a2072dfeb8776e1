// Sampling kernels: cell selection (searchsorted on the fp64 CDF), corner
// gather, k-D multilinear inverse CDF, affine map — one fused kernel per
// (cell source, uniform source, dim, precision).  C ABI in
// include/lintsampler_b200.h.
//
// Reference citations are relative to /root/reference/src/lintsampler.
#include <algorithm>
#include "lint_device.cuh"

extern "C" size_t lint_multi_scratch_bytes(int dim, int ndomains);

namespace lint {

int validate_grid(const lint_grid_t *g, bool need_cdf, uint64_t *ncells_out, uint64_t *nverts_out);
template <int K> GridView<K> make_grid_view(const lint_grid_t *g, uint64_t ncells, bool stage_edges);

// ---------------------------------------------------------------------------
// cell-source adaptors
// ---------------------------------------------------------------------------
template <int K, typename DensT>
struct GridSource {
    GridView<K> g;
    __host__ size_t smem_bytes(size_t elem) const { return (size_t)g.edge_total * elem; }
    __device__ __forceinline__ size_t smem_bytes_dev(size_t elem) const { return (size_t)g.edge_total * elem; }
    template <typename T>
    __device__ __forceinline__ void prepare(T *smem) const { grid_stage_edges<K, T>(g, smem); }
    __device__ __forceinline__ bool prenormalised() const { return g.rec != nullptr; }
    __device__ __forceinline__ void bracket(double &u, uint32_t &lo, uint32_t &hi) const {
        if (g.table.n == 1) { lo = hi = 0; return; }             // grid.py:427-428 shortcut
        guide_bracket(g.table, u, lo, hi);
    }
    __device__ __forceinline__ uint32_t resolve(double u, uint32_t lo, uint32_t hi) const {
        return search_bracket(g.table, u, lo, hi);
    }
    template <typename T>
    __device__ __forceinline__ void fetch(const T *smem, uint32_t idx, T (&c)[1 << K], T (&lo)[K],
                                          T (&hi)[K]) const {
        grid_fetch<K, DensT, T>(g, smem, idx, c, lo, hi);
    }
};

template <int K, typename DensT>
struct ListSource {
    CellsView s;
    __host__ size_t smem_bytes(size_t) const { return 0; }
    __device__ __forceinline__ size_t smem_bytes_dev(size_t) const { return 0; }
    template <typename T>
    __device__ __forceinline__ void prepare(T *) const {}
    __device__ __forceinline__ bool prenormalised() const { return s.rec != nullptr; }
    __device__ __forceinline__ void bracket(double &u, uint32_t &lo, uint32_t &hi) const {
        guide_bracket(s.table, u, lo, hi);
    }
    __device__ __forceinline__ uint32_t resolve(double u, uint32_t lo, uint32_t hi) const {
        return search_bracket(s.table, u, lo, hi);
    }
    template <typename T>
    __device__ __forceinline__ void fetch(const T *, uint32_t idx, T (&c)[1 << K], T (&lo)[K],
                                          T (&hi)[K]) const {
        cells_fetch<K, DensT, T>(s, idx, c, lo, hi);
    }
};

// A list of grids as ONE domain  [lintsampler.py:281-307]: the cell-selecting uniform first picks
// the grid (`_choice` on the normalised cumulative grid masses, :286-290), is then rescaled into
// that grid's interval exactly as the reference does, (u - start) / (end - start) (:306), and
// finally picks the cell inside the grid.  Cells are numbered through the concatenation of the
// grids (cell_off), which is also what cell_idx reports.
constexpr int LINT_MULTI_MAX = 16;
template <int K, typename DensT>
struct MultiGridSource {
    const GridView<K> *views;              // device table, one view per grid
    int nd;
    double choice_cdf[LINT_MULTI_MAX];     // c / c[-1], c = p.cumsum()          (utils.py:195-196)
    double starts[LINT_MULTI_MAX], ends[LINT_MULTI_MAX];      // np.append(0, p.cumsum())   (:293-297)
    uint32_t cell_off[LINT_MULTI_MAX + 1];
    bool has_records;
    __host__ size_t smem_bytes(size_t) const { return 0; }
    __device__ __forceinline__ size_t smem_bytes_dev(size_t) const { return 0; }
    template <typename T>
    __device__ __forceinline__ void prepare(T *) const {}
    __device__ __forceinline__ bool prenormalised() const { return has_records; }
    __device__ __forceinline__ int domain_of(uint32_t cell) const {
        int d = 0;
        while (d + 1 < nd && cell >= cell_off[d + 1]) ++d;
        return d;
    }
    __device__ __forceinline__ void bracket(double &u, uint32_t &lo, uint32_t &hi) const {
        int d = 0;
        while (d + 1 < nd && !(choice_cdf[d] > u)) ++d;           // searchsorted(side='right')
        u = __ddiv_rn(__dsub_rn(u, starts[d]), __dsub_rn(ends[d], starts[d]));
        const GridView<K> &g = views[d];
        if (g.table.n == 1) { lo = hi = cell_off[d]; return; }
        guide_bracket(g.table, u, lo, hi);
        lo += cell_off[d];
        hi += cell_off[d];
    }
    __device__ __forceinline__ uint32_t resolve(double u, uint32_t lo, uint32_t hi) const {
        const int d = domain_of(lo);
        return cell_off[d] + search_bracket(views[d].table, u, lo - cell_off[d], hi - cell_off[d]);
    }
    template <typename T>
    __device__ __forceinline__ void fetch(const T *, uint32_t idx, T (&c)[1 << K], T (&lo)[K], T (&hi)[K]) const {
        const int d = domain_of(idx);
        grid_fetch<K, DensT, T>(views[d], (const T *)nullptr, idx - cell_off[d], c, lo, hi);
    }
};

// ---------------------------------------------------------------------------
// the fused per-sample kernel   [sampling.py:97-127]
// ---------------------------------------------------------------------------
constexpr int SAMPLE_THREADS = 256;

// Rows per thread per iteration.  The chain guide -> CDF window -> cell record is
// three dependent memory round trips per row; issuing it for several independent
// rows at once is what hides the latency (the tables are L2/HBM resident and
// randomly addressed, so there is no locality to exploit instead).
template <typename T> struct RowsPerThread { static constexpr int value = 1; };

template <int K, typename T, typename Source, typename Uniforms>
__global__ void __launch_bounds__(SAMPLE_THREADS)
sample_kernel(Source src, Uniforms unif, int64_t N, T *__restrict__ out,
              int64_t *__restrict__ cell_idx, const uint32_t *__restrict__ first_row_dev) {
    constexpr int R = RowsPerThread<T>::value;
    // rows [first, N): first = 0 normally; the cell-major path tops up its tail
    // from a device-resident row count
    const int64_t first = first_row_dev ? (int64_t)*first_row_dev : 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    src.prepare(smem);                                   // edge arrays
    const uint32_t *utab = unif.stage(                   // uniform source's tables, after the edges
        reinterpret_cast<uint32_t *>(smem_raw + ((src.smem_bytes_dev(sizeof(T)) + 15) & ~(size_t)15)));
    const int64_t tile = (int64_t)SAMPLE_THREADS * R;
    for (int64_t base = first + (int64_t)blockIdx.x * tile; base < N; base += (int64_t)gridDim.x * tile) {
        double ucell[R];
        T u[R][K];
        uint32_t lo[R], hi[R];
        bool live[R];
        // stage 1: uniforms and guide brackets for all rows (independent loads in flight)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = base + (int64_t)r * SAMPLE_THREADS + threadIdx.x;
            live[r] = row < N;
            unif.get(live[r] ? row : 0, ucell[r], u[r], utab);
            src.bracket(ucell[r], lo[r], hi[r]);           // grid.choose_cells(u[..., -1])
        }
        // stage 2: resolve the brackets that span more than one cell
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (lo[r] < hi[r]) lo[r] = src.resolve(ucell[r], lo[r], hi[r]);
        // stage 3: cell boxes and corner densities
        T c[R][1 << K], xlo[R][K], xhi[R][K];
#pragma unroll
        for (int r = 0; r < R; ++r) src.fetch(smem, lo[r], c[r], xlo[r], xhi[r]);
        // stage 4: in-cell inverse CDF, affine map, streaming stores
#pragma unroll
        for (int r = 0; r < R; ++r) {
            T z[K];
            incell<K>(c[r], u[r], z, src.prenormalised());  // _unitsample_kd(*corners, u=u[..., :-1])
            const int64_t row = base + (int64_t)r * SAMPLE_THREADS + threadIdx.x;
            if (live[r]) {
#pragma unroll
                for (int d = 0; d < K; ++d)
                    store_stream(out + row * K + d, affine(xlo[r][d], xhi[r][d], z[d]));
                if (cell_idx) __stcs(reinterpret_cast<long long *>(cell_idx) + row, (long long)lo[r]);
            }
        }
    }
}

// Rows [*first_row_dev, N) instead of [0, N), at most max_rows of them (the grid is
// sized from max_rows because the first row is only known on the device).
struct RowWindow {
    const uint32_t *first_row_dev = nullptr;
    int64_t max_rows = 0;
    const uint32_t *cell_range_dev = nullptr;   // slab of the table that is valid (lint_grid_build)
};

template <int K, typename T, typename Source, typename Uniforms>
int launch_sample(const Source &src, const Uniforms &unif, int64_t N, void *out, int64_t *cell_idx,
                  cudaStream_t st, const char *what, const RowWindow &win) {
    if (N == 0) return LINT_OK;
    const int64_t rows = win.first_row_dev ? std::min(N, win.max_rows) : N;
    if (rows == 0) return LINT_OK;
    const int64_t tile = (int64_t)SAMPLE_THREADS * RowsPerThread<T>::value;
    const int64_t want = (rows + tile - 1) / tile;
    const int blocks = (int)std::min<int64_t>(want, 148 * 32);
    const size_t smem = ((src.smem_bytes(sizeof(T)) + 15) & ~(size_t)15) + unif.smem_words() * 4;
    sample_kernel<K, T, Source, Uniforms><<<blocks, SAMPLE_THREADS, smem, st>>>(
        src, unif, N, (T *)out, cell_idx, win.first_row_dev);
    return check_launch(what);
}

// ---------------------------------------------------------------------------
// dispatch helpers
// ---------------------------------------------------------------------------
template <int K, typename UnifMaker>
int grid_dispatch_k(const lint_grid_t *g, uint64_t ncells, int64_t N, void *out, int64_t *cell_idx,
                    cudaStream_t st, const UnifMaker &mk, const char *what, const RowWindow &win) {
    if (g->dtype == LINT_F32) {
        GridSource<K, float> src{make_grid_view<K>(g, ncells, true)};
        src.g.table.range = win.cell_range_dev;
        return launch_sample<K, float>(src, mk.template make<K>(), N, out, cell_idx, st, what, win);
    }
    GridSource<K, double> src{make_grid_view<K>(g, ncells, true)};
    src.g.table.range = win.cell_range_dev;
    return launch_sample<K, double>(src, mk.template make<K>(), N, out, cell_idx, st, what, win);
}

template <typename UnifMaker>
int grid_dispatch(const lint_grid_t *g, int64_t N, void *out, int64_t *cell_idx, lint_stream_t stream,
                  const UnifMaker &mk, const char *what, const RowWindow &win = RowWindow{}) {
    uint64_t nc, nv;
    if (int rc = validate_grid(g, true, &nc, &nv)) return rc;
    if (N < 0 || (N > 0 && !out)) return fail(LINT_ERR_BAD_ARG, "%s: bad N / out_dev", what);
    cudaStream_t st = (cudaStream_t)stream;
    switch (g->dim) {
        case 1: return grid_dispatch_k<1>(g, nc, N, out, cell_idx, st, mk, what, win);
        case 2: return grid_dispatch_k<2>(g, nc, N, out, cell_idx, st, mk, what, win);
        case 3: return grid_dispatch_k<3>(g, nc, N, out, cell_idx, st, mk, what, win);
        case 4: return grid_dispatch_k<4>(g, nc, N, out, cell_idx, st, mk, what, win);
        case 5: return grid_dispatch_k<5>(g, nc, N, out, cell_idx, st, mk, what, win);
        default: return grid_dispatch_k<6>(g, nc, N, out, cell_idx, st, mk, what, win);
    }
}

static int validate_cells(const lint_cells_t *c) {
    if (!c) return fail(LINT_ERR_BAD_ARG, "cell-list descriptor is NULL");
    if (c->dim < 1 || c->dim > LINT_MAX_DIM)
        return fail(LINT_ERR_BAD_ARG, "cell-list dim %d outside 1..%d", c->dim, LINT_MAX_DIM);
    if (c->dtype != LINT_F32 && c->dtype != LINT_F64)
        return fail(LINT_ERR_BAD_ARG, "cell-list dtype %d is not LINT_F32/LINT_F64", c->dtype);
    if (c->ncells < 1 || c->ncells >= (1ll << 31))
        return fail(LINT_ERR_BAD_ARG, "cell-list ncells outside 1..2^31-1");
    if (!c->mins_dev || !c->widths_dev || !c->corners_dev || !c->cdf_dev)
        return fail(LINT_ERR_BAD_ARG, "cell-list device pointer is NULL");
    if (c->guide_bits < 0 || c->guide_bits > 30 || (c->guide_bits > 0 && !c->guide_dev))
        return fail(LINT_ERR_BAD_ARG, "cell-list guide table inconsistent (bits=%d)", c->guide_bits);
    return LINT_OK;
}

static CellsView make_cells_view(const lint_cells_t *c) {
    CellsView v;
    v.mins = c->mins_dev;
    v.widths = c->widths_dev;
    v.corners = c->corners_dev;
    v.rec = c->cell_records_dev;
    v.table.cdf = c->cdf_dev;
    v.table.guide = c->guide_dev;
    v.table.guide_bits = c->guide_bits;
    v.table.n = (uint32_t)c->ncells;
    return v;
}

template <int K, typename UnifMaker>
int cells_dispatch_k(const lint_cells_t *c, int64_t N, void *out, int64_t *cell_idx, cudaStream_t st,
                     const UnifMaker &mk, const char *what, const RowWindow &win) {
    if (c->dtype == LINT_F32) {
        ListSource<K, float> src{make_cells_view(c)};
        return launch_sample<K, float>(src, mk.template make<K>(), N, out, cell_idx, st, what, win);
    }
    ListSource<K, double> src{make_cells_view(c)};
    return launch_sample<K, double>(src, mk.template make<K>(), N, out, cell_idx, st, what, win);
}

template <typename UnifMaker>
int cells_dispatch(const lint_cells_t *c, int64_t N, void *out, int64_t *cell_idx, lint_stream_t stream,
                   const UnifMaker &mk, const char *what, const RowWindow &win = RowWindow{}) {
    if (int rc = validate_cells(c)) return rc;
    if (N < 0 || (N > 0 && !out)) return fail(LINT_ERR_BAD_ARG, "%s: bad N / out_dev", what);
    cudaStream_t st = (cudaStream_t)stream;
    switch (c->dim) {
        case 1: return cells_dispatch_k<1>(c, N, out, cell_idx, st, mk, what, win);
        case 2: return cells_dispatch_k<2>(c, N, out, cell_idx, st, mk, what, win);
        case 3: return cells_dispatch_k<3>(c, N, out, cell_idx, st, mk, what, win);
        case 4: return cells_dispatch_k<4>(c, N, out, cell_idx, st, mk, what, win);
        case 5: return cells_dispatch_k<5>(c, N, out, cell_idx, st, mk, what, win);
        default: return cells_dispatch_k<6>(c, N, out, cell_idx, st, mk, what, win);
    }
}

struct ArrayMaker {
    const double *u;
    template <int K> UniformsFromArray<K> make() const { return UniformsFromArray<K>{u}; }
};
struct PhiloxMaker {
    uint64_t seed, offset;
    double u_lo = 0.0, u_hi = 1.0;
    template <int K> UniformsPhilox<K> make() const {
        return UniformsPhilox<K>{make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                                 StreamPos{offset, offset_counter()}, u_lo, u_hi - u_lo};
    }
};

static int check_stratum(double u_lo, double u_hi, const char *what) {
    if (!(u_lo >= 0.0) || !(u_hi <= 1.0) || !(u_lo < u_hi))
        return fail(LINT_ERR_BAD_ARG, "%s: stratum [%g, %g) is not inside [0, 1]", what, u_lo, u_hi);
    return LINT_OK;
}
struct SobolMaker {
    const lint_sobol_t *s;
    template <int K> UniformsSobol<K> make() const {
        UniformsSobol<K> u;
        for (int j = 0; j <= K; ++j) {
            for (int b = 0; b < 32; ++b) u.sv[j][b] = s->sv[j][b];
            u.shift[j] = s->shift[j];
        }
        for (int b = 0; b < 32; ++b) u.ainv[b] = s->sorted_ainv[b];
        u.first = s->first_index;
        u.sorted_bits = s->sorted_bits;
        return u;
    }
};

// ---------------------------------------------------------------------------
// in-cell stage only (generic DensityStructure)   [sampling.py:123-126]
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(SAMPLE_THREADS)
incell_kernel(int64_t N, const double *mins, const double *maxs, const double *corners,
              const double *u, double *out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < N; row += stride) {
        double c[1 << K], uu[K], z[K];
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) c[j] = corners[(int64_t)j * N + row];
#pragma unroll
        for (int d = 0; d < K; ++d) uu[d] = u[row * (K + 1) + d];
        incell_exact<K>(c, uu, z);
#pragma unroll
        for (int d = 0; d < K; ++d)
            out[row * K + d] = affine(mins[row * K + d], maxs[row * K + d], z[d]);
    }
}

template <int K, typename T, typename Uniforms>
__global__ void __launch_bounds__(SAMPLE_THREADS)
uniforms_kernel(Uniforms unif, int64_t N, double *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t *utab = unif.stage(reinterpret_cast<uint32_t *>(smem_raw));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < N; row += stride) {
        double ucell;
        T u[K];
        unif.get(row, ucell, u, utab);
#pragma unroll
        for (int d = 0; d < K; ++d) out[row * (K + 1) + d] = (double)u[d];
        out[row * (K + 1) + K] = ucell;
    }
}

template <int K, typename Maker>
int uniforms_dispatch_k(int dtype, int64_t N, const Maker &mk, double *out, cudaStream_t st) {
    if (N == 0) return LINT_OK;
    const int blocks = (int)std::min<int64_t>((N + SAMPLE_THREADS - 1) / SAMPLE_THREADS, 148 * 32);
    auto u = mk.template make<K>();
    if (dtype == LINT_F32)
        uniforms_kernel<K, float, decltype(u)><<<blocks, SAMPLE_THREADS, u.smem_words() * 4, st>>>(u, N, out);
    else
        uniforms_kernel<K, double, decltype(u)><<<blocks, SAMPLE_THREADS, u.smem_words() * 4, st>>>(u, N, out);
    return check_launch("uniforms");
}

template <typename Maker>
int uniforms_dispatch(int dim, int dtype, int64_t N, const Maker &mk, double *out, lint_stream_t stream) {
    if (dim < 1 || dim > LINT_MAX_DIM) return fail(LINT_ERR_BAD_ARG, "dim %d outside 1..%d", dim, LINT_MAX_DIM);
    if (dtype != LINT_F32 && dtype != LINT_F64) return fail(LINT_ERR_BAD_ARG, "bad dtype %d", dtype);
    if (N < 0 || (N > 0 && !out)) return fail(LINT_ERR_BAD_ARG, "bad N / u_dev");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dim) {
        case 1: return uniforms_dispatch_k<1>(dtype, N, mk, out, st);
        case 2: return uniforms_dispatch_k<2>(dtype, N, mk, out, st);
        case 3: return uniforms_dispatch_k<3>(dtype, N, mk, out, st);
        case 4: return uniforms_dispatch_k<4>(dtype, N, mk, out, st);
        case 5: return uniforms_dispatch_k<5>(dtype, N, mk, out, st);
        default: return uniforms_dispatch_k<6>(dtype, N, mk, out, st);
    }
}

// list of grids: one launch
template <int K, typename UnifMaker>
int multi_dispatch_k(const lint_grid_t *grids, int nd, const double *choice_cdf, const double *bounds, int64_t N,
                     void *views_dev, void *out, int64_t *cell_idx, cudaStream_t st, const UnifMaker &mk,
                     const char *what) {
    GridView<K> host_views[LINT_MULTI_MAX];
    uint32_t off = 0;
    bool all_rec = true;
    uint32_t cell_off[LINT_MULTI_MAX + 1];
    for (int d = 0; d < nd; ++d) {
        uint64_t nc, nv;
        if (int rc = validate_grid(&grids[d], true, &nc, &nv)) return rc;
        if (grids[d].dim != K || grids[d].dtype != grids[0].dtype)
            return fail(LINT_ERR_BAD_ARG, "%s: the grids of a list must share dimension and dtype", what);
        host_views[d] = make_grid_view<K>(&grids[d], nc, false);
        all_rec = all_rec && grids[d].cell_records_dev != nullptr;
        cell_off[d] = off;
        if ((uint64_t)off + nc >= (1ull << 31)) return fail(LINT_ERR_UNSUPPORTED, "%s: >= 2^31 cells in total", what);
        off += (uint32_t)nc;
    }
    cell_off[nd] = off;
    if (!all_rec)
        for (int d = 0; d < nd; ++d) host_views[d].rec = nullptr;     // one convention for the whole list
    if (cudaMemcpyAsync(views_dev, host_views, sizeof(GridView<K>) * nd, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail(LINT_ERR_CUDA, "%s: copying the grid table failed", what);
    auto fill = [&](auto &src) {
        src.views = reinterpret_cast<const GridView<K> *>(views_dev);
        src.nd = nd;
        src.has_records = all_rec;
        for (int d = 0; d < nd; ++d) {
            src.choice_cdf[d] = choice_cdf[d];
            src.starts[d] = bounds[d];
            src.ends[d] = bounds[d + 1];
        }
        for (int d = 0; d <= nd; ++d) src.cell_off[d] = cell_off[d];
    };
    if (grids[0].dtype == LINT_F32) {
        MultiGridSource<K, float> src;
        fill(src);
        return launch_sample<K, float>(src, mk.template make<K>(), N, out, cell_idx, st, what, RowWindow{});
    }
    MultiGridSource<K, double> src;
    fill(src);
    return launch_sample<K, double>(src, mk.template make<K>(), N, out, cell_idx, st, what, RowWindow{});
}

template <typename UnifMaker>
int multi_dispatch(const lint_grid_t *grids, int nd, const double *choice_cdf, const double *bounds, int64_t N,
                   void *views_dev, size_t views_bytes, void *out, int64_t *cell_idx, lint_stream_t stream,
                   const UnifMaker &mk, const char *what) {
    if (!grids || nd < 1 || nd > LINT_MULTI_MAX) return fail(LINT_ERR_BAD_ARG, "%s: 1..%d grids", what, LINT_MULTI_MAX);
    if (!choice_cdf || !bounds) return fail(LINT_ERR_BAD_ARG, "%s: choice_cdf / bounds is NULL", what);
    if (N < 0 || (N > 0 && !out)) return fail(LINT_ERR_BAD_ARG, "%s: bad N / out_dev", what);
    if (N == 0) return LINT_OK;
    if (!views_dev || views_bytes < lint_multi_scratch_bytes(grids[0].dim, nd))
        return fail(LINT_ERR_BAD_ARG, "%s: scratch too small", what);
    cudaStream_t st = (cudaStream_t)stream;
    switch (grids[0].dim) {
        case 1: return multi_dispatch_k<1>(grids, nd, choice_cdf, bounds, N, views_dev, out, cell_idx, st, mk, what);
        case 2: return multi_dispatch_k<2>(grids, nd, choice_cdf, bounds, N, views_dev, out, cell_idx, st, mk, what);
        case 3: return multi_dispatch_k<3>(grids, nd, choice_cdf, bounds, N, views_dev, out, cell_idx, st, mk, what);
        case 4: return multi_dispatch_k<4>(grids, nd, choice_cdf, bounds, N, views_dev, out, cell_idx, st, mk, what);
        case 5: return multi_dispatch_k<5>(grids, nd, choice_cdf, bounds, N, views_dev, out, cell_idx, st, mk, what);
        case 6: return multi_dispatch_k<6>(grids, nd, choice_cdf, bounds, N, views_dev, out, cell_idx, st, mk, what);
        default: return fail(LINT_ERR_BAD_ARG, "%s: dim outside 1..%d", what, LINT_MAX_DIM);
    }
}

// packed leaf records (lint_cells_records)
template <int K, typename DensT>
__global__ void cells_records_kernel(CellsView s, uint32_t ncells, DensT *rec) {
    constexpr int RS = cells_record_stride<K, DensT>();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += gridDim.x * blockDim.x) {
        DensT *dst = rec + (size_t)i * RS;
        const DensT *C = reinterpret_cast<const DensT *>(s.corners) + (size_t)i * (1 << K);
        double scale = 1.0;                       // fp32 records: corners / largest corner
        if (sizeof(DensT) == 4) {
            double m = 0.0;
            for (int j = 0; j < (1 << K); ++j) m = fmax(m, (double)C[j]);
            scale = m > 0.0 ? 1.0 / m : 0.0;
        }
        for (int j = 0; j < (1 << K); ++j) dst[j] = (DensT)((double)C[j] * scale);
        for (int d = 0; d < K; ++d) {
            const double x = s.mins[(size_t)i * K + d], dx = s.widths[(size_t)i * K + d];
            dst[(1 << K) + d] = (DensT)x;
            dst[(1 << K) + K + d] = (DensT)__dadd_rn(x, dx);
        }
        for (int j = (1 << K) + 2 * K; j < RS; ++j) dst[j] = (DensT)0;
    }
}

template <int K>
int cells_records_k(const lint_cells_t *c, void *rec, cudaStream_t st) {
    CellsView v = make_cells_view(c);
    const int blocks = (int)std::min<int64_t>((c->ncells + 255) / 256, 148 * 16);
    if (c->dtype == LINT_F32)
        cells_records_kernel<K, float><<<blocks, 256, 0, st>>>(v, (uint32_t)c->ncells, (float *)rec);
    else
        cells_records_kernel<K, double><<<blocks, 256, 0, st>>>(v, (uint32_t)c->ncells, (double *)rec);
    return check_launch("lint_cells_records");
}

template <int K> static size_t cells_record_bytes_k(int dtype) {
    return dtype == LINT_F32 ? cells_record_stride<K, float>() * sizeof(float)
                             : cells_record_stride<K, double>() * sizeof(double);
}

// u-driven top-up of rows [*first_row_dev, N) for the cell-major path
int topup_grid(const lint_grid_t *g, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
               uint64_t seed, uint64_t offset, double u_lo, double u_hi, const uint32_t *cell_range_dev, void *out,
               int64_t *cell_idx, cudaStream_t st) {
    return grid_dispatch(g, N, out, cell_idx, (lint_stream_t)st, PhiloxMaker{seed, offset, u_lo, u_hi},
                         "cell-major top-up", RowWindow{first_row_dev, max_rows, cell_range_dev});
}

int topup_cells(const lint_cells_t *c, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
                uint64_t seed, uint64_t offset, double u_lo, double u_hi, void *out, int64_t *cell_idx,
                cudaStream_t st) {
    return cells_dispatch(c, N, out, cell_idx, (lint_stream_t)st, PhiloxMaker{seed, offset, u_lo, u_hi},
                          "cell-major top-up", RowWindow{first_row_dev, max_rows});
}

}  // namespace lint

using namespace lint;

extern "C" {

size_t lint_multi_scratch_bytes(int dim, int ndomains) {
    if (ndomains < 1 || ndomains > LINT_MULTI_MAX) return 0;
    switch (dim) {
        case 1: return sizeof(GridView<1>) * ndomains;
        case 2: return sizeof(GridView<2>) * ndomains;
        case 3: return sizeof(GridView<3>) * ndomains;
        case 4: return sizeof(GridView<4>) * ndomains;
        case 5: return sizeof(GridView<5>) * ndomains;
        case 6: return sizeof(GridView<6>) * ndomains;
        default: return 0;
    }
}

size_t lint_cells_record_bytes(int dim, int dtype) {
    switch (dim) {
        case 1: return cells_record_bytes_k<1>(dtype);
        case 2: return cells_record_bytes_k<2>(dtype);
        case 3: return cells_record_bytes_k<3>(dtype);
        case 4: return cells_record_bytes_k<4>(dtype);
        case 5: return cells_record_bytes_k<5>(dtype);
        case 6: return cells_record_bytes_k<6>(dtype);
        default: return 0;
    }
}

int lint_cells_records(const lint_cells_t *c, void *records_dev, lint_stream_t stream) {
    if (!c || c->dim < 1 || c->dim > LINT_MAX_DIM || c->ncells < 1 || c->ncells >= (1ll << 31) ||
        !c->mins_dev || !c->widths_dev || !c->corners_dev || !records_dev ||
        (c->dtype != LINT_F32 && c->dtype != LINT_F64))
        return fail(LINT_ERR_BAD_ARG, "lint_cells_records: bad cell-list descriptor or NULL records_dev");
    cudaStream_t st = (cudaStream_t)stream;
    switch (c->dim) {
        case 1: return cells_records_k<1>(c, records_dev, st);
        case 2: return cells_records_k<2>(c, records_dev, st);
        case 3: return cells_records_k<3>(c, records_dev, st);
        case 4: return cells_records_k<4>(c, records_dev, st);
        case 5: return cells_records_k<5>(c, records_dev, st);
        default: return cells_records_k<6>(c, records_dev, st);
    }
}

int lint_grid_sample_u(const lint_grid_t *g, const double *u_dev, int64_t N, void *out_dev,
                       int64_t *cell_idx_dev, lint_stream_t stream) {
    if (N > 0 && !u_dev) return fail(LINT_ERR_BAD_ARG, "lint_grid_sample_u: u_dev is NULL");
    return grid_dispatch(g, N, out_dev, cell_idx_dev, stream, ArrayMaker{u_dev}, "lint_grid_sample_u");
}

int lint_cells_sample_u(const lint_cells_t *c, const double *u_dev, int64_t N, void *out_dev,
                        int64_t *cell_idx_dev, lint_stream_t stream) {
    if (N > 0 && !u_dev) return fail(LINT_ERR_BAD_ARG, "lint_cells_sample_u: u_dev is NULL");
    return cells_dispatch(c, N, out_dev, cell_idx_dev, stream, ArrayMaker{u_dev}, "lint_cells_sample_u");
}

int lint_incell_sample_u(int dim, int64_t N, const double *mins_dev, const double *maxs_dev,
                         const double *corners_dev, const double *u_dev, double *out_dev,
                         lint_stream_t stream) {
    if (dim < 1 || dim > LINT_MAX_DIM)
        return fail(LINT_ERR_BAD_ARG, "lint_incell_sample_u: dim %d outside 1..%d", dim, LINT_MAX_DIM);
    if (N < 0) return fail(LINT_ERR_BAD_ARG, "lint_incell_sample_u: N < 0");
    if (N == 0) return LINT_OK;
    if (!mins_dev || !maxs_dev || !corners_dev || !u_dev || !out_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_incell_sample_u: NULL device pointer");
    const int blocks = (int)std::min<int64_t>((N + SAMPLE_THREADS - 1) / SAMPLE_THREADS, 148 * 32);
    cudaStream_t st = (cudaStream_t)stream;
#define LINT_CASE(KK) case KK: incell_kernel<KK><<<blocks, SAMPLE_THREADS, 0, st>>>(N, mins_dev, maxs_dev, corners_dev, u_dev, out_dev); break;
    switch (dim) { LINT_CASE(1) LINT_CASE(2) LINT_CASE(3) LINT_CASE(4) LINT_CASE(5) LINT_CASE(6) }
#undef LINT_CASE
    return check_launch("lint_incell_sample_u");
}

int lint_grid_sample_philox(const lint_grid_t *g, int64_t N, uint64_t seed, uint64_t offset,
                            double u_lo, double u_hi, void *out_dev, int64_t *cell_idx_dev,
                            lint_stream_t stream) {
    if (int rc = check_stratum(u_lo, u_hi, "lint_grid_sample_philox")) return rc;
    return grid_dispatch(g, N, out_dev, cell_idx_dev, stream, PhiloxMaker{seed, offset, u_lo, u_hi},
                         "lint_grid_sample_philox");
}

int lint_cells_sample_philox(const lint_cells_t *c, int64_t N, uint64_t seed, uint64_t offset,
                             double u_lo, double u_hi, void *out_dev, int64_t *cell_idx_dev,
                             lint_stream_t stream) {
    if (int rc = check_stratum(u_lo, u_hi, "lint_cells_sample_philox")) return rc;
    return cells_dispatch(c, N, out_dev, cell_idx_dev, stream, PhiloxMaker{seed, offset, u_lo, u_hi},
                          "lint_cells_sample_philox");
}

int lint_philox_uniforms(int dim, int dtype, int64_t N, uint64_t seed, uint64_t offset,
                         double *u_dev, lint_stream_t stream) {
    return uniforms_dispatch(dim, dtype, N, PhiloxMaker{seed, offset, 0.0, 1.0}, u_dev, stream);
}

static int check_sobol(const lint_sobol_t *sobol, int64_t N, const char *what) {
    if (!sobol) return fail(LINT_ERR_BAD_ARG, "%s: sobol is NULL", what);
    if (sobol->sorted_bits) {
        if (sobol->sorted_bits > 32 || sobol->first_index != 0 || N != (1ll << sobol->sorted_bits))
            return fail(LINT_ERR_BAD_ARG, "%s: sorted enumeration needs first_index == 0 and N == 2^sorted_bits",
                        what);
    } else if ((uint64_t)N + sobol->first_index > (1ull << 32)) {
        return fail(LINT_ERR_BAD_ARG, "%s: a 32-bit Sobol sequence has 2^32 points", what);
    }
    return LINT_OK;
}

int lint_grid_sample_sobol(const lint_grid_t *g, int64_t N, const lint_sobol_t *sobol, void *out_dev,
                           int64_t *cell_idx_dev, lint_stream_t stream) {
    if (int rc = check_sobol(sobol, N, "lint_grid_sample_sobol")) return rc;
    return grid_dispatch(g, N, out_dev, cell_idx_dev, stream, SobolMaker{sobol}, "lint_grid_sample_sobol");
}

int lint_cells_sample_sobol(const lint_cells_t *c, int64_t N, const lint_sobol_t *sobol, void *out_dev,
                            int64_t *cell_idx_dev, lint_stream_t stream) {
    if (int rc = check_sobol(sobol, N, "lint_cells_sample_sobol")) return rc;
    return cells_dispatch(c, N, out_dev, cell_idx_dev, stream, SobolMaker{sobol}, "lint_cells_sample_sobol");
}

int lint_sobol_uniforms(int dim, int64_t N, const lint_sobol_t *sobol, double *u_dev,
                        lint_stream_t stream) {
    if (int rc = check_sobol(sobol, N, "lint_sobol_uniforms")) return rc;
    return uniforms_dispatch(dim, LINT_F64, N, SobolMaker{sobol}, u_dev, stream);
}

int lint_multi_grid_sample_u(const lint_grid_t *grids, int ndomains, const double *choice_cdf, const double *bounds,
                             const double *u_dev, int64_t N, void *scratch_dev, size_t scratch_bytes, void *out_dev,
                             int64_t *cell_idx_dev, lint_stream_t stream) {
    if (N > 0 && !u_dev) return fail(LINT_ERR_BAD_ARG, "lint_multi_grid_sample_u: u_dev is NULL");
    return multi_dispatch(grids, ndomains, choice_cdf, bounds, N, scratch_dev, scratch_bytes, out_dev, cell_idx_dev,
                          stream, ArrayMaker{u_dev}, "lint_multi_grid_sample_u");
}

int lint_multi_grid_sample_philox(const lint_grid_t *grids, int ndomains, const double *choice_cdf,
                                  const double *bounds, int64_t N, uint64_t seed, uint64_t offset, void *scratch_dev,
                                  size_t scratch_bytes, void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream) {
    return multi_dispatch(grids, ndomains, choice_cdf, bounds, N, scratch_dev, scratch_bytes, out_dev, cell_idx_dev,
                          stream, PhiloxMaker{seed, offset, 0.0, 1.0}, "lint_multi_grid_sample_philox");
}

int lint_multi_grid_sample_sobol(const lint_grid_t *grids, int ndomains, const double *choice_cdf,
                                 const double *bounds, int64_t N, const lint_sobol_t *sobol, void *scratch_dev,
                                 size_t scratch_bytes, void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream) {
    if (int rc = check_sobol(sobol, N, "lint_multi_grid_sample_sobol")) return rc;
    return multi_dispatch(grids, ndomains, choice_cdf, bounds, N, scratch_dev, scratch_bytes, out_dev, cell_idx_dev,
                          stream, SobolMaker{sobol}, "lint_multi_grid_sample_sobol");
}

}  // extern "C"

// ===========================================================================
// row permutation: the device counterpart of `rng.shuffle(X, axis=0)`
// [lintsampler.py:309-311]
// ===========================================================================
namespace lint {

// Bijection of [0, 2^bits) : alternating unbalanced Feistel network whose round
// function is a 32-bit integer mix; restricted to [0, N) by cycle walking
// (expected < 2 evaluations since 2^bits < 2N).
struct FeistelPerm {
    uint32_t key[8];
    uint32_t lb, rb;                  // bit widths of the two halves, lb + rb = bits, rb >= lb
    uint64_t n;
    __device__ __forceinline__ static uint32_t mix(uint32_t x, uint32_t k) {
        x = x * 0x9E3779B1u + k;
        x ^= x >> 15; x *= 0x85EBCA77u;
        x ^= x >> 13; x *= 0xC2B2AE3Du;
        x ^= x >> 16;
        return x;
    }
    __device__ __forceinline__ uint64_t once(uint64_t x) const {
        uint32_t wl = lb, wr = rb;
        uint32_t l = (uint32_t)(x >> wr), r = (uint32_t)(x & ((1ull << wr) - 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t t = (l ^ mix(r, key[i])) & (uint32_t)((1ull << wl) - 1);
            l = r;                    // widths swap every round; 8 rounds restore them
            r = t;
            const uint32_t w = wl; wl = wr; wr = w;
        }
        return ((uint64_t)l << wr) | r;
    }
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const {
        uint64_t x = once(i);
        while (x >= n) x = once(x);
        return x;
    }
};

template <typename E>
__global__ void __launch_bounds__(256)
permute_rows_kernel(const E *__restrict__ in, E *__restrict__ out, uint64_t N, int k, FeistelPerm perm) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = perm(i);
        for (int d = 0; d < k; ++d) __stcs(out + i * k + d, __ldg(in + j * k + d));
    }
}

}  // namespace lint

extern "C" int lint_permute_rows(const void *in_dev, void *out_dev, int64_t N, int32_t k, int32_t dtype,
                                 uint64_t seed, lint_stream_t stream) {
    using namespace lint;
    if (N < 0 || k < 1 || k > LINT_MAX_DIM || (dtype != LINT_F32 && dtype != LINT_F64))
        return fail(LINT_ERR_BAD_ARG, "lint_permute_rows: bad N / k / dtype");
    if (N == 0) return LINT_OK;
    if (!in_dev || !out_dev || in_dev == out_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_permute_rows: in/out must be distinct, non-NULL buffers");
    FeistelPerm p;
    uint32_t bits = 2;
    while ((1ull << bits) < (uint64_t)N) ++bits;
    p.lb = bits / 2;
    p.rb = bits - p.lb;
    p.n = (uint64_t)N;
    uint4 a = philox4x32_10(make_uint4(0x50455231u, 0, 0, 0), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint4 b = philox4x32_10(make_uint4(0x50455232u, 0, 0, 0), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t ks[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    for (int i = 0; i < 8; ++i) p.key[i] = ks[i];
    const int blocks = (int)std::min<int64_t>((N + 255) / 256, 148 * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LINT_F32)
        permute_rows_kernel<float><<<blocks, 256, 0, st>>>((const float *)in_dev, (float *)out_dev, (uint64_t)N, k, p);
    else
        permute_rows_kernel<double><<<blocks, 256, 0, st>>>((const double *)in_dev, (double *)out_dev, (uint64_t)N, k, p);
    return check_launch("lint_permute_rows");
}
