// lintsampler_b200 — tree construction on the device for fused pdf families.
//
// Reference: DensityTree / _TreeCell, density_structures/tree.py:228-292 (construction),
// 357-488 (refine), 593-613 (_evaluate_cell), 615-647 (open), 649-703 (Romberg estimates).
//
// The tree is a table of cells in structure-of-arrays form, appended to as cells are opened.
// `lint_tree_open` creates the 2^k children of a batch of cells: one thread per child evaluates the
// pdf on its 2^k corners, the trapezoid mass and the chain of Richardson-extrapolated masses from
// the corner tables of its ancestors — the arithmetic of tree.py:608 and 649-703 in NumPy's
// operation order (the sums over corners are NumPy's pairwise add.reduce), so that a tree built
// here and one built by the host code from the same corner densities agree bit for bit.  Which
// cells to open, and in which order (it fixes the leaf order and with it the leaf CDF,
// tree.py:335-338), stays with the caller: it needs only the three numbers per new cell returned in
// `summary_dev`.
#include <algorithm>
#include "lint_device.cuh"
#include "lint_pdf.cuh"
#include "lint_pairwise.cuh"

namespace lint {

constexpr int TREE_MAX_LEVEL = LINT_TREE_MAX_LEVEL;       // int32 lattice coordinates
constexpr int TREE_EST = LINT_TREE_MAX_LEVEL + 2;         // Romberg estimates kept per cell (stride)

template <int K>
struct TreeTable {
    double mins[K], span[K];
    int32_t *origin;        // (cap, K) lattice coordinates at the cell's level
    int32_t *level;         // (cap)
    int32_t *parent;        // (cap), -1 for the root
    double *corners;        // (cap, 2^K) corner densities, corner order (dim 0 = most significant bit)
    double *mass_raw;       // (cap)
    double *est;            // (cap, TREE_EST) Romberg estimates e[0..level]
};

// NumPy's add.reduce over N contiguous doubles (np.sum(a, axis=1) of an (n, N) array):
// below 8 elements a running sum from -0.0, otherwise eight accumulators combined pairwise.
template <int N>
__device__ __forceinline__ double numpy_sum(const double (&a)[N]) {
    if constexpr (N < 8) {
        double r = -0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) r = __dadd_rn(r, a[i]);
        return r;
    } else {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
#pragma unroll
        for (int i = 8; i < N; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
        }
        return __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    }
}

// dx = span / 2^level per dimension, vol = np.prod(dx) (a running product from 1.0)
template <int K>
__device__ __forceinline__ double cell_volume(const double (&span)[K], int level) {
    double vol = 1.0;
#pragma unroll
    for (int d = 0; d < K; ++d) vol = __dmul_rn(vol, scalbn(span[d], -level));
    return vol;
}

// Lattice origin and level of child `c` (corner order) of cell `par`, or of the root (par < 0).
template <int K>
__device__ __forceinline__ int tree_child_origin(const TreeTable<K> &tb, int par, uint32_t c, int32_t (&org)[K]) {
    if (par < 0) {
#pragma unroll
        for (int d = 0; d < K; ++d) org[d] = 0;
        return 0;
    }
#pragma unroll
    for (int d = 0; d < K; ++d)                                              // tree.py:632-640
        org[d] = 2 * tb.origin[(size_t)par * K + d] + (int32_t)((c >> (K - 1 - d)) & 1u);
    return tb.level[par] + 1;
}

// Density at corner `cc` of the cell at (org, level): mins + coords * (span / 2^level)
// [tree.py:593-596, 866-867]; ORs the findings of tree.py:598-605 into `bad`.
template <int K, typename Pdf>
__device__ __forceinline__ double tree_corner_density(const TreeTable<K> &tb, const Pdf &pdf, const int32_t (&org)[K],
                                                      int level, uint32_t cc, uint32_t &bad) {
    double x[K];
#pragma unroll
    for (int d = 0; d < K; ++d) {
        const int32_t coord = org[d] + (int32_t)((cc >> (K - 1 - d)) & 1u);
        x[d] = __dadd_rn(tb.mins[d], __dmul_rn((double)coord, scalbn(tb.span[d], -level)));
    }
    const double f = pdf(x);
    if (f < 0.0) bad |= LINT_FLAG_NEGATIVE;
    if (!isfinite(f)) bad |= LINT_FLAG_NONFINITE;
    return f;
}

// Enters cell `id` (child of `par` at (org, level), corner densities f) into the table: trapezoid
// mass and Romberg chain.  Returns raw mass, Romberg mass and last Romberg correction.
template <int K>
__device__ void tree_finish_cell(const TreeTable<K> &tb, uint32_t id, int par, const int32_t (&org)[K], int level,
                                 const double (&f)[1 << K], double &raw_out, double &rom_out, double &err_out) {
    constexpr int NC = 1 << K;
    const double vol = cell_volume<K>(tb.span, level);
    const double raw = __dmul_rn(vol, __ddiv_rn(numpy_sum<NC>(f), (double)NC));         // tree.py:608
    tb.level[id] = level;
    tb.parent[id] = par;
#pragma unroll
    for (int d = 0; d < K; ++d) tb.origin[(size_t)id * K + d] = org[d];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) tb.corners[(size_t)id * NC + cc] = f[cc];
    tb.mass_raw[id] = raw;

    // Romberg chain [tree.py:649-703]: cur[0] the cell's own trapezoid mass, cur[j] the mass its
    // j-th ancestor's multilinear interpolant gives it (value at the cell's midpoint x volume)
    double cur[TREE_MAX_LEVEL + 1];
    cur[0] = raw;
    int anc = par;
    for (int j = 1; j <= level; ++j) {
        double w[NC];
        w[0] = 1.0;
        // midpoint of the cell in the ancestor's unit cube (dyadic, exact); weights built dimension
        // by dimension, ((w0 * w1) * w2) ..., corner index with dim 0 as the most significant bit
#pragma unroll
        for (int d = 0; d < K; ++d) {
            const double mid = scalbn((double)org[d] + 0.5, -j) - (double)tb.origin[(size_t)anc * K + d];
            const int have = 1 << d;
#pragma unroll
            for (int cc = have - 1; cc >= 0; --cc) {
                const double base = w[cc];
                w[2 * cc + 1] = d == 0 ? mid : __dmul_rn(base, mid);
                w[2 * cc] = d == 0 ? 1.0 - mid : __dmul_rn(base, 1.0 - mid);
            }
        }
        double prod[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) prod[cc] = __dmul_rn(tb.corners[(size_t)anc * NC + cc], w[cc]);
        const double fmid = numpy_sum<NC>(prod);
        const double avol = cell_volume<K>(tb.span, level - j);
        cur[j] = scalbn(__dmul_rn(fmid, avol), -(j * K));                    // fmid * anc.vol / 2^(j k)
        anc = tb.parent[anc];
    }
    double *est = tb.est + (size_t)id * TREE_EST;
    est[0] = raw;
    double prev = raw, last = raw;
    for (int i = 0; i < level; ++i) {                                        // Richardson extrapolation
        const double denom = scalbn(1.0, 2 * (i + 1)) - 1.0;                 // 4^(i+1) - 1
        for (int j = 0; j < level - i; ++j)
            cur[j] = __dsub_rn(cur[j], __ddiv_rn(__dsub_rn(cur[j + 1], cur[j]), denom));
        prev = last;
        last = cur[0];
        est[i + 1] = last;
    }
    raw_out = raw;
    rom_out = last;                                                          // mass_romberg
    err_out = fabs(__dsub_rn(last, prev));                                   // |e[-1] - e[-2]| (0 for the root)
}

// Creates cell `id`: child `c` (corner order) of cell `par`, or the root (par < 0), by one thread.
template <int K, typename Pdf>
__device__ void tree_new_cell(const TreeTable<K> &tb, const Pdf &pdf, uint32_t id, int par, uint32_t c,
                              double &raw_out, double &rom_out, double &err_out, uint32_t &bad) {
    constexpr int NC = 1 << K;
    int32_t org[K];
    const int level = tree_child_origin<K>(tb, par, c, org);
    double f[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) f[cc] = tree_corner_density<K>(tb, pdf, org, level, cc, bad);
    tree_finish_cell<K>(tb, id, par, org, level, f, raw_out, rom_out, err_out);
}

// One thread per new cell.  Cell `first + t` is child `t mod 2^K` of open_ids[t / 2^K]
// (root: the single cell 0 with no parent, n_open == 0).
template <int K, typename Pdf>
__global__ void __launch_bounds__(128)
tree_open_kernel(TreeTable<K> tb, Pdf pdf, const int32_t *__restrict__ open_ids, uint32_t n_new, uint32_t first,
                 bool root, double *__restrict__ summary, uint32_t *flags) {
    constexpr int NC = 1 << K;
    uint32_t bad = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_new; t += gridDim.x * blockDim.x) {
        double raw, rom, err;
        tree_new_cell<K>(tb, pdf, first + t, root ? -1 : open_ids[t >> K], t & (NC - 1), raw, rom, err, bad);
        summary[(size_t)t * 3 + 0] = raw;
        summary[(size_t)t * 3 + 1] = rom;
        summary[(size_t)t * 3 + 2] = err;
    }
    if (bad) atomicOr(flags, bad);
}

// Stage 2 of refine [tree.py:448-488] as one launch: while the tree-wide fractional error
// sqrt(sum((m_romberg - m_raw)^2)) / sum(m_romberg) is not below tree_tol, open the single leaf with
// the largest squared difference (np.argmax: the first of equals), drop it from the leaf list and
// append its children.  One CTA; the two sums are NumPy's pairwise np.sum over the current leaf
// arrays (recomputed every iteration, as the reference does), so the loop opens exactly the cells
// the host loop would.  Stops early, to be relaunched or finished by the caller, when a table
// runs out of room, after max_iters openings, or at the level limit.
// state: [0] leaves, [1] cells, [2] openings done by this launch, [3] status (TREE_* below)
enum { TREE_CONVERGED = 0, TREE_OUT_OF_ROOM = 1, TREE_LEVEL_LIMIT = 2, TREE_MORE = 3 };
constexpr int REFINE_THREADS = 1024;
constexpr uint32_t REFINE_MAX_LEAVES = 128u * 1024u;      // pairwise tree of at most 1024 nodes per array

template <int K, typename Pdf>
__global__ void __launch_bounds__(REFINE_THREADS)
tree_refine_worst_kernel(TreeTable<K> tb, Pdf pdf, int32_t *leaf_ids, double *m_rom, double *m_raw, double *sq,
                         uint32_t leaf_cap, uint32_t cell_cap, double tree_tol, uint32_t max_iters,
                         uint32_t *state, uint32_t *flags) {
    constexpr int NC = 1 << K;
    // the threads that create the children of the leaf being opened: one per (child, corner) for
    // the densities (up to four dimensions; one per child beyond), one per child for the rest
    constexpr bool WIDE = K <= 4;
    constexpr uint32_t NMAKE = WIDE ? (NC * NC > 32 ? NC * NC : 32) : (NC > 32 ? NC : 32);
    constexpr uint32_t MAKERS = REFINE_THREADS - NMAKE;
    constexpr int SHIFT = 4;                               // elements per thread and round of the shift
    __shared__ double part[2][2][REFINE_THREADS];         // [ping-pong][array][node]
    __shared__ double s_val[REFINE_THREADS / 32];
    __shared__ uint32_t s_idx[REFINE_THREADS / 32];
    __shared__ uint32_t s_worst, s_stop, s_bad;
    __shared__ int s_par;
    __shared__ double s_new[2][NC];
    __shared__ double s_f[WIDE ? NC : 1][WIDE ? NC : 1];   // corner densities of the children
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t n = state[0], ncell = state[1], done = 0, status = TREE_MORE, bad = 0;
    while (true) {
        // (1) squared differences and the largest of them (np.argmax: the first of equals)
        double best = -1.0;
        uint32_t best_i = 0xffffffffu;
        for (uint32_t i = tid; i < n; i += REFINE_THREADS) {
            const double d = __dsub_rn(m_rom[i], m_raw[i]);
            const double q = __dmul_rn(d, d);
            sq[i] = q;
            if (q > best) { best = q; best_i = i; }       // ascending i per thread: keeps the first
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = best_i; }
        __syncthreads();
        if (tid == 0) {
            double bv = s_val[0];
            uint32_t bi = s_idx[0];
            for (int w = 1; w < REFINE_THREADS / 32; ++w)
                if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
            // can that leaf be opened at all?  (decided before the sums: its children are made beside them)
            uint32_t stop = TREE_MORE;
            int par = -1;
            if (done >= max_iters || n + NC - 1 > leaf_cap || ncell + NC > cell_cap || n + NC - 1 > REFINE_MAX_LEAVES ||
                bi >= n)
                stop = TREE_OUT_OF_ROOM;
            else {
                par = leaf_ids[bi];
                if (tb.level[par] >= TREE_MAX_LEVEL) stop = TREE_LEVEL_LIMIT;
            }
            s_worst = bi;
            s_stop = stop;
            s_par = par;
            s_bad = 0;
        }
        __syncthreads();
        const uint32_t worst = s_worst;
        const bool can_open = s_stop == TREE_MORE;
        // (2) np.sum(m_rom) and np.sum(sq): nodes of the pairwise tree at depth `depth` (one thread per
        // node and array), folded below; beside them the last warps make the children of the worst
        // leaf — speculatively: the sums may say the tree has converged, then they are never adopted
        const int depth = pw_depth(n);
        const uint32_t nnodes = 1u << depth;
        const bool two = 2 * nnodes <= MAKERS;             // one thread per (node, array)
        if (tid < (two ? 2 * nnodes : nnodes)) {
            const uint32_t p = tid < nnodes ? tid : tid - nnodes;
            uint64_t start = 0, len = n;
            for (int b = depth - 1; b >= 0; --b) {
                const uint64_t n2 = pw_left(len);
                if ((p >> b) & 1) { start += n2; len -= n2; } else { len = n2; }
            }
            if (tid < nnodes) part[0][0][p] = pw_node(m_rom + start, len);
            if (!two || tid >= nnodes) part[0][1][p] = pw_node(sq + start, len);
        } else if (can_open && tid >= MAKERS && tid < MAKERS + (WIDE ? NC * NC : NC)) {
            uint32_t my_bad = 0;
            if constexpr (WIDE) {
                const uint32_t c = (tid - MAKERS) / NC, cc = (tid - MAKERS) % NC;
                int32_t org[K];
                const int level = tree_child_origin<K>(tb, s_par, c, org);
                s_f[c][cc] = tree_corner_density<K>(tb, pdf, org, level, cc, my_bad);
            } else {
                const uint32_t c = tid - MAKERS;
                double raw, rom, err;
                tree_new_cell<K>(tb, pdf, ncell + c, s_par, c, raw, rom, err, my_bad);
                s_new[0][c] = raw;
                s_new[1][c] = rom;
            }
            if (my_bad) atomicOr(&s_bad, my_bad);
        }
        __syncthreads();
        if constexpr (WIDE) {                              // masses and Romberg chains, beside the fold
            if (can_open && tid >= MAKERS && tid < MAKERS + NC) {
                const uint32_t c = tid - MAKERS;
                int32_t org[K];
                const int level = tree_child_origin<K>(tb, s_par, c, org);
                double f[NC], raw, rom, err;
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) f[cc] = s_f[c][cc];
                tree_finish_cell<K>(tb, ncell + c, s_par, org, level, f, raw, rom, err);
                s_new[0][c] = raw;
                s_new[1][c] = rom;
            }
        }
        int src = 0;
        for (uint32_t w = nnodes >> 1; w >= 1; w >>= 1) {
            if (tid < w) {
                part[src ^ 1][0][tid] = __dadd_rn(part[src][0][2 * tid], part[src][0][2 * tid + 1]);
                part[src ^ 1][1][tid] = __dadd_rn(part[src][1][2 * tid], part[src][1][2 * tid + 1]);
            }
            src ^= 1;
            __syncthreads();
        }
        const double m_tot = part[src][0][0], err_tot = sqrt(part[src][1][0]);
        if (__ddiv_rn(err_tot, m_tot) < tree_tol) { status = TREE_CONVERGED; break; }     // uniform: shared values
        if (!can_open) { status = s_stop; break; }
        bad |= s_bad;
        // (3) np.delete(a, worst) — the tail moves left by one, SHIFT elements per thread and round
        // (loads of a round, barrier, stores) — and np.concatenate((a, children))
        for (uint32_t base = worst + 1; base < n; base += REFINE_THREADS * SHIFT) {
            int32_t vi[SHIFT];
            double vr[SHIFT], vw[SHIFT];
#pragma unroll
            for (int j = 0; j < SHIFT; ++j) {
                const uint32_t i = base + j * REFINE_THREADS + tid;
                if (i < n) { vi[j] = leaf_ids[i]; vr[j] = m_rom[i]; vw[j] = m_raw[i]; }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < SHIFT; ++j) {
                const uint32_t i = base + j * REFINE_THREADS + tid;
                if (i < n) { leaf_ids[i - 1] = vi[j]; m_rom[i - 1] = vr[j]; m_raw[i - 1] = vw[j]; }
            }
        }
        __syncthreads();
        if (tid < NC) {
            leaf_ids[n - 1 + tid] = (int32_t)(ncell + tid);
            m_raw[n - 1 + tid] = s_new[0][tid];
            m_rom[n - 1 + tid] = s_new[1][tid];
        }
        __syncthreads();
        n += NC - 1;
        ncell += NC;
        ++done;
    }
    if (tid == 0) {
        if (bad) atomicOr(flags, bad);
        state[0] = n;
        state[1] = ncell;
        state[2] = done;
        state[3] = status;
    }
}

// pdf on a list of points (the host tree builder's vectorised pdf call; parity tests)
template <int K, typename Pdf>
__global__ void __launch_bounds__(256)
points_eval_kernel(Pdf pdf, const double *__restrict__ pts, uint32_t n, double *__restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x[K];
#pragma unroll
        for (int d = 0; d < K; ++d) x[d] = pts[(size_t)i * K + d];
        out[i] = pdf(x);
    }
}

template <int K>
TreeTable<K> make_tree_table(const lint_tree_t *t) {
    TreeTable<K> tb;
    for (int d = 0; d < K; ++d) {
        tb.mins[d] = t->mins[d];
        tb.span[d] = t->maxs[d] - t->mins[d];
    }
    tb.origin = t->origin_dev;
    tb.level = t->level_dev;
    tb.parent = t->parent_dev;
    tb.corners = t->corners_dev;
    tb.mass_raw = t->mass_raw_dev;
    tb.est = t->est_dev;
    return tb;
}

template <int K, typename Pdf>
int tree_open_launch(const lint_tree_t *t, const Pdf &pdf, const int32_t *open_ids, int64_t n_open,
                     int64_t first, double *summary, uint32_t *flags, cudaStream_t st) {
    TreeTable<K> tb = make_tree_table<K>(t);
    const bool root = n_open == 0;
    const uint32_t n_new = root ? 1u : (uint32_t)(n_open << K);
    const int blocks = (int)std::min<uint32_t>((n_new + 127) / 128, 148 * 8);
    tree_open_kernel<K, Pdf><<<blocks, 128, 0, st>>>(tb, pdf, open_ids, n_new, (uint32_t)first, root, summary, flags);
    return check_launch("lint_tree_open");
}

template <int K>
int tree_open_k(const lint_tree_t *t, const lint_gmm_t *gmm, const lint_halos_t *halos, const int32_t *open_ids,
                int64_t n_open, int64_t first, double *summary, uint32_t *flags, cudaStream_t st) {
    if (gmm) return tree_open_launch<K>(t, fill_gmm<K>(gmm), open_ids, n_open, first, summary, flags, st);
    return tree_open_launch<K>(t, fill_halos<K>(halos), open_ids, n_open, first, summary, flags, st);
}

template <int K, typename Pdf>
int tree_refine_launch(const lint_tree_t *t, const Pdf &pdf, int32_t *leaf_ids, double *m_rom, double *m_raw,
                       double *sq, int64_t leaf_cap, double tree_tol, int64_t max_iters, uint32_t *state,
                       uint32_t *flags, cudaStream_t st) {
    TreeTable<K> tb = make_tree_table<K>(t);
    tree_refine_worst_kernel<K, Pdf><<<1, REFINE_THREADS, 0, st>>>(
        tb, pdf, leaf_ids, m_rom, m_raw, sq, (uint32_t)std::min<int64_t>(leaf_cap, REFINE_MAX_LEAVES),
        (uint32_t)t->capacity, tree_tol, (uint32_t)std::min<int64_t>(max_iters, 0x7fffffff), state, flags);
    return check_launch("lint_tree_refine_worst");
}

template <int K>
int tree_refine_k(const lint_tree_t *t, const lint_gmm_t *gmm, const lint_halos_t *halos, int32_t *leaf_ids,
                  double *m_rom, double *m_raw, double *sq, int64_t leaf_cap, double tree_tol, int64_t max_iters,
                  uint32_t *state, uint32_t *flags, cudaStream_t st) {
    if (gmm)
        return tree_refine_launch<K>(t, fill_gmm<K>(gmm), leaf_ids, m_rom, m_raw, sq, leaf_cap, tree_tol, max_iters,
                                     state, flags, st);
    return tree_refine_launch<K>(t, fill_halos<K>(halos), leaf_ids, m_rom, m_raw, sq, leaf_cap, tree_tol, max_iters,
                                 state, flags, st);
}

template <int K>
int points_eval_k(const lint_gmm_t *gmm, const lint_halos_t *halos, const double *pts, int64_t n, double *out,
                  cudaStream_t st) {
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    if (gmm) points_eval_kernel<K><<<blocks, 256, 0, st>>>(fill_gmm<K>(gmm), pts, (uint32_t)n, out);
    else points_eval_kernel<K><<<blocks, 256, 0, st>>>(fill_halos<K>(halos), pts, (uint32_t)n, out);
    return check_launch("lint_points_eval");
}

static int check_pdf(const char *what, int dim, const lint_gmm_t *gmm, const lint_halos_t *halos) {
    if ((gmm != nullptr) == (halos != nullptr))
        return fail(LINT_ERR_BAD_ARG, "%s: exactly one of gmm / halos must be given", what);
    if (dim < 1 || dim > LINT_MAX_DIM) return fail(LINT_ERR_BAD_ARG, "%s: dim %d outside 1..%d", what, dim, LINT_MAX_DIM);
    const int pdim = gmm ? gmm->dim : halos->dim, nc = gmm ? gmm->ncomp : halos->ncomp;
    if (pdim != dim) return fail(LINT_ERR_BAD_ARG, "%s: pdf has %d dimensions, the domain %d", what, pdim, dim);
    if (nc < 1 || nc > LINT_GMM_MAX_COMP)
        return fail(LINT_ERR_BAD_ARG, "%s: %d components outside 1..%d", what, nc, LINT_GMM_MAX_COMP);
    return LINT_OK;
}

}  // namespace lint

using namespace lint;

extern "C" {

int lint_tree_est_stride(void) { return TREE_EST; }

int lint_tree_open(const lint_tree_t *t, const lint_gmm_t *gmm, const lint_halos_t *halos,
                   const int32_t *open_ids_dev, int64_t n_open, int64_t first_new, double *summary_dev,
                   uint32_t *flags_dev, lint_stream_t stream) {
    if (!t) return fail(LINT_ERR_BAD_ARG, "lint_tree_open: NULL tree");
    if (int rc = check_pdf("lint_tree_open", t->dim, gmm, halos)) return rc;
    if (!t->origin_dev || !t->level_dev || !t->parent_dev || !t->corners_dev || !t->mass_raw_dev || !t->est_dev ||
        !summary_dev || !flags_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_tree_open: NULL table / summary / flags pointer");
    if (n_open < 0 || first_new < 0 || (n_open > 0 && !open_ids_dev) || (n_open == 0 && first_new != 0))
        return fail(LINT_ERR_BAD_ARG, "lint_tree_open: bad batch (n_open=%lld, first_new=%lld)", (long long)n_open,
                    (long long)first_new);
    const int64_t n_new = n_open == 0 ? 1 : n_open << t->dim;
    if (first_new + n_new > t->capacity || first_new + n_new >= (1ll << 31))
        return fail(LINT_ERR_BAD_ARG, "lint_tree_open: %lld new cells do not fit the table (capacity %lld)",
                    (long long)n_new, (long long)t->capacity);
    for (int d = 0; d < t->dim; ++d)
        if (!(t->maxs[d] > t->mins[d])) return fail(LINT_ERR_BAD_ARG, "lint_tree_open: empty box along dimension %d", d);
    cudaStream_t st = (cudaStream_t)stream;
    switch (t->dim) {
        case 1: return tree_open_k<1>(t, gmm, halos, open_ids_dev, n_open, first_new, summary_dev, flags_dev, st);
        case 2: return tree_open_k<2>(t, gmm, halos, open_ids_dev, n_open, first_new, summary_dev, flags_dev, st);
        case 3: return tree_open_k<3>(t, gmm, halos, open_ids_dev, n_open, first_new, summary_dev, flags_dev, st);
        case 4: return tree_open_k<4>(t, gmm, halos, open_ids_dev, n_open, first_new, summary_dev, flags_dev, st);
        case 5: return tree_open_k<5>(t, gmm, halos, open_ids_dev, n_open, first_new, summary_dev, flags_dev, st);
        default: return tree_open_k<6>(t, gmm, halos, open_ids_dev, n_open, first_new, summary_dev, flags_dev, st);
    }
}

int lint_tree_refine_worst(const lint_tree_t *t, const lint_gmm_t *gmm, const lint_halos_t *halos,
                           int32_t *leaf_ids_dev, double *m_romberg_dev, double *m_raw_dev, double *scratch_dev,
                           int64_t leaf_capacity, double tree_tol, int64_t max_openings, uint32_t *state_dev,
                           uint32_t *flags_dev, lint_stream_t stream) {
    if (!t) return fail(LINT_ERR_BAD_ARG, "lint_tree_refine_worst: NULL tree");
    if (int rc = check_pdf("lint_tree_refine_worst", t->dim, gmm, halos)) return rc;
    if (!t->origin_dev || !t->level_dev || !t->parent_dev || !t->corners_dev || !t->mass_raw_dev || !t->est_dev ||
        !leaf_ids_dev || !m_romberg_dev || !m_raw_dev || !scratch_dev || !state_dev || !flags_dev)
        return fail(LINT_ERR_BAD_ARG, "lint_tree_refine_worst: NULL pointer");
    if (leaf_capacity < 1 || max_openings < 0 || t->capacity >= (1ll << 31) || !(tree_tol == tree_tol))
        return fail(LINT_ERR_BAD_ARG, "lint_tree_refine_worst: bad capacity / tolerance");
    cudaStream_t st = (cudaStream_t)stream;
    switch (t->dim) {
        case 1: return tree_refine_k<1>(t, gmm, halos, leaf_ids_dev, m_romberg_dev, m_raw_dev, scratch_dev, leaf_capacity, tree_tol, max_openings, state_dev, flags_dev, st);
        case 2: return tree_refine_k<2>(t, gmm, halos, leaf_ids_dev, m_romberg_dev, m_raw_dev, scratch_dev, leaf_capacity, tree_tol, max_openings, state_dev, flags_dev, st);
        case 3: return tree_refine_k<3>(t, gmm, halos, leaf_ids_dev, m_romberg_dev, m_raw_dev, scratch_dev, leaf_capacity, tree_tol, max_openings, state_dev, flags_dev, st);
        case 4: return tree_refine_k<4>(t, gmm, halos, leaf_ids_dev, m_romberg_dev, m_raw_dev, scratch_dev, leaf_capacity, tree_tol, max_openings, state_dev, flags_dev, st);
        case 5: return tree_refine_k<5>(t, gmm, halos, leaf_ids_dev, m_romberg_dev, m_raw_dev, scratch_dev, leaf_capacity, tree_tol, max_openings, state_dev, flags_dev, st);
        default: return tree_refine_k<6>(t, gmm, halos, leaf_ids_dev, m_romberg_dev, m_raw_dev, scratch_dev, leaf_capacity, tree_tol, max_openings, state_dev, flags_dev, st);
    }
}

int lint_tree_max_refine_leaves(void) { return (int)REFINE_MAX_LEAVES; }

int lint_points_eval(int dim, const lint_gmm_t *gmm, const lint_halos_t *halos, const double *points_dev, int64_t n,
                     double *out_dev, lint_stream_t stream) {
    if (int rc = check_pdf("lint_points_eval", dim, gmm, halos)) return rc;
    if (n < 0 || n >= (1ll << 32) || (n > 0 && (!points_dev || !out_dev)))
        return fail(LINT_ERR_BAD_ARG, "lint_points_eval: bad point list");
    if (n == 0) return LINT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dim) {
        case 1: return points_eval_k<1>(gmm, halos, points_dev, n, out_dev, st);
        case 2: return points_eval_k<2>(gmm, halos, points_dev, n, out_dev, st);
        case 3: return points_eval_k<3>(gmm, halos, points_dev, n, out_dev, st);
        case 4: return points_eval_k<4>(gmm, halos, points_dev, n, out_dev, st);
        case 5: return points_eval_k<5>(gmm, halos, points_dev, n, out_dev, st);
        default: return points_eval_k<6>(gmm, halos, points_dev, n, out_dev, st);
    }
}

}  // extern "C"
