/*
 * lintsampler_b200 — C ABI of the B200-native linear-interpolant sampler.
 *
 * The reference (aneeshnaik/lintsampler) is pure Python/NumPy and has no FFI
 * of its own.  Each entry point below replaces one NumPy
 * whole-array step of the reference's `LintSampler.sample` hot path and cites
 * it as `file:line` relative to /root/reference/src/lintsampler.  The Python
 * host layer (lintsampler_b200/*.py) binds these with ctypes; INTEGRATION.md
 * shows the equivalent stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions across the boundary.
 *   - Every `*_dev` / descriptor pointer is a DEVICE pointer owned by the
 *     caller; the library never allocates or frees device memory.  Scratch is
 *     caller-provided, sized by the matching `*_scratch_bytes` query.
 *   - `stream` is a cudaStream_t passed as void*.  All work is enqueued on it;
 *     no call synchronises the host.  Device-side findings (negative or
 *     non-finite densities, zero total mass) are written to caller-provided
 *     device words the caller reads after synchronising.
 *   - Return value: LINT_OK or an error code; `lint_last_error()` returns a
 *     thread-local message for the last non-OK return.
 *   - Row-major (C-order) everywhere, as NumPy in the reference.
 *   - Corner order: corner c of a k-cube is offset along dim d by bit (k-1-d)
 *     of c (dim 0 = MSB)  [density_structures/grid.py:456-459, utils.py:37-74].
 *   - Uniform layout: a row of k+1 uniforms; columns 0..k-1 drive the in-cell
 *     coordinates, column k selects the cell  [lintsampler.py:269-270].
 */
#ifndef LINTSAMPLER_B200_H
#define LINTSAMPLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LINT_MAX_DIM 6          /* reference advises <= 6 dims (paper/paper.md:84) */
#define LINT_ABI_VERSION 2

typedef void *lint_stream_t;    /* cudaStream_t */

enum lint_status {
    LINT_OK = 0,
    LINT_ERR_BAD_ARG = 1,
    LINT_ERR_UNSUPPORTED = 2,
    LINT_ERR_CUDA = 3
};

/* precision mode of a domain: arithmetic the in-cell solve runs in, dtype of the
 * stored densities and dtype of the emitted samples.  The cell CDF and the
 * cell-selecting uniform are ALWAYS fp64 / 53-bit. */
enum lint_dtype {
    LINT_F32 = 0,   /* densities fp32, samples fp32, stable fp32 in-cell solve   */
    LINT_F64 = 1    /* densities fp64, samples fp64, the reference's literal fp64
                       expression order (bit-identical to NumPy)                 */
};

/* bits of the device status word written by lint_grid_masses / lint_cells_check */
#define LINT_FLAG_NEGATIVE   1u  /* a density < 0       [grid.py:401-402, tree.py:598-601] */
#define LINT_FLAG_NONFINITE  2u  /* a NaN/Inf density   [grid.py:403-404, tree.py:602-605] */

/* CDF construction modes for lint_cdf_build */
enum lint_cdf_mode {
    LINT_CDF_PARALLEL = 0,   /* deterministic tiled reduce-then-scan (throughput)       */
    LINT_CDF_SEQUENTIAL = 1  /* NumPy-identical: pairwise m.sum(), p=m/sum, strictly
                                sequential cumsum, /= last  [grid.py:431, utils.py:195-196] */
};

/* Rectilinear grid domain == reference DensityGrid  [grid.py:253-307]. */
typedef struct lint_grid {
    int32_t dim;                          /* k, 1..LINT_MAX_DIM                          */
    int32_t dtype;                        /* enum lint_dtype                             */
    int64_t ncells[LINT_MAX_DIM];         /* cells per dim (grid.shape)                  */
    const void *vertex_densities_dev;     /* (n0+1,..,nk-1+1), fp32 or fp64 per dtype     */
    const double *edges_dev[LINT_MAX_DIM];/* edges_dev[d]: n_d+1 fp64 edge coordinates   */
    const double *cdf_dev;                /* prod(ncells) fp64, normalised inclusive CDF */
    const uint32_t *guide_dev;            /* optional 2^guide_bits+1 entry guide table   */
    int32_t guide_bits;                   /* 0 = no guide table                          */
    int32_t reserved;
    const void *cell_records_dev;         /* optional (prod(ncells), 2^k) per-cell corner
                                             densities in corner order (lint_grid_records);
                                             NULL = gather corners from vertex_densities   */
} lint_grid_t;

/* Flat cell list == the leaves of a reference DensityTree in list order
 * [tree.py:334-355], or any user DensityStructure flattened by the host. */
typedef struct lint_cells {
    int32_t dim;
    int32_t dtype;
    int64_t ncells;                       /* L                                           */
    const double *mins_dev;               /* (L,k)  leaf.x                               */
    const double *widths_dev;             /* (L,k)  leaf.dx                              */
    const void *corners_dev;              /* (L,2^k) leaf.corner_densities, per dtype    */
    const double *cdf_dev;                /* (L) normalised inclusive CDF of leaf masses */
    const uint32_t *guide_dev;
    int32_t guide_bits;
    int32_t reserved;
    const void *cell_records_dev;         /* optional packed per-cell records built by
                                             lint_cells_records (corners + box in one aligned
                                             record); NULL = read the three arrays above     */
} lint_cells_t;

/* Scrambled Sobol state read from a host scipy.stats.qmc.Sobol(d=k+1, bits=32)
 * engine (`_sv`, `_shift`, `num_generated`)  [lintsampler.py:465,515]. */
typedef struct lint_sobol {
    uint32_t sv[LINT_MAX_DIM + 1][32];    /* HOST: scrambled direction numbers           */
    uint32_t shift[LINT_MAX_DIM + 1];     /* HOST: digital shift                         */
    uint64_t first_index;                 /* index of the first point to emit            */
    uint32_t sorted_bits;                 /* 0: rows are points first_index, first_index+1, ...
                                             m: rows are the first 2^m points enumerated with the
                                             cell-selecting coordinate increasing (needs N == 2^m,
                                             first_index == 0 and sorted_ainv)             */
    uint32_t sorted_ainv[32];             /* HOST: column r = A^-1 e_r over GF(2), A = top m x m
                                             block of the cell coordinate's generator matrix */
} lint_sobol_t;

/* Gaussian mixture sum_j w_j N(x; mu_j, Sigma_j), passed as HOST arrays:
 *   log_norm[j] = log w_j - 0.5 log det Sigma_j - 0.5 k log(2 pi)
 *   mean[j*k + d]
 *   prec_chol[j*k*k + r*k + c] = U_j[r][c], upper triangular, with
 *   Sigma_j^-1 = U_j^T U_j, so the Mahalanobis term is |U_j (x - mu_j)|^2. */
typedef struct lint_gmm {
    int32_t dim;
    int32_t ncomp;                        /* 1..LINT_GMM_MAX_COMP                        */
    const double *log_norm;
    const double *mean;
    const double *prec_chol;
} lint_gmm_t;
#define LINT_GMM_MAX_COMP 16

/* Sum of spherical double-power-law profiles (the NFW family of the reference's
 * docs/source/example_notebooks/3_dark_matter.ipynb, cells 11 and 16), HOST arrays:
 *   rho(x) = sum_h rho0[h] / ((q + eps)^gamma * (1 + q^alpha)^((beta - gamma)/alpha)),
 *   q = |x - centre[h]| / r_s[h];   a halo contributes 0 where q > q_cut (q_cut <= 0: never).
 * NFW as in the notebook: alpha 1, beta 3, gamma 1, eps 0.1, q_cut 10. */
typedef struct lint_halos {
    int32_t dim;
    int32_t ncomp;                        /* 1..LINT_GMM_MAX_COMP                        */
    const double *rho0;                   /* [ncomp]                                     */
    const double *centre;                 /* [ncomp * dim]                               */
    const double *r_s;                    /* [ncomp], > 0                                */
    double alpha, beta, gamma;
    double eps;
    double q_cut;
} lint_halos_t;

/* ---- library ---------------------------------------------------------------- */
int lint_abi_version(void);
const char *lint_last_error(void);

/* Device-resident stream position (for CUDA graphs).  While a counter is set, every sampling call
 * draws at Philox offset `offset + *counter_dev` (read on the device when the kernels run), so a
 * captured graph that contains lint_advance_counter draws fresh samples at every replay.
 * Process-wide setting, not thread-safe; NULL clears it. */
int lint_set_offset_counter(const uint64_t *counter_dev);
int lint_advance_counter(uint64_t *counter_dev, uint64_t by, lint_stream_t stream);

/* ---- grid build ------------------------------------------------------------- */

/* Fused vertex evaluation of a built-in analytic pdf on the grid vertices;
 * replaces the meshgrid + `pdf(edgegrid)` of grid.py:384-398 without ever
 * materialising the (npts,k) coordinate array.  Writes g->vertex_densities_dev
 * (cast away const) in g->dtype; densities are multiplied by `scale` first. */
int lint_grid_eval_gmm(const lint_grid_t *g, const lint_gmm_t *gmm, double scale,
                       lint_stream_t stream);
/* The same for a sum of double-power-law halos (3_dark_matter.ipynb cells 11, 16). */
int lint_grid_eval_halos(const lint_grid_t *g, const lint_halos_t *halos, double scale,
                         lint_stream_t stream);

/* Trapezoid cell masses: mean of the 2^k corner densities (summed in corner
 * order from zero, one division by 2^k) times the cell volume ((w0*w1)*w2...)
 * [grid.py:469-498, 501-518, 411].  masses_dev: prod(ncells) fp64.
 * Also performs the density validity check of grid.py:401-404 on every vertex:
 * ORs LINT_FLAG_* into *flags_dev (caller zeroes it first). */
int lint_grid_masses(const lint_grid_t *g, double *masses_dev, uint32_t *flags_dev,
                     lint_stream_t stream);

/* Per-cell corner records: for every cell, its 2^k corner densities contiguous in
 * corner order — the gather of grid.py:440-467 done once per build instead of
 * once per sample, so a sample reads one record (one 32-byte sector at k=3 fp32)
 * instead of 2^(k-1) scattered vertex pairs.  records_dev: prod(ncells)*2^k
 * values in g->dtype.  LINT_F64 records hold the densities themselves; LINT_F32
 * records hold them divided by the cell's largest corner (the in-cell draw depends
 * on ratios only), which spares the sampler a per-sample normalisation.
 * Optional: set g->cell_records_dev to use them. */
int lint_grid_records(const lint_grid_t *g, void *records_dev, lint_stream_t stream);

/* Packed per-cell records of a cell list: [2^k corners][k lower bounds][k upper bounds =
 * fl(x + dx), tree.py:348] in c->dtype (LINT_F32: corners divided by the cell's largest, as
 * for lint_grid_records), each record padded to a multiple of 32 bytes
 * (lint_cells_record_bytes) — the per-sample copies of tree.py:345-350 done once per
 * build.  records_dev: ncells * lint_cells_record_bytes(dim, dtype) bytes. */
size_t lint_cells_record_bytes(int dim, int dtype);
int lint_cells_records(const lint_cells_t *c, void *records_dev, lint_stream_t stream);

/* Same validity check for a cell list's corner densities [tree.py:598-605]. */
int lint_cells_check(const lint_cells_t *c, uint32_t *flags_dev, lint_stream_t stream);

/* Scratch needed by lint_cdf_build / lint_sum_numpy_f64 for n values. */
size_t lint_cdf_scratch_bytes(int64_t n);

/* Bit-exact NumPy `np.sum` of a contiguous fp64 array (pairwise summation with
 * 8 accumulators, 128-element leaves)  [grid.py:412 `np.sum(self.masses)`]. */
int lint_sum_numpy_f64(const double *x_dev, int64_t n, double *out_dev,
                       void *scratch_dev, size_t scratch_bytes, lint_stream_t stream);

/* Normalised inclusive CDF of `masses`  [grid.py:431 + utils.py:195-196; also
 * tree.py:335 and lintsampler.py:286-290].  total_dev receives the total mass
 * (NumPy-identical in SEQUENTIAL mode).  cdf_dev[n-1] is exactly 1.0. */
int lint_cdf_build(const double *masses_dev, int64_t n, int mode, double *cdf_dev,
                   double *total_dev, void *scratch_dev, size_t scratch_bytes,
                   lint_stream_t stream);

/* Parallel build for a grid: cell masses [grid.py:469-518, 411] -> normalised inclusive CDF
 * [grid.py:431 + utils.py:195-196] -> guide table, hierarchical and without ever storing the
 * masses (masses_dev == NULL).  The cells are cut into groups of consecutive cell rows (~2 K
 * cells); group totals come from one streaming pass over the vertex table (separable trapezoid
 * weights; it also performs the validity check of grid.py:401-404 on every vertex), and inside
 * a group the CDF is the exact chained scan of the cell masses laid on the group's offset:
 *     cdf_i = min(min(Off[g] + L_i, Off[g + 1]) * (1 / total), 1),   cdf[n - 1] = 1.
 * Non-decreasing, equal to its predecessor on zero-mass cells (utils.py:197 never selects them),
 * bit-reproducible, within ~1e-15 of the NumPy-sequential table.
 * [u_lo, u_hi) is the stratum of the cell-selecting uniform the caller will draw from ((0, 1) =
 * whole domain): every rank of a sharded run does the cheap vertex pass (total mass and group
 * offsets, identical on all ranks, no communication) but scans, and writes cdf_dev / masses_dev /
 * guide_dev for, only the groups its stratum can touch plus one group of margin.  range_dev
 * (4 words, device) receives that slab: [0] first cell, [1] end cell, [2] first group, [3] end
 * group; hand it to lint_grid_sample_cellmajor.  Outside the slab (and the one CDF entry before
 * it) the tables are left untouched.  total_dev receives the total mass. */
size_t lint_grid_build_scratch_bytes(const lint_grid_t *g);
/* cells per group of lint_grid_build (whole cell rows, about 2 K cells): group b starts at cell
 * b * lint_grid_group_cells(g).  A stratum cut at a group boundary has no straddling cell. */
int64_t lint_grid_group_cells(const lint_grid_t *g);
int lint_grid_build(const lint_grid_t *g, double u_lo, double u_hi, double *cdf_dev, double *total_dev,
                    uint32_t *guide_dev, int guide_bits, double *masses_dev, uint32_t *flags_dev,
                    uint32_t *range_dev, void *scratch_dev, size_t scratch_bytes, lint_stream_t stream);

/* Guide (index) table that shortens the searchsorted of utils.py:197 without
 * changing its result: guide[j] & 0x7fffffff = first i with cdf[i] > j / 2^guide_bits
 * for j = 0..2^guide_bits (inclusive; last entry n-1); bit 31 is set when
 * guide[j+1] names the same cell, i.e. the entry alone answers every query in the
 * bin.  guide_dev: 2^guide_bits+1 entries. */
int lint_guide_build(const double *cdf_dev, int64_t n, int guide_bits,
                     uint32_t *guide_dev, lint_stream_t stream);

/* ---- sampling: uniforms supplied by the caller (parity mode) ------------------ */

/* == sampling._grid_sample(grid, u)  [sampling.py:97-127] on a DensityGrid:
 * searchsorted(side='right') on u[:,k] [utils.py:197], C-order unravel
 * [grid.py:437], 2^k corner gather [grid.py:440-467], k sequential 1-D inverse
 * CDFs [sampling.py:41-94, 6-38], affine map [sampling.py:126].
 * u_dev: (N,k+1) fp64.  out_dev: (N,k) in g->dtype.  cell_idx_dev: optional
 * (N) int64 flat C-order cell index. */
int lint_grid_sample_u(const lint_grid_t *g, const double *u_dev, int64_t N,
                       void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);

/* Same over a cell list  [tree.py:314-355 + sampling.py:123-126]. */
int lint_cells_sample_u(const lint_cells_t *c, const double *u_dev, int64_t N,
                        void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);

/* In-cell stage only, for user-defined DensityStructure subclasses whose
 * choose_cells ran on the host  [base.py:11-30, sampling.py:123-126].
 * mins/maxs: (N,k) fp64; corners: (2^k,N) fp64; u: (N,k+1) fp64 (column k
 * unused); out: (N,k) fp64. */
int lint_incell_sample_u(int dim, int64_t N, const double *mins_dev, const double *maxs_dev,
                         const double *corners_dev, const double *u_dev, double *out_dev,
                         lint_stream_t stream);

/* ---- sampling: device-generated uniforms --------------------------------------- */

/* Philox4x32-10 stream replacing `rng.random((N,k+1))` [lintsampler.py:517]:
 * row i of the call uses counter (offset+i) under key `seed`; successive calls
 * continue the stream by advancing `offset` by N.  The cell-selecting uniform
 * has 53 bits; in-cell uniforms have 53 bits (F64) or 24 bits (F32).
 * [u_lo, u_hi) is the stratum of the cell-selecting uniform this call draws
 * from: (0, 1) for the whole domain; a rank of a sharded run passes its own
 * slice of the CDF, which is the u-rescaling of lintsampler.py:306 inverted. */
int lint_grid_sample_philox(const lint_grid_t *g, int64_t N, uint64_t seed, uint64_t offset,
                            double u_lo, double u_hi, void *out_dev, int64_t *cell_idx_dev,
                            lint_stream_t stream);
int lint_cells_sample_philox(const lint_cells_t *c, int64_t N, uint64_t seed, uint64_t offset,
                             double u_lo, double u_hi, void *out_dev, int64_t *cell_idx_dev,
                             lint_stream_t stream);

/* The uniforms the Philox samplers consume, materialised as (N,k+1) fp64 — for
 * tests and for callers that want to replay a draw on the host. */
int lint_philox_uniforms(int dim, int dtype, int64_t N, uint64_t seed, uint64_t offset,
                         double *u_dev, lint_stream_t stream);

/* Scrambled Sobol points by direct Gray-code index, bit-identical to
 * scipy.stats.qmc.Sobol(d=k+1, bits=32).random(N)  [lintsampler.py:465,515]. */
int lint_grid_sample_sobol(const lint_grid_t *g, int64_t N, const lint_sobol_t *sobol,
                           void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);
int lint_cells_sample_sobol(const lint_cells_t *c, int64_t N, const lint_sobol_t *sobol,
                            void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);
int lint_sobol_uniforms(int dim, int64_t N, const lint_sobol_t *sobol, double *u_dev,
                        lint_stream_t stream);

/* ---- sampling: a list of grids as one domain, one launch ------------------------------ */

/* == the list-of-domains branch of LintSampler.sample [lintsampler.py:281-307] for grids of one
 * dimension and dtype: per row, the cell-selecting uniform picks the grid (`_choice` on the
 * normalised cumulative grid masses, :286-290 — choice_cdf[i] = c[i] / c[-1], c = p.cumsum()), is
 * rescaled into that grid's interval exactly as :306 does, (u - bounds[i]) / (bounds[i+1] - bounds[i])
 * with bounds = np.append(0, p.cumsum()) (:293-297), and then selects the cell and the in-cell
 * position as lint_grid_sample_*.  `grids`, `choice_cdf` (ndomains), `bounds` (ndomains + 1) are HOST
 * arrays; 1 <= ndomains <= 16.  Rows come out in the order of the uniforms (draw order), not
 * grouped by grid as the reference's vstack leaves them: cell_idx_dev numbers the cells through
 * the concatenation of the grids, so the caller can regroup (a stable partition by grid
 * reproduces the reference's row order exactly).  scratch_dev: lint_multi_scratch_bytes. */
size_t lint_multi_scratch_bytes(int dim, int ndomains);
int lint_multi_grid_sample_u(const lint_grid_t *grids, int ndomains, const double *choice_cdf, const double *bounds,
                             const double *u_dev, int64_t N, void *scratch_dev, size_t scratch_bytes, void *out_dev,
                             int64_t *cell_idx_dev, lint_stream_t stream);
int lint_multi_grid_sample_philox(const lint_grid_t *grids, int ndomains, const double *choice_cdf,
                                  const double *bounds, int64_t N, uint64_t seed, uint64_t offset, void *scratch_dev,
                                  size_t scratch_bytes, void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);
int lint_multi_grid_sample_sobol(const lint_grid_t *grids, int ndomains, const double *choice_cdf,
                                 const double *bounds, int64_t N, const lint_sobol_t *sobol, void *scratch_dev,
                                 size_t scratch_bytes, void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);

/* ---- sampling: cell-major throughput mode ----------------------------------------- */

/* N i.i.d. draws as a multiset, emitted grouped by cell (C order of the flat cell
 * index) instead of in draw order.  The per-cell counts are an exact
 * Multinomial(N, p) draw over the CDF of grid.py:431 / utils.py:195-196 — the
 * distribution `_choice` realises one row at a time, utils.py:179-198: independent
 * Poisson counts of total mean N - 8 sqrt(N) (given their total T they are
 * Multinomial(T, p)) plus N - T rows drawn one by one, the "top-up" (T > N is an
 * 8-sigma event and is clamped); each cell's rows are then drawn from its k-linear
 * interpolant (sampling.py:41-94, 126) with the box and corners loaded once per
 * cell.  No per-sample search or gather.  Row order is NOT exchangeable: apply a
 * row permutation (lintsampler.py:309-311) when the caller relies on it.
 * Philox4x32-10 keyed by `seed`; `offset` separates successive calls.
 * [u_lo, u_hi) restricts the draw to a stratum of the cell CDF exactly as in
 * lint_grid_sample_philox ((0, 1) = whole domain; a cell straddling a stratum
 * boundary is shared in proportion to the overlap).
 * cell_range_dev (grid only): NULL, or the slab written by lint_grid_build for this stratum — the
 * count and scan passes then visit those cells only, and the top-up rows search that slab.  The
 * guide table is optional here (guide_bits 0): only the top-up rows search the CDF at all.
 * Row order: the flat parts of the split cells grouped by cell, then all other counted rows
 * grouped by cell, then the top-up rows in draw order (see csrc/lint_cellmajor.cuh).
 * N <= 2^32 - 2^16, fewer than 2^29 cells, out_dev 16-byte aligned.
 * Scratch: lint_cellmajor_scratch_bytes. */
size_t lint_cellmajor_scratch_bytes(int64_t ncells, int64_t N);
int lint_grid_sample_cellmajor(const lint_grid_t *g, int64_t N, uint64_t seed, uint64_t offset,
                               double u_lo, double u_hi, const uint32_t *cell_range_dev, void *scratch_dev,
                               size_t scratch_bytes, void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream);
int lint_cells_sample_cellmajor(const lint_cells_t *c, int64_t N, uint64_t seed, uint64_t offset,
                                double u_lo, double u_hi, void *scratch_dev, size_t scratch_bytes, void *out_dev,
                                int64_t *cell_idx_dev, lint_stream_t stream);

/* Measurement aid: with timing enabled every cell-major call records CUDA events on its stream
 * around the flat emitter and around the general emitter.  lint_cellmajor_last_times waits for
 * the last timed call and returns [0] ms of the flat emitter, [1] ms of the general emitter,
 * [2] rows of the flat run T_f, [3] rows of the general run T_r, [4] T_f + T_r (first top-up
 * row).  Process-wide, not thread-safe: for benchmarks. */
int lint_cellmajor_timing(int enable);
int lint_cellmajor_last_times(double *out5);

/* ---- tree build (fused pdf families) ------------------------------------------------ */

/* The cells of a DensityTree as a device table in structure-of-arrays form (caller-owned buffers
 * of `capacity` cells; cell 0 is the root).  Replaces the _TreeCell objects of
 * density_structures/tree.py:544-703 for pdfs the library can evaluate itself. */
#define LINT_TREE_MAX_LEVEL 30            /* lattice coordinates are int32 */
typedef struct lint_tree {
    int32_t dim;
    int32_t reserved;
    double mins[LINT_MAX_DIM], maxs[LINT_MAX_DIM];   /* the root box (HOST values) */
    int64_t capacity;
    int32_t *origin_dev;    /* (capacity, dim) integer lattice coordinates of the cell at its level */
    int32_t *level_dev;     /* (capacity) */
    int32_t *parent_dev;    /* (capacity) parent cell, -1 for the root */
    double *corners_dev;    /* (capacity, 2^dim) corner densities, corner order (dim 0 = most significant bit) */
    double *mass_raw_dev;   /* (capacity) trapezoid mass, tree.py:608 */
    double *est_dev;        /* (capacity, lint_tree_est_stride()) Romberg estimates e[0..level], tree.py:649-703 */
} lint_tree_t;

int lint_tree_est_stride(void);

/* == _TreeCell.open + _evaluate_cell + _calculate_romberg_estimates for a batch [tree.py:615-647,
 * 593-613, 649-703]: creates the 2^dim children of the n_open cells open_ids_dev[] (device int32,
 * in the order the caller opens them) as cells first_new, first_new + 1, ... (children of one cell
 * contiguous, in corner order); n_open == 0 creates the root (first_new must be 0).  Per new cell:
 * pdf on its corners at mins + coords * ((maxs - mins) / 2^level), mass_raw = vol * mean(corners),
 * the Richardson chain over its ancestors' corner tables — NumPy's operation order throughout.
 * summary_dev: (new cells, 3) fp64 = mass_raw, mass_romberg (= e[-1]), |e[-1] - e[-2]|, all a caller
 * needs to pick the next cells to open (tree.py:413-488).  Exactly one of gmm / halos is non-NULL
 * (HOST structs).  ORs LINT_FLAG_* into *flags_dev for negative / non-finite densities
 * (tree.py:598-605).  Levels beyond LINT_TREE_MAX_LEVEL are the caller's to refuse. */
int lint_tree_open(const lint_tree_t *t, const lint_gmm_t *gmm, const lint_halos_t *halos,
                   const int32_t *open_ids_dev, int64_t n_open, int64_t first_new, double *summary_dev,
                   uint32_t *flags_dev, lint_stream_t stream);

/* == stage 2 of DensityTree.refine [tree.py:448-488] in one launch: while
 * sqrt(sum((m_romberg - m_raw)^2)) / sum(m_romberg) >= tree_tol over the leaves (both sums NumPy's
 * pairwise np.sum of the current leaf arrays), open the leaf with the largest squared difference
 * (np.argmax: first of equals), delete it from the leaf arrays and append its children — the cells
 * and the order of the reference's loop.  leaf_ids_dev / m_romberg_dev / m_raw_dev: the leaves in
 * leaf order, device arrays of leaf_capacity entries; scratch_dev: leaf_capacity doubles.
 * state_dev (4 x uint32): in  [0] leaves, [1] cells in the table;
 *                         out [0], [1] updated, [2] openings done, [3] 0 converged, 1 stopped for room
 * (a table or leaf array full, max_openings reached, more than lint_tree_max_refine_leaves() leaves:
 * grow and call again, or finish on the host), 2 level limit reached.  New cells are appended to the
 * table as by lint_tree_open. */
int lint_tree_refine_worst(const lint_tree_t *t, const lint_gmm_t *gmm, const lint_halos_t *halos,
                           int32_t *leaf_ids_dev, double *m_romberg_dev, double *m_raw_dev, double *scratch_dev,
                           int64_t leaf_capacity, double tree_tol, int64_t max_openings, uint32_t *state_dev,
                           uint32_t *flags_dev, lint_stream_t stream);
int lint_tree_max_refine_leaves(void);

/* The same pdf at a list of points, points_dev (n, dim) fp64 -> out_dev (n) fp64: what a
 * vectorised `pdf(x)` call of the reference returns (grid.py:398, tree.py:816), with the
 * arithmetic of lint_grid_eval_* / lint_tree_open. */
int lint_points_eval(int dim, const lint_gmm_t *gmm, const lint_halos_t *halos, const double *points_dev, int64_t n,
                     double *out_dev, lint_stream_t stream);

/* ---- row order ------------------------------------------------------------------- */

/* out[i] = in[pi(i)] for a pseudo-random bijection pi of [0, N) keyed by `seed`
 * (8-round Feistel network + cycle walking): the device counterpart of the
 * reference's final `rng.shuffle(X, axis=0)`  [lintsampler.py:309-311].  Rows
 * are k values of `dtype`.  in and out must not overlap. */
int lint_permute_rows(const void *in_dev, void *out_dev, int64_t N, int32_t k, int32_t dtype,
                      uint64_t seed, lint_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LINTSAMPLER_B200_H */
