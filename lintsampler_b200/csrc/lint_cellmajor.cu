// C ABI of the cell-major ("grouped") sampling path.  The kernels live in
// lint_cellmajor.cuh and are compiled per dimension by lint_cellmajor_inst.cu.
#include <cstdlib>
#include "lint_device.cuh"
#include "lint_cellmajor_plan.h"

namespace lint {

int validate_grid(const lint_grid_t *g, bool need_cdf, uint64_t *ncells_out, uint64_t *nverts_out);

EmitTimers &emit_timers() {
    static EmitTimers t;
    return t;
}

Lanes *side_lanes() {
    static Lanes table[64];
    static const bool off = [] { const char *v = getenv("LINT_SIDE"); return v && *v == '0'; }();
    if (off || emit_timers().enabled) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    Lanes &l = table[dev];
    if (!l.side[0]) {
        for (auto &s : l.side)
            if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (auto &e : l.ev)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &l;
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
    if (N > 0 && ((uintptr_t)out & 15)) return fail(LINT_ERR_BAD_ARG, "%s: out_dev must be 16-byte aligned", what);
    if (!scratch || scratch_bytes < carve(nullptr, ncells, (uint64_t)N).bytes)
        return fail(LINT_ERR_BAD_ARG, "%s: scratch too small", what);
    return LINT_OK;
}

// defined in lint_sample.cu: u-driven rows [*first_row_dev, N), at most max_rows of them
int topup_grid(const lint_grid_t *g, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
               uint64_t seed, uint64_t offset, double u_lo, double u_hi, const uint32_t *cell_range_dev, void *out,
               int64_t *cell_idx, cudaStream_t st);
int topup_cells(const lint_cells_t *c, int64_t N, const uint32_t *first_row_dev, int64_t max_rows,
                uint64_t seed, uint64_t offset, double u_lo, double u_hi, void *out, int64_t *cell_idx,
                cudaStream_t st);

}  // namespace lint

using namespace lint;

extern "C" {

int lint_cellmajor_timing(int enable) {
    EmitTimers &t = emit_timers();
    if (enable && !t.ev[0]) {
        for (auto &e : t.ev)
            if (cudaEventCreate(&e) != cudaSuccess) return fail(LINT_ERR_CUDA, "lint_cellmajor_timing: cudaEventCreate failed");
    }
    t.enabled = enable != 0;
    return LINT_OK;
}

int lint_cellmajor_last_times(double *out5) {
    EmitTimers &t = emit_timers();
    if (!out5) return fail(LINT_ERR_BAD_ARG, "lint_cellmajor_last_times: out is NULL");
    if (!t.enabled || !t.plan_dev) return fail(LINT_ERR_BAD_ARG, "lint_cellmajor_last_times: no timed call yet");
    if (cudaEventSynchronize(t.ev[3]) != cudaSuccess) return fail(LINT_ERR_CUDA, "lint_cellmajor_last_times: sync failed");
    float a = 0.f, b = 0.f;
    if (t.flat_ran) cudaEventElapsedTime(&a, t.ev[0], t.ev[1]);
    cudaEventElapsedTime(&b, t.ev[2], t.ev[3]);
    uint32_t plan[3] = {0, 0, 0};
    if (cudaMemcpy(plan, t.plan_dev, sizeof plan, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(LINT_ERR_CUDA, "lint_cellmajor_last_times: reading the row budget failed");
    out5[0] = a; out5[1] = b; out5[2] = plan[0]; out5[3] = plan[1]; out5[4] = plan[2];
    return LINT_OK;
}

size_t lint_cellmajor_scratch_bytes(int64_t ncells, int64_t N) {
    if (ncells < 1 || N < 0) return 0;
    return carve(nullptr, (uint64_t)ncells, (uint64_t)N).bytes;
}

int lint_grid_sample_cellmajor(const lint_grid_t *g, int64_t N, uint64_t seed, uint64_t offset,
                               double u_lo, double u_hi, const uint32_t *cell_range_dev, void *scratch_dev,
                               size_t scratch_bytes, void *out_dev, int64_t *cell_idx_dev, lint_stream_t stream) {
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
        case 1: rc = grid_cellmajor_k<1>(g, nc, N, key, offset, u, cell_range_dev, s, out_dev, cell_idx_dev, st); break;
        case 2: rc = grid_cellmajor_k<2>(g, nc, N, key, offset, u, cell_range_dev, s, out_dev, cell_idx_dev, st); break;
        case 3: rc = grid_cellmajor_k<3>(g, nc, N, key, offset, u, cell_range_dev, s, out_dev, cell_idx_dev, st); break;
        case 4: rc = grid_cellmajor_k<4>(g, nc, N, key, offset, u, cell_range_dev, s, out_dev, cell_idx_dev, st); break;
        case 5: rc = grid_cellmajor_k<5>(g, nc, N, key, offset, u, cell_range_dev, s, out_dev, cell_idx_dev, st); break;
        default: rc = grid_cellmajor_k<6>(g, nc, N, key, offset, u, cell_range_dev, s, out_dev, cell_idx_dev, st); break;
    }
    if (rc) return rc;
    // rows [T, N): the u-driven sampler on the same stratum, beside the emitters
    Lanes *L = side_lanes();
    cudaStream_t st_top = st;
    if (L) {
        cudaStreamWaitEvent(L->side[1], L->ev[4], 0);           // the row budget is known (grid_cellmajor_k)
        st_top = L->side[1];
    }
    rc = topup_grid(g, N, s.plan + 2, (int64_t)topup_bound((uint64_t)N), seed, offset, u_lo, u_hi, cell_range_dev,
                    out_dev, cell_idx_dev, st_top);
    if (L) {
        cudaEventRecord(L->ev[5], st_top);
        cudaStreamWaitEvent(st, L->ev[5], 0);
    }
    return rc;
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
    Lanes *L = side_lanes();
    cudaStream_t st_top = st;
    if (L) {
        cudaStreamWaitEvent(L->side[1], L->ev[4], 0);
        st_top = L->side[1];
    }
    rc = topup_cells(c, N, s.plan + 2, (int64_t)topup_bound((uint64_t)N), seed, offset, u_lo, u_hi, out_dev,
                     cell_idx_dev, st_top);
    if (L) {
        cudaEventRecord(L->ev[5], st_top);
        cudaStreamWaitEvent(st, L->ev[5], 0);
    }
    return rc;
}

}  // extern "C"
