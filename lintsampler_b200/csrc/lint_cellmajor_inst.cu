// One translation unit per dimension (compiled with -DLINT_K=1..6, see _build.py): the
// cell-major kernels of that dimension for both precisions and both cell sources.
#ifndef LINT_K
#error "compile with -DLINT_K=<dimension>"
#endif

#ifdef LINT_STUB
// development builds (LINT_DIMS=3 python -m lintsampler_b200._build) leave the other
// dimensions out: their entry points report LINT_ERR_UNSUPPORTED
#include "lint_device.cuh"
#include "lint_cellmajor_plan.h"
namespace lint {
template <int K>
int grid_cellmajor_k(const lint_grid_t *, uint64_t, uint64_t, uint2, uint64_t, Stratum, const uint32_t *,
                     const Scratch &, void *, int64_t *, cudaStream_t) {
    return fail(LINT_ERR_UNSUPPORTED, "cell-major kernels for dim %d were left out of this development build", K);
}
template <int K>
int cells_cellmajor_k(const lint_cells_t *, uint64_t, uint2, uint64_t, Stratum, const Scratch &, void *,
                      int64_t *, cudaStream_t) {
    return fail(LINT_ERR_UNSUPPORTED, "cell-major kernels for dim %d were left out of this development build", K);
}
}  // namespace lint
#else
#include "lint_cellmajor.cuh"
#endif

namespace lint {
template int grid_cellmajor_k<LINT_K>(const lint_grid_t *, uint64_t, uint64_t, uint2, uint64_t, Stratum,
                                      const uint32_t *, const Scratch &, void *, int64_t *, cudaStream_t);
template int cells_cellmajor_k<LINT_K>(const lint_cells_t *, uint64_t, uint2, uint64_t, Stratum,
                                       const Scratch &, void *, int64_t *, cudaStream_t);
}  // namespace lint
