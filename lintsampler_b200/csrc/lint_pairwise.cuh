// NumPy's float64 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum)
// as device code: the recursion splits n into a left part of n/2 rounded down to a multiple of 8
// and the rest until at most 128 elements remain, which are summed with eight accumulators.
// Shared by lint_sum_numpy_f64 (lint_build.cu) and the tree refinement loop (lint_tree.cu).
#pragma once
#include <cstdint>

namespace lint {

__host__ __device__ inline uint64_t pw_left(uint64_t n) {
    uint64_t n2 = n / 2;
    return n2 - (n2 % 8);
}

__device__ inline double pw_leaf(const double *a, uint32_t n) {
    if (n < 8) {
        double r = -0.0;
        for (uint32_t i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    uint32_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__device__ inline double pw_node(const double *a, uint64_t n) {
    // explicit stack; subtree sizes here are a few hundred elements at most
    struct Frame { const double *a; uint64_t n; int state; double left; };
    Frame st[12];
    int sp = 0;
    st[0] = {a, n, 0, 0.0};
    double ret = 0.0;
    while (sp >= 0) {
        Frame &f = st[sp];
        if (f.n <= 128) { ret = pw_leaf(f.a, (uint32_t)f.n); --sp; continue; }
        const uint64_t n2 = pw_left(f.n);
        if (f.state == 0) { f.state = 1; st[++sp] = {f.a, n2, 0, 0.0}; }
        else if (f.state == 1) { f.left = ret; f.state = 2; st[++sp] = {f.a + n2, f.n - n2, 0, 0.0}; }
        else { ret = __dadd_rn(f.left, ret); --sp; }
    }
    return ret;
}

__host__ __device__ inline int pw_depth(uint64_t n) {
    // deepest level whose smallest node (always the left-most) still exceeds 128
    int depth = 0;
    uint64_t smallest = n;
    while (smallest > 128) {
        const uint64_t child = pw_left(smallest);
        if (child <= 128) break;
        smallest = child;
        ++depth;
    }
    return depth;
}

}  // namespace lint
