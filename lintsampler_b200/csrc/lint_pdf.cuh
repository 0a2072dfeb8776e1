// Fused pdf families evaluated on the device: parameter blocks (kernel arguments), the density
// functions, and the host code that fills the blocks from the C-ABI structs.  Shared by the grid
// vertex kernels (lint_build.cu) and the tree builder (lint_tree.cu): one arithmetic, so a grid
// vertex and a tree corner at the same position get the same density bit for bit.
#pragma once
#include "lint_device.cuh"

namespace lint {

template <int K>
struct GmmParams {
    int ncomp;
    double log_norm[LINT_GMM_MAX_COMP];
    double mean[LINT_GMM_MAX_COMP][K];
    double U[LINT_GMM_MAX_COMP][K * (K + 1) / 2];   // upper triangle, row-major
    __device__ __forceinline__ double operator()(const double (&x)[K]) const {
        double f = 0.0;
        for (int j = 0; j < ncomp; ++j) {
            double y[K];
#pragma unroll
            for (int d = 0; d < K; ++d) y[d] = x[d] - mean[j][d];
            double q = 0.0;
            int t = 0;
#pragma unroll
            for (int rr = 0; rr < K; ++rr) {
                double s = 0.0;
#pragma unroll
                for (int cc = rr; cc < K; ++cc) s = fma(U[j][t++], y[cc], s);
                q = fma(s, s, q);
            }
            f += exp(log_norm[j] - 0.5 * q);
        }
        return f;
    }
};

template <int K>
inline GmmParams<K> fill_gmm(const lint_gmm_t *gmm) {
    GmmParams<K> p;
    p.ncomp = gmm->ncomp;
    for (int j = 0; j < gmm->ncomp; ++j) {
        p.log_norm[j] = gmm->log_norm[j];
        int t = 0;
        for (int r = 0; r < K; ++r) {
            p.mean[j][r] = gmm->mean[j * K + r];
            for (int c = r; c < K; ++c) p.U[j][t++] = gmm->prec_chol[(j * K + r) * K + c];
        }
    }
    return p;
}

template <int K>
struct HaloParams {
    int ncomp;
    double rho0[LINT_GMM_MAX_COMP], inv_rs[LINT_GMM_MAX_COMP], centre[LINT_GMM_MAX_COMP][K];
    double alpha, outer, gamma, eps, q_cut;      // outer = (beta - gamma) / alpha
    bool nfw;                                    // alpha 1, beta 3, gamma 1: no pow()
    __device__ __forceinline__ double operator()(const double (&x)[K]) const {
        double f = 0.0;
        for (int j = 0; j < ncomp; ++j) {
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < K; ++d) { const double y = x[d] - centre[j][d]; r2 = fma(y, y, r2); }
            const double q = sqrt(r2) * inv_rs[j];
            if (q_cut > 0.0 && q > q_cut) continue;
            const double den = nfw ? (q + eps) * (1.0 + q) * (1.0 + q)
                                   : pow(q + eps, gamma) * pow(1.0 + pow(q, alpha), outer);
            f += rho0[j] / den;
        }
        return f;
    }
};

template <int K>
inline HaloParams<K> fill_halos(const lint_halos_t *h) {
    HaloParams<K> p;
    p.ncomp = h->ncomp;
    for (int j = 0; j < h->ncomp; ++j) {
        p.rho0[j] = h->rho0[j];
        p.inv_rs[j] = 1.0 / h->r_s[j];
        for (int d = 0; d < K; ++d) p.centre[j][d] = h->centre[j * K + d];
    }
    p.alpha = h->alpha;
    p.outer = (h->beta - h->gamma) / h->alpha;
    p.gamma = h->gamma;
    p.eps = h->eps;
    p.q_cut = h->q_cut;
    p.nfw = h->alpha == 1.0 && h->beta == 3.0 && h->gamma == 1.0;
    return p;
}

}  // namespace lint
