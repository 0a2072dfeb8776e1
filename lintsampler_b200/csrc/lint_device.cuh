// Device-side building blocks of the B200 linear-interpolant sampler.
// Everything here is header-only and shared by lint_build.cu / lint_sample.cu.
//
// Reference citations are relative to /root/reference/src/lintsampler.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include "../../include/lintsampler_b200.h"

namespace lint {

// ---------------------------------------------------------------------------
// host-side error plumbing (no exceptions cross the C ABI)
// ---------------------------------------------------------------------------
std::string &last_error_slot();
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);

// ---------------------------------------------------------------------------
// Programmatic dependent launch: the small kernels of a step form a chain in which every kernel
// needs the whole output of the one before it.  Launched with the stream-serialization attribute,
// a kernel's CTAs are scheduled as soon as the kernel before it lets them (pdl_release, called at
// its start) and park at pdl_wait until that kernel has finished and its writes are visible: the
// launch latency and the ramp of the grid are hidden behind the predecessor.  Without the
// attribute both calls do nothing.  Measured on the 8-way sharded step (0.46 ms, eleven kernels in
// a CUDA graph): within +-3 us of the ordinary launches, i.e. no gain inside a graph, whose
// kernel-to-kernel gaps are already short — so it is off unless LINT_PDL=1 is set.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: one 128-bit block per call
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// The same block function with the ten round keys formed beforehand (on the host, passed as a
// kernel parameter): inside a kernel every key is then a constant-bank operand of the round's
// XOR instead of two additions per round.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
    // mantissa mask and exponent bits of the 21-bit uniform fields (see FlatTraits<float>): kept
    // beside the keys so that the compiler sees run-time values and fuses mask and OR into one
    // three-input logic instruction instead of two with an immediate each
    uint32_t field_mask, field_one;
};
inline PhiloxKeys philox_round_keys(uint2 key) {
    PhiloxKeys s;
    s.field_mask = 0x007ffffcu;
    s.field_one = 0x3f800002u;
    for (int r = 0; r < 10; ++r) {
        s.k0[r] = key.x + (uint32_t)r * 0x9E3779B9u;
        s.k1[r] = key.y + (uint32_t)r * 0xBB67AE85u;
    }
    return s;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, const PhiloxKeys &k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ k.k0[r], lo1, hi0 ^ ctr.w ^ k.k1[r], lo0);
    }
    return ctr;
}

// Position in the Philox stream: the `offset` argument of a call plus, optionally, a counter that
// lives in device memory (lint_set_offset_counter) — a captured CUDA graph can then draw fresh
// samples at every replay by advancing the counter on the device (lint_advance_counter).
struct StreamPos {
    uint64_t base;
    const uint64_t *extra;
    __device__ __forceinline__ uint64_t get() const { return base + (extra ? *extra : 0ull); }
};
const uint64_t *offset_counter();          // the process-wide setting (lint_build.cu)

__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
    // same construction as NumPy's Generator.random: top 53 bits * 2^-53
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------
// searchsorted(side='right'): first i in [0, n) with cdf[i] > u   [utils.py:197]
// The guide table only narrows the bracket; the comparisons, and therefore the
// result, are exactly those of a full binary search on the same fp64 table.
// ---------------------------------------------------------------------------
struct CdfView {
    const double *cdf;
    const uint32_t *guide;
    uint32_t n;
    int guide_bits;
    const uint32_t *range = nullptr;    // device [first cell, end cell) of a slab build; brackets the search when there is no guide table
};

// Guide entries carry a flag in bit 31: set when the whole bin lies inside one
// cell (guide[j] == guide[j+1]), in which case the entry IS the answer and no
// second load is needed — the common case for every cell wider than a bin.
constexpr uint32_t GUIDE_SAME = 0x80000000u;

// 256-bit non-coherent loads (sm_100+): one request, one 32-byte sector
__device__ __forceinline__ void ldg256(const double *p, double (&v)[4]) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ldg256(const float *p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
                   "=f"(v[7]) : "l"(p));
}

// First i in [lo, hi] with cdf[i] > u, given that cdf[hi] > u.  Long brackets are
// halved by binary search; short ones are resolved by 32-byte windows of four
// consecutive CDF entries (one sector, one request) so the common case costs a
// single memory round trip instead of log2(range) dependent ones.
__device__ __forceinline__ uint32_t search_bracket(const CdfView &t, double u, uint32_t lo, uint32_t hi) {
    while (hi - lo > 8) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(t.cdf + mid) > u) hi = mid; else lo = mid + 1;
    }
    while (lo < hi) {
        const uint32_t a = lo & ~3u;
        if (a + 4 > t.n) {                       // ragged tail of the table: scalar probes
            if (__ldg(t.cdf + lo) > u) return lo;
            ++lo;
            continue;
        }
        double w[4];
        ldg256(t.cdf + a, w);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = a + q;
            if (i >= lo && i < hi && w[q] > u) return i;
        }
        lo = a + 4 < hi ? a + 4 : hi;
    }
    return lo;
}

// Guide lookup: bracket [lo, hi] of the answer, lo == hi when the bin is flagged.
__device__ __forceinline__ void guide_bracket(const CdfView &t, double u, uint32_t &lo, uint32_t &hi) {
    lo = 0;
    hi = t.n - 1;
    if (t.guide_bits > 0) {
        // u * 2^bits is exact in fp64; u < 1 keeps j < 2^bits
        const uint32_t G = 1u << t.guide_bits;
        uint32_t j = (uint32_t)(u * (double)G);
        j = j < G ? j : G - 1;
        const uint32_t e = __ldg(t.guide + j);
        lo = e & ~GUIDE_SAME;
        hi = (e & GUIDE_SAME) ? lo : (__ldg(t.guide + j + 1) & ~GUIDE_SAME);
    } else if (t.range) {
        lo = __ldg(t.range);
        hi = __ldg(t.range + 1) - 1;
    }
}

// ---------------------------------------------------------------------------
// 1-D inverse CDF of a linear density   [sampling.py:6-38]
// ---------------------------------------------------------------------------
// fp64: the reference's literal expression, one IEEE rounding per NumPy ufunc
// call, so the result is bit-identical to sampling.py:32-37.
__device__ __forceinline__ double quantile_exact(double f0, double f1, double u) {
    if (f0 == f1) return u;
    const double a2 = __dmul_rn(f0, f0);
    const double b2 = __dmul_rn(f1, f1);
    const double rad = __dadd_rn(a2, __dmul_rn(__dsub_rn(b2, a2), u));
    return __ddiv_rn(__dadd_rn(-f0, __dsqrt_rn(rad)), __dsub_rn(f1, f0));
}

// fp64, cancellation-free: the conjugate form of the same root.  Used where rows are not
// tied to a reference uniform stream (the cell-major emitter); the literal form collapses
// z onto 0 when f0 and f1 differ by rounding noise only.
__device__ __forceinline__ double quantile_conjugate(double f0, double f1, double u) {
    const double a2 = f0 * f0;
    const double den = f0 + sqrt(a2 + (f1 * f1 - a2) * u);
    return den > 0.0 ? (f0 + f1) * u / den : u;
}

// fp32: algebraically equal conjugate form (no cancellation when f0 ~ f1,
// SURVEY.md §7.3-3); inputs are pre-normalised so the squares cannot underflow.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float quantile_fast(float f0, float f1, float u) {
    // rad >= 0 without a clamp: a = fl(f0^2) >= 0, b = fl(f1^2 - a) >= -a, 0 <= u <= 1, so
    // a + b u >= 0 exactly and the single rounding of the fma keeps the sign
    const float a = f0 * f0;
    const float rad = fmaf(u, fmaf(f1, f1, -a), a);
    // den == 0 only where the segment carries no density at all (or f0 == 0 and u == 0):
    // the numerator is 0 there too, so flooring den gives z = 0 instead of 0/0
    const float den = fmaxf(f0 + sqrt_approx(rad), 1e-30f);
    return (f0 + f1) * u * rcp_approx(den);
}

// ---------------------------------------------------------------------------
// k-D multilinear-interpolant inverse CDF in the unit cube   [sampling.py:41-94]
// ---------------------------------------------------------------------------
// fp64 parity version: operation order identical to the oracle's
// unit_cube_sample (= the order NumPy executes for the reference).
template <int K, bool LITERAL = true>
__device__ __forceinline__ void incell_exact(const double (&c)[1 << K], const double (&u)[K],
                                             double (&z)[K]) {
#pragma unroll
    for (int d = 0; d < K; ++d) {
        const int nlater = 1 << (K - 1 - d);   // corner bits of dims > d
        const int nearly = 1 << d;             // corner bits of dims < d
        double g[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            double total = 0.0;
#pragma unroll
            for (int hi = 0; hi < nearly; ++hi) {
                const int base = (hi * 2 + b) * nlater;
                double acc = c[base];
#pragma unroll
                for (int lo = 1; lo < nlater; ++lo) acc = __dadd_rn(acc, c[base + lo]);
                if (nlater > 1) acc = __dmul_rn(acc, 1.0 / nlater);   // == division by 2^j
#pragma unroll
                for (int i = 0; i < d; ++i) {
                    const int bit = (hi >> (d - 1 - i)) & 1;
                    acc = __dmul_rn(acc, bit ? z[i] : __dsub_rn(1.0, z[i]));
                }
                total = (hi == 0) ? acc : __dadd_rn(total, acc);
            }
            g[b] = total;
        }
        z[d] = LITERAL ? quantile_exact(g[0], g[1], u[d]) : quantile_conjugate(g[0], g[1], u[d]);
    }
}

// fp32 throughput version: condition the corner array on each coordinate as it
// is drawn (2^K - 1 lerps in total); the 1/2^j averaging factors are dropped
// because the 1-D quantile is scale invariant.
//
// normalise_corners: divide by the largest corner so squares stay in range
// (densities may be ~1e-30 in physical units; g^2 would underflow fp32).
template <int NC>
__device__ __forceinline__ void normalise_corners(float *c) {
    float m = c[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) m = fmaxf(m, c[j]);
    const float inv = m > 0.f ? rcp_approx(m) : 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) c[j] *= inv;
}

// K sequential conditional draws on already-normalised corners; `c` is consumed.
template <int K>
__device__ __forceinline__ void incell_fast_core(float *c, const float *u, float *z) {
#pragma unroll
    for (int d = 0; d < K; ++d) {
        const int half = 1 << (K - 1 - d);     // live corners: 2*half
        float g0 = c[0], g1 = c[half];
#pragma unroll
        for (int j = 1; j < half; ++j) { g0 += c[j]; g1 += c[half + j]; }
        const float zd = quantile_fast(g0, g1, u[d]);
        z[d] = zd;
#pragma unroll
        for (int j = 0; j < half; ++j) c[j] = fmaf(zd, c[half + j] - c[j], c[j]);
    }
}

// Per-cell constants of the fp32 draw, for callers that emit many rows per cell:
// everything about dimension 0 that does not depend on the row is hoisted
// (normalisation, the two marginal sums, the quantile's coefficients and the
// first lerp's differences).
template <int K>
struct CellFast {
    static constexpr int H = 1 << (K - 1);
    float c[H], dc[H];             // corners with x0 = 0, and (x0 = 1) - (x0 = 0)
    float g0, a0, b0, s0;          // z0 = s0 u / (g0 + sqrt(a0 + b0 u))
    __device__ __forceinline__ void set(float (&corners)[1 << K], bool prenormalised) {
        if (!prenormalised) normalise_corners<(1 << K)>(corners);
        float g1 = corners[H];
        g0 = corners[0];
#pragma unroll
        for (int j = 1; j < H; ++j) { g0 += corners[j]; g1 += corners[H + j]; }
        a0 = g0 * g0;
        b0 = fmaf(g1, g1, -a0);
        s0 = g0 + g1;
#pragma unroll
        for (int j = 0; j < H; ++j) { c[j] = corners[j]; dc[j] = corners[H + j] - corners[j]; }
    }
    // flat image of the constants (for staging them in shared memory)
    static constexpr int WORDS = 2 * H + 4;
    __device__ __forceinline__ void pack(float *w) const {
#pragma unroll
        for (int j = 0; j < H; ++j) { w[j] = c[j]; w[H + j] = dc[j]; }
        w[2 * H] = g0; w[2 * H + 1] = a0; w[2 * H + 2] = b0; w[2 * H + 3] = s0;
    }
    __device__ __forceinline__ void unpack(const float *w) {
#pragma unroll
        for (int j = 0; j < H; ++j) { c[j] = w[j]; dc[j] = w[H + j]; }
        g0 = w[2 * H]; a0 = w[2 * H + 1]; b0 = w[2 * H + 2]; s0 = w[2 * H + 3];
    }
    __device__ __forceinline__ void draw(const float (&u)[K], float (&z)[K]) const {
        const float den = fmaxf(g0 + sqrt_approx(fmaf(b0, u[0], a0)), 1e-30f);   // see quantile_fast
        const float z0 = s0 * u[0] * rcp_approx(den);
        z[0] = z0;
        if constexpr (K > 1) {
            float w[H];
#pragma unroll
            for (int j = 0; j < H; ++j) w[j] = fmaf(z0, dc[j], c[j]);
            incell_fast_core<K - 1>(w, &u[1], &z[1]);
        }
    }
};

// fp64: no hoisting — the reference's order of the marginal sums, with the
// cancellation-free quantile
template <int K>
struct CellExact {
    double c[1 << K];
    static constexpr int WORDS = 0;                                  // never staged
    __device__ __forceinline__ void pack(float *) const {}
    __device__ __forceinline__ void unpack(const float *) {}
    __device__ __forceinline__ void set(double (&corners)[1 << K], bool) {
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) c[j] = corners[j];
    }
    __device__ __forceinline__ void draw(const double (&u)[K], double (&z)[K]) const {
        incell_exact<K, false>(c, u, z);
    }
};

template <int K, typename T> struct CellCoefs;
template <int K> struct CellCoefs<K, float> : CellFast<K> {};
template <int K> struct CellCoefs<K, double> : CellExact<K> {};

// ---------------------------------------------------------------------------
// cell sources
// ---------------------------------------------------------------------------
// Rectilinear grid (reference DensityGrid).  Cell counts and vertex counts are
// limited to < 2^31 so all index arithmetic is 32-bit.
template <int K>
struct GridView {
    const void *V;                 // vertex densities, C order
    const void *rec;               // optional per-cell corner records (ncells, 2^K), or NULL
    const double *edges[K];
    CdfView table;
    uint32_t n[K];                 // cells per dim
    uint32_t div_m[K], div_s[K];   // q = umulhi(r, div_m) >> div_s  ==  r / n[d]  for r < 2^31
    uint32_t vstride[K];           // vertex strides (elements)
    uint32_t edge_off[K];          // offset of dim d's edges in the shared-memory copy
    uint32_t edge_total;           // sum(n[d] + 1); 0 = do not stage edges in shared memory
    uint32_t corner_off[1 << K];   // vertex offset of corner c from the cell origin
};

// C-order unravel of a flat cell index  [grid.py:437]; exact for flat < 2^31
template <int K>
__device__ __forceinline__ void grid_unravel(const GridView<K> &g, uint32_t flat, uint32_t (&i)[K]) {
    uint32_t r = flat;
#pragma unroll
    for (int d = K - 1; d > 0; --d) {
        const uint32_t q = g.n[d] == 1 ? r : (__umulhi(r, g.div_m[d]) >> g.div_s[d]);
        i[d] = r - q * g.n[d];
        r = q;
    }
    i[0] = r;
}

template <int K, typename DensT, typename T>
__device__ __forceinline__ void grid_corners(const GridView<K> &g, uint32_t flat, const uint32_t (&i)[K],
                                             T (&c)[1 << K]);

// Cell box + 2^K corner densities of cell `flat`.  `sedges` is the CTA's
// shared-memory copy of the edge arrays in arithmetic type T (or unused when
// g.edge_total == 0).
template <int K, typename DensT, typename T>
__device__ __forceinline__ void grid_fetch(const GridView<K> &g, const T *sedges, uint32_t flat,
                                           T (&c)[1 << K], T (&lo)[K], T (&hi)[K]) {
    uint32_t i[K];
    grid_unravel<K>(g, flat, i);
    if (g.edge_total) {
#pragma unroll
        for (int d = 0; d < K; ++d) {
            lo[d] = sedges[g.edge_off[d] + i[d]];          // _cell_mins gather [grid.py:353]
            hi[d] = sedges[g.edge_off[d] + i[d] + 1];      // _cell_maxs gather [grid.py:354]
        }
    } else {
#pragma unroll
        for (int d = 0; d < K; ++d) {
            lo[d] = (T)__ldg(g.edges[d] + i[d]);
            hi[d] = (T)__ldg(g.edges[d] + i[d] + 1);
        }
    }
    grid_corners<K, DensT, T>(g, flat, i, c);
}

// 2^K corner densities of cell `flat` (= multi-index i), corner order = MSB-first bit
// pattern [grid.py:440-467]
template <int K, typename DensT, typename T>
__device__ __forceinline__ void grid_corners(const GridView<K> &g, uint32_t flat, const uint32_t (&i)[K],
                                             T (&c)[1 << K]) {
    if (g.rec) {
        // one contiguous record per cell: a single 32 B sector for K=3 fp32
        constexpr int NB = (1 << K) * (int)sizeof(DensT);
        const DensT *p = reinterpret_cast<const DensT *>(g.rec) + (size_t)flat * (1 << K);
        DensT v[1 << K];
        if constexpr (NB % 32 == 0) {
            constexpr int PER = 32 / (int)sizeof(DensT);
#pragma unroll
            for (int q = 0; q < NB / 32; ++q)
                ldg256(p + q * PER, *reinterpret_cast<DensT(*)[PER]>(v + q * PER));
        } else if constexpr (NB % 16 == 0) {
#pragma unroll
            for (int q = 0; q < NB / 16; ++q)
                reinterpret_cast<uint4 *>(v)[q] = __ldg(reinterpret_cast<const uint4 *>(p) + q);
        } else {
#pragma unroll
            for (int q = 0; q < NB / 8; ++q)
                reinterpret_cast<uint2 *>(v)[q] = __ldg(reinterpret_cast<const uint2 *>(p) + q);
        }
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) c[j] = (T)v[j];
    } else {
        uint32_t base = 0;
#pragma unroll
        for (int d = 0; d < K; ++d) base += i[d] * g.vstride[d];
        const DensT *V = reinterpret_cast<const DensT *>(g.V);
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) c[j] = (T)__ldg(V + base + g.corner_off[j]);
    }
}

// Stage the edge arrays into shared memory (converted to T) once per CTA.
template <int K, typename T>
__device__ __forceinline__ void grid_stage_edges(const GridView<K> &g, T *sedges) {
    if (!g.edge_total) return;
#pragma unroll
    for (int d = 0; d < K; ++d)
        for (uint32_t j = threadIdx.x; j <= g.n[d]; j += blockDim.x)
            sedges[g.edge_off[d] + j] = (T)__ldg(g.edges[d] + j);
    __syncthreads();
}

// Flat cell list (reference DensityTree leaves, list order)  [tree.py:334-355]
struct CellsView {
    const double *mins;
    const double *widths;
    const void *corners;
    const void *rec;               // optional packed records (see cells_record_stride), or NULL
    CdfView table;
};

// Packed per-cell record, in the domain's element type: 2^K corner densities, K lower
// bounds, K upper bounds (= fl(x + dx), tree.py:348), padded to a multiple of 32 bytes
// so a record is read with aligned 256-bit loads.
template <int K, typename DensT>
__host__ __device__ constexpr int cells_record_stride() {
    constexpr int per32 = 32 / (int)sizeof(DensT);
    return (((1 << K) + 2 * K + per32 - 1) / per32) * per32;
}

// corner densities only
template <int K, typename DensT, typename T>
__device__ __forceinline__ void cells_corners(const CellsView &s, uint32_t idx, T (&c)[1 << K]) {
    const DensT *C = s.rec ? reinterpret_cast<const DensT *>(s.rec) + (size_t)idx * cells_record_stride<K, DensT>()
                           : reinterpret_cast<const DensT *>(s.corners) + (size_t)idx * (1 << K);
#pragma unroll
    for (int j = 0; j < (1 << K); ++j) c[j] = (T)__ldg(C + j);
}

template <int K, typename DensT, typename T>
__device__ __forceinline__ void cells_fetch(const CellsView &s, uint32_t idx, T (&c)[1 << K],
                                            T (&lo)[K], T (&hi)[K]) {
    if (s.rec) {
        constexpr int RS = cells_record_stride<K, DensT>();
        constexpr int PER = 32 / (int)sizeof(DensT);
        const DensT *p = reinterpret_cast<const DensT *>(s.rec) + (size_t)idx * RS;
        DensT v[RS];
#pragma unroll
        for (int q = 0; q < RS / PER; ++q) ldg256(p + q * PER, *reinterpret_cast<DensT(*)[PER]>(v + q * PER));
#pragma unroll
        for (int j = 0; j < (1 << K); ++j) c[j] = (T)v[j];
#pragma unroll
        for (int d = 0; d < K; ++d) { lo[d] = (T)v[(1 << K) + d]; hi[d] = (T)v[(1 << K) + K + d]; }
        return;
    }
#pragma unroll
    for (int d = 0; d < K; ++d) {
        const double x = __ldg(s.mins + (size_t)idx * K + d);
        const double dx = __ldg(s.widths + (size_t)idx * K + d);
        lo[d] = (T)x;
        hi[d] = (T)__dadd_rn(x, dx);              // maxs = leaf.x + leaf.dx [tree.py:348]
    }
    const DensT *C = reinterpret_cast<const DensT *>(s.corners) + (size_t)idx * (1 << K);
#pragma unroll
    for (int j = 0; j < (1 << K); ++j) c[j] = (T)__ldg(C + j);
}

// ---------------------------------------------------------------------------
// uniform sources: each yields, for row i, the cell-selecting fp64 uniform and
// the K in-cell uniforms in arithmetic type T   [lintsampler.py:269-270]
// ---------------------------------------------------------------------------
template <int K>
struct UniformsFromArray {        // caller-supplied (N, K+1) fp64  [lintsampler.py:517]
    const double *u;
    __host__ size_t smem_words() const { return 0; }
    __device__ __forceinline__ const uint32_t *stage(uint32_t *) const { return nullptr; }
    template <typename T>
    __device__ __forceinline__ void get(int64_t row, double &ucell, T (&uc)[K], const uint32_t *) const {
        const double *p = u + row * (K + 1);
#pragma unroll
        for (int d = 0; d < K; ++d) uc[d] = (T)__ldg(p + d);
        ucell = __ldg(p + K);
    }
};

template <int K>
struct UniformsPhilox {
    uint2 key;
    StreamPos pos;
    double u_lo, u_span;           // cell-uniform stratum [u_lo, u_lo + u_span); (0, 1) = whole domain
    __host__ size_t smem_words() const { return 0; }
    __device__ __forceinline__ const uint32_t *stage(uint32_t *) const { return nullptr; }
    __device__ __forceinline__ double stratum(double r) const {
        return __dadd_rn(u_lo, __dmul_rn(u_span, r));   // exact identity for (0, 1)
    }
    // fp64 rows: cell uniform from words (0,1), coordinate d from words (2+2d, 3+2d)
    __device__ __forceinline__ void get(int64_t row, double &ucell, double (&uc)[K], const uint32_t *) const {
        const uint64_t i = pos.get() + (uint64_t)row;
        constexpr int NCALL = (K + 2) / 2;
        uint32_t w[4 * NCALL];
#pragma unroll
        for (int q = 0; q < NCALL; ++q) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), q, 0), key);
            w[4 * q] = r.x; w[4 * q + 1] = r.y; w[4 * q + 2] = r.z; w[4 * q + 3] = r.w;
        }
        ucell = stratum(u53(w[0], w[1]));
#pragma unroll
        for (int d = 0; d < K; ++d) uc[d] = u53(w[2 + 2 * d], w[3 + 2 * d]);
    }
    // fp32 rows: cell uniform from words (0,1); 24-bit coordinates: d=0,1 from the
    // top bits of words 2,3; d=2 from the otherwise unused low bytes of words
    // 2, 3 and 0; d=3..5 from a second block.
    __device__ __forceinline__ void get(int64_t row, double &ucell, float (&uc)[K], const uint32_t *) const {
        const uint64_t i = pos.get() + (uint64_t)row;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0, 0), key);
        ucell = stratum(u53(r.x, r.y));
        constexpr float S = 1.0f / 16777216.0f;
        uc[0] = (float)(r.z >> 8) * S;
        if (K > 1) uc[1 % K] = (float)(r.w >> 8) * S;
        if (K > 2) uc[2 % K] = (float)(((r.z & 0xffu) << 16) | ((r.w & 0xffu) << 8) | (r.x & 0xffu)) * S;
        if (K > 3) {
            const uint4 s = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1, 0), key);
            uc[3 % K] = (float)(s.x >> 8) * S;
            if (K > 4) uc[4 % K] = (float)(s.y >> 8) * S;
            if (K > 5) uc[5 % K] = (float)(s.z >> 8) * S;
        }
    }
};

template <int K>
struct UniformsSobol {            // scipy Sobol(d=K+1, bits=32)
    uint32_t sv[K + 1][32];       // scrambled direction numbers
    uint32_t shift[K + 1];        // digital shift
    uint32_t ainv[32];            // sorted mode: columns of the GF(2) inverse (see below)
    uint64_t first;               // direct mode: index of the first point
    uint32_t sorted_bits;         // 0 = direct Gray-code index; m = sorted enumeration of the first 2^m points

    // tables are staged in shared memory: they are indexed by data-dependent bit
    // positions, which serialises in the constant bank the parameters live in
    static constexpr int TABLE_WORDS = (K + 1) * 33 + 32;
    __host__ size_t smem_words() const { return TABLE_WORDS; }
    __device__ __forceinline__ const uint32_t *stage(uint32_t *s) const {
        for (int i = threadIdx.x; i < (K + 1) * 32; i += blockDim.x) s[i] = sv[i / 32][i % 32];
        for (int i = threadIdx.x; i <= K; i += blockDim.x) s[(K + 1) * 32 + i] = shift[i];
        for (int i = threadIdx.x; i < 32; i += blockDim.x) s[(K + 1) * 33 + i] = ainv[i];
        __syncthreads();
        return s;
    }

    // Direct mode: row i is point n = first + i, x_n[j] = shift[j] ^ XOR_{b in gray(n)} sv[j][b],
    // bit-identical to scipy's recurrence.
    // Sorted mode (N == 2^m, first == 0): row i is THE point of the first 2^m whose
    // cell-selecting coordinate has top m bits == i.  The first 2^m points of one
    // Sobol' coordinate visit every dyadic interval of width 2^-m exactly once (it is a
    // (0,m,1)-net: the top m x m block A of its generator matrix is invertible over
    // GF(2)), so gray(n) = A^-1 (i ^ top_m(shift)).  The same point set as scipy, but
    // enumerated with the cell uniform increasing — consecutive rows land in the same
    // or neighbouring cells, which turns the sampler's table accesses from random into
    // streaming.  Row order is restored to random by the final row permutation that
    // the reference applies to every draw anyway (lintsampler.py:309-311).
    template <typename T>
    __device__ __forceinline__ void get(int64_t row, double &ucell, T (&uc)[K], const uint32_t *tab) const {
        uint32_t gray;
        if (sorted_bits) {
            uint32_t t = (uint32_t)row ^ (tab[(K + 1) * 32 + K] >> (32 - sorted_bits));
            gray = 0;
            while (t) {
                const int r = __ffs(t) - 1;
                t &= t - 1;
                gray ^= tab[(K + 1) * 33 + r];
            }
        } else {
            const uint64_t n = first + (uint64_t)row;
            gray = (uint32_t)(n ^ (n >> 1));
        }
        uint32_t x[K + 1];
#pragma unroll
        for (int j = 0; j <= K; ++j) x[j] = tab[(K + 1) * 32 + j];
        while (gray) {
            const int b = __ffs(gray) - 1;
            gray &= gray - 1;
#pragma unroll
            for (int j = 0; j <= K; ++j) x[j] ^= tab[j * 32 + b];
        }
        // scipy scales the 32-bit integers by 2^-32 in fp64
#pragma unroll
        for (int d = 0; d < K; ++d) uc[d] = (T)((double)x[d] * (1.0 / 4294967296.0));
        ucell = (double)x[K] * (1.0 / 4294967296.0);
    }
};

// ---------------------------------------------------------------------------
// affine map into the cell   [sampling.py:126]
// ---------------------------------------------------------------------------
__device__ __forceinline__ double affine(double lo, double hi, double z) {
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), z));
}
__device__ __forceinline__ float affine(float lo, float hi, float z) {
    return fmaf(hi - lo, z, lo);
}
// the same with the width w = hi - lo formed once per cell
__device__ __forceinline__ double affine_w(double lo, double w, double z) {
    return __dadd_rn(lo, __dmul_rn(w, z));
}
__device__ __forceinline__ float affine_w(float lo, float w, float z) { return fmaf(w, z, lo); }

// streaming (evict-first) stores for the sample rows: 12 GB of output must not
// evict the guide table and cell records from L2
__device__ __forceinline__ void store_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void store_stream(double *p, double v) { __stcs(p, v); }
// two consecutive values, p aligned to twice the element size
__device__ __forceinline__ void store_stream2(float *p, float a, float b) {
    __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b));
}
__device__ __forceinline__ void store_stream2(double *p, double a, double b) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
}

// `prenormalised`: the corners come from fp32 records, which are stored divided by the
// cell's largest corner (the draw depends on ratios only), so the per-sample
// normalisation is skipped.  fp64 records hold the raw densities.
template <int K> __device__ __forceinline__ void incell(double (&c)[1 << K], const double (&u)[K], double (&z)[K], bool) { incell_exact<K>(c, u, z); }
template <int K> __device__ __forceinline__ void incell(float (&c)[1 << K], const float (&u)[K], float (&z)[K], bool prenormalised) {
    if (!prenormalised) normalise_corners<(1 << K)>(c);
    incell_fast_core<K>(c, u, z);
}

}  // namespace lint
