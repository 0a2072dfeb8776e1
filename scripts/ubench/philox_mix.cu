// Micro-benchmark 2: Philox4x32-10 at the occupancy and instruction mix of the flat emitter.
//   ./philox_mix
// For each (CTAs per SM, variant): cycles per warp-block per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Keys { uint32_t k0[10], k1[10]; uint32_t m, o; };

template <bool CONVERT>
__global__ void __launch_bounds__(256) mix_kernel(float *out, int iters, const __grid_constant__ Keys key, float lo, float w) {
    float acc = 0.f;
    uint32_t base = blockIdx.x * 1000003u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint32_t c[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) { c[j][0] = base + it; c[j][1] = 0x46000000u | j; c[j][2] = 17; c[j][3] = 99; }
#pragma unroll
        for (int r = 0; r < 10; ++r) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint64_t p0 = (uint64_t)0xD2511F53u * c[j][0], p1 = (uint64_t)0xCD9E8D57u * c[j][2];
                const uint32_t y = c[j][1], z = c[j][3];
                c[j][0] = (uint32_t)(p1 >> 32) ^ y ^ key.k0[r]; c[j][1] = (uint32_t)p1;
                c[j][2] = (uint32_t)(p0 >> 32) ^ z ^ key.k1[r]; c[j][3] = (uint32_t)p0;
            }
        }
        if (CONVERT) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t s[6] = {c[j][0] << 2, __funnelshift_r(c[j][0], c[j][1], 19), __funnelshift_r(c[j][1], c[j][2], 8),
                                 __funnelshift_r(c[j][1], c[j][2], 29), __funnelshift_r(c[j][2], c[j][3], 18), c[j][3] >> 7};
#pragma unroll
                for (int f = 0; f < 6; ++f)
                    acc += fmaf(w, __uint_as_float((s[f] & key.m) | key.o) - 1.0f, lo);
            }
        } else {
            acc += __uint_as_float((c[0][0] ^ c[0][1] ^ c[0][2] ^ c[0][3] ^ c[1][0] ^ c[1][1] ^ c[1][2] ^ c[1][3]) >> 9);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double hz = khz * 1e3;
    float *out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    Keys key;
    for (int r = 0; r < 10; ++r) { key.k0[r] = 1 + r * 0x9E3779B9u; key.k1[r] = 2 + r * 0xBB67AE85u; }
    key.m = 0x7ffffc; key.o = 0x3f800002;
    const int iters = 4000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int conv = 0; conv < 2; ++conv)
        for (int cps : {1, 2, 3, 4, 6, 8}) {
            const int blocks = sms * cps;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (conv) mix_kernel<true><<<blocks, 256>>>(out, iters, key, -4.0f, 0.03125f);
                else mix_kernel<false><<<blocks, 256>>>(out, iters, key, -4.0f, 0.03125f);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
            }
            const double warp_blocks_per_smsp = cps * 2.0 * iters * 2.0;      // 8 warps per CTA / 4 sub-partitions, 2 blocks per iteration
            printf("convert %d, %d CTAs/SM (%d warps per sub-partition): %.3f ms, %.1f cycles per warp-block per sub-partition\n",
                   conv, cps, cps * 2, ms, ms * 1e-3 * hz / warp_blocks_per_smsp);
        }
    return 0;
}
