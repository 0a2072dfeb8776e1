// Micro-benchmark 3: which instructions hide in the shadow of Philox's wide multiplies?
// Philox4x32-10 (2 blocks in flight per thread) plus NA integer-logic and NF fp32 (and NI I2F)
// instructions per block that depend on the block's output.  Prints cycles per warp-block per
// SM sub-partition at 8 warps per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Keys { uint32_t k0[10], k1[10]; };

template <int NA, int NF, int NI>
__global__ void __launch_bounds__(256) shadow_kernel(float *out, int iters, const __grid_constant__ Keys key, float lo, float w) {
    float facc[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t iacc[4] = {1, 2, 3, 4};
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
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int a = 0; a < NA; ++a)        // integer logic: independent 3-input ops
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iacc[a & 3]) : "r"(c[j][a & 3]), "r"(c[j][(a + 1) & 3]));
#pragma unroll
            for (int f = 0; f < NF; ++f)        // fp32: independent fmas on the block's words
                facc[f & 3] = fmaf(__uint_as_float(c[j][f & 3] | 0x3f800000u), w, facc[f & 3]);
#pragma unroll
            for (int i = 0; i < NI; ++i)
                facc[i & 3] += __uint2float_rz(c[j][i & 3]);
        }
        if (NA == 0 && NF == 0 && NI == 0) iacc[0] ^= c[0][0] ^ c[1][1] ^ c[0][2] ^ c[1][3];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = facc[0] + facc[1] + facc[2] + facc[3] + lo +
                                                 __uint_as_float((iacc[0] ^ iacc[1] ^ iacc[2] ^ iacc[3]) >> 9);
}

template <int NA, int NF, int NI>
void run(float *out, int sms, double hz, const Keys &key) {
    const int iters = 4000, cps = 4, blocks = sms * cps;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        shadow_kernel<NA, NF, NI><<<blocks, 256>>>(out, iters, key, -4.0f, 0.03125f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    const double warp_blocks_per_smsp = cps * 2.0 * iters * 2.0;
    printf("per block: +%2d logic +%2d fp32 +%2d i2f(+fadd): %.1f cycles per warp-block per sub-partition\n", NA, NF, NI,
           ms * 1e-3 * hz / warp_blocks_per_smsp);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    Keys key;
    for (int r = 0; r < 10; ++r) { key.k0[r] = 1 + r * 0x9E3779B9u; key.k1[r] = 2 + r * 0xBB67AE85u; }
    const int sms = p.multiProcessorCount;
    const double hz = khz * 1e3;
    run<0, 0, 0>(out, sms, hz, key);
    run<6, 0, 0>(out, sms, hz, key);
    run<12, 0, 0>(out, sms, hz, key);
    run<24, 0, 0>(out, sms, hz, key);
    run<0, 6, 0>(out, sms, hz, key);
    run<0, 12, 0>(out, sms, hz, key);
    run<0, 24, 0>(out, sms, hz, key);
    run<12, 12, 0>(out, sms, hz, key);
    run<0, 0, 6>(out, sms, hz, key);
    run<12, 6, 6>(out, sms, hz, key);
    return 0;
}
