// Micro-benchmark: how fast can an SM evaluate Philox4x32-10, and which pipe limits it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o philox_rate philox_rate.cu && ./philox_rate
// Prints cycles per warp-instruction-equivalent per SM sub-partition for a few instruction mixes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    if (MODE == 0) {                       // one 32x32 -> 64 multiply (IMAD.WIDE.U32)
        const uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    } else {                               // separate low and high multiplies
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    }
}

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) philox_kernel(uint32_t *out, int iters, uint32_t k0, uint32_t k1) {
    uint32_t c[ILP][4];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = blockIdx.x; c[j][2] = j; c[j][3] = 7; }
    for (int it = 0; it < iters; ++it) {
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                uint32_t h0, l0, h1, l1;
                mulhilo<MODE>(0xD2511F53u, c[j][0], h0, l0);
                mulhilo<MODE>(0xCD9E8D57u, c[j][2], h1, l1);
                const uint32_t y = c[j][1], w = c[j][3];
                c[j][0] = h1 ^ y ^ ka; c[j][1] = l1; c[j][2] = h0 ^ w ^ kb; c[j][3] = l0;
            }
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc ^= c[j][0] ^ c[j][1] ^ c[j][2] ^ c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// pure pipes: chains of one instruction kind
template <int KIND, int ILP>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t *out, int iters, uint32_t m) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { x[j] = threadIdx.x + j; y[j] = blockIdx.x + 3 * j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (KIND == 0) { const uint64_t p = (uint64_t)x[j] * m + y[j]; x[j] = (uint32_t)p; y[j] = (uint32_t)(p >> 32); }   // IMAD.WIDE
                else if (KIND == 1) { x[j] = x[j] * m + y[j]; }                                  // IMAD
                else if (KIND == 2) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(y[j]), "r"(m)); }   // LOP3
                else if (KIND == 3) { x[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.0001f, __uint_as_float(y[j]))); } // FFMA
                else if (KIND == 4) { x[j] = __umulhi(x[j], m) + y[j]; }                             // IMAD.HI
                else if (KIND == 5) { x[j] = __funnelshift_r(x[j], y[j], 19); }                     // SHF
                else if (KIND == 6) { x[j] = __float_as_uint(__uint2float_rz(x[j])); }             // I2F
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc ^= x[j] ^ y[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double hz = khz * 1e3;
    uint32_t *out;
    const int blocks = sms * 8, threads = 256;      // 64 warps per SM, 16 per sub-partition
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    const int iters = 2000;
    const double warps_per_smsp = 8.0 * (threads / 32) / 4.0;
    printf("SMs %d, clock %.0f MHz (nominal; cycles below assume it)\n", sms, hz / 1e6);
#define PHILOX(MODE, ILP) { float ms = time_ms([&] { philox_kernel<MODE, ILP><<<blocks, threads>>>(out, iters, 1u, 2u); }); \
        double blocks_per_smsp = warps_per_smsp * iters * ILP; \
        printf("philox mode %d ilp %d: %.3f ms, %.1f cycles per warp-block per sub-partition, %.3e blocks/s/GPU\n", MODE, ILP, ms, \
               ms * 1e-3 * hz / blocks_per_smsp, (double)blocks * threads * iters * ILP / (ms * 1e-3)); }
    PHILOX(0, 1) PHILOX(0, 2) PHILOX(0, 4) PHILOX(1, 1) PHILOX(1, 2) PHILOX(1, 4)
    const char *names[] = {"IMAD.WIDE", "IMAD", "LOP3", "FFMA", "IMAD.HI", "SHF", "I2F"};
#define PIPE(KIND, ILP) { float ms = time_ms([&] { pipe_kernel<KIND, ILP><<<blocks, threads>>>(out, iters, 0x9E3779B9u); }); \
        double n = warps_per_smsp * iters * 16.0 * ILP; \
        printf("%-10s ilp %d: %.3f ms, %.2f cycles per warp-instruction per sub-partition\n", names[KIND], ILP, ms, ms * 1e-3 * hz / n); }
    PIPE(0, 4) PIPE(1, 4) PIPE(2, 4) PIPE(3, 4) PIPE(4, 4) PIPE(5, 4) PIPE(6, 4) PIPE(0, 8) PIPE(2, 8)
    return 0;
}
