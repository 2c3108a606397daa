// Microbenchmark: random 32-byte sector gathers (one 256-bit load per thread) over a buffer of X MB.
// Reports sectors/s vs footprint: shows the effective L2 capacity for the candidate-window access pattern.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void gather(const uint4 *buf, uint64_t nsec, uint32_t iters, uint32_t seed, uint64_t *out) {
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + seed;
    uint64_t acc = 0;
    for (uint32_t i = 0; i < iters; i += 4) {
        uint64_t idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; idx[k] = ((x * 0x2545F4914F6CDD1Dull) >> 11) % nsec; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t a, b, c, d;
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(buf + 2 * idx[k]));
            acc += a ^ b ^ c ^ d;
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}
int main() {
    const size_t maxb = 1024ull << 20;
    uint4 *buf; uint64_t *out; cudaMalloc(&buf, maxb); cudaMalloc(&out, 8); cudaMemset(buf, 1, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int mbs[] = {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256, 384, 512, 1024};
    for (int mb : mbs) {
        uint64_t nsec = ((uint64_t)mb << 20) / 32;
        const int blocks = 148 * 8, threads = 256; const uint32_t iters = 2048;
        gather<<<blocks, threads>>>(buf, nsec, 256, 1, out);   // warm
        cudaEventRecord(e0);
        gather<<<blocks, threads>>>(buf, nsec, iters, 7, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)blocks * threads * iters;
        printf("footprint %5d MB: %.1f G sectors/s = %.2f TB/s of 32-byte sectors\n", mb, n / ms / 1e6, n * 32 / ms / 1e9);
    }
    return 0;
}
