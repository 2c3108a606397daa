// Microbenchmark: what keeps a gather table L2 resident while data streams past it?
// Every thread does, per step: one random 32-byte sector gather from a table of T MB, plus a coalesced read of S bytes from a
// large buffer that is touched once (the loc / header / stream traffic of screen_bits). Variants: no hints, or createpolicy
// evict_last on the gathers + evict_first on the stream. Prints time per launch; run under
//   ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
// to see which variant keeps the table out of DRAM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_mix gather_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int HINT>
__global__ void mix(const uint4 *tab, uint64_t nsec, const uint4 *stream, uint64_t stream_vec, uint32_t vec_per_step, uint32_t steps, uint32_t seed, uint64_t *out) {
    uint64_t keep, strm;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(strm));
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + seed, acc = 0;
    for (uint32_t i = 0; i < steps; i++) {
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
        const uint64_t idx = ((x * 0x2545F4914F6CDD1Dull) >> 11) % nsec;
        uint64_t a, b, c, d;
        if (HINT) asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(tab + 2 * idx), "l"(keep));
        else asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(tab + 2 * idx));
        acc += a ^ b ^ c ^ d;
        for (uint32_t v = 0; v < vec_per_step; v++) {
            const uint64_t s = (((uint64_t)i * vec_per_step + v) * nthr + tid) % stream_vec;      // coalesced, every vector read once
            uint32_t p, q, r, t;
            if (HINT) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(p), "=r"(q), "=r"(r), "=r"(t) : "l"(stream + s), "l"(strm));
            else asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(p), "=r"(q), "=r"(r), "=r"(t) : "l"(stream + s));
            acc += p ^ q ^ r ^ t;
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}
int main() {
    const size_t tab_max = 256ull << 20, stream_bytes = 4096ull << 20;
    uint4 *tab, *stream; uint64_t *out;
    cudaMalloc(&tab, tab_max); cudaMalloc(&stream, stream_bytes); cudaMalloc(&out, 8);
    cudaMemset(tab, 1, tab_max); cudaMemset(stream, 2, stream_bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256; const uint32_t steps = 96;      // 29 M gathers per launch, like one big screen_bits launch
    const double mbs[] = {31.25, 62.5, 125.0};
    for (double mb : mbs) for (int vec : {0, 3}) for (int hint = 0; hint < 2; hint++) {
        const uint64_t nsec = (uint64_t)(mb * 1048576.0) / 32;
        for (int rep = 0; rep < 2; rep++) {                                     // rep 0 warms the table
            cudaEventRecord(e0);
            if (hint) mix<1><<<blocks, threads>>>(tab, nsec, stream, stream_bytes / 16, vec, steps, 7 + rep, out);
            else mix<0><<<blocks, threads>>>(tab, nsec, stream, stream_bytes / 16, vec, steps, 7 + rep, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double n = (double)blocks * threads * steps;
        printf("table %6.2f MB, stream %2d B/gather, hints %d: %.3f ms, %.1f G gathers/s\n", mb, vec * 16, hint, ms, n / ms / 1e6);
    }
    return 0;
}
