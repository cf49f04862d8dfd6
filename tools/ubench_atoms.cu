// ubench_atoms.cu -- measures on THIS GPU the two instruction classes DESIGN.md 6.1 says bind the partition kernels:
// a shared-memory atomic with spread addresses (rank = atomicAdd(&cnt[bin], 1) over 1024 bins) and a spread global
// reduction (atomicAdd without result into a 2 MB L2-resident histogram).  Prints SM cycles per warp instruction.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_atoms tools/ubench_atoms.cu && /tmp/ubench_atoms
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13; return x; }

template <int MODE>   // 0: spread shared atomic with result, 1: spread global reduction, 2: the hash alone (baseline)
__global__ void __launch_bounds__(512) k(uint32_t *hist, unsigned iters, uint32_t *sink) {
    __shared__ uint32_t cnt[1024];
    for (int i = threadIdx.x; i < 1024; i += 512) cnt[i] = 0;
    __syncthreads();
    uint32_t x = blockIdx.x * 512 + threadIdx.x, acc = 0;
    for (unsigned it = 0; it < iters; ++it) {
        x = mix(x + it);
        if (MODE == 0) acc += atomicAdd(&cnt[x & 1023], 1u);
        if (MODE == 1) atomicAdd(hist + (x & ((1u << 19) - 1)), 1u);
        if (MODE == 2) acc += x;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE>
float run(uint32_t *hist, uint32_t *sink, int grid, unsigned iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 512>>>(hist, iters, sink);
    cudaEventRecord(a);
    k<MODE><<<grid, 512>>>(hist, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t *hist, *sink;
    cudaMalloc(&hist, 4u << 19); cudaMemset(hist, 0, 4u << 19);
    cudaMalloc(&sink, 4);
    const int grid = p.multiProcessorCount * 3;                  // 3 CTAs of 512 threads per SM, like the partition kernels
    const unsigned iters = 4096;
    const float base = run<2>(hist, sink, grid, iters), ats = run<0>(hist, sink, grid, iters), red = run<1>(hist, sink, grid, iters);
    const double warp_instr_per_sm = 3.0 * 16 * iters;           // warps per SM x iterations
    auto cyc = [&](float ms) { return ms * 1e-3 * clk_khz * 1e3 / warp_instr_per_sm; };
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"hash_only_cyc_per_warp_instr\": %.2f, \"shared_atomic_spread_cyc_per_warp_instr\": %.2f, "
           "\"global_red_spread_cyc_per_warp_instr\": %.2f}\n", p.name, p.multiProcessorCount, clk_khz / 1e3, cyc(base), cyc(ats), cyc(red));
    return 0;
}
