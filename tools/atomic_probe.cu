// Micro-probe: throughput of RETURNING atomicAdd (one lane per warp, as a warp-aggregated queue append issues it) when the
// whole grid targets 1, 2, 6, 64 or 4096 distinct addresses — the append counters of k_finish / k_shade are 2-6 words that
// every warp of the GPU hits once per 32 rays. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/atomic_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void hammer(unsigned* ctr, unsigned n_ctr, unsigned stride_words, int iters, unsigned* sink)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(ctr + size_t((warp + i) % n_ctr) * stride_words, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        acc += base;                                   // the append needs the returned position
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
}
__global__ void hammer_noret(unsigned* ctr, unsigned n_ctr, unsigned stride_words, int iters)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i)
        if (lane == 0) atomicAdd(ctr + size_t((warp + i) % n_ctr) * stride_words, 32u);      // RED: fire and forget
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    unsigned* ctr; cudaMalloc(&ctr, 4096 * 64 * 4); cudaMemset(ctr, 0, 4096 * 64 * 4);
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 256;
    printf("{\"gpu\": \"%s\", \"rows\": [", p.name);
    bool first = true;
    for (int wpsm : { 16, 32, 64 })
        for (unsigned n_ctr : { 1u, 2u, 6u, 64u, 4096u }) {
            const int blocks = p.multiProcessorCount * wpsm / 4, threads = 128;
            float ms_ret = 0, ms_red = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0); hammer<<<blocks, threads>>>(ctr, n_ctr, 64, iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms_ret, e0, e1);
                cudaEventRecord(e0); hammer_noret<<<blocks, threads>>>(ctr, n_ctr, 64, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms_red, e0, e1);
            }
            const double n = double(blocks) * (threads / 32) * iters;
            printf("%s{\"warps_per_sm\": %d, \"addresses\": %u, \"returning_Matomics_per_s\": %.1f, \"red_Matomics_per_s\": %.1f}", first ? "" : ", ",
                   wpsm, n_ctr, n / (ms_ret * 1e-3) / 1e6, n / (ms_red * 1e-3) / 1e6);
            first = false;
        }
    printf("]}\n");
    return cudaGetLastError() != cudaSuccess;
}
