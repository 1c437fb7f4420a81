// Micro-probe: FP64 vs FP32 FMA issue rate and double min/max rate on the current GPU.
// The extend kernel's design (double-precision slab + primitive tests) assumes B200 issues DFMA at half the
// FFMA rate; this measures it. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_loop(T* out, int iters)
{
    T a = T(threadIdx.x) * T(1e-3), b = T(1.000001), c = T(1e-6);
    T x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * b + c; x1 = x1 * b + c; x2 = x2 * b + c; x3 = x3 * b + c;
        x4 = x4 * b + c; x5 = x5 * b + c; x6 = x6 * b + c; x7 = x7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void dminmax_loop(double* out, int iters)
{
    double a = threadIdx.x * 1e-3;
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, lo = 0.5, hi = 1e9;
    for (int i = 0; i < iters; ++i) {
        x0 = fmax(fmin(x0 + 1.0, hi), lo); x1 = fmax(fmin(x1 + 1.0, hi), lo);
        x2 = fmax(fmin(x2 + 1.0, hi), lo); x3 = fmax(fmin(x3 + 1.0, hi), lo);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    void* buf; cudaMalloc(&buf, size_t(blocks) * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); fma_loop<float><<<blocks, threads>>>((float*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double f32 = 2.0 * 8 * iters * double(blocks) * threads / (ms * 1e-3) / 1e12;
        cudaEventRecord(e0); fma_loop<double><<<blocks, threads>>>((double*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double f64 = 2.0 * 8 * iters * double(blocks) * threads / (ms * 1e-3) / 1e12;
        cudaEventRecord(e0); dminmax_loop<<<blocks, threads>>>((double*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double mm = 3.0 * 4 * iters * double(blocks) * threads / (ms * 1e-3) / 1e12;
        if (rep) printf("{\"gpu\": \"%s\", \"sms\": %d, \"ffma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"dfma_over_ffma\": %.3f, \"dadd_min_max_tops\": %.2f}\n",
                        p.name, p.multiProcessorCount, f32, f64, f64 / f32, mm);
    }
    return cudaGetLastError() != cudaSuccess;
}
