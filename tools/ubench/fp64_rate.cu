// micro-benchmark: DFMA / FFMA / LDS.64 issue rate per SM sub-partition on this GPU (answers "how expensive is float64?")
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu ; run: ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_chain(T* out, int iters, long long* cyc)
{
    T a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i);
    const T b = (T)1.0000001, c = (T)0.5;
    __syncthreads();
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
    }
    const long long t1 = clock64();
    T s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void dep_chain(double* out, int iters, long long* cyc)
{
    double a = threadIdx.x;
    const double b = 1.0000001, c = 0.5;
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) a = a * b + c;
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    double* d; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&c, 1 << 16);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        long long h;
        fma_chain<double><<<1, 32 * warps>>>(d, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        double fp64 = (double)h / (iters * 8.0);
        fma_chain<float><<<1, 32 * warps>>>((float*)d, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        double fp32 = (double)h / (iters * 8.0);
        printf("warps/SM %2d (per SMSP %.2f): cycles per warp-DFMA %.2f -> %.1f DFMA lanes/clk/SM | per warp-FFMA %.2f -> %.1f lanes/clk/SM\n",
               warps, warps / 4.0, fp64, 32.0 * warps / fp64, fp32, 32.0 * warps / fp32);
    }
    long long h;
    dep_chain<<<1, 32>>>(d, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.1f cycles\n", (double)h / iters);
    return 0;
}
