// Micro-benchmark: FP64 FMA issue rate and shared-memory broadcast load rate on one GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_chain(double* out, int iters, double a, double b) {
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_chain(float* out, int iters, float a, float b) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// broadcast LDS.128 + 2 DFMA per load (the stencil's inner loop shape)
__global__ void lds_dfma(double* out, int iters) {
    __shared__ __align__(16) double k[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) k[i] = 1.0 / (1 + i);
    __syncthreads();
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double2 kk = *reinterpret_cast<const double2*>(&k[((it * 8 + i) * 2) & 511]);
            v[i] = fma(v[i], kk.x, kk.y);
        }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* o;
    cudaMalloc(&o, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dfma_chain<<<blocks, threads>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  : %.2f TFLOP/s (%.3f ms)\n", 2.0 * blocks * threads * 8.0 * iters / ms / 1e9, ms);
        cudaEventRecord(e0); ffma_chain<<<blocks, threads>>>((float*)o, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA  : %.2f TFLOP/s (%.3f ms)\n", 2.0 * blocks * threads * 8.0 * iters / ms / 1e9, ms);
        cudaEventRecord(e0); lds_dfma<<<blocks, threads>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("LDS128+DFMA: %.2f TFLOP/s (%.3f ms)\n", 2.0 * blocks * threads * 8.0 * iters / ms / 1e9, ms);
    }
    printf("SMs %d clock %d kHz\n", p.multiProcessorCount, p.clockRate);
    return 0;
}
