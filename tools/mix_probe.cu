// Does non-FP64 work issue in the shadow of the FP64 pipe on sm_100?  Times a DFMA-only loop against the same loop
// with K extra independent FP32 FMAs / integer ops per DFMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int KF, int KI>
__global__ void __launch_bounds__(256) mix(double* out, int iters, double a, double b, float fa, float fb, int ia) {
    double x[8]; float f[8]; int n[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-3f + i; n[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], a, b);
#pragma unroll
                for (int k = 0; k < KF; ++k) f[(i + k) & 7] = fmaf(f[(i + k) & 7], fa, fb);
#pragma unroll
                for (int k = 0; k < KI; ++k) n[(i + k) & 7] = n[(i + k) & 7] * ia + 12345;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + f[i] + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KF, int KI> float run(double* d, int grid) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<KF, KI><<<grid, 256>>>(d, 2000, 0.999999, 1e-9, 0.9999f, 1e-5f, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) mix<KF, KI><<<grid, 256>>>(d, 2000, 0.999999, 1e-9, 0.9999f, 1e-5f, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int grid = p.multiProcessorCount * 8;
    double* d; cudaMalloc(&d, sizeof(double) * grid * 256);
    printf("DFMA only            : %.3f ms\n", run<0, 0>(d, grid));
    printf("DFMA + 1 FFMA each   : %.3f ms\n", run<1, 0>(d, grid));
    printf("DFMA + 2 FFMA each   : %.3f ms\n", run<2, 0>(d, grid));
    printf("DFMA + 1 IMAD each   : %.3f ms\n", run<0, 1>(d, grid));
    printf("DFMA + 1 FFMA+1 IMAD : %.3f ms\n", run<1, 1>(d, grid));
    printf("DFMA + 2 FFMA+2 IMAD : %.3f ms\n", run<2, 2>(d, grid));
    return 0;
}
