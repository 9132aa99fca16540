// Micro-benchmark: DFMA dependent latency and throughput per SM sub-partition on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_micro fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double v[CHAINS];
    for (int c = 0; c < CHAINS; c++) v[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) v[c] = fma(v[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
void run(int threads, int blocks) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<CHAINS><<<blocks, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CHAINS><<<blocks, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters * 8 * CHAINS;   // DFMA per thread
    double tf = n * threads * blocks * 2 / (ms * 1e-3) / 1e12;
    printf("chains=%2d threads/blk=%4d blocks=%4d: %.2f cycles per DFMA per warp, %.2f TFLOP/s total\n", CHAINS, threads, blocks,
           (double)c / n, tf);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(32, 1); run<2>(32, 1); run<4>(32, 1); run<8>(32, 1); run<16>(32, 1);
    run<8>(128, 1); run<8>(256, 1); run<8>(512, 1);
    run<8>(128, 148); run<8>(256, 148); run<8>(512, 148); run<8>(1024, 148); run<4>(256, 148 * 2);
    return 0;
}
