// FP64 / INT32 issue-rate microbenchmark (measurement aid): many warps, independent dependency chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) { a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c; a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_popc(unsigned* out, int iters) {
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, s = 0;
    for (int i = 0; i < iters; i++) { s += __popc(a0) + __popc(a1) + __popc(a2) + __popc(a3); a0 += 0x9E3779B9u; a1 += 0x7F4A7C15u; a2 ^= a0; a3 += a1; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        int iters = 20000;
        cudaEventRecord(e0); k_dfma<<<148 * 4, 512>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * 148.0 * 4 * 512;
        printf("DFMA: %.3f ms -> %.2f TFLOP/s FP64 (%.1f DFMA lanes/clk/SM at 1.965 GHz)\n", ms, flops / ms / 1e9, flops / 2 / (ms * 1e-3) / 148 / 1.965e9);
        cudaEventRecord(e0); k_popc<<<148 * 4, 512>>>((unsigned*)d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = 4.0 * iters * 148.0 * 4 * 512;
        printf("POPC: %.3f ms -> %.1f popc lanes/clk/SM\n", ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
