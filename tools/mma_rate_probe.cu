// Measures the issue rate of the legacy warp-level MMA paths on this GPU (IMMA m16n8k32 u8, HMMA m16n8k16 f16 with f16 / f32 accumulate):
// decides whether a tensor-pipe pre-filter can replace the xor/popc filter of the brute-force matcher. nvcc -arch=sm_100a, run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, int* out)
{
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11, b1 = threadIdx.x * 13;
    int c[4][4] = {};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {           // 4 independent accumulator chains
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(c[j][0]), "+r"(c[j][1]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0;
    for (int j = 0; j < 4; j++) for (int q = 0; q < 4; q++) s += c[j][q];
    if (s == 123456789) out[0] = s;
}

template <int MODE>
void run(const char* name, int k_per_mma)
{
    int* d; cudaMalloc(&d, 4);
    const int iters = 4096, grid = 148 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(iters, d);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(iters, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)grid * 8 * iters * 4;                 // warp-level MMAs issued
    const double per_smsp_clk = ms * 1e-3 * 1.965e9 / (mmas / (148.0 * 4));
    printf("%-28s %s  %.3f ms  %.2f clk per MMA per SM sub-partition (at 1965 MHz)  %.1f Tmac/s\n", name, cudaGetErrorString(e), ms, per_smsp_clk,
           mmas * 16 * 8 * k_per_mma / (ms * 1e-3) / 1e12);
    cudaFree(d);
}

int main()
{
    run<0>("IMMA m16n8k32 u8 s32", 32);
    run<1>("HMMA m16n8k16 f16 acc f16", 16);
    run<2>("HMMA m16n8k16 f16 acc f32", 16);
    run<3>("QMMA m16n8k32 e4m3 acc f32", 32);
    return 0;
}
