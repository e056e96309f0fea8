// Issue-rate probe for the packed 16-bit min/max and add instructions k_fast is built from (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o pipe_probe tools/pipe_probe.cu ; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed)
{
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + i + 1); b[i] = seed ^ (threadIdx.x * 77 + i); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __vimax3_s16x2(a[i], b[i], a[(i + 1) & 7]);
            if (MODE == 1) { __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b[i])); a[i] = *reinterpret_cast<uint32_t*>(&r); b[i] ^= a[(i + 1) & 7]; }
            if (MODE == 2) {
                a[i] = __vimax3_s16x2(a[i], b[i], a[(i + 1) & 7]);
                __half2 r = __hmax2(*reinterpret_cast<__half2*>(&b[i]), *reinterpret_cast<__half2*>(&a[(i + 3) & 7])); b[i] = *reinterpret_cast<uint32_t*>(&r);
            }
            if (MODE == 3) a[i] = __vadd2(a[i], b[i]);
            if (MODE == 4) asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
            if (MODE == 5) { a[i] = __vimax3_s16x2(a[i], b[i], a[(i + 1) & 7]); asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(b[i]) : "r"(b[i]), "r"(a[(i + 2) & 7])); }
            if (MODE == 6) a[i] = __vmaxs2(a[i], b[i]);
            if (MODE == 7) { a[i] = __vmaxs2(a[i], b[i]); b[i] = __vmins2(b[i], a[(i + 1) & 7]); }
            if (MODE == 8) { a[i] = max(max((int)a[i], (int)b[i]), (int)a[(i + 1) & 7]); }
            if (MODE == 9) { a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = __funnelshift_r(b[i], a[i], 16); }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int per_iter)
{
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(d, 12345u);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(d, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winstr = 148.0 * 8 * 8 * ITERS * 8 * per_iter;      // CTAs x warps x iters x unroll x instr
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, winstr / cycles / 148.0, clk / 1000);
    cudaFree(d);
}

int main()
{
    run<0>("VIMNMX3.S16x2", 1);
    run<1>("HMNMX2 (+LOP3)", 2);
    run<2>("VIMNMX3.S16x2 + HMNMX2", 2);
    run<3>("VIADD.16x2", 1);
    run<4>("IMAD (mad.lo x1)", 1);
    run<5>("VIMNMX3.S16x2 + IMAD", 2);
    run<6>("VIMNMX.S16x2 (2-input)", 1);
    run<7>("VIMNMX.S16x2 max+min", 2);
    run<8>("VIMNMX3.S32", 1);
    run<9>("VIMNMX3.U16x2 + SHF", 2);
    return 0;
}
