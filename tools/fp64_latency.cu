// Single-warp latency / issue probe of the FP64 instructions the dense solver's serial parts are made of (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/fp64_latency tools/fp64_latency.cu ; prints clocks per operation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 512
template <int MODE>
__global__ void probe(double* out, long long* clk, double seed)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = seed + lane; sm[lane + 32] = seed - lane;
    __syncwarp();
    double a = seed + lane * 1e-3, b = 1.0000001, c = 1e-9;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a + i;
    int iv = (int)seed + lane;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (MODE == 0) a = fma(a, b, c);                                   // dependent DFMA
        if (MODE == 1) { 
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = fma(x[j], b, c); }           // 8 independent DFMA chains
        if (MODE == 2) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);   // dependent 64-bit shuffle (2 SHFL)
        if (MODE == 3) a = sm[(__double2loint(a) + i) & 63];             // dependent LDS.64
        if (MODE == 4) a = __drcp_rn(a) + 0.5;                             // reciprocal + DADD
        if (MODE == 5) { iv = (int)(a); a = (double)(iv + i) ; }          // F2I + I2F
        if (MODE == 6) a = a + b;                                          // dependent DADD
        if (MODE == 7) a = 1.0 / a + 0.5;                                  // IEEE division + DADD
        if (MODE == 8) { a = __hiloint2double(0x43300000, iv ^ 0x80000000) - 4503601774854144.0; iv = __double2loint(a) + i; }   // magic int->double (+ dependent int op)
        if (MODE == 9) { 
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = (double)(__double2loint(x[j]) + j); }    // 8 independent (F2I.lo, IADD, I2F.F64)
        if (MODE == 10) a = rsqrt(a) + 1.5;                                // FP64 rsqrt + DADD
    }
    long long t1 = clock64();
    double r = a + iv;
#pragma unroll
    for (int i = 0; i < 8; i++) r += x[i];
    out[blockIdx.x * 32 + lane] = r;
    if (lane == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops)
{
    double* d; long long* c; cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 8);
    probe<MODE><<<1, 32>>>(d, c, 1.25);
    probe<MODE><<<1, 32>>>(d, c, 1.25);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %7.1f clk per op (one warp)\n", name, (double)h / N / ops);
    cudaFree(d); cudaFree(c);
}

int main()
{
    run<0>("DFMA, dependent chain", 1);
    run<6>("DADD, dependent chain", 1);
    run<1>("DFMA, 8 independent chains (per instruction)", 8);
    run<2>("64-bit shuffle (2 SHFL), dependent", 1);
    run<3>("LDS.64, dependent (address from the value)", 1);
    run<4>("__drcp_rn + DADD, dependent", 1);
    run<7>("1.0 / x + DADD, dependent", 1);
    run<10>("rsqrt(double) + DADD, dependent", 1);
    run<5>("F2I.F64 + IADD + I2F.F64, dependent", 1);
    run<9>("I2F.F64 (+ lo-word extract + IADD), 8 independent", 8);
    run<8>("magic int->double (LOP3 + DADD) + lo-word + IADD", 1);
    return 0;
}
