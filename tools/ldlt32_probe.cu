// One-warp 32 x 32 LDL^T variants (the serial core of the dense solver's diagonal blocks), timed with clock64 on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldlt32_probe tools/ldlt32_probe.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

constexpr int P = 130;

// V0: column through shared memory, pivot read back from it
__device__ void v0(double* Lt, double* colb, int lane)
{
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = c <= lane ? Lt[c * P + lane] : 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const double ck = a[k];
        double* cb = colb + (k & 1) * 32;
        cb[lane] = ck;
        __syncwarp();
        const double d = cb[k];
        double cj[32];
        if ((k + 1) & 1) { if (k + 1 < 32) cj[k + 1] = cb[k + 1]; }
#pragma unroll
        for (int j = (k + 2) & ~1; j < 32; j += 2) { const double2 c2 = *reinterpret_cast<const double2*>(cb + j); cj[j] = c2.x; cj[j + 1] = c2.y; }
        const double inv = fabs(d) > 0 ? __drcp_rn(d) : 0.0;
        const double l = ck * inv;
#pragma unroll
        for (int j = k + 1; j < 32; j++) a[j] -= l * cj[j];
        a[k] = lane == k ? d : l;
    }
#pragma unroll
    for (int c = 0; c < 32; c++) if (c <= lane) Lt[c * P + lane] = a[c];
}

// V1: as V0, but every lane also learns the next pivot's inputs (the diagonal entry and the column entry of row k + 1) from the
// exchange of column k and computes d_{k+1} and its reciprocal itself -- the reciprocal no longer waits for the next exchange
__device__ void v1(double* Lt, double* colb, int lane)
{
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = c <= lane ? Lt[c * P + lane] : 0.0;
    double dg = 0.0;                                              // this lane's current diagonal entry
#pragma unroll
    for (int c = 0; c < 32; c++) if (c == lane) dg = a[c];
    double* dgb = colb + 64;                                      // [2][32] diagonals
    // column 0 and the diagonals
    colb[lane] = a[0]; dgb[lane] = dg;
    __syncwarp();
    double d = colb[0];
    double inv = fabs(d) > 0 ? __drcp_rn(d) : 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const double* cb = colb + (k & 1) * 32;
        const double* db = dgb + (k & 1) * 32;
        double cj[32];
        if ((k + 1) & 1) { if (k + 1 < 32) cj[k + 1] = cb[k + 1]; }
#pragma unroll
        for (int j = (k + 2) & ~1; j < 32; j += 2) { const double2 c2 = *reinterpret_cast<const double2*>(cb + j); cj[j] = c2.x; cj[j + 1] = c2.y; }
        const double ck = a[k];
        const double l = ck * inv;
        double dn = 1.0, invn = 0.0;
        if (k + 1 < 32) {
            // next column of this lane first, published at once
            a[k + 1] -= l * cj[k + 1];
            dg -= l * ck;                                         // a(i, i) -= l_i c_i  (j = i term; harmless for lanes <= k)
            double* cbn = colb + ((k + 1) & 1) * 32;
            double* dbn = dgb + ((k + 1) & 1) * 32;
            cbn[lane] = a[k + 1]; dbn[lane] = dg;
            // the next pivot, computed by every lane: d_{k+1} = a(k+1, k+1) - c(k+1)^2 / d_k
            const double cn = cj[k + 1];
            dn = db[k + 1] - cn * cn * inv;
            invn = fabs(dn) > 0 ? __drcp_rn(dn) : 0.0;
        }
#pragma unroll
        for (int j = k + 2; j < 32; j++) a[j] -= l * cj[j];
        a[k] = lane == k ? d : l;
        d = dn; inv = invn;
        __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 32; c++) if (c <= lane) Lt[c * P + lane] = a[c];
}

// V2: registers + shuffles only
__device__ void v2(double* Lt, double*, int lane)
{
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = c <= lane ? Lt[c * P + lane] : 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const double ck = a[k];
        const double d = __shfl_sync(0xffffffffu, ck, k);
        const double inv = fabs(d) > 0 ? __drcp_rn(d) : 0.0;
        const double l = ck * inv;
#pragma unroll
        for (int j = k + 1; j < 32; j++) a[j] -= l * __shfl_sync(0xffffffffu, ck, j);
        a[k] = lane == k ? d : l;
    }
#pragma unroll
    for (int c = 0; c < 32; c++) if (c <= lane) Lt[c * P + lane] = a[c];
}


// V3: rolled loops, the block in a row-major shared scratch (pitch 34: a quarter-warp's 16-byte accesses to its own rows cover all
// banks once): small code -- the unrolled variants are 40 KB of straight-line instructions that miss the instruction cache every time
// they run inside the big kernel
constexpr int RP = 34;
__device__ void v3(double* Lt, double* colb, int lane)
{
    double* R = colb + 128;                                       // [32][RP]
    for (int c = 0; c < 32; c++) R[lane * RP + c] = c <= lane ? Lt[c * P + lane] : 0.0;
    __syncwarp();
    double* row = R + lane * RP;
#pragma unroll 1
    for (int k = 0; k < 32; k++) {
        const double ck = row[k];
        colb[lane] = ck;
        __syncwarp();
        const double d = colb[k];
        const double inv = fabs(d) > 0 ? __drcp_rn(d) : 0.0;
        const double l = ck * inv;
#pragma unroll 4
        for (int j = (k + 1) & ~1; j < 32; j += 2) {
            double2 c2 = *reinterpret_cast<const double2*>(colb + j);
            double2 a2 = *reinterpret_cast<double2*>(row + j);
            if (j <= k) c2.x = 0.0;
            a2.x -= l * c2.x; a2.y -= l * c2.y;
            *reinterpret_cast<double2*>(row + j) = a2;
        }
        if (lane > k) row[k] = l;
        __syncwarp();
    }
    for (int c = 0; c < 32; c++) if (c <= lane) Lt[c * P + lane] = row[c];
}

template <int V>
__global__ void probe(const double* A, double* out, long long* clk)
{
    extern __shared__ double sm[];
    double* Lt = sm; double* colb = sm + 32 * P;
    const int lane = threadIdx.x;
    for (int i = lane; i < 32 * 32; i += 32) { const int r = i / 32, c = i % 32; if (r >= c) Lt[c * P + r] = A[r * 32 + c]; }
    __syncwarp();
    long long t0 = clock64();
    if (V == 0) v0(Lt, colb, lane);
    if (V == 1) v1(Lt, colb, lane);
    if (V == 2) v2(Lt, colb, lane);
    if (V == 3) v3(Lt, colb, lane);
    long long t1 = clock64();
    __syncwarp();
    for (int i = lane; i < 32 * 32; i += 32) { const int r = i / 32, c = i % 32; out[r * 32 + c] = r >= c ? Lt[c * P + r] : 0.0; }
    if (lane == 0) *clk = t1 - t0;
}

int main()
{
    double hA[1024], ref[1024];
    srand(1);
    double M[32][40];
    for (auto& r : M) for (auto& v : r) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 32; i++) for (int j = 0; j < 32; j++) { double s = i == j ? 0.5 : 0; for (int k = 0; k < 40; k++) s += M[i][k] * M[j][k]; hA[i * 32 + j] = s; }
    for (int i = 0; i < 1024; i++) ref[i] = hA[i];
    for (int k = 0; k < 32; k++) {                                 // reference LDL^T (right-looking)
        const double d = ref[k * 32 + k];
        for (int i = k + 1; i < 32; i++) { const double l = ref[i * 32 + k] / d; for (int j = k + 1; j <= i; j++) ref[i * 32 + j] -= l * ref[j * 32 + k]; }
        for (int i = k + 1; i < 32; i++) ref[i * 32 + k] /= d;
    }
    double *dA, *dO; long long* dC;
    cudaMalloc(&dA, 8192); cudaMalloc(&dO, 8192); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, hA, 8192, cudaMemcpyHostToDevice);
    const char* names[] = {"V0 shared-memory column exchange", "V1 + look-ahead pivot (redundant d, 1/d)", "V2 shuffles only", "V3 rolled loops, row-major shared scratch"};
    for (int v = 0; v < 4; v++) {
        long long c = 0, c0 = 0; double o[1024];
        for (int rep = 0; rep < 2; rep++) {
            if (v == 0) probe<0><<<1, 32, (32 * P + 256 + 32 * 34) * 8>>>(dA, dO, dC);
            if (v == 1) probe<1><<<1, 32, (32 * P + 256 + 32 * 34) * 8>>>(dA, dO, dC);
            if (v == 2) probe<2><<<1, 32, (32 * P + 256 + 32 * 34) * 8>>>(dA, dO, dC);
            if (v == 3) probe<3><<<1, 32, (32 * P + 256 + 32 * 34) * 8>>>(dA, dO, dC);
            if (rep == 0) { cudaDeviceSynchronize(); cudaMemcpy(&c0, dC, 8, cudaMemcpyDeviceToHost); }
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o, dO, 8192, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < 32; i++) for (int j = 0; j <= i; j++) err = fmax(err, fabs(o[i * 32 + j] - ref[i * 32 + j]));
        printf("%-44s %6lld clk  (%.0f per column; first launch, cold instruction cache: %lld)  max |err| %.2e  %s\n", names[v], c, c / 32.0, c0, err, cudaGetErrorString(e));
    }
    return 0;
}
