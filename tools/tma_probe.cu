// Probe: 3-D TMA tile load of a byte image (x, y, frame) with out-of-bounds start coordinates, descriptor passed three ways.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
struct Maps { CUtensorMap m[4]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ Maps maps, int idx, const CUtensorMap* gmap, int mode, int x, int y, int z, uint8_t* out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* d = mode == 0 ? &maps.m[0] : mode == 1 ? &maps.m[idx] : gmap;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * 72) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sm)),
                     "l"(reinterpret_cast<uint64_t>(d)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n .reg .pred p;\n W:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 128 * 72; i += blockDim.x) out[i] = sm[i];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    const int W = 640, H = 480, P = 640, F = 3;
    std::vector<uint8_t> img((size_t)P * H * F);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d, *o; cudaMalloc(&d, img.size()); cudaMalloc(&o, 128 * 72);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Maps maps;
    cuuint64_t dim[3] = {W, H, F}, str[2] = {P, (cuuint64_t)P * H}; cuuint32_t box[3] = {128, 72, 1}, es[3] = {1, 1, 1};
    for (int i = 0; i < 4; i++) {
        CUresult r = ((Fn)fn)(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &maps.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 72);
    for (int mode = 0; mode < 3; mode++) {
        for (int t = 0; t < 2; t++) {
            int x = t ? 596 : -4, y = t ? 444 : -4, z = t ? 2 : 1;
            cudaMemset(o, 0xEE, 128 * 72);
            k<<<1, 256, 128 * 72>>>(maps, 2, gm, mode, x, y, z, o);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d test %d: %s\n", mode, t, cudaGetErrorString(e)); return 2; }
            std::vector<uint8_t> h(128 * 72); cudaMemcpy(h.data(), o, h.size(), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 72; r++) for (int c = 0; c < 128; c++) {
                int yy = y + r, xx = x + c;
                uint8_t want = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(size_t)z * P * H + (size_t)yy * P + xx] : 0;
                bad += want != h[r * 128 + c];
            }
            printf("mode %d (x=%d,y=%d,z=%d): %d mismatches\n", mode, x, y, z, bad);
        }
    }
    return 0;
}
