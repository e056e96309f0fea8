// Probe 2: libcu++ wrappers, 2-D and 3-D maps, in-bounds and out-of-bounds coordinates
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int z, uint8_t* out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (RANK == 3) cde::cp_async_bulk_tensor_3d_global_to_shared(sm, &map, x, y, z, bar);
        else cde::cp_async_bulk_tensor_2d_global_to_shared(sm, &map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, 128 * 72);
    } else token = bar.arrive();
    bar.wait(std::move(token));
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
    printf("entry point %p query %d\n", fn, (int)q);
    cuuint64_t dim[3] = {W, H, F}, str[2] = {P, (cuuint64_t)P * H}; cuuint32_t box[3] = {128, 72, 1}, es[3] = {1, 1, 1};
    alignas(64) CUtensorMap m2, m3;
    CUresult r2 = ((Fn)fn)(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r3 = ((Fn)fn)(&m3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 2d %d 3d %d\n", (int)r2, (int)r3);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 72);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 72);
    for (int rank = 2; rank <= 3; rank++)
        for (int t = 0; t < 5; t++) {
            const int xs[5] = {16, -16, 592, 624, 20}, ys[5] = {8, -4, 444, 470, 8};
            int x = xs[t], y = ys[t], z = rank == 3 ? 2 : 0;
            cudaMemset(o, 0xEE, 128 * 72);
            if (rank == 2) k<2><<<1, 256, 128 * 72>>>(m2, x, y, z, o); else k<3><<<1, 256, 128 * 72>>>(m3, x, y, z, o);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rank %d (x=%d,y=%d): %s\n", rank, x, y, cudaGetErrorString(e)); return 2; }
            std::vector<uint8_t> h(128 * 72); cudaMemcpy(h.data(), o, h.size(), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 72; r++) for (int c = 0; c < 128; c++) {
                int yy = y + r, xx = x + c;
                uint8_t want = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(size_t)z * P * H + (size_t)yy * P + xx] : 0;
                bad += want != h[r * 128 + c];
            }
            printf("rank %d (x=%d,y=%d,z=%d): %d mismatches\n", rank, x, y, z, bad);
        }
    return 0;
}
