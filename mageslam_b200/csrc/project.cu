// Map-point projection + candidate culling + octave prediction for sm_100a.
//
// Replaces, for a whole local map at once, the per-point prologue of TrackLocalMap::ProjectMapPointIntoCurrentFrame
// (ref Core/MAGESLAM/Source/Tracking/TrackLocalMap.cpp:325-388): ProjectUndistorted (ref Tracking/Reprojection.cpp:26-43),
// IsGoodCandidate (ref TrackLocalMap.cpp:519-554: behind-camera / image-border test, viewing-angle test, scale-invariance
// range test) and ComputeOctave (ref Map/MappingMath.h:13-16). The output keypoints are the `mapPointKp` the reference hands
// to RadiusMatch (ref TrackLocalMap.cpp:371), so they feed mage_radius_match directly.
//
// f32 throughout, one thread per map point, 32 B in / 36 B out per point: a streaming kernel. Every product and sum is an
// explicitly rounded intrinsic in the reference's evaluation order (cv::Matx product: s = 0; s += a(i,k) * b(k); Point3f::dot:
// (x*x' + y*y') + z*z'), i.e. no FMA contraction, so the results equal a scalar C++ build without contraction bit for bit.
// The one exception is log2f in ComputeOctave: CUDA's and the host libm's log2f may differ in the last place, which can move
// the predicted octave only when log2(d/dmin)/log2(scale) lies within a few 1e-7 of an integer (documented in DESIGN.md).
#include "common.cuh"

namespace mage {

struct ProjConst {
    float v[12];
    float fx, fy, cx, cy;
    float px, py, pz;        // frame position
    float fwx, fwy, fwz;     // frame forward
    float min_cos, border, xmax, ymax, log2_scale;
    int num_levels;
};

__device__ __forceinline__ float row_dot4(const float* r, float x, float y, float z)
{
    // cv::Matx<float,3,4> * cv::Matx<float,4,1>: s = 0; for k: s += a(i,k) * b(k)   (0 + a*b is exact)
    float s = __fmul_rn(r[0], x);
    s = __fadd_rn(s, __fmul_rn(r[1], y));
    s = __fadd_rn(s, __fmul_rn(r[2], z));
    s = __fadd_rn(s, r[3]);                       // b(3) = 1: a * 1 is exact
    return s;
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

__global__ void __launch_bounds__(256) k_project_map_points(const __grid_constant__ ProjConst c, const mage_map_point* __restrict__ pts, int n,
                                                            mage_keypoint* __restrict__ out_kps, float* __restrict__ out_depth,
                                                            uint8_t* __restrict__ out_flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // 32-byte record = two 16-byte loads
    const float4 a = __ldg(reinterpret_cast<const float4*>(pts + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(pts + i) + 1);
    const float X = a.x, Y = a.y, Z = a.z, nx = a.w, ny = b.x, nz = b.y, dmin = b.z, dmax = b.w;

    // ProjectUndistorted
    const float cs0 = row_dot4(c.v, X, Y, Z), cs1 = row_dot4(c.v + 4, X, Y, Z), depth = row_dot4(c.v + 8, X, Y, Z);
    const float div = depth != 0.f ? depth : 1.f;
    const float u = __fadd_rn(__fmul_rn(__fdiv_rn(cs0, div), c.fx), c.cx);
    const float v = __fadd_rn(__fmul_rn(__fdiv_rn(cs1, div), c.fy), c.cy);

    // IsGoodCandidate
    bool good = !(depth < 0.f) && c.border <= u && c.border <= v && u < c.xmax && v < c.ymax;
    const float dotv = dot3(nx, ny, nz, c.fwx, c.fwy, c.fwz);
    good = good && !(dotv < c.min_cos);
    const float dx = __fsub_rn(X, c.px), dy = __fsub_rn(Y, c.py), dz = __fsub_rn(Z, c.pz);
    const float d2 = dot3(dx, dy, dz, dx, dy, dz);
    good = good && !(d2 < __fmul_rn(dmin, dmin)) && !(__fmul_rn(dmax, dmax) < d2);

    // ComputeOctave: static_cast<int>(roundf(log2f(distance / dmin) / log2f(scaleFactor) - 0.5f))
    int octave = 0;
    bool predicted = false;
    if (good) {
        const float dist = __fsqrt_rn(d2);
        const float t = __fsub_rn(__fdiv_rn(log2f(__fdiv_rn(dist, dmin)), c.log2_scale), 0.5f);
        octave = (int)roundf(t);
        predicted = octave >= 0 && octave <= c.num_levels;
    }
    mage_keypoint kp;                               // cv::KeyPoint(projected.Point, -1.0f, 0.0f, 0.0f, octave, -1)
    kp.x = u; kp.y = v; kp.size = -1.f; kp.angle = 0.f; kp.response = 0.f; kp.octave = octave; kp.class_id = -1;
    out_kps[i] = kp;
    if (out_depth) out_depth[i] = depth;
    out_flags[i] = (uint8_t)((good ? MAGE_PROJ_GOOD_CANDIDATE : 0) | (predicted ? MAGE_PROJ_PREDICTED : 0));
}

} // namespace mage

using namespace mage;

static int make_const(const mage_projection_params* p, ProjConst& c)
{
    MAGE_REQUIRE(p != nullptr, MAGE_ERR_INVALID, "mage_project_map_points: params is NULL");
    MAGE_REQUIRE(p->pyramid_scale > 0.f && p->pyramid_scale != 1.f, MAGE_ERR_INVALID, "mage_project_map_points: pyramid_scale must be > 0 and != 1");
    memcpy(c.v, p->view, sizeof(c.v));
    c.fx = p->fx; c.fy = p->fy; c.cx = p->cx; c.cy = p->cy;
    c.px = p->frame_position[0]; c.py = p->frame_position[1]; c.pz = p->frame_position[2];
    c.fwx = p->frame_forward[0]; c.fwy = p->frame_forward[1]; c.fwz = p->frame_forward[2];
    c.min_cos = p->min_cos_view_angle;
    c.border = p->image_border;
    c.xmax = (float)p->width - p->image_border;     // uint32_t GetWidth() - float imageBorder (ref Image/AnalyzedImage.h:118-122)
    c.ymax = (float)p->height - p->image_border;
    c.log2_scale = log2f(p->pyramid_scale);         // the host libm, as the reference evaluates it
    c.num_levels = (int)p->num_levels;
    return MAGE_OK;
}

extern "C" int mage_project_map_points_device(const mage_projection_params* params, const mage_map_point* d_points, int n,
                                              mage_keypoint* d_out_kps, float* d_out_depth, uint8_t* d_out_flags, void* cuda_stream)
{
    ProjConst c;
    int rc = make_const(params, c);
    if (rc != MAGE_OK) return rc;
    MAGE_REQUIRE(n >= 0, MAGE_ERR_INVALID, "mage_project_map_points: n < 0");
    if (n == 0) return MAGE_OK;
    MAGE_REQUIRE(d_points && d_out_kps && d_out_flags, MAGE_ERR_INVALID, "mage_project_map_points: NULL buffer");
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    k_project_map_points<<<div_up(n, 256), 256, 0, s>>>(c, d_points, n, d_out_kps, d_out_depth, d_out_flags);
    MAGE_CUDA_TRY(cudaGetLastError());
    return MAGE_OK;
}

extern "C" int mage_project_map_points(const mage_projection_params* params, const mage_map_point* points, int n, mage_keypoint* out_kps,
                                       float* out_depth, uint8_t* out_flags, void* cuda_stream)
{
    MAGE_REQUIRE(n >= 0, MAGE_ERR_INVALID, "mage_project_map_points: n < 0");
    if (n == 0) return MAGE_OK;
    MAGE_REQUIRE(points && out_kps && out_flags, MAGE_ERR_INVALID, "mage_project_map_points: NULL buffer");
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t o_pts = 0, o_kps = align_up(o_pts + sizeof(mage_map_point) * n, 256), o_dep = align_up(o_kps + sizeof(mage_keypoint) * n, 256);
    const size_t o_flg = align_up(o_dep + sizeof(float) * n, 256), total = o_flg + (size_t)n;
    DevStage dsg = dev_stage_acquire(total);                 // pooled device scratch: no allocation per call
    MAGE_REQUIRE(dsg.p, MAGE_ERR_CUDA, "mage_project_map_points: no device memory for %zu bytes", total);
    uint8_t* d = dsg.p;
    int rc = MAGE_OK;
    // the map points go up from, and the three result arrays come back into, ONE pinned staging buffer laid out like the device scratch
    // (a copy between the device and a caller's pageable array is a staged, synchronous transfer of the driver: four of them per call before)
    PinnedStage st = stage_acquire(total);
    if (!st.p) { dev_stage_release(dsg); MAGE_REQUIRE(false, MAGE_ERR_CUDA, "mage_project_map_points: no pinned staging memory"); }
    memcpy(st.p + o_pts, points, sizeof(mage_map_point) * n);
    cudaError_t e = cudaMemcpyAsync(d + o_pts, st.p + o_pts, sizeof(mage_map_point) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        rc = mage_project_map_points_device(params, reinterpret_cast<const mage_map_point*>(d + o_pts), n, reinterpret_cast<mage_keypoint*>(d + o_kps),
                                            reinterpret_cast<float*>(d + o_dep), d + o_flg, cuda_stream);
    }
    if (e == cudaSuccess && rc == MAGE_OK) e = cudaMemcpyAsync(st.p + o_kps, d + o_kps, total - o_kps, cudaMemcpyDeviceToHost, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    dev_stage_release(dsg);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess && rc == MAGE_OK) {
        memcpy(out_kps, st.p + o_kps, sizeof(mage_keypoint) * n);
        if (out_depth) memcpy(out_depth, st.p + o_dep, sizeof(float) * n);
        memcpy(out_flags, st.p + o_flg, (size_t)n);
    }
    stage_release(st);
    if (rc != MAGE_OK) return rc;
    MAGE_CUDA_TRY(e);
    return MAGE_OK;
}
