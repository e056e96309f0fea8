// Keypoint undistortion for sm_100a: OrbFeatureDetector::UndistortKeypoints (ref Core/MAGESLAM/Source/Image/OrbFeatureDetector.cpp:30-62)
// = cv::undistortPoints(pts, K_distorted, distCoeffs, noArray(), K_undistorted) applied to every keypoint's pt.
// One thread per keypoint, FP64 like OpenCV, every operation an explicitly rounded intrinsic (no FMA contraction): the result is
// bit-identical to the CPU evaluation. Runs on the detector's output while it is still in HBM (28-byte cv::KeyPoint records).
#include "common.cuh"

namespace mage {

struct UndistortConst {
    double k[8];
    double fx, fy, cx, cy, ifx, ify;
    double P[9];
    int has_dist;
};

__global__ void __launch_bounds__(256) k_undistort_keypoints(const __grid_constant__ UndistortConst c, mage_keypoint* __restrict__ kps,
                                                             const int* __restrict__ counts, int per_frame, int n_total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    if (counts && (i % per_frame) >= counts[i / per_frame]) return;          // slot beyond the frame's keypoint count
    double x = (double)kps[i].x, y = (double)kps[i].y;
    const double u = x, v = y;
    x = __dmul_rn(__dsub_rn(x, c.cx), c.ifx);
    y = __dmul_rn(__dsub_rn(y, c.cy), c.ify);
    if (c.has_dist) {
        const double x0 = x, y0 = y;
#pragma unroll 1
        for (int j = 0; j < 5; j++) {                                         // TermCriteria(MAX_ITER, 5, 0.01): five fixed-point steps
            const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), r2 = __dadd_rn(xx, yy);
            const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k[7], r2), c.k[6]), r2), c.k[5]), r2));
            const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k[4], r2), c.k[1]), r2), c.k[0]), r2));
            const double icdist = __ddiv_rn(num, den);
            if (icdist < 0) { x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx); y = __dmul_rn(__dsub_rn(v, c.cy), c.ify); break; }
            // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x) + k8*r2 + k9*r2*r2 with k8..k11 = 0 (thin-prism terms are not part of the reference's
            // models): the two trailing additions of +0.0 are kept, they can turn a -0.0 into +0.0
            const double dX = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.k[2]), x), y),
                                                           __dmul_rn(c.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))), __dmul_rn(0.0, r2)),
                                        __dmul_rn(__dmul_rn(0.0, r2), r2));
            const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                                           __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.k[3]), x), y)), __dmul_rn(0.0, r2)),
                                        __dmul_rn(__dmul_rn(0.0, r2), r2));
            x = __dmul_rn(__dsub_rn(x0, dX), icdist);
            y = __dmul_rn(__dsub_rn(y0, dY), icdist);
        }
    }
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(c.P[0], x), __dmul_rn(c.P[1], y)), c.P[2]);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(c.P[3], x), __dmul_rn(c.P[4], y)), c.P[5]);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(c.P[6], x), __dmul_rn(c.P[7], y)), c.P[8]));
    kps[i].x = (float)__dmul_rn(xx, ww);
    kps[i].y = (float)__dmul_rn(yy, ww);
}

} // namespace mage

using namespace mage;

static int make_undistort_const(const mage_camera_calibration* d, const mage_camera_calibration* u, UndistortConst& c)
{
    MAGE_REQUIRE(d && u, MAGE_ERR_INVALID, "mage_undistort_keypoints: NULL calibration");
    MAGE_REQUIRE(d->n_dist_coeffs == 0 || d->n_dist_coeffs == 4 || d->n_dist_coeffs == 5 || d->n_dist_coeffs == 8, MAGE_ERR_INVALID,
                 "n_dist_coeffs must be 0, 4, 5 (Poly3k) or 8 (Rational6k), got %d", d->n_dist_coeffs);
    for (int i = 0; i < 8; i++) c.k[i] = i < d->n_dist_coeffs ? (double)d->dist_coeffs[i] : 0.0;
    c.fx = d->camera_matrix[0]; c.fy = d->camera_matrix[4]; c.cx = d->camera_matrix[2]; c.cy = d->camera_matrix[5];
    MAGE_REQUIRE(c.fx != 0 && c.fy != 0, MAGE_ERR_INVALID, "distorted camera matrix has a zero focal length");
    c.ifx = 1. / c.fx; c.ify = 1. / c.fy;
    for (int i = 0; i < 9; i++) c.P[i] = (double)u->camera_matrix[i];
    c.has_dist = d->n_dist_coeffs > 0;
    return MAGE_OK;
}

extern "C" int mage_undistort_keypoints_device(mage_keypoint* d_keypoints, const int* d_counts, int n_frames, int per_frame,
                                               const mage_camera_calibration* distorted, const mage_camera_calibration* undistorted, void* cuda_stream)
{
    UndistortConst c;
    int rc = make_undistort_const(distorted, undistorted, c);
    if (rc != MAGE_OK) return rc;
    MAGE_REQUIRE(n_frames >= 0 && per_frame >= 0, MAGE_ERR_INVALID, "mage_undistort_keypoints: negative size");
    const long long total = (long long)n_frames * per_frame;
    if (total == 0) return MAGE_OK;
    MAGE_REQUIRE(d_keypoints && total < (1ll << 31), MAGE_ERR_INVALID, "mage_undistort_keypoints: bad buffer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: keypoint undistortion has no CPU fallback"); return MAGE_ERR_CUDA; }
    k_undistort_keypoints<<<div_up((int)total, 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(c, d_keypoints, d_counts, per_frame, (int)total);
    MAGE_CUDA_TRY(cudaGetLastError());
    return MAGE_OK;
}

extern "C" int mage_undistort_keypoints(mage_keypoint* keypoints, int n, const mage_camera_calibration* distorted,
                                        const mage_camera_calibration* undistorted, void* cuda_stream)
{
    UndistortConst c;
    int rc = make_undistort_const(distorted, undistorted, c);
    if (rc != MAGE_OK) return rc;
    MAGE_REQUIRE(n >= 0, MAGE_ERR_INVALID, "mage_undistort_keypoints: n < 0");
    if (n == 0) return MAGE_OK;                          // ref :39-42
    MAGE_REQUIRE(keypoints, MAGE_ERR_INVALID, "mage_undistort_keypoints: NULL buffer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: keypoint undistortion has no CPU fallback"); return MAGE_ERR_CUDA; }
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    // pooled device scratch and pinned staging (one DMA each way instead of two staged copies between the device and the caller's
    // pageable array; no allocation per call): this runs once per analysed frame, right after DetectAndCompute
    const size_t bytes = sizeof(mage_keypoint) * (size_t)n;
    DevStage dsg = dev_stage_acquire(bytes);
    MAGE_REQUIRE(dsg.p, MAGE_ERR_CUDA, "mage_undistort_keypoints: no device memory for %zu bytes", bytes);
    PinnedStage st = stage_acquire(bytes);
    if (!st.p) { dev_stage_release(dsg); MAGE_REQUIRE(false, MAGE_ERR_CUDA, "mage_undistort_keypoints: no pinned staging memory"); }
    mage_keypoint* d = reinterpret_cast<mage_keypoint*>(dsg.p);
    memcpy(st.p, keypoints, bytes);
    cudaError_t e = cudaMemcpyAsync(d, st.p, bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) rc = mage_undistort_keypoints_device(d, nullptr, 1, n, distorted, undistorted, cuda_stream);
    if (e == cudaSuccess && rc == MAGE_OK) e = cudaMemcpyAsync(st.p, d, bytes, cudaMemcpyDeviceToHost, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess && rc == MAGE_OK) memcpy(keypoints, st.p, bytes);
    stage_release(st);
    dev_stage_release(dsg);
    if (rc != MAGE_OK) return rc;
    MAGE_CUDA_TRY(e);
    return MAGE_OK;
}
