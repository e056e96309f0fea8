/*
 * oracle/tracking_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement; never linked into the product library).
 *
 * Map-point projection + candidate culling + octave prediction, restated from the reference:
 *   ProjectUndistorted   Core/MAGESLAM/Source/Tracking/Reprojection.cpp:26-43
 *   IsGoodCandidate      Core/MAGESLAM/Source/Tracking/TrackLocalMap.cpp:519-554
 *   PointWithinImageBorder  Core/MAGESLAM/Source/Image/AnalyzedImage.h:118-122
 *   ComputeOctave        Core/MAGESLAM/Source/Map/MappingMath.h:13-16
 *   the prologue of ProjectMapPointIntoCurrentFrame   TrackLocalMap.cpp:325-366
 * cv::Matx / cv::Point3f arithmetic (un-vendored OpenCV) is restated from its published definition: Matx product
 * s = 0; s += a(i,k) * b(k,j) in k order; Vec::dot likewise; Point3_::dot = x*x' + y*y' + z*z' left to right. Built with
 * -ffp-contract=off (scalar IEEE f32, as an x64 MSVC /fp:precise build evaluates it).
 *
 * Parity status: Reprojection.cpp cannot be compiled here (OpenCV C++ headers absent) => the projection is "parity
 * unpinned by the reference", cross-checked against a float32 numpy restatement (tests/test_tracking_oracle.py);
 * ComputeOctave is pinned against the reference's own header compiled in oracle/_ref/libtracking_ref.so.
 */
#include <math.h>
#include <stdint.h>

extern "C" {

typedef struct { float position[3]; float mean_view_dir[3]; float dmin, dmax; } trk_map_point;
typedef struct {
    float view[12];
    float fx, fy, cx, cy;
    float frame_position[3];
    float frame_forward[3];
    float min_cos_view_angle;
    float image_border;
    uint32_t width, height;
    float pyramid_scale;
    uint32_t num_levels;
} trk_projection_params;
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } trk_keypoint;

int trk_compute_octave(float distance, float dmin, float scale_factor)
{
    return static_cast<int>(roundf(log2f(distance / dmin) / log2f(scale_factor) - 0.5f));
}

static float matx_row(const float* r, float x, float y, float z)
{
    float s = 0;
    s += r[0] * x; s += r[1] * y; s += r[2] * z; s += r[3] * 1.f;
    return s;
}

void trk_project_map_points(const trk_projection_params* p, const trk_map_point* pts, int n, trk_keypoint* out_kps, float* out_depth,
                            uint8_t* out_flags)
{
    for (int i = 0; i < n; i++) {
        const trk_map_point& m = pts[i];
        const float X = m.position[0], Y = m.position[1], Z = m.position[2];
        // ProjectUndistorted
        const float c0 = matx_row(p->view, X, Y, Z), c1 = matx_row(p->view + 4, X, Y, Z), depth = matx_row(p->view + 8, X, Y, Z);
        const float div = depth != 0 ? depth : 1;
        const float u = (c0 / div) * p->fx + p->cx;
        const float v = (c1 / div) * p->fy + p->cy;
        // IsGoodCandidate
        bool good = true;
        const float border = p->image_border;
        if (depth < 0 || !(border <= u && border <= v && u < p->width - border && v < p->height - border)) good = false;
        if (good) {
            float dotv = 0;
            for (int k = 0; k < 3; k++) dotv += m.mean_view_dir[k] * p->frame_forward[k];
            if (dotv < p->min_cos_view_angle) good = false;
        }
        const float dx = X - p->frame_position[0], dy = Y - p->frame_position[1], dz = Z - p->frame_position[2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (good && (d2 < (m.dmin * m.dmin) || (m.dmax * m.dmax) < d2)) good = false;
        int octave = 0;
        bool predicted = false;
        if (good) {
            octave = trk_compute_octave(sqrtf(d2), m.dmin, p->pyramid_scale);
            predicted = !(octave < 0 || octave > static_cast<int>(p->num_levels));
        }
        trk_keypoint kp = {u, v, -1.0f, 0.0f, 0.0f, octave, -1};
        out_kps[i] = kp;
        if (out_depth) out_depth[i] = depth;
        out_flags[i] = (uint8_t)((good ? 1 : 0) | (predicted ? 2 : 0));
    }
}

/* pre-rounding value of ComputeOctave, for the tests' "near a rounding boundary" classification */
float trk_octave_real(float distance, float dmin, float scale_factor)
{
    return log2f(distance / dmin) / log2f(scale_factor) - 0.5f;
}



/*
 * Keypoint undistortion: mage::OrbFeatureDetector::UndistortKeypoints (Core/MAGESLAM/Source/Image/OrbFeatureDetector.cpp:30-62)
 * = cv::undistortPoints(src, dst, distorted.GetCameraMatrix(), distorted.GetCVDistortionCoeffs(), noArray(), undistorted.GetCameraMatrix())
 * on the keypoints' pt. OpenCV is not vendored: restated from its published algorithm (calib3d undistort, cvUndistortPointsInternal):
 * everything in double, TermCriteria(MAX_ITER, 5, 0.01) => exactly five fixed-point iterations, the icdist < 0 escape, then the
 * projective map with P = the new camera matrix. Pinned bit-for-bit against cv2 4.13 (tests/test_tracking_oracle.py).
 * Coefficient order k1 k2 p1 p2 k3 [k4 k5 k6] (CameraCalibration.h:53); 5 (Poly3k) or 8 (Rational6k) are used by the reference.
 */
typedef struct { float camera_matrix[9]; float dist_coeffs[8]; int32_t n_dist_coeffs; } trk_calibration;

void trk_undistort_keypoints(trk_keypoint* kps, int n, const trk_calibration* dist, const trk_calibration* undist)
{
    double k[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < dist->n_dist_coeffs && i < 8; i++) k[i] = (double)dist->dist_coeffs[i];
    const double fx = dist->camera_matrix[0], fy = dist->camera_matrix[4], cx = dist->camera_matrix[2], cy = dist->camera_matrix[5];
    const double ifx = 1. / fx, ify = 1. / fy;
    double RR[9];
    for (int i = 0; i < 9; i++) RR[i] = (double)undist->camera_matrix[i];
    for (int i = 0; i < n; i++) {
        double x = kps[i].x, y = kps[i].y;
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        if (dist->n_dist_coeffs > 0) {
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        const double xx = RR[0] * x + RR[1] * y + RR[2];
        const double yy = RR[3] * x + RR[4] * y + RR[5];
        const double ww = 1. / (RR[6] * x + RR[7] * y + RR[8]);
        kps[i].x = (float)(xx * ww);
        kps[i].y = (float)(yy * ww);
    }
}

} // extern "C"
