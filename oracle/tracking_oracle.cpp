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

} // extern "C"
