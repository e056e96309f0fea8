/*
 * mage_b200.h -- C ABI of the B200-native (sm_100a) hot paths of MAGE-SLAM.
 *
 * One shared library (libmage_b200.so), plain pointers and sizes, status codes instead of exceptions.
 * Each entry point names the reference interface it replaces ("ref" = microsoft/mageslam source tree):
 *
 *   ORB extract   ref Core/MAGESLAM/Source/Image/OpenCVModified.h:68-88   OrbDetector ctor + DetectAndCompute
 *                 ref Core/MAGESLAM/Source/Image/OrbFeatureDetector.h:37  OrbFeatureDetector::Process
 *   Match         ref Core/MAGESLAM/Source/Tracking/FeatureMatcher.h:68-77 Match
 *                 ref Core/MAGESLAM/Source/Tracking/FeatureMatcher.h:134-136 GetDescriptorDistance
 *   Bundle adj.   ref Dependencies/BundlerLib/Include/BundlerLib.h:20-66   class mage::BundlerLib
 *
 * Threading: a handle owns its device scratch and is used by one thread at a time (same contract as the
 * reference objects: OrbDetector is re-entrant only with distinct thread_memory; one BundlerLib per thread).
 * All functions return MAGE_OK (0) or a negative status; mage_last_error() returns a thread-local message.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns MAGE_ERR_CUDA.
 */
#ifndef MAGE_B200_H
#define MAGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGE_OK               0
#define MAGE_ERR_INVALID     -1   /* bad argument (null pointer, size out of range, capacity too small) */
#define MAGE_ERR_UNSUPPORTED -2   /* configuration the reference itself asserts on, or not built */
#define MAGE_ERR_CUDA        -3   /* CUDA runtime error / no device */
#define MAGE_ERR_OVERFLOW    -4   /* an internal candidate list overflowed (cannot happen with default sizing) */

const char* mage_last_error(void);
/* library / device probe: returns number of CUDA devices visible (>= 0) or MAGE_ERR_CUDA */
int mage_device_count(void);
const char* mage_version(void);
/* synchronous device -> host copy for bindings that hold raw device pointers */
int mage_memcpy_d2h(void* dst, const void* src, size_t bytes);

/* Optional per-kernel timing with CUDA events on the launching stream (measurement aid for bench.py; off by default).
 * Slots are kernel names (mage_profile_name); a "group" is one timed launch (or the chained pyramid launches). */
int         mage_profile_enable(int on);
int         mage_profile_collect(void);      /* device sync + fold the recorded event pairs into the totals */
int         mage_profile_reset(void);
int         mage_profile_slots(void);
const char* mage_profile_name(int slot);
int         mage_profile_get(int slot, double* total_ms, long long* groups);

/* ------------------------------------------------------------------------------------------------ ORB extract */

/* cv::KeyPoint memory layout, 28 bytes (ref Image/ImageData.h keypoint buffer) */
typedef struct mage_keypoint {
    float   x, y;        /* level-0 pixel coordinates (pt) */
    float   size;        /* patchSize * layerScale[octave] */
    float   angle;       /* degrees [0,360], 0 when orientation is disabled */
    float   response;    /* FAST score */
    int32_t octave;
    int32_t class_id;    /* always -1 */
} mage_keypoint;

/* The 14 OrbDetector constructor scalars in ctor order (ref Image/OpenCVModified.h:68-82,
 * defaults ref MageSettings.h:151-167). */
typedef struct mage_orb_params {
    uint32_t gaussian_kernel_size;   /* odd, 1 (= no blur) .. 15; sigma is fixed at 2 as in the reference */
    uint32_t nfeatures;
    float    scale_factor;
    uint32_t nlevels;
    uint32_t patch_size;             /* 2..127: 31 / 15 use the pre-rotated BRIEF tables, any other size the cv::RNG pattern rotated at run time (ref :866-885) */
    uint32_t fast_threshold;
    int32_t  use_orientation;
    float    feature_factor;         /* featureFactorANMS */
    float    feature_strength;       /* featureStrengthANMS */
    int32_t  strong_response;        /* strongResponseANMS, must be > fast_threshold */
    float    min_robust_factor;
    float    max_robust_factor;
    int32_t  num_cells_x;
    int32_t  num_cells_y;
} mage_orb_params;

typedef struct mage_orb_s* mage_orb_t;

/* Replaces the OrbDetector constructor. The handle owns device scratch for up to max_batch frames of
 * width x height pixels (pyramids, candidate lists, tables). */
int  mage_orb_create(const mage_orb_params* params, int width, int height, int max_batch, mage_orb_t* out);
void mage_orb_destroy(mage_orb_t h);

/* The arithmetic of the reference's cv::GaussianBlur call (ref OpenCVModified.cpp:863) is a property of the OpenCV build behind it,
 * not of the reference source. Default = MAGE_BLUR_AUTO. Call between extractions, not during one.
 *   MAGE_BLUR_AUTO           OpenCV 4.13: Q8.8 fixed point when the level view is the whole packed buffer (one level, width % 16 == 0),
 *                            else the separable float path with fused multiply-adds (AVX2/FMA build, the stock cv2 wheel);
 *   MAGE_BLUR_FLOAT_FUSED    the float path with FMAs everywhere;
 *   MAGE_BLUR_FLOAT_UNFUSED  the float path with separately rounded products (SSE2-baseline build, e.g. MSVC x64);
 *   MAGE_BLUR_FIXED          the Q8.8 bit-exact path everywhere. */
enum { MAGE_BLUR_AUTO = 0, MAGE_BLUR_FLOAT_FUSED = 1, MAGE_BLUR_FLOAT_UNFUSED = 2, MAGE_BLUR_FIXED = 3 };
int  mage_orb_set_blur_mode(mage_orb_t h, int mode);

/* Replaces OrbDetector::DetectAndCompute for ONE host image (CV_8UC1, row stride in bytes).
 * kps/desc are host buffers with room for `capacity` features (ImageData::maxFeatures); desc is 32 bytes each
 * (ref Image/ORBDescriptor.h). Synchronous on return. stream may be NULL. */
int  mage_orb_detect_and_compute(mage_orb_t h, const uint8_t* image, int width, int height, int stride,
                                 mage_keypoint* kps, uint8_t* desc, int capacity, int* count, void* cuda_stream);

/* Batched host variant: n frames (n <= max_batch), frame i at images + i*frame_stride. Outputs are
 * [n][capacity] arrays, counts[n]. Host buffers should be pinned for full copy bandwidth. Synchronous. */
int  mage_orb_detect_and_compute_batch(mage_orb_t h, const uint8_t* images, int n, int width, int height, int stride,
                                       size_t frame_stride, mage_keypoint* kps, uint8_t* desc, int capacity,
                                       int* counts, void* cuda_stream);

/* Device-resident variant: inputs and outputs are device pointers, nothing leaves HBM, asynchronous on
 * cuda_stream. d_images rows must be 4-byte aligned (stride % 4 == 0, base 16-byte aligned). */
int  mage_orb_extract_device(mage_orb_t h, const uint8_t* d_images, int n, int width, int height, int stride,
                             size_t frame_stride, mage_keypoint* d_kps, uint8_t* d_desc, int capacity,
                             int* d_counts, void* cuda_stream);

/* Level geometry the handle derived (ref OpenCVModified.cpp:795-811, :660-670); arrays of nlevels entries. */
int  mage_orb_level_info(mage_orb_t h, int* widths, int* heights, float* scales, int* nfeatures_per_level);

/* Debug/inspection taps used by the stage parity tests (device -> host copies of internal buffers of frame f). */
int  mage_orb_debug_get_level(mage_orb_t h, int frame, int level, int blurred, uint8_t* out /* w*h, tight */);
/* candidates after FAST+NMS+border cull of one level, sorted in raster order: packed (score<<24 | y*w+x) */
int  mage_orb_debug_get_candidates(mage_orb_t h, int frame, int level, uint32_t* out, int capacity, int* count);

/* ------------------------------------------------------------------------------------------------------ Match */

/* cv::DMatch {queryIdx, trainIdx, distance}; imgIdx is not carried (always 0 in the reference's use). */
typedef struct mage_dmatch { int32_t query_idx, train_idx; float distance; } mage_dmatch;

typedef struct mage_matcher_s* mage_matcher_t;
int  mage_matcher_create(int max_descriptors, int max_pairs, mage_matcher_t* out);
void mage_matcher_destroy(mage_matcher_t m);

/* Replaces Match(): two-way brute-force Hamming match with max distance, min best/second-best difference
 * and cross-check; matches are emitted in ascending A index. masks may be NULL (= all true), else one byte
 * per descriptor. `out` needs room for nA entries. Host buffers, synchronous. */
int  mage_match_bf(mage_matcher_t m, const uint8_t* descA, int nA, const uint8_t* maskA,
                   const uint8_t* descB, int nB, const uint8_t* maskB, int max_hamming, int min_hamming_diff,
                   mage_dmatch* out, int* count, void* cuda_stream);

/* Device-resident batched variant: pair p matches A = d_desc + a_index[p]*slot_stride (d_counts[a_index[p]]
 * descriptors) against B likewise; a_index/b_index are host arrays of n_pairs entries. Outputs
 * d_matches[p][capacity], d_match_counts[p]. d_desc and slot_stride must be 16-byte aligned. Asynchronous. */
int  mage_match_bf_device(mage_matcher_t m, const uint8_t* d_desc, const int* d_counts, size_t slot_stride,
                          const int* a_index, const int* b_index, int n_pairs, int max_hamming, int min_hamming_diff,
                          mage_dmatch* d_matches, int capacity, int* d_match_counts, void* cuda_stream);

/* Same, split in two: register the pair table once, then launch any sub-range of it (no host work per launch). */
int  mage_matcher_set_jobs_device(mage_matcher_t m, const uint8_t* d_desc, const int* d_counts, size_t slot_stride,
                                  const int* a_index, const int* b_index, int n_pairs);
int  mage_match_run_jobs(mage_matcher_t m, int first_pair, int n_pairs, int max_hamming, int min_hamming_diff,
                         mage_dmatch* d_matches, int capacity, int* d_match_counts, void* cuda_stream);

/* IndexedMatch (ref Tracking/FeatureMatcher.h:30-45, .cpp:192-268): the vocabulary-gated two-way match used for new map
 * point creation and loop closure. The BoW lookups stay with the caller (BaseBow::QueryFeatures(desc, keyframeId, closeMatches)
 * / BaseFeatureMatcher::QueryFeatures, ref .cpp:222-229, :251-258): their results arrive as CSR candidate lists, in the order
 * QueryFeatures returned them -- a2b_offsets[nA + 1] / a2b_candidates (indices into B) for every A feature, b2a_offsets[nB + 1] /
 * b2a_candidates (indices into A) for every B feature (only the lists of B features that end up matched are consulted, as
 * in the reference). Masks may be NULL (= all true). out needs room for nA matches (ascending query index = ref order).
 * Host buffers, synchronous. */
int  mage_indexed_match(mage_matcher_t m, const uint8_t* descA, int nA, const uint8_t* maskA, const uint8_t* descB, int nB,
                        const uint8_t* maskB, const int* a2b_offsets, const int* a2b_candidates, const int* b2a_offsets,
                        const int* b2a_candidates, int max_hamming, int min_hamming_diff, mage_dmatch* out, int* count,
                        void* cuda_stream);

/* GetDescriptorDistance for n pairs of device-resident descriptors (a[i] vs b[i]) -> d_out[i]. */
int  mage_descriptor_distance_device(const uint8_t* d_a, const uint8_t* d_b, int n, int* d_out, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------ RadiusMatch */

/* KeypointSpatialIndex (ref Image/KeypointSpatialIndex.h:22-38, .cpp:45-58): built once per analysed image from its
 * keypoints. Instead of an R*-tree the handle stores every keypoint's rank in the enumeration order of the reference's
 * packed boost R*-tree, which is all RadiusMatch's result depends on. */
typedef struct mage_spatial_index_s* mage_spatial_index_t;
int  mage_spatial_index_create(const mage_keypoint* keypoints, int n, mage_spatial_index_t* out);
void mage_spatial_index_destroy(mage_spatial_index_t ix);
int  mage_spatial_index_rank(mage_spatial_index_t ix, int* rank_out /* n */);
/* RadiusMatch, multi-query overload (ref Tracking/FeatureMatcher.h:92-106, .cpp:294-376); nq = 1 gives the single-query
 * overload (ref .h:119-132, .cpp:384-446: count is 0 or 1, out[0].query_idx = 0). query_pos_override (2 floats per query),
 * query_mask and target_mask (one byte per keypoint) may be NULL. Target keypoints are the ones the index was built from;
 * target_desc holds their 32-byte descriptors. out needs room for nq matches. Host buffers, synchronous. */
int  mage_radius_match(mage_spatial_index_t ix, const mage_keypoint* query_kps, int nq, const float* query_pos_override,
                       const uint8_t* query_mask, const uint8_t* query_desc, const uint8_t* target_mask,
                       const uint8_t* target_desc, float radius, int max_hamming, int min_hamming_diff,
                       mage_dmatch* out, int* count, void* cuda_stream);

/* ------------------------------------------------------------------- Map-point projection + candidate culling */

/* One local-map point as TrackLocalMap reads it through its proxy (ref Tracking/TrackLocalMap.cpp:519-554):
 * GetPosition(), GetMeanViewingDirection() (unit length), GetDMin(), GetDMax(). 32 bytes. */
typedef struct {
    float position[3];
    float mean_view_dir[3];
    float dmin, dmax;
} mage_map_point;

typedef struct {
    float view[12];              /* cv::Matx34f, row-major: Pose::GetViewMatrix() (ref TrackLocalMap.cpp:167) */
    float fx, fy, cx, cy;        /* calibration(0,0), (1,1), (0,2), (1,2) (ref Tracking/Reprojection.cpp:35-40) */
    float frame_position[3];     /* Pose::GetWorldSpacePosition() (ref TrackLocalMap.cpp:152) */
    float frame_forward[3];      /* Pose::GetWorldSpaceForward(), unit length (ref TrackLocalMap.cpp:153) */
    float min_cos_view_angle;    /* std::cos(mira::deg2rad(MinDegreesBetweenCurrentViewAndMapPointView)) (ref :541) */
    float image_border;          /* AnalyzedImage::GetImageBorder() */
    uint32_t width, height;      /* AnalyzedImage::GetWidth()/GetHeight() */
    float pyramid_scale;         /* AnalyzedImage::GetPyramidScale() */
    uint32_t num_levels;         /* AnalyzedImage::GetNumLevels() */
} mage_projection_params;

enum { MAGE_PROJ_GOOD_CANDIDATE = 1,   /* IsGoodCandidate() returned true */
       MAGE_PROJ_PREDICTED = 2 };      /* ... and the predicted octave is inside [0, num_levels]: `predicted` of ref :360 */

/* For every map point: ProjectUndistorted (ref Tracking/Reprojection.cpp:26-43) -> IsGoodCandidate (ref
 * Tracking/TrackLocalMap.cpp:519-554) -> ComputeOctave (ref Map/MappingMath.h:13-16), i.e. the part of
 * TrackLocalMap::ProjectMapPointIntoCurrentFrame (ref :325-388) that precedes the RadiusMatch call. out_kps[i] is the
 * cv::KeyPoint(projected.Point, -1, 0, 0, octave, -1) the reference passes to RadiusMatch (octave is 0 unless the point is
 * a good candidate); out_depth[i] (nullable) is Projection::Distance; out_flags[i] is a MAGE_PROJ_* bit set. Also covers
 * ProjectPoints (ref Reprojection.cpp:16-24): read out_kps[i].x/.y and out_depth[i], ignore the flags.
 * Host buffers, synchronous; the _device variant takes device pointers and only enqueues on cuda_stream. */
int  mage_project_map_points(const mage_projection_params* params, const mage_map_point* points, int n, mage_keypoint* out_kps,
                             float* out_depth, uint8_t* out_flags, void* cuda_stream);
int  mage_project_map_points_device(const mage_projection_params* params, const mage_map_point* d_points, int n,
                                    mage_keypoint* d_out_kps, float* d_out_depth, uint8_t* d_out_flags, void* cuda_stream);

/* --------------------------------------------------------------------------------------- Keypoint undistortion */

/* A CameraCalibration as UndistortKeypoints reads it (ref Device/CameraCalibration.h:44-71): GetCameraMatrix() (cv::Matx33f,
 * row-major) and GetCVDistortionCoeffs() in OpenCV order k1 k2 p1 p2 k3 [k4 k5 k6] -- 5 values for Poly3k, 8 for Rational6k,
 * 0 for DistortionType::None. */
typedef struct {
    float camera_matrix[9];
    float dist_coeffs[8];
    int32_t n_dist_coeffs;
} mage_camera_calibration;

/* OrbFeatureDetector::UndistortKeypoints (ref Image/OrbFeatureDetector.cpp:30-62): every keypoint's pt is replaced by
 * cv::undistortPoints(pt, distorted.K, distorted.coeffs, noArray(), undistorted.K) -- five fixed-point iterations in double,
 * then the projective map by the undistorted camera matrix; bit-identical to OpenCV's result. The other KeyPoint fields are
 * untouched. Host buffer, synchronous. The _device variant works in place on the detector's device output: n_frames slots of
 * per_frame keypoints each, of which the first d_counts[f] are valid (d_counts NULL = all); it only enqueues on cuda_stream. */
int  mage_undistort_keypoints(mage_keypoint* keypoints, int n, const mage_camera_calibration* distorted,
                              const mage_camera_calibration* undistorted, void* cuda_stream);
int  mage_undistort_keypoints_device(mage_keypoint* d_keypoints, const int* d_counts, int n_frames, int per_frame,
                                     const mage_camera_calibration* distorted, const mage_camera_calibration* undistorted, void* cuda_stream);

/* ------------------------------------------------------------------------------- Front-end for a video stream */

/* ORB extract of a batch of frames + Match of every frame (query) against its predecessor (train): the reference's
 * per-frame call sequence (ref Tasks/ImageAnalyzer.cpp:119 -> OrbFeatureDetector::Process; Tracking/MapInitialization.cpp:585
 * -> Match) for `batch` frames per call. The last frame of a call is kept on the device as predecessor of the next call.
 * chunk (<= batch, 0 = batch) is the pipelining granularity of the host variant: upload of chunk k+1, compute of chunk k
 * and download of chunk k-1 overlap on three streams. */
typedef struct mage_frontend_s* mage_frontend_t;
int  mage_frontend_create(const mage_orb_params* params, int width, int height, int batch, int chunk, int max_hamming,
                          int min_hamming_diff, mage_frontend_t* out);
void mage_frontend_destroy(mage_frontend_t f);
int  mage_frontend_reset(mage_frontend_t f);                 /* start a new sequence (no predecessor) */
int  mage_frontend_capacity(mage_frontend_t f);              /* features per frame slot in the output arrays */
/* Host buffers (pinned for overlap), synchronous: kps/desc/matches are [n][capacity] arrays, counts/match_counts [n]. */
int  mage_frontend_process(mage_frontend_t f, const uint8_t* images, int n, int stride, size_t frame_stride,
                           mage_keypoint* kps, uint8_t* desc, int* counts, mage_dmatch* matches, int* match_counts);
/* Pipelined form of mage_frontend_process for a continuous stream: _submit enqueues upload -> compute -> download of one call and
 * returns; _wait blocks until the OLDEST submitted call has delivered its results into the host buffers given to its _submit. Two
 * calls may be in flight (frames are staged in two device buffers), so the upload of call j+1 and the download of call j-1 run under
 * the compute of call j. The host buffers of a call must stay valid (and, for overlap, pinned) until its _wait returns.
 * mage_frontend_process == _submit + _wait. */
int  mage_frontend_submit(mage_frontend_t f, const uint8_t* images, int n, int stride, size_t frame_stride,
                          mage_keypoint* kps, uint8_t* desc, int* counts, mage_dmatch* matches, int* match_counts);
int  mage_frontend_wait(mage_frontend_t f);
/* Device-resident frames, results stay in the handle's device buffers (mage_frontend_device_buffers). Asynchronous: the extraction
 * is enqueued on cuda_stream, the matches on an internal stream so that they run under the extraction of the next call;
 * mage_frontend_join makes a stream wait for everything enqueued so far (a device synchronisation does too). */
int  mage_frontend_process_device(mage_frontend_t f, const uint8_t* d_images, int n, int stride, size_t frame_stride,
                                  void* cuda_stream);
int  mage_frontend_join(mage_frontend_t f, void* cuda_stream);
int  mage_frontend_device_buffers(mage_frontend_t f, mage_keypoint** d_kps, uint8_t** d_desc, int** d_counts,
                                  mage_dmatch** d_matches, int** d_match_counts, int* capacity);

/* ------------------------------------------------------------------------------------------ Bundle adjustment */

typedef struct mage_ba_s* mage_ba_t;

/* BundlerLib(const BundlerParameters&) -- ref BundlerLib.cpp:184-196 */
int  mage_ba_create(int are_points_fixed, mage_ba_t* out);
void mage_ba_destroy(mage_ba_t h);
/* AllocateCameras / AllocateMapPoints / AllocateObservations -- ref BundlerLib.cpp:198-230 (allocate once) */
int  mage_ba_alloc_cameras(mage_ba_t h, int count);
int  mage_ba_alloc_points(mage_ba_t h, int count);
int  mage_ba_alloc_observations(mage_ba_t h, int count);
/* SetCameraPose -- ref BundlerLib.cpp:261-276. position[3], orientation 3x3 column-major (world->camera),
 * intrinsics (cx, cy, fx, fy); only fx is used as the focal length, exactly like the reference. */
int  mage_ba_set_camera(mage_ba_t h, int idx, const float position[3], const float orientation_colmajor[9],
                        const float intrinsics_cxcyfxfy[4], int is_fixed);
int  mage_ba_fix_camera(mage_ba_t h, int idx, int value);                                   /* ref :278-281 */
int  mage_ba_set_point(mage_ba_t h, int idx, const float xyz[3]);                           /* ref :283-292 */
int  mage_ba_set_observation(mage_ba_t h, int idx, const float uv[2], int camera_idx, int point_idx,
                             float information_scalar);                                     /* ref :294-309 */
/* bulk SoA uploads (same semantics as calling the setters for idx = 0..n-1 in order) */
int  mage_ba_set_cameras_bulk(mage_ba_t h, int n, const float* positions /*n*3*/, const float* orientations /*n*9*/,
                              const float* intrinsics /*n*4*/, const int32_t* is_fixed /*n*/);
int  mage_ba_set_points_bulk(mage_ba_t h, int n, const float* xyz /*n*3*/);
int  mage_ba_set_observations_bulk(mage_ba_t h, int n, const float* uv /*n*2*/, const int32_t* camera_idx,
                                   const int32_t* point_idx, const float* information /*n*/);

/* Tether edges between two cameras (ref BundlerLib.h:41-48; BundlerLib.cpp:243-259 pools, :311-350 setters; edge types :24-90 and
 * g2o EdgeSE3Expmap). Each pool is allocated once and filled slot by slot; a setter adds the edge (and dirties the problem).
 *   fixed distance     : error = (distance - |t2 - t1|) * weight on the translations of the two view transforms
 *   relative rotation  : error = angularDistance((T1^-1 T2).rotation(), q) * weight; q = (x, y, z, w), used as given
 *   relative transform : error = log(T2^-1 C T1), C = (q normalised, delta_position), information = weight * I6
 * The two scalar edges have no analytic Jacobian in the reference (g2o differentiates BaseMultiEdge numerically, central
 * differences with delta 1e-9); the kernel evaluates the same differences. Tether edges never appear in the outlier list and do
 * not enter the returned mean error (the reference's closing loop reads their camera vertex as a point: undefined behaviour,
 * see DESIGN.md). Supported by the single-CTA kernel: MAGE_ERR_UNSUPPORTED on problems whose reduced system exceeds it. */
int  mage_ba_alloc_fixed_distance_constraints(mage_ba_t h, int count);
int  mage_ba_set_fixed_distance_constraint(mage_ba_t h, int idx, int cam1, int cam2, float distance, float weight);
int  mage_ba_alloc_relative_rotation_constraints(mage_ba_t h, int count);
int  mage_ba_set_relative_rotation_constraint(mage_ba_t h, int idx, int cam1, int cam2, const float* q_xyzw, float weight);
int  mage_ba_alloc_relative_transform_constraints(mage_ba_t h, int count);
int  mage_ba_set_relative_transform_constraint(mage_ba_t h, int idx, int cam1, int cam2, const float* delta_position,
                                               const float* q_xyzw, float weight);

int  mage_ba_set_lambda(mage_ba_t h, float user_lambda);                                    /* ref :352-355 */
int  mage_ba_get_lambda(mage_ba_t h, float* lambda);                                        /* ref :357-360 */
/* StepBundleAdjustment -- ref BundlerLib.cpp:364-447. Runs one LM step per Huber width, then classifies every
 * active observation (behind camera, or squared error > max_error_square) as an outlier, removes it from later
 * steps and reports its index. *mean_sq_error = sum of inlier squared errors / inlier count (NaN if none). */
int  mage_ba_step(mage_ba_t h, const float* huber_width_per_iteration, int n_iterations, float max_error_square,
                  unsigned int* outliers, int outlier_capacity, int* n_outliers, float* mean_sq_error);
int  mage_ba_get_pose(mage_ba_t h, int idx, float position[3], float orientation_colmajor[9]);   /* ref :457-465 */
int  mage_ba_get_point(mage_ba_t h, int idx, float xyz[3]);                                      /* ref :467-471 */
/* bulk read-back + double precision state for parity tests */
int  mage_ba_get_poses_bulk(mage_ba_t h, float* positions /*n*3*/, float* orientations /*n*9*/);
int  mage_ba_get_points_bulk(mage_ba_t h, float* xyz /*n*3*/);
int  mage_ba_get_state_f64(mage_ba_t h, double* cam_qxyzw_t /*K*7*/, double* points /*P*3*/);
/* counters: [0]=LM iterations run, [1]=lambda trials, [2]=kernel launches, [3]=structure rebuilds */
int  mage_ba_get_stats(mage_ba_t h, int64_t stats[4]);
/* diagnostics: accumulated nanoseconds per phase of the cooperative kernel as seen by block 0 */
int  mage_ba_debug_phase_ns(mage_ba_t h, long long phase_ns[16]);
/* TrackLocalMap::OptimizeCameraPose in one call (ref Tracking/TrackLocalMap.cpp:421-501; it runs twice per frame on the tracking thread):
 * equivalent to a new BundlerLib with ArePointsFixed, AllocateCameras(1) + SetCameraPose(0, position, rotation, intrinsics, false),
 * n map points with one observation each (SetMapPoint(i), SetObservation(i, projection_i, 0, i, information_i)), ONE
 * StepBundleAdjustment(n_iters x huber_width, max_outlier_err_sq, outliers) and GetPose(0) -- same kernel and results as that sequence
 * through mage_ba_*, without the instance, its structure build and its copies (one upload, one launch, one read-back).
 * position / rotation (column-major 3x3) = the view transform, intrinsics = (cx, cy, fx, fy); outliers receives the indices of the
 * observations the reference would remove (at most `capacity` are written, *n_outliers is the full count); mean_sq_error is nullable.
 * n_iters <= 64. Thread-safe (contexts are pooled per process). */
int  mage_optimize_camera_pose(const float position[3], const float rotation_colmajor[9], const float intrinsics[4], int n,
                               const float* map_points /*n*3*/, const float* projections /*n*2*/, const float* information /*n*/,
                               int n_iters, float huber_width, float max_outlier_err_sq, float out_position[3],
                               float out_rotation_colmajor[9], unsigned int* outliers, int capacity, int* n_outliers,
                               float* mean_sq_error);
/* Sharded global bundle adjustment (SURVEY 8(e) "next": observations partitioned by landmark over the ranks, the all-reduce of the reduced
 * camera system is the one exchange step; ref block_solver.hpp:331-422 builds that system, linear_solver_dense.h:65-113 solves it). Every
 * rank creates the problem with ALL cameras and its own points / observations; mage_ba_shard_prepare builds the structure and returns the
 * device buffers to all-reduce (S: n x n, bs: n, xchg: 16 + 2 n doubles), mage_ba_shard_stage runs one stage of the LM iteration
 * (1 linearise, 2 Schur complement of this rank's landmarks, 3 damp + solve + update + trial chi2, 4 restore, 5 error sums); the host
 * loop around them is mageslam_b200/sharded.py. lead = 1 on exactly one rank (it adds the camera part of the gain-ratio denominator). */
int  mage_ba_shard_prepare(mage_ba_t h, int* n, double** d_S, double** d_bs, double** d_xchg);
int  mage_ba_shard_stage(mage_ba_t h, int stage, double huber_delta, double lambda, int lead);
/* diagnostics: a work array of the built problem (0 x = the last increment, 1 Hpp, 2 bp, 3 bl, 4 Hll, 5 bs, 6 S of a large system) */
int  mage_ba_debug_get_array(mage_ba_t h, int which, double* out, long long capacity, long long* count);
/* The dense solver of the reduced camera system on its own (test entry; replaces Eigen::LDLT of ref
 * Dependencies/g2o/g2o/solvers/dense/linear_solver_dense.h:65-113): A is n x n row-major symmetric positive definite (lower triangle
 * read), b the right-hand side; x receives the solution, factor (nullable, n x n) L below the diagonal and D on it, *positive 0 when a
 * pivot was not positive, phase_ns (nullable) the time of the diagonal blocks / panel rows / tensor-core update / grid barriers /
 * backward substitution as CTA 0 sees them, then finer splits (dense_ldlt.cuh). */
int  mage_dense_debug_solve(int n, const double* A, const double* b, double* x, double* factor, int* positive, long long phase_ns[16]);
/* Many independent problems stepped concurrently (one CTA group per problem): same result per handle as calling
 * mage_ba_step on each. means/outlier outputs are per handle. */
int  mage_ba_step_many(mage_ba_t* handles, int n_handles, const float* huber_width_per_iteration, int n_iterations,
                       float max_error_square, float* mean_sq_errors);
/* The observation indices the last mage_ba_step / mage_ba_step_many call removed from this problem (the `outliers` vector of
 * StepBundleAdjustment, ref BundlerLib.cpp:385-441); returns the total count in *n_outliers even when capacity is smaller. */
int  mage_ba_last_outliers(mage_ba_t h, unsigned int* outliers, int capacity, int* n_outliers);
int  mage_ba_last_outlier_counts(mage_ba_t* handles, int n_handles, int* counts);     /* the counts of many problems in one call */

#ifdef __cplusplus
}
#endif
#endif /* MAGE_B200_H */
