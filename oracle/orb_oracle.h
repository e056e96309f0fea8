/*
 * oracle/orb_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference ORB front-end + Match).
 *
 * Nothing under mageslam_b200/ may include, link or call this. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * Parity status: PINNED against the reference's own code. Core/MAGESLAM/Source/Image/OpenCVModified.cpp is compiled
 * UNMODIFIED from /root/reference behind a small cv:: shim (oracle/cvshim/, oracle/Makefile target _ref/liborb_ref.so; both
 * with and without its CV_SSE2 branches) and tests/test_orb_ref.py checks that order mode 0 of this restatement reproduces it
 * bit for bit -- key points (order included), angles and descriptors -- over the tier configurations, the reference
 * defaults, patch 15 / generic patch sizes, degenerate selections and randomised settings. The three OpenCV primitives that
 * are NOT in the reference tree (cv::resize, cv::GaussianBlur, cv::fastAtan2) are restated here and pinned bit for bit
 * against OpenCV 4.13 (tests/test_oracle_vs_cv2.py); the shim forwards to those restatements. The reference ships no tests
 * or golden vectors of its own for this path.
 */
#ifndef MAGE_ORB_ORACLE_H
#define MAGE_ORB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint memory layout (28 bytes) */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

typedef struct { int32_t query, train; float distance; } orc_dmatch;

/* The 14 OrbDetector constructor scalars, reference Image/OpenCVModified.h:68-82, same order. */
typedef struct {
    uint32_t gaussian_kernel_size;
    uint32_t nfeatures;
    float    scale_factor;
    uint32_t nlevels;
    uint32_t patch_size;
    uint32_t fast_threshold;
    int32_t  use_orientation;
    float    feature_factor;
    float    feature_strength;
    int32_t  strong_response;
    float    min_robust_factor;
    float    max_robust_factor;
    int32_t  num_cells_x;
    int32_t  num_cells_y;
} orc_orb_params;

enum { ORC_ORDER_LIBSTDCXX = 0, ORC_ORDER_CANONICAL = 1 };

/* DetectAndCompute. order_mode: 0 = literal std::nth_element (libstdc++) ordering, 1 = canonical ordering
 * (stable raster order in RetainBestFeatures; full sort (r desc, strength desc, raster asc) in ANMS).
 * Returns 0 on success, <0 on unsupported configuration. */
int orc_orb_detect_and_compute(const orc_orb_params* p, const uint8_t* img, int w, int h, int stride,
                               int order_mode, orc_keypoint* kps, uint8_t* desc, int capacity, int* count);

/* Same with the Gaussian-blur arithmetic chosen explicitly:
 *   ORC_BLUR_AUTO          what OpenCV 4.13 does on the reference's call: fixed-point Q8.8 when the level view is the whole
 *                          packed buffer (one level, cols % 16 == 0), the fused float path for a proper submatrix;
 *   ORC_BLUR_FLOAT_FUSED   sepFilter2D float path with every multiply-add as one FMA (an AVX2/FMA build of OpenCV) everywhere;
 *   ORC_BLUR_FLOAT_UNFUSED the same path with separately rounded products (an SSE2-baseline build, e.g. the reference's MSVC x64);
 *   ORC_BLUR_FIXED         the Q8.8 bit-exact path everywhere. */
enum { ORC_BLUR_AUTO = 0, ORC_BLUR_FLOAT_FUSED = 1, ORC_BLUR_FLOAT_UNFUSED = 2, ORC_BLUR_FIXED = 3 };
int orc_orb_detect_and_compute_ex(const orc_orb_params* p, const uint8_t* img, int w, int h, int stride,
                                  int order_mode, int blur_mode, orc_keypoint* kps, uint8_t* desc, int capacity, int* count);

/* --- stage-level entry points, used by the cv2 cross-check tests and by the stage parity tests --- */
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int   orc_gaussian_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize);
/* cv::GaussianBlur of a SUBMATRIX source (generic separable float path); fused = every multiply-add as one FMA */
int   orc_gaussian_blur_submatrix_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize, int fused);
float orc_fast_atan2(float y, float x);
/* MakeRandomPattern (ref OpenCVModified.cpp:551-560): npoints (x, y) pairs from cv::RNG(0x34985739) */
void  orc_random_pattern(int patch_size, int* xy, int npoints);
void  orc_generic_descriptors(const uint8_t* img, int stride, const float* xya, int n, int patch_size, uint8_t* desc);
int   orc_cv_round_f(float v);
/* FAST-9/16 + score + 3x3 NMS in raster order; returns number found (may exceed capacity; only capacity written) */
int   orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold, orc_keypoint* out, int capacity);
/* full score map (uint8, 0 for non-corners), no NMS */
void  orc_fast9_score_map(const uint8_t* img, int w, int h, int stride, int threshold, uint8_t* score, int score_stride);
/* level geometry: sizes[2*l]=w, sizes[2*l+1]=h, scales[l], nfeat[l]; returns 0 */
int   orc_level_layout(const orc_orb_params* p, int w, int h, int* sizes, float* scales, int* nfeat);
/* builds the chained pyramid into separately allocated level images (tightly packed, stride = level width);
 * level_ptrs[l] must have room for w_l*h_l bytes. */
int   orc_build_pyramid(const orc_orb_params* p, const uint8_t* img, int w, int h, int stride, uint8_t** level_ptrs);
/* keypoint selection for ONE level given raster-ordered candidates (x, y, response); returns kept count.
 * Applies RetainBestFeatures + ANMS exactly as ComputeKeyPoints does when n_in > n_keep. */
int   orc_select_level(const orc_orb_params* p, orc_keypoint* kps, int n_in, int n_keep, int order_mode);
/* ANMS suppression radii (r) for every input item, in input order (diagnostics / tie statistics). */
int   orc_anms_radii(const orc_orb_params* p, const orc_keypoint* kps, int n_in, int n_keep, int* r_out);
/* 30 x 1024 int8 pre-rotated pattern regenerated from row 0; patch 31 or 15. returns 0 or -1 */
int   orc_brief_pattern(int patch_size, int8_t* out30x1024);
int   orc_umax(int half_patch, int* umax /* half_patch+2 */);

/* Match (reference Tracking/FeatureMatcher.cpp:61-190). masks may be NULL (= all true). out capacity >= nA. */
int   orc_match(const uint8_t* descA, int nA, const uint8_t* maskA, const uint8_t* descB, int nB, const uint8_t* maskB,
                int max_hamming, int min_diff, orc_dmatch* out, int* count);
int   orc_descriptor_distance(const uint8_t* a, const uint8_t* b);

/* IndexedMatch (reference Tracking/FeatureMatcher.cpp:192-268): BoW-gated two-way match; the vocabulary lookups are given as CSR
 * candidate lists (a2b_off[nA+1]/a2b, b2a_off[nB+1]/b2a) in QueryFeatures order. out capacity >= nA; returns the count. */
int   orc_indexed_match(const uint8_t* descA, int nA, const uint8_t* maskA, const uint8_t* descB, int nB, const uint8_t* maskB,
                        const int* a2b_off, const int* a2b, const int* b2a_off, const int* b2a, int max_hamming, int min_diff, orc_dmatch* out);

/* RadiusMatch (reference Tracking/FeatureMatcher.cpp:294-446) over the enumeration order of the packed boost R*-tree
 * (reference Image/KeypointSpatialIndex.cpp). order_out receives the depth-first value order of the tree built from kps. */
int   orc_rtree_order(const orc_keypoint* kps, int n, int* order_out);
int   orc_radius_match(const orc_keypoint* qk, int nq, const float* qpos_override, const uint8_t* qmask, const uint8_t* qdesc,
                       const orc_keypoint* tk, int nt, const uint8_t* tmask, const uint8_t* tdesc, float radius, int max_hamming,
                       int min_diff, orc_dmatch* out);

#ifdef __cplusplus
}
#endif
#endif
