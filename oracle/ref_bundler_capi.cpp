/*
 * oracle/ref_bundler_capi.cpp -- TEST INFRASTRUCTURE ONLY.
 * A C wrapper around the reference's own, unmodified mage::BundlerLib (compiled from
 * /root/reference/Dependencies/BundlerLib/Source/BundlerLib.cpp + vendored g2o + Eigen by oracle/Makefile into
 * oracle/_ref/libbundler_ref.so). It is the BA oracle ("kind": "reference") and the CPU arm of bench.py.
 * Nothing under mageslam_b200/ links or calls this.
 */
#include <BundlerLib.h>

#include <cstdint>
#include <vector>

extern "C" {

void* refba_create(int are_points_fixed)
{
    mage::BundlerParameters p;
    p.ArePointsFixed = are_points_fixed != 0;
    return new mage::BundlerLib(p);
}
void refba_destroy(void* h) { delete static_cast<mage::BundlerLib*>(h); }

void refba_alloc(void* h, int cams, int points, int obs)
{
    auto* b = static_cast<mage::BundlerLib*>(h);
    b->AllocateCameras((size_t)cams);
    b->AllocateMapPoints((size_t)points);
    b->AllocateObservations((size_t)obs);
}
void refba_set_camera(void* h, int idx, const float* pos, const float* rot_colmajor, const float* intr, int fixed)
{
    static_cast<mage::BundlerLib*>(h)->SetCameraPose((size_t)idx, Eigen::Map<const Eigen::Vector3f>(pos),
                                                     Eigen::Map<const Eigen::Matrix3f>(rot_colmajor), Eigen::Map<const Eigen::Vector4f>(intr), fixed != 0);
}
void refba_set_point(void* h, int idx, const float* xyz)
{
    static_cast<mage::BundlerLib*>(h)->SetMapPoint((size_t)idx, Eigen::Map<const Eigen::Vector3f>(xyz));
}
void refba_set_observation(void* h, int idx, const float* uv, int cam, int pt, float info)
{
    static_cast<mage::BundlerLib*>(h)->SetObservation((size_t)idx, Eigen::Map<const Eigen::Vector2f>(uv), (size_t)cam, (size_t)pt, info);
}
// tether edges (BundlerLib.h:41-48). q = Eigen coefficient order (x, y, z, w)
void refba_alloc_tethers(void* h, int n_distance, int n_rotation, int n_transform)
{
    auto* b = static_cast<mage::BundlerLib*>(h);
    if (n_distance >= 0) b->AllocateFixedDistanceConstraints((size_t)n_distance);
    if (n_rotation >= 0) b->AllocateRelativeRotationConstraints((size_t)n_rotation);
    if (n_transform >= 0) b->AllocateRelativeTransformConstraints((size_t)n_transform);
}
void refba_set_fixed_distance(void* h, int idx, int cam1, int cam2, float distance, float weight)
{
    static_cast<mage::BundlerLib*>(h)->SetFixedDistanceConstraint((size_t)idx, (size_t)cam1, (size_t)cam2, distance, weight);
}
void refba_set_relative_rotation(void* h, int idx, int cam1, int cam2, const float* q_xyzw, float weight)
{
    Eigen::Quaternionf q(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]);
    static_cast<mage::BundlerLib*>(h)->SetRelativeRotationConstraint((size_t)idx, (size_t)cam1, (size_t)cam2, q, weight);
}
void refba_set_relative_transform(void* h, int idx, int cam1, int cam2, const float* dpos, const float* q_xyzw, float weight)
{
    Eigen::Quaternionf q(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]);
    static_cast<mage::BundlerLib*>(h)->SetRelativeTransformConstraint((size_t)idx, (size_t)cam1, (size_t)cam2, Eigen::Map<const Eigen::Vector3f>(dpos), q, weight);
}
void refba_fix_camera(void* h, int idx, int value) { static_cast<mage::BundlerLib*>(h)->FixCameraPose((size_t)idx, value != 0); }
void refba_set_lambda(void* h, float l) { static_cast<mage::BundlerLib*>(h)->SetCurrentLambda(l); }
float refba_get_lambda(void* h) { return static_cast<mage::BundlerLib*>(h)->GetCurrentLambda(); }

float refba_step(void* h, const float* huber, int n, float max_err_sq, unsigned* outliers, int cap, int* n_out)
{
    std::vector<unsigned> out;
    float r = static_cast<mage::BundlerLib*>(h)->StepBundleAdjustment(gsl::span<const float>(huber, n), max_err_sq, out);
    int m = 0;
    for (unsigned o : out) { if (m < cap) outliers[m] = o; m++; }
    *n_out = m;
    return r;
}
void refba_get_pose(void* h, int idx, float* pos, float* rot_colmajor)
{
    static_cast<mage::BundlerLib*>(h)->GetPose((size_t)idx, Eigen::Map<Eigen::Vector3f>(pos), Eigen::Map<Eigen::Matrix3f>(rot_colmajor));
}
void refba_get_point(void* h, int idx, float* xyz)
{
    static_cast<mage::BundlerLib*>(h)->GetPoint((size_t)idx, Eigen::Map<Eigen::Vector3f>(xyz));
}

} // extern "C"
