// Header-compatible drop-in for the reference's Dependencies/BundlerLib/Include/BundlerLib.h:20-66.
// Same namespace, class name, method names and signatures; every method forwards to the C ABI of mage_b200.h
// (libmage_b200.so), so Core/MAGESLAM/Source/BundleAdjustment/BundleAdjust.cpp and Tracking/TrackLocalMap.cpp compile
// and link against it unchanged (replace the BundlerLib include directory and link libmage_b200 instead of
// BundlerLib + g2o). The three tether constraint pools (reference BundlerLib.cpp:243-259, :311-350) forward too.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <gsl/span>
#include <map>

#include <Eigen/Geometry>

#include "../mage_b200.h"

namespace mage
{
    struct BundlerParameters
    {
        bool  ArePointsFixed{ false };        // True if map points should not be optimized
    };

    class BundlerLib
    {
    public:
        BundlerLib(const BundlerParameters& bundlerParameters) : m_bundlerParameters(bundlerParameters)
        {
            Check(mage_ba_create(bundlerParameters.ArePointsFixed ? 1 : 0, &m_handle));
        }
        ~BundlerLib() { mage_ba_destroy(m_handle); }
        BundlerLib(const BundlerLib&) = delete;
        BundlerLib& operator=(const BundlerLib&) = delete;

        void AllocateCameras(size_t count) { Check(mage_ba_alloc_cameras(m_handle, static_cast<int>(count))); }

        void SetCameraPose(size_t idx,
            Eigen::Map<const Eigen::Vector3f> position,
            Eigen::Map<const Eigen::Matrix3f> orientation,
            Eigen::Map<const Eigen::Vector4f> intrinsics, bool isFixed)
        {
            Check(mage_ba_set_camera(m_handle, static_cast<int>(idx), position.data(), orientation.data(), intrinsics.data(), isFixed ? 1 : 0));
        }

        void FixCameraPose(size_t idx, bool value) { Check(mage_ba_fix_camera(m_handle, static_cast<int>(idx), value ? 1 : 0)); }

        void AllocateMapPoints(size_t count) { Check(mage_ba_alloc_points(m_handle, static_cast<int>(count))); }
        void SetMapPoint(size_t idx, Eigen::Map<const Eigen::Vector3f> point) { Check(mage_ba_set_point(m_handle, static_cast<int>(idx), point.data())); }

        void AllocateObservations(size_t count) { m_observations = count; Check(mage_ba_alloc_observations(m_handle, static_cast<int>(count))); }
        void SetObservation(size_t idx, Eigen::Map<const Eigen::Vector2f> position, size_t cameraIndex, size_t mapPointIndex, float informationMatrixScalar)
        {
            Check(mage_ba_set_observation(m_handle, static_cast<int>(idx), position.data(), static_cast<int>(cameraIndex),
                                          static_cast<int>(mapPointIndex), informationMatrixScalar));
        }

        void AllocateFixedDistanceConstraints(size_t count) { Check(mage_ba_alloc_fixed_distance_constraints(m_handle, static_cast<int>(count))); }
        void SetFixedDistanceConstraint(size_t idx, size_t cameraIndex1, size_t cameraIndex2, float distance = 1.0f, float weight = 1.0f)
        {
            Check(mage_ba_set_fixed_distance_constraint(m_handle, static_cast<int>(idx), static_cast<int>(cameraIndex1), static_cast<int>(cameraIndex2), distance, weight));
        }

        void AllocateRelativeRotationConstraints(size_t count) { Check(mage_ba_alloc_relative_rotation_constraints(m_handle, static_cast<int>(count))); }
        void SetRelativeRotationConstraint(size_t idx, size_t cameraIndex1, size_t cameraIndex2, const Eigen::Quaternionf& deltaRotation, float weight = 1.0f)
        {
            const float q[4] = { deltaRotation.x(), deltaRotation.y(), deltaRotation.z(), deltaRotation.w() };
            Check(mage_ba_set_relative_rotation_constraint(m_handle, static_cast<int>(idx), static_cast<int>(cameraIndex1), static_cast<int>(cameraIndex2), q, weight));
        }

        void AllocateRelativeTransformConstraints(size_t count) { Check(mage_ba_alloc_relative_transform_constraints(m_handle, static_cast<int>(count))); }
        void SetRelativeTransformConstraint(size_t idx, size_t cameraIndex1, size_t cameraIndex2, Eigen::Map<const Eigen::Vector3f> deltaPosition, const Eigen::Quaternionf& deltaRotation, float weight)
        {
            const float q[4] = { deltaRotation.x(), deltaRotation.y(), deltaRotation.z(), deltaRotation.w() };
            Check(mage_ba_set_relative_transform_constraint(m_handle, static_cast<int>(idx), static_cast<int>(cameraIndex1), static_cast<int>(cameraIndex2),
                                                            deltaPosition.data(), q, weight));
        }

        void SetCurrentLambda(float userLambda) { Check(mage_ba_set_lambda(m_handle, userLambda)); }
        float GetCurrentLambda() const { float l = 0; Check(mage_ba_get_lambda(m_handle, &l)); return l; }

        // Runs an iteration of the solver for each provided Huber width.
        // Return the average square error.
        float StepBundleAdjustment(gsl::span<const float> huberWidthPerIteration, float maxErrorSquare, std::vector<unsigned int>& outliers)
        {
            std::vector<unsigned int> buf(m_observations ? m_observations : 1);
            int n = 0; float mean = 0;
            Check(mage_ba_step(m_handle, huberWidthPerIteration.data(), static_cast<int>(huberWidthPerIteration.size()), maxErrorSquare,
                               buf.data(), static_cast<int>(buf.size()), &n, &mean));
            outliers.insert(outliers.end(), buf.begin(), buf.begin() + n);
            return mean;
        }

        void GetPose(size_t idx, Eigen::Map<Eigen::Vector3f> position, Eigen::Map<Eigen::Matrix3f> orientation) const
        {
            Check(mage_ba_get_pose(m_handle, static_cast<int>(idx), position.data(), orientation.data()));
        }
        void GetPoint(size_t idx, Eigen::Map<Eigen::Vector3f> position) const { Check(mage_ba_get_point(m_handle, static_cast<int>(idx), position.data())); }

    private:
        static void Check(int rc) { if (rc != MAGE_OK) throw std::runtime_error(std::string("mage_b200: ") + mage_last_error()); }

        mage_ba_t m_handle{ nullptr };
        size_t m_observations{ 0 };

        // Description of the problem to optimize from the calling code.
        BundlerParameters m_bundlerParameters;
    };
}
