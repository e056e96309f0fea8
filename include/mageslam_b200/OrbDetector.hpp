// C++ host-side mirror of the reference's class OrbDetector (Core/MAGESLAM/Source/Image/OpenCVModified.h:64-172) and
// of the free function Match (Tracking/FeatureMatcher.h:68-77) over the C ABI of mage_b200.h. OpenCV-free: images are
// (pointer, width, height, stride) and keypoints are mage_keypoint, which has cv::KeyPoint's 28-byte layout, so an
// adapter inside the reference tree is a reinterpret_cast away.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mage_b200.h"

namespace mage_b200
{
    inline void Check(int rc) { if (rc != MAGE_OK) throw std::runtime_error(std::string("mage_b200: ") + mage_last_error()); }

    class OrbDetector
    {
    public:
        // same 14 scalars, same order as the reference constructor (OpenCVModified.h:68-82)
        OrbDetector(unsigned int gaussianKernelSize, unsigned int nfeatures, float scaleFactor, unsigned int nlevels, unsigned int patchSize,
                    unsigned int fastThreshold, bool useOrientation, float featureFactorANMS, float featureStrengthANMS, int strongResponseANMS,
                    float minRobustFactor, float maxRobustFactor, int numCellsX, int numCellsY)
            : m_params{ gaussianKernelSize, nfeatures, scaleFactor, nlevels, patchSize, fastThreshold, useOrientation ? 1 : 0, featureFactorANMS,
                        featureStrengthANMS, strongResponseANMS, minRobustFactor, maxRobustFactor, numCellsX, numCellsY }
        {
        }
        ~OrbDetector() { mage_orb_destroy(m_handle); }
        OrbDetector(const OrbDetector&) = delete;
        OrbDetector& operator=(const OrbDetector&) = delete;

        // DetectAndCompute(memory, imageData, image): writes up to `capacity` (= ImageData::maxFeatures) keypoints and
        // 32-byte descriptors in place and returns the feature count (ImageData::GetFeatureCount()).
        int DetectAndCompute(const std::uint8_t* image, int width, int height, int stride, mage_keypoint* keypoints, std::uint8_t* descriptors, int capacity)
        {
            if (!m_handle || width != m_width || height != m_height) {
                mage_orb_destroy(m_handle); m_handle = nullptr;
                Check(mage_orb_create(&m_params, width, height, 1, &m_handle));
                m_width = width; m_height = height;
            }
            int count = 0;
            Check(mage_orb_detect_and_compute(m_handle, image, width, height, stride, keypoints, descriptors, capacity, &count, nullptr));
            return count;
        }

    private:
        mage_orb_params m_params;
        mage_orb_t m_handle{ nullptr };
        int m_width{ 0 }, m_height{ 0 };
    };

    // OrbFeatureDetector::UndistortKeypoints(inoutKeypoints, distortedCalibration, undistortedCalibration, memory)
    // (Image/OrbFeatureDetector.cpp:30-62): calibrations as mage_camera_calibration = GetCameraMatrix() + GetCVDistortionCoeffs()
    inline void UndistortKeypoints(mage_keypoint* inoutKeypoints, int count, const mage_camera_calibration& distortedCalibration,
                                   const mage_camera_calibration& undistortedCalibration)
    {
        Check(mage_undistort_keypoints(inoutKeypoints, count, &distortedCalibration, &undistortedCalibration, nullptr));
    }

    // Match(imageA, imageB, imageAMask, imageBMask, ..., maxHammingDist, minHammingDifference, goodMatches) -> count
    inline unsigned int Match(mage_matcher_t matcher, const std::uint8_t* descA, int nA, const std::uint8_t* maskA, const std::uint8_t* descB, int nB,
                              const std::uint8_t* maskB, int maxHammingDist, int minHammingDifference, std::vector<mage_dmatch>& goodMatches)
    {
        goodMatches.resize(nA > 0 ? nA : 0);
        int count = 0;
        Check(mage_match_bf(matcher, descA, nA, maskA, descB, nB, maskB, maxHammingDist, minHammingDifference, goodMatches.data(), &count, nullptr));
        goodMatches.resize(count);
        return static_cast<unsigned int>(count);
    }
}
