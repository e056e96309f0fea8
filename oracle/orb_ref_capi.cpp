/*
 * oracle/orb_ref_capi.cpp -- TEST INFRASTRUCTURE ONLY: a C entry point around the REFERENCE's own OrbDetector
 * (Core/MAGESLAM/Source/Image/OpenCVModified.{h,cpp}, compiled unmodified from /root/reference by oracle/Makefile,
 * target _ref/liborb_ref.so) with the reference's own ImageData / thread_memory / allocator headers. OpenCV types come
 * from oracle/cvshim/cvshim.hpp (OpenCV is not vendored); see that header for what is the reference's and what is ours.
 */
#include "OpenCVModified.h"

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

int g_cvshim_blur_mode = 0;

extern "C" {

struct ref_orb_params {                 /* the 14 OrbDetector ctor scalars, ref OpenCVModified.h:68-82, same order (= orc_orb_params) */
    uint32_t gaussian_kernel_size, nfeatures; float scale_factor; uint32_t nlevels, patch_size, fast_threshold; int32_t use_orientation;
    float feature_factor, feature_strength; int32_t strong_response; float min_robust_factor, max_robust_factor; int32_t num_cells_x, num_cells_y;
};

/* returns 0, or -1 when the reference threw (cv::Exception from CV_Assert) */
int ref_orb_detect_and_compute(const ref_orb_params* p, const uint8_t* img, int w, int h, int stride, int blur_mode,
                               void* kps /* n x 28 B cv::KeyPoint */, uint8_t* desc, int capacity, int* count)
{
    using ImageDataT = mage::ImageData<mage::ImageAllocator>;
    g_cvshim_blur_mode = blur_mode;
    *count = 0;
    try {
        OrbDetector det(p->gaussian_kernel_size, p->nfeatures, p->scale_factor, p->nlevels, p->patch_size, p->fast_threshold,
                        p->use_orientation != 0, p->feature_factor, p->feature_strength, p->strong_response,
                        p->min_robust_factor, p->max_robust_factor, p->num_cells_x, p->num_cells_y);
        /* image storage exactly as Image/ImageFactory.h:74-76 builds it: one block split by a non-owning stack allocator */
        size_t maxFeatures = (size_t)capacity;
        size_t bytes = ImageDataT::AllocationSizeInBytes<ImageDataT>(maxFeatures) + 64;
        std::vector<uint8_t> storage(bytes);
        mage::block_splitting_allocation_strategy strategy{ mage::memory::block{ storage.data(), storage.size() } };
        mage::ImageAllocator alloc{ strategy };
        ImageDataT data(mage::CameraIdentity::MONO, maxFeatures, p->scale_factor, p->nlevels, 0.f, alloc);
        /* scratch: the reference sizes these pools per thread; give them room for the largest tier image */
        mage::temp_memory scratch(64u << 20, 256u << 20);
        cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
        det.DetectAndCompute(scratch, data, image);
        int n = (int)data.GetFeatureCount();
        *count = n;
        if (n > 0) {
            memcpy(kps, data.GetKeypoints().data(), (size_t)n * sizeof(cv::KeyPoint));
            memcpy(desc, data.GetDescriptors().data(), (size_t)n * 32);
        }
        return 0;
    } catch (const cv::Exception&) {
        return -1;
    }
}

int ref_orb_sse2(void) { return CV_SSE2; }

}
