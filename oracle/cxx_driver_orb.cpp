/*
 * oracle/cxx_driver_orb.cpp -- TEST INFRASTRUCTURE: one C++ program that extracts ORB features from a generated frame the way
 * OrbFeatureDetector::Process does (ref Core/MAGESLAM/Source/Image/OrbFeatureDetector.cpp:64-93 -> OrbDetector::DetectAndCompute,
 * ref OpenCVModified.cpp:771-886) and prints every key point and descriptor. Built twice by oracle/Makefile:
 *   _ref/orb_driver_ref   -DDRIVER_REFERENCE: the reference's own OrbDetector (OpenCVModified.cpp compiled unmodified, cv:: types
 *                         from oracle/cvshim) filling the reference's own ImageData;
 *   _ref/orb_driver_b200  mage_b200::OrbDetector of include/mageslam_b200/OrbDetector.hpp over libmage_b200.so.
 * Only Extract() differs between the two builds -- it is the adapter INTEGRATION.md section 2 describes.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Params { unsigned gaussian, nfeatures; float scale; unsigned nlevels, patch, thr; bool orient; float factor, strength; int strong; float minrf, maxrf; int cx, cy; };
struct Record { float x, y, size, angle, response; int octave, class_id; uint8_t desc[32]; };
static_assert(sizeof(Record) == 60, "cv::KeyPoint (28 bytes) + ORBDescriptor (32 bytes)");

#ifdef DRIVER_REFERENCE
#include "OpenCVModified.h"
int g_cvshim_blur_mode = 0;
static int Extract(const Params& p, const uint8_t* img, int w, int h, int capacity, std::vector<Record>& out)
{
    using ImageDataT = mage::ImageData<mage::ImageAllocator>;
    OrbDetector det(p.gaussian, p.nfeatures, p.scale, p.nlevels, p.patch, p.thr, p.orient, p.factor, p.strength, p.strong, p.minrf, p.maxrf, p.cx, p.cy);
    std::vector<uint8_t> storage(ImageDataT::AllocationSizeInBytes<ImageDataT>((size_t)capacity) + 64);      // as Image/ImageFactory.h:74-76
    mage::block_splitting_allocation_strategy strategy{ mage::memory::block{ storage.data(), storage.size() } };
    mage::ImageAllocator alloc{ strategy };
    ImageDataT data(mage::CameraIdentity::MONO, (size_t)capacity, p.scale, p.nlevels, 0.f, alloc);
    mage::temp_memory scratch(64u << 20, 256u << 20);
    cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)w);
    det.DetectAndCompute(scratch, data, image);
    const int n = (int)data.GetFeatureCount();
    out.resize(n);
    for (int i = 0; i < n; i++) {
        memcpy(&out[i], &data.GetKeypoints()[i], 28);
        memcpy(out[i].desc, &data.GetDescriptors()[i], 32);
    }
    return n;
}
#else
#include "mageslam_b200/OrbDetector.hpp"
static int Extract(const Params& p, const uint8_t* img, int w, int h, int capacity, std::vector<Record>& out)
{
    mage_b200::OrbDetector det(p.gaussian, p.nfeatures, p.scale, p.nlevels, p.patch, p.thr, p.orient, p.factor, p.strength, p.strong, p.minrf, p.maxrf, p.cx, p.cy);
    std::vector<mage_keypoint> kps(capacity);
    std::vector<uint8_t> desc((size_t)capacity * 32);
    const int n = det.DetectAndCompute(img, w, h, w, kps.data(), desc.data(), capacity);
    out.resize(n);
    for (int i = 0; i < n; i++) {
        memcpy(&out[i], &kps[i], 28);
        memcpy(out[i].desc, &desc[(size_t)i * 32], 32);
    }
    return n;
}
#endif

int main(int argc, char** argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 640, h = argc > 2 ? atoi(argv[2]) : 480;
    const int which = argc > 3 ? atoi(argv[3]) : 0;
    // frame: rectangles and discs on a gradient + box-filtered noise, all integer arithmetic (identical in both builds)
    std::vector<uint8_t> img((size_t)w * h);
    uint64_t s = 0x2545F4914F6CDD1Dull + (uint64_t)which;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    std::vector<int> acc((size_t)w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) acc[(size_t)y * w + x] = 60 + 80 * x / w + 40 * y / h;
    for (int k = 0; k < 90; k++) {
        const int val = 10 + (int)(rnd() % 236), x0 = (int)(rnd() % (unsigned)w), y0 = (int)(rnd() % (unsigned)h), ww = 8 + (int)(rnd() % 90), hh = 8 + (int)(rnd() % 90);
        const bool disc = rnd() % 5 < 2;
        for (int y = y0; y < std::min(h, y0 + hh); y++) for (int x = x0; x < std::min(w, x0 + ww); x++) {
            if (disc) { const int dx = 2 * (x - x0) - ww, dy = 2 * (y - y0) - hh; if (dx * dx * hh * hh + dy * dy * ww * ww > ww * ww * hh * hh) continue; }
            acc[(size_t)y * w + x] = val;
        }
    }
    std::vector<int> noise((size_t)w * h);
    for (auto& v : noise) v = (int)(rnd() % 41) - 20;
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        int t = 0;
        for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) t += noise[(size_t)std::min(std::max(y + dy, 0), h - 1) * w + std::min(std::max(x + dx, 0), w - 1)];
        img[(size_t)y * w + x] = (uint8_t)std::min(std::max(acc[(size_t)y * w + x] + t / 3, 0), 255);
    }
    const Params tier{ 7, 2000, 1.2f, 8, 31, 10, true, 1.5f, 0.9f, 20, 1.1f, 2.0f, 32, 32 };
    const Params dflt{ 7, 440, 1.5f, 1, 15, 4, false, 1.5f, 0.9f, 20, 1.1f, 2.0f, 32, 32 };           // MageSettings.h:151-167
    std::vector<Record> rec;
    const int n = Extract(which % 2 ? dflt : tier, img.data(), w, h, 2000, rec);
    printf("count %d\n", n);
    for (const Record& r : rec) {
        uint32_t f[5]; memcpy(f, &r, 20);
        printf("kp %08x %08x %08x %08x %08x %d %d ", f[0], f[1], f[2], f[3], f[4], r.octave, r.class_id);
        for (int i = 0; i < 32; i++) printf("%02x", r.desc[i]);
        printf("\n");
    }
    return 0;
}
