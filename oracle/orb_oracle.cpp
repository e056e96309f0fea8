/*
 * oracle/orb_oracle.cpp -- TEST INFRASTRUCTURE ONLY. See orb_oracle.h for the usage rules and parity status.
 *
 * CPU restatement of the reference's ORB front-end and brute-force matcher. Every function cites the reference
 * lines it follows ("ref" = /root/reference/Core/MAGESLAM/Source). Arithmetic types (float vs double, truncation
 * vs round-half-even, strict vs non-strict compares) follow the reference; the OpenCV primitives the reference
 * calls (resize / GaussianBlur / fastAtan2 / cvRound / BFMatcher::radiusMatch) are restated from the arithmetic
 * recovered in SURVEY.md appendix A and cross-checked against cv2 4.13 by tests/test_oracle_vs_cv2.py.
 *
 * Build: g++ -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: fastAtan2 / ANMS float math must not fuse).
 */
#include "orb_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

#include "../mageslam_b200/csrc/brief_base_patterns.inc"   // data only: row 0 of the two pattern tables

// ---- OpenCV scalar helpers -------------------------------------------------------------------------------------
// cvRound: round-half-to-even (SSE cvtss2si / cvtsd2si under the default rounding mode).
inline int cvRoundF(float v) { return (int)lrintf(v); }
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeilF(float v) { int i = (int)v; return i + (i < v); }

// cv::fastAtan2 (degrees), float32, no FMA.
float fastAtan2Impl(float y, float x)
{
    const float scale = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// cv::resize INTER_LINEAR on CV_8UC1 (11-bit fixed-point coefficients).
void resizeLinear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> xa(2 * dw), ya(2 * dh);
    auto axis = [](int ssize, int dsize, std::vector<int>& ofs, std::vector<short>& ab) {
        double scale = (double)ssize / dsize;
        for (int d = 0; d < dsize; d++) {
            float f = (float)((d + 0.5) * scale - 0.5);
            int s = (int)std::floor(f);
            f -= s;
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
            ofs[d] = s;
            ab[2 * d] = (short)cvRoundF((1.f - f) * 2048.f);
            ab[2 * d + 1] = (short)cvRoundF(f * 2048.f);
        }
    };
    axis(sw, dw, xofs, xa);
    axis(sh, dh, yofs, ya);
    std::vector<int> row0(dw), row1(dw);
    int cached0 = -1, cached1 = -1;
    auto hrow = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            out[dx] = S[sx] * xa[2 * dx] + S[sx1] * xa[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; dy++) {
        int sy = yofs[dy], sy1 = std::min(sy + 1, sh - 1);
        if (cached0 != sy) { if (cached1 == sy) { row0.swap(row1); std::swap(cached0, cached1); } else { hrow(sy, row0); cached0 = sy; } }
        if (cached1 != sy1) { if (sy1 == sy) { row1 = row0; cached1 = sy1; } else { hrow(sy1, row1); cached1 = sy1; } }
        int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

// cv::GaussianBlur(ksize x ksize, sigma 2, BORDER_REFLECT_101) on CV_8UC1, bit-exact
// fixed-point path of OpenCV 4.x. Q8.8 kernels (sum 256) for sigma = 2 read off cv2 4.13 impulse responses.
const int* gaussKernelQ8(int ksize)
{
    static const int k3[] = {82, 92, 82};
    static const int k5[] = {39, 57, 64, 57, 39};
    static const int k7[] = {18, 34, 48, 56, 48, 34, 18};
    static const int k9[] = {7, 17, 32, 46, 52, 46, 32, 17, 7};
    static const int k11[] = {2, 7, 17, 31, 45, 52, 45, 31, 17, 7, 2};
    static const int k13[] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
    static const int k15[] = {0, 1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1, 0};
    switch (ksize) {
    case 3: return k3; case 5: return k5; case 7: return k7; case 9: return k9;
    case 11: return k11; case 13: return k13; case 15: return k15;
    default: return nullptr;
    }
}
inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}
int gaussianBlur(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize)
{
    const int* k = gaussKernelQ8(ksize);
    if (!k) return -1;
    int r = ksize / 2;
    std::vector<uint16_t> T((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            unsigned acc = 0;
            for (int i = 0; i < ksize; i++) acc += (unsigned)k[i] * S[reflect101(x + i - r, w)];
            T[(size_t)y * w + x] = (uint16_t)acc;       // <= 255*256 = 65280
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            unsigned acc = 0;
            for (int j = 0; j < ksize; j++) acc += (unsigned)k[j] * T[(size_t)reflect101(y + j - r, h) * w + x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    return 0;
}

// cv::GaussianBlur on a SUBMATRIX source -- what the reference's in-place blur of a level ROI of the packed pyramid buffer is
// (ref :853-865; every level is a proper submatrix unless nlevels == 1 and cols % 16 == 0). OpenCV only takes the bit-exact
// fixed-point path above for non-submatrix 8-bit sources; a submatrix goes createGaussianKernels(CV_32F) -> sepFilter2D:
//   row pass    RowFilter<uchar, float, RowNoVec>:            s = kx[0]*x[0];  s += kx[k]*x[k], k = 1 .. ksize-1   (left to right)
//   column pass SymmColumnFilter<Cast<float, uchar>, NoVec>:  s = ky[c]*r[c] + 0;  s += ky[c+k]*(r[c+k] + r[c-k]), k = 1 .. ksize/2
//   saturate_cast<uchar>(float) = clamp(cvRound(s)).
// Float kernels = getGaussianKernel(ksize, 2, CV_32F) of cv2 4.13 (first half; symmetric). `fused` evaluates every
// multiply-add as one FMA -- what a -mfma build of OpenCV (the AVX2 dispatch unit of the cv2 wheel) makes of the same
// source; the reference semantics (MSVC /fp:precise, SSE2 baseline) is the unfused one. The two differ in ~1e-5 of the pixels.
const float* gaussKernelF32(int ksize)
{
    static const float k3[] = {0x1.46d3eap-2f, 0x1.72582cp-2f};
    static const float k5[] = {0x1.3841bep-3f, 0x1.c654bap-3f, 0x1.016988p-2f};
    static const float k7[] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95c0p-3f};
    static const float k9[] = {0x1.c4b2eep-6f, 0x1.0f7df8p-4f, 0x1.fb36c8p-4f, 0x1.70fefap-3f, 0x1.a22092p-3f};
    static const float k11[] = {0x1.20c256p-7f, 0x1.bcb86ap-6f, 0x1.0ab50ap-4f, 0x1.f2464cp-4f, 0x1.6a7e1ep-3f, 0x1.9ac20ap-3f};
    static const float k13[] = {0x1.22be4ep-9f, 0x1.1f7a64p-7f, 0x1.babf56p-6f, 0x1.098622p-4f, 0x1.f01066p-4f, 0x1.68e26cp-3f, 0x1.98ef8ap-3f};
    static const float k15[] = {0x1.c99b3ap-12f, 0x1.227d56p-9f, 0x1.1f3a28p-7f, 0x1.ba5c6ap-6f, 0x1.094acep-4f, 0x1.efa190p-4f, 0x1.6891cap-3f, 0x1.98942ap-3f};
    switch (ksize) {
    case 3: return k3; case 5: return k5; case 7: return k7; case 9: return k9;
    case 11: return k11; case 13: return k13; case 15: return k15;
    default: return nullptr;
    }
}
// The loops run tap-outer / pixel-inner over a REFLECT_101-padded row so that they vectorise; per pixel the sequence of rounded
// operations is exactly the one written above (the CPU-baseline timings of bench.py run through this function, so it should not be
// slower than it has to be). FUSED needs hardware FMA to be fast: the same template is compiled a second time for AVX2+FMA and
// picked at run time; without FMA hardware fmaf() falls back to libm (correct, slow).
template <bool FUSED>
static inline __attribute__((always_inline)) void blurSubmatrixImpl(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride,
                                                                     int ksize, const float* kh)
{
    const int r = ksize / 2;
    float K[16];
    for (int i = 0; i < ksize; i++) K[i] = kh[i <= r ? i : 2 * r - i];      // full symmetric kernel, index 0 .. ksize-1
    std::vector<float> T((size_t)w * h), rb((size_t)w + 2 * r);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = -r; x < w + r; x++) rb[(size_t)(x + r)] = (float)S[reflect101(x, w)];
        float* t = T.data() + (size_t)y * w;
        const float* b = rb.data();
        for (int x = 0; x < w; x++) t[x] = K[0] * b[x];
        for (int k = 1; k < ksize; k++) {
            const float kk = K[k];
            const float* bk = b + k;
            if (FUSED) for (int x = 0; x < w; x++) t[x] = __builtin_fmaf(kk, bk[x], t[x]);
            else       for (int x = 0; x < w; x++) t[x] = t[x] + kk * bk[x];
        }
    }
    std::vector<float> acc((size_t)w);
    for (int y = 0; y < h; y++) {
        const float* c = T.data() + (size_t)y * w;
        float* a = acc.data();
        if (FUSED) for (int x = 0; x < w; x++) a[x] = __builtin_fmaf(kh[r], c[x], 0.f);
        else       for (int x = 0; x < w; x++) a[x] = kh[r] * c[x] + 0.f;
        for (int k = 1; k <= r; k++) {
            const float* p = T.data() + (size_t)reflect101(y + k, h) * w;
            const float* m = T.data() + (size_t)reflect101(y - k, h) * w;
            const float kk = kh[r - k];
            if (FUSED) for (int x = 0; x < w; x++) a[x] = __builtin_fmaf(kk, p[x] + m[x], a[x]);
            else       for (int x = 0; x < w; x++) a[x] = a[x] + kk * (p[x] + m[x]);
        }
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            int iv = cvRoundF(a[x]);
            D[x] = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        }
    }
}
#if defined(__x86_64__)
__attribute__((target("avx2,fma"))) static void blurSubmatrixFusedFma(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride,
                                                                        int ksize, const float* kh)
{
    blurSubmatrixImpl<true>(src, w, h, sstride, dst, dstride, ksize, kh);
}
#endif
int gaussianBlurSubmatrix(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize, bool fused)
{
    const float* kh = gaussKernelF32(ksize);
    if (!kh) return -1;
    if (!fused) { blurSubmatrixImpl<false>(src, w, h, sstride, dst, dstride, ksize, kh); return 0; }
#if defined(__x86_64__)
    if (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2")) { blurSubmatrixFusedFma(src, w, h, sstride, dst, dstride, ksize, kh); return 0; }
#endif
    blurSubmatrixImpl<true>(src, w, h, sstride, dst, dstride, ksize, kh);
    return 0;
}

// ---- FAST-9/16 ----------------------------------------------------------------------------------------------------
// ref Image/OpenCVModified.cpp:890-921 (makeOffsets), ring order k = 0..15
const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                          {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// ref :926-1071 cornerScore<16>: max threshold for which the pixel is still a 9-contiguous-arc corner.
inline int cornerScore16(const uint8_t* ptr, const int pixel[25])
{
    int v = ptr[0];
    short d[25];
    for (int k = 0; k < 25; k++) d[k] = (short)(v - ptr[pixel[k]]);
    int q0 = -1000, q1 = 1000;
    for (int k = 0; k < 16; k++) {
        int a = d[k], b = d[k];
        for (int j = 1; j < 9; j++) { a = std::min(a, (int)d[k + j]); b = std::max(b, (int)d[k + j]); }
        q0 = std::max(q0, a);
        q1 = std::min(q1, b);
    }
    return std::max(q0, -q1) - 1;
}

// ref :1224-1512 FAST_t<16> with nonmax_suppression = true. Emits (x, y, score) in raster order.
// Detection follows the reference's scalar path (:1415-1479): threshold_tab quick test, then contiguous-arc count.
template <class Emit>
void fast9(const uint8_t* img, int w, int h, int stride, int threshold, bool nms, uint8_t* scoreMap, int scoreStride, Emit emit)
{
    threshold = std::min(std::max(threshold, 0), 255);
    int pixel[25];
    for (int k = 0; k < 16; k++) pixel[k] = kRing[k][0] + kRing[k][1] * stride;
    for (int k = 16; k < 25; k++) pixel[k] = pixel[k - 16];
    uint8_t tab[512];                                                       // ref :694-698
    for (int t = -255; t <= 255; t++) tab[t + 255] = (uint8_t)(t < -threshold ? 1 : t > threshold ? 2 : 0);

    std::vector<uint8_t> buf((size_t)w * 3, 0);
    std::vector<short> cp((size_t)(w + 1) * 3, 0);
    uint8_t* rows[3] = {buf.data(), buf.data() + w, buf.data() + 2 * w};
    short* cprow[3] = {cp.data() + 1, cp.data() + 1 + (w + 1), cp.data() + 1 + 2 * (w + 1)};

    for (int i = 3; i < h - 2; i++) {
        const uint8_t* ptr = img + (size_t)i * stride + 3;
        uint8_t* curr = rows[(i - 3) % 3];
        short* cornerpos = cprow[(i - 3) % 3];
        std::memset(curr, 0, w);
        int ncorners = 0;
        if (i < h - 3) {
            for (int j = 3; j < w - 3; j++, ptr++) {
                int v = ptr[0];
                const uint8_t* t = &tab[0] - v + 255;
                int d = t[ptr[pixel[0]]] | t[ptr[pixel[8]]];
                if (d == 0) continue;
                d &= t[ptr[pixel[2]]] | t[ptr[pixel[10]]];
                d &= t[ptr[pixel[4]]] | t[ptr[pixel[12]]];
                d &= t[ptr[pixel[6]]] | t[ptr[pixel[14]]];
                if (d == 0) continue;
                d &= t[ptr[pixel[1]]] | t[ptr[pixel[9]]];
                d &= t[ptr[pixel[3]]] | t[ptr[pixel[11]]];
                d &= t[ptr[pixel[5]]] | t[ptr[pixel[13]]];
                d &= t[ptr[pixel[7]]] | t[ptr[pixel[15]]];
                bool corner = false;
                if (d & 1) {
                    int vt = v - threshold, count = 0;
                    for (int k = 0; k < 25; k++) {
                        if (ptr[pixel[k]] < vt) { if (++count > 8) { corner = true; break; } } else count = 0;
                    }
                }
                if (!corner && (d & 2)) {
                    int vt = v + threshold, count = 0;
                    for (int k = 0; k < 25; k++) {
                        if (ptr[pixel[k]] > vt) { if (++count > 8) { corner = true; break; } } else count = 0;
                    }
                }
                if (corner) {
                    cornerpos[ncorners++] = (short)j;
                    curr[j] = (uint8_t)cornerScore16(ptr, pixel);
                }
            }
        }
        if (scoreMap && i < h - 3) std::memcpy(scoreMap + (size_t)i * scoreStride, curr, w);
        cornerpos[-1] = (short)ncorners;
        if (i == 3) continue;
        const uint8_t* prev = rows[(i - 4 + 3) % 3];
        const uint8_t* pprev = rows[(i - 5 + 3) % 3];
        const short* cpos = cprow[(i - 4 + 3) % 3];
        int nc = cpos[-1];
        for (int k = 0; k < nc; k++) {
            int j = cpos[k];
            int score = prev[j];
            if (!nms || (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] && score > pprev[j] &&
                         score > pprev[j + 1] && score > curr[j - 1] && score > curr[j] && score > curr[j + 1]))
                emit(j, i - 1, score);
        }
    }
}

// ---- level geometry ------------------------------------------------------------------------------------------------
struct Levels {
    int n = 0;
    std::vector<int> w, h, nfeat;
    std::vector<float> scale;
};
// ref :564-567 getScale, :795-811 level sizes, :660-670 per-level feature budget
Levels levelLayout(const orc_orb_params& p, int cols, int rows)
{
    Levels L;
    L.n = (int)p.nlevels;
    L.w.resize(L.n); L.h.resize(L.n); L.nfeat.resize(L.n); L.scale.resize(L.n);
    for (int l = 0; l < L.n; l++) {
        float scale = (float)std::pow((double)p.scale_factor, (double)l);
        L.scale[l] = scale;
        L.w[l] = cvRoundF(cols / scale);
        L.h[l] = cvRoundF(rows / scale);
    }
    int nfeatures = (int)p.nfeatures;
    float factor = 1.0f / p.scale_factor;
    float ndesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)L.n));
    int sum = 0;
    for (int l = 0; l < L.n - 1; l++) {
        L.nfeat[l] = cvRoundF(ndesired);
        sum += L.nfeat[l];
        ndesired *= factor;
    }
    L.nfeat[L.n - 1] = std::max(nfeatures - sum, 0);
    return L;
}

// ref :673-689
void computeUmax(int halfPatch, std::vector<int>& umax)
{
    umax.assign(halfPatch + 2, 0);
    int v, v0, vmax = cvFloorF(halfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = cvCeilF(halfPatch * std::sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(std::sqrt((double)halfPatch * halfPatch - v * v));
    for (v = halfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// ---- selection: RetainBestFeatures + ANMS -------------------------------------------------------------------------
// ref :571-617. Returns the number kept; reorders kps so that the kept ones come first.
int retainBest(std::vector<orc_keypoint>& kps, int minThreshold, int maxNum, int minNum, float responseFactor, int mode)
{
    unsigned hist[256] = {0};
    size_t minNumThreshold = (size_t)minThreshold;
    for (auto& k : kps) {
        float r = std::min(std::max(k.response, 0.0f), 255.0f);     // mira::clamp
        hist[(uint8_t)r]++;
    }
    int numFeatures = 0;
    for (int i = 255; i >= minThreshold; i--) {
        numFeatures += hist[i];
        if (numFeatures >= minNum) { minNumThreshold = i; break; }
    }
    numFeatures = 0;
    int lo = std::max((int)(minNumThreshold * responseFactor), minThreshold);
    int stop = lo;
    for (int i = 255; i >= lo; i--) {
        numFeatures += hist[i];
        if (numFeatures >= maxNum) { stop = i; break; }
    }
    if (mode == ORC_ORDER_LIBSTDCXX) {
        std::nth_element(kps.begin(), kps.begin() + numFeatures, kps.end(),
                         [](const orc_keypoint& a, const orc_keypoint& b) { return a.response > b.response; });
    } else {
        // canonical: the kept set is whole histogram bins {response >= stop}; a stable (raster-order preserving)
        // partition is a conforming std::nth_element outcome.
        auto it = std::stable_partition(kps.begin(), kps.end(), [stop](const orc_keypoint& a) {
            return (int)(uint8_t)std::min(std::max(a.response, 0.0f), 255.0f) >= stop; });
        if ((int)(it - kps.begin()) != numFeatures) abort();
    }
    kps.resize(numFeatures);
    return numFeatures;
}

struct AnmsItem { int x, y; float strength; int r; int idx; int next; };

// ref :144-326: suppression radii via the cell-ring search (literal control flow, arrays instead of pointers).
void anmsRadii(const orc_orb_params& p, const std::vector<orc_keypoint>& kps, unsigned numToKeep, int threshold,
               std::vector<AnmsItem>& items)
{
    const int numX = p.num_cells_x, numY = p.num_cells_y;
    const float ROBUST_EPS = 0.002f;
    unsigned n = (unsigned)kps.size();
    items.assign(n, AnmsItem());
    int minX, maxX, minY, maxY;
    minX = maxX = (int)kps[0].x;
    minY = maxY = (int)kps[0].y;
    float minStrength = kps[0].response;
    for (unsigned i = 0; i < n; i++) {
        items[i].idx = (int)i;
        items[i].strength = kps[i].response;
        items[i].x = (int)kps[i].x;
        items[i].y = (int)kps[i].y;
        items[i].r = 0;
        items[i].next = -1;
        minX = std::min(minX, items[i].x); minY = std::min(minY, items[i].y);
        maxX = std::max(maxX, items[i].x); maxY = std::max(maxY, items[i].y);
        minStrength = std::min(minStrength, items[i].strength);
    }
    float robustnessFactor, robustnessFactorInv;
    {
        float hi = p.strong_response - static_cast<float>(threshold);
        float val = std::min(std::max(minStrength - threshold, 0.0f), hi);            // mira::clamp<float>
        float range = std::max<float>(0.0f, p.max_robust_factor - p.min_robust_factor);
        robustnessFactor = p.max_robust_factor - (val / (p.strong_response - threshold)) * range;
        robustnessFactorInv = 1.0f / robustnessFactor;
    }
    std::vector<int> cells((size_t)numX * numY, -1);
    for (unsigned i = 0; i < n; i++) {
        int cellX = (items[i].x - minX) * numX / (maxX + 1 - minX);
        int cellY = (items[i].y - minY) * numY / (maxY + 1 - minY);
        int b = cellY * numX + cellX;
        if (cells[b] < 0) { cells[b] = (int)i; continue; }
        if (items[cells[b]].strength < items[i].strength) { items[i].next = cells[b]; cells[b] = (int)i; continue; }
        int after = cells[b];
        while (items[after].next >= 0 && items[items[after].next].strength > items[i].strength) after = items[after].next;
        items[i].next = items[after].next;
        items[after].next = (int)i;
    }
    int globalMaxR2 = (int)(((double)(maxX - minX)) * ((double)(maxY - minY)) / (double)numToKeep);
    int minCellDelta2;
    {
        int dx = std::max((maxX - minX) / numX, 1);
        int dy = std::max((maxY - minY) / numY, 1);
        minCellDelta2 = std::min(dy, dx) * std::min(dy, dx);
    }
    for (int cy = 0; cy < numY; cy++)
        for (int cx = 0; cx < numX; cx++)
            for (int it = cells[cy * numX + cx]; it >= 0; it = items[it].next) {
                AnmsItem& item = items[it];
                int minR2 = globalMaxR2;
                float s = (item.strength >= 0) ? (item.strength * robustnessFactor + ROBUST_EPS)
                                               : (item.strength * robustnessFactorInv + ROBUST_EPS);
                for (int d = 0; std::max(0, d - 1) * std::max(0, d - 1) * minCellDelta2 < minR2; d++) {
                    for (int yy = -d; yy <= d; yy++) {
                        int cYY = yy + cy;
                        if (cYY < 0 || cYY >= numY) continue;
                        for (int xx = -d; xx <= d; xx++) {
                            int cXX = xx + cx;
                            if (cXX < 0 || cXX >= numX || std::max(std::abs(xx), std::abs(yy)) != d) continue;
                            int o = cells[cYY * numX + cXX];
                            while (o >= 0 && items[o].strength > s) {
                                int ddx = item.x - items[o].x, ddy = item.y - items[o].y;
                                int r = ddx * ddx + ddy * ddy;
                                if (r < minR2) minR2 = r;
                                o = items[o].next;
                            }
                        }
                    }
                    if (d > numX + numY) break;   // rings beyond the grid hold no cells: semantically a no-op, bounds the loop
                }
                item.r = minR2;
            }
}

inline bool anmsLess(const AnmsItem& l, const AnmsItem& r)
{
    if (l.r > r.r) return true; else if (l.r < r.r) return false;
    if (l.strength > r.strength) return true; else if (l.strength < r.strength) return false;
    return l.idx < r.idx;
}

// ref :144-360
void anms(const orc_orb_params& p, std::vector<orc_keypoint>& kps, unsigned numToKeep, int threshold, int mode)
{
    unsigned n = (unsigned)kps.size();
    if (numToKeep > n) return;
    std::vector<AnmsItem> items;
    anmsRadii(p, kps, numToKeep, threshold, items);
    if (mode == ORC_ORDER_LIBSTDCXX)
        std::nth_element(items.begin(), items.begin() + numToKeep, items.end(), anmsLess);
    else
        std::sort(items.begin(), items.end(), anmsLess);      // a fully sorted prefix is a conforming nth_element outcome
    std::vector<orc_keypoint> out;
    out.reserve(numToKeep);
    for (unsigned i = 0; i < numToKeep; i++) out.push_back(kps[items[i].idx]);
    kps.swap(out);
}

// the "if (keypoints.size() > n_l)" branch of ref :721-727
void selectLevel(const orc_orb_params& p, std::vector<orc_keypoint>& kps, int nKeep, int mode)
{
    if (kps.size() > (size_t)nKeep) {
        int maxNum = (int)(nKeep * p.feature_factor);
        retainBest(kps, (int)p.fast_threshold, maxNum, nKeep, p.feature_strength, mode);
        anms(p, kps, (unsigned)nKeep, (int)p.fast_threshold, mode);
    }
}

// ---- pattern ------------------------------------------------------------------------------------------------------
// rows 1..29 = row 0 rotated by 12r degrees in float32 with round-half-even.
// ref :551-560 MakeRandomPattern: cv::RNG(0x34985739), x then y, each rng.uniform(-patchSize / 2, patchSize / 2 + 1).
// cv::RNG is OpenCV's multiply-with-carry generator (core.hpp / operations.hpp, unchanged since 2.x):
//   next(): state = (uint64)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state
//   uniform(int a, int b) = a == b ? a : (int)(next() % (b - a) + a)
// Pinned through stock cv::ORB, which builds the same pattern for patchSize != 31 (tests/test_oracle_vs_cv2.py).
void makeRandomPattern(int patchSize, int* xy, int npoints)
{
    uint64_t state = 0x34985739u;
    auto next = [&]() { state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32); return (uint32_t)state; };
    const int lo = -patchSize / 2, hi = patchSize / 2 + 1;
    for (int i = 0; i < npoints; i++) {
        xy[2 * i] = lo == hi ? lo : (int)(next() % (uint32_t)(hi - lo) + lo);
        xy[2 * i + 1] = lo == hi ? lo : (int)(next() % (uint32_t)(hi - lo) + lo);
    }
}

int briefPattern(int patch, int8_t* out)
{
    const signed char* base = patch == 31 ? kBriefBase31 : patch == 15 ? kBriefBase15 : nullptr;
    if (!base) return -1;
    for (int r = 0; r < 30; r++) {
        double ang = (double)(12 * r) * M_PI / 180.0;
        float c = (float)std::cos(ang), s = (float)std::sin(ang);
        for (int i = 0; i < 512; i++) {
            float x = base[2 * i], y = base[2 * i + 1];
            float xr = x * c - y * s;
            float yr = x * s + y * c;
            out[r * 1024 + 2 * i] = (int8_t)lrintf(xr);
            out[r * 1024 + 2 * i + 1] = (int8_t)lrintf(yr);
        }
    }
    return 0;
}

} // namespace

// ======================================================================================================================
extern "C" {

float orc_fast_atan2(float y, float x) { return fastAtan2Impl(y, x); }
int orc_cv_round_f(float v) { return cvRoundF(v); }

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{
    resizeLinear(src, sw, sh, sstride, dst, dw, dh, dstride);
}

int orc_gaussian_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize)
{
    return gaussianBlur(src, w, h, sstride, dst, dstride, ksize);
}
int orc_gaussian_blur_submatrix_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize, int fused)
{
    return gaussianBlurSubmatrix(src, w, h, sstride, dst, dstride, ksize, fused != 0);
}

int orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int threshold, orc_keypoint* out, int capacity)
{
    int n = 0;
    fast9(img, w, h, stride, threshold, true, nullptr, 0, [&](int x, int y, int score) {
        if (n < capacity) out[n] = orc_keypoint{(float)x, (float)y, 7.f, -1.f, (float)score, 0, -1};
        n++;
    });
    return n;
}

void orc_fast9_score_map(const uint8_t* img, int w, int h, int stride, int threshold, uint8_t* score, int score_stride)
{
    for (int y = 0; y < h; y++) std::memset(score + (size_t)y * score_stride, 0, w);
    fast9(img, w, h, stride, threshold, false, score, score_stride, [](int, int, int) {});
}

int orc_level_layout(const orc_orb_params* p, int w, int h, int* sizes, float* scales, int* nfeat)
{
    Levels L = levelLayout(*p, w, h);
    for (int l = 0; l < L.n; l++) {
        sizes[2 * l] = L.w[l]; sizes[2 * l + 1] = L.h[l];
        scales[l] = L.scale[l]; nfeat[l] = L.nfeat[l];
    }
    return 0;
}

// ref :819-842: level 0 = copy, level l = resize(level l-1) (chained, unblurred)
int orc_build_pyramid(const orc_orb_params* p, const uint8_t* img, int w, int h, int stride, uint8_t** level_ptrs)
{
    Levels L = levelLayout(*p, w, h);
    for (int y = 0; y < h; y++) std::memcpy(level_ptrs[0] + (size_t)y * w, img + (size_t)y * stride, w);
    for (int l = 1; l < L.n; l++)
        resizeLinear(level_ptrs[l - 1], L.w[l - 1], L.h[l - 1], L.w[l - 1], level_ptrs[l], L.w[l], L.h[l], L.w[l]);
    return 0;
}

int orc_select_level(const orc_orb_params* p, orc_keypoint* kps, int n_in, int n_keep, int order_mode)
{
    std::vector<orc_keypoint> v(kps, kps + n_in);
    selectLevel(*p, v, n_keep, order_mode);
    std::copy(v.begin(), v.end(), kps);
    return (int)v.size();
}

int orc_anms_radii(const orc_orb_params* p, const orc_keypoint* kps, int n_in, int n_keep, int* r_out)
{
    std::vector<orc_keypoint> v(kps, kps + n_in);
    std::vector<AnmsItem> items;
    anmsRadii(*p, v, (unsigned)n_keep, (int)p->fast_threshold, items);
    for (int i = 0; i < n_in; i++) r_out[i] = items[i].r;
    return 0;
}

int orc_brief_pattern(int patch_size, int8_t* out) { return briefPattern(patch_size, out); }
void orc_random_pattern(int patch_size, int* xy, int npoints) { makeRandomPattern(patch_size, xy, npoints); }
/* ComputeOrbDescriptors (ref :452-492) on ONE already blurred level with scale 1: x, y = integer pixel centres, angle in degrees */
void orc_generic_descriptors(const uint8_t* img, int stride, const float* xya, int n, int patch_size, uint8_t* desc)
{
    std::vector<int> pat(1024);
    makeRandomPattern(patch_size, pat.data(), 512);
    for (int j = 0; j < n; j++) {
        float angle = xya[3 * j + 2];
        angle *= (float)(3.1415926535897932384626433832795 / 180.0f);
        const float a = (float)cos((double)angle), b = (float)sin((double)angle);
        const uint8_t* center = img + (size_t)cvRoundF(xya[3 * j + 1]) * stride + cvRoundF(xya[3 * j]);
        const int* pp = pat.data();
        auto value = [&](int idx) {
            const float px = (float)pp[2 * idx], py = (float)pp[2 * idx + 1];
            const float x = px * a - py * b, y = px * b + py * a;
            return (int)center[cvRoundF(y) * stride + cvRoundF(x)];
        };
        for (int i = 0; i < 32; ++i, pp += 32) {
            int val = 0;
            for (int bit = 0; bit < 8; bit++) val |= (value(2 * bit) < value(2 * bit + 1)) << bit;
            desc[32 * j + i] = (uint8_t)val;
        }
    }
}

int orc_umax(int half_patch, int* umax)
{
    std::vector<int> u;
    computeUmax(half_patch, u);
    std::copy(u.begin(), u.end(), umax);
    return 0;
}

// ref :771-886 DetectAndCompute (+ :642-761 ComputeKeyPoints, :399-437 ICAngles, :502-549 descriptors)
int orc_orb_detect_and_compute(const orc_orb_params* pp, const uint8_t* img, int w, int h, int stride,
                               int order_mode, orc_keypoint* kps, uint8_t* desc, int capacity, int* count)
{
    return orc_orb_detect_and_compute_ex(pp, img, w, h, stride, order_mode, ORC_BLUR_AUTO, kps, desc, capacity, count);
}

int orc_orb_detect_and_compute_ex(const orc_orb_params* pp, const uint8_t* img, int w, int h, int stride,
                                  int order_mode, int blur_mode, orc_keypoint* kps, uint8_t* desc, int capacity, int* count)
{
    const orc_orb_params& p = *pp;
    *count = 0;
    if (p.patch_size < 2 || p.nlevels < 1) return -1;                       // CV_Assert(m_patchSize >= 2)
    if (p.patch_size > 255) return -2;                                      // pattern coordinates are kept in 8 bits on the device
    if (p.gaussian_kernel_size > 1 && !gaussKernelQ8((int)p.gaussian_kernel_size)) return -3;

    Levels L = levelLayout(p, w, h);
    std::vector<std::vector<uint8_t>> pyr(L.n);
    std::vector<uint8_t*> ptrs(L.n);
    for (int l = 0; l < L.n; l++) { pyr[l].resize((size_t)std::max(L.w[l], 1) * std::max(L.h[l], 1)); ptrs[l] = pyr[l].data(); }
    orc_build_pyramid(pp, img, w, h, stride, ptrs.data());

    const int patchSize = (int)p.patch_size, halfPatch = patchSize / 2;
    std::vector<int> umax;
    computeUmax(halfPatch, umax);

    // ---- ComputeKeyPoints
    std::vector<orc_keypoint> all;
    std::vector<orc_keypoint> lvl;
    const int border = p.use_orientation ? cvCeilF(halfPatch * std::sqrt(2.0f)) : halfPatch;
    for (int l = 0; l < L.n; l++) {
        lvl.clear();
        const int lw = L.w[l], lh = L.h[l];
        fast9(ptrs[l], lw, lh, lw, (int)p.fast_threshold, true, nullptr, 0, [&](int x, int y, int score) {
            lvl.push_back(orc_keypoint{(float)x, (float)y, 7.f, -1.f, (float)score, 0, -1});
        });
        // RunByImageBorder ref :619-639
        if (border > 0) {
            if (lh <= border * 2 || lw <= border * 2) lvl.clear();
            else lvl.erase(std::remove_if(lvl.begin(), lvl.end(), [&](const orc_keypoint& k) {
                               int x = cvRoundF(k.x), y = cvRoundF(k.y);
                               return !(border <= x && x < lw - border && border <= y && y < lh - border);
                           }), lvl.end());
        }
        for (auto& k : lvl) { k.octave = l; k.size = patchSize * L.scale[l]; }
        selectLevel(p, lvl, L.nfeat[l], order_mode);
        // ImageData::Insert (ref Image/ImageData.h:65-70): copy while capacity remains
        int room = capacity - (int)all.size();
        int toCopy = std::min((int)lvl.size(), room);
        all.insert(all.end(), lvl.begin(), lvl.begin() + toCopy);
        if (!lvl.empty() && toCopy == 0) break;
    }
    if (all.empty()) return 0;

    // ---- orientation (ref :399-437) on the UNBLURRED pyramid
    if (p.use_orientation) {
        for (auto& k : all) {
            const int lw = L.w[k.octave];
            const uint8_t* center = ptrs[k.octave] + (size_t)cvRoundF(k.y) * lw + cvRoundF(k.x);
            int m_01 = 0, m_10 = 0;
            for (int u = -halfPatch; u <= halfPatch; ++u) m_10 += u * center[u];
            for (int v = 1; v <= halfPatch; ++v) {
                int v_sum = 0, d = umax[v];
                for (int u = -d; u <= d; ++u) {
                    int val_plus = center[u + v * lw], val_minus = center[u - v * lw];
                    v_sum += (val_plus - val_minus);
                    m_10 += u * (val_plus + val_minus);
                }
                m_01 += v * v_sum;
            }
            k.angle = fastAtan2Impl((float)m_01, (float)m_10);
        }
    } else {
        for (auto& k : all) k.angle = 0;
    }
    for (auto& k : all) { float s = L.scale[k.octave]; k.x *= s; k.y *= s; }

    // ---- blur (ref :853-865). Per-level REFLECT_101; see DESIGN.md for when this equals the reference's ROI blur.
    // The source is imagePyramid(layerInfo[level]): a proper submatrix of the (cols + 15) & -16 wide packed buffer (ref :792-812)
    // unless a single level fills it, and cv::GaussianBlur routes submatrices through its generic float path.
    if (p.gaussian_kernel_size > 1) {
        const bool submatrix = L.n > 1 || (w & 15) != 0;
        const bool floatPath = blur_mode == ORC_BLUR_FLOAT_FUSED || blur_mode == ORC_BLUR_FLOAT_UNFUSED || (blur_mode == ORC_BLUR_AUTO && submatrix);
        for (int l = 0; l < L.n; l++) {
            if (L.w[l] < 1 || L.h[l] < 1) continue;
            std::vector<uint8_t> tmp(pyr[l].size());
            if (floatPath) gaussianBlurSubmatrix(ptrs[l], L.w[l], L.h[l], L.w[l], tmp.data(), L.w[l], (int)p.gaussian_kernel_size, blur_mode != ORC_BLUR_FLOAT_UNFUSED);
            else gaussianBlur(ptrs[l], L.w[l], L.h[l], L.w[l], tmp.data(), L.w[l], (int)p.gaussian_kernel_size);
            pyr[l].swap(tmp);
            ptrs[l] = pyr[l].data();
        }
    }

    // ---- descriptors: generic pattern (ref :452-492) for patch sizes without a pre-rotated table (ref :878-885)
    if (patchSize != 31 && patchSize != 15) {
        std::vector<int> pat(1024);
        makeRandomPattern(patchSize, pat.data(), 512);
        for (size_t j = 0; j < all.size(); j++) {
            const orc_keypoint& k = all[j];
            const int lw = L.w[k.octave];
            float scale = 1.f / L.scale[k.octave];
            float angle = k.angle;
            angle *= (float)(3.1415926535897932384626433832795 / 180.0f);          // (float)(CV_PI / HALF_CIRCLE_DEGREES<float>)
            const float a = (float)cos((double)angle), b = (float)sin((double)angle);
            const uint8_t* center = ptrs[k.octave] + (size_t)cvRoundF(k.y * scale) * lw + cvRoundF(k.x * scale);
            uint8_t* d = desc + j * 32;
            const int* pp = pat.data();
            auto value = [&](int idx) {
                const float px = (float)pp[2 * idx], py = (float)pp[2 * idx + 1];
                const float x = px * a - py * b, y = px * b + py * a;                 // float32, products rounded separately
                return (int)center[cvRoundF(y) * lw + cvRoundF(x)];
            };
            for (int i = 0; i < 32; ++i, pp += 32) {
                int val = 0;
                for (int bit = 0; bit < 8; bit++) val |= (value(2 * bit) < value(2 * bit + 1)) << bit;
                d[i] = (uint8_t)val;
            }
        }
        std::copy(all.begin(), all.end(), kps);
        *count = (int)all.size();
        return 0;
    }
    // ---- descriptors (ref :502-549)
    std::vector<int8_t> pattern(30 * 1024);
    briefPattern(patchSize, pattern.data());
    for (size_t j = 0; j < all.size(); j++) {
        const orc_keypoint& k = all[j];
        const int lw = L.w[k.octave];
        float scale = 1.f / L.scale[k.octave];
        int angleIncrement = cvRoundF(k.angle / 12) % 30;
        const uint8_t* center = ptrs[k.octave] + (size_t)cvRoundF(k.y * scale) * lw + cvRoundF(k.x * scale);
        const int8_t* pat = pattern.data() + angleIncrement * 1024;
        uint8_t* d = desc + j * 32;
        for (int i = 0; i < 32; ++i, pat += 32) {
            int val = 0;
            for (int bit = 0; bit < 8; bit++) {
                int t0 = center[pat[4 * bit + 1] * lw + pat[4 * bit]];
                int t1 = center[pat[4 * bit + 3] * lw + pat[4 * bit + 2]];
                val |= (t0 < t1) << bit;
            }
            d[i] = (uint8_t)val;
        }
    }
    std::copy(all.begin(), all.end(), kps);
    *count = (int)all.size();
    return 0;
}

// ---- Match ------------------------------------------------------------------------------------------------------------
// ref Tracking/FeatureMatcher.cpp:453-504 (non-ARM branch): 8 x 32-bit SWAR popcount
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int result = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4);
        std::memcpy(&y, b + 4 * i, 4);
        uint32_t bits = x ^ y;
        bits = bits - ((bits >> 1) & 0x55555555u);
        bits = (bits & 0x33333333u) + ((bits >> 2) & 0x33333333u);
        result += (int)((((bits + (bits >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return result;
}

// ref Tracking/FeatureMatcher.cpp:61-190. cv::BFMatcher(NORM_HAMMING,false).radiusMatch(A,B,maxDist,noArray(),true):
// per query keep every train with d <= maxDist, sorted ascending by distance. The reference's
// bestBackwardsMatch lookup is sized by the COMPACTED row count but indexed by queryIdx (:123-137,158) -- a latent
// out-of-range access whenever some B row has no candidate; the evident intent (lookup sized nB) is restated here.
// With min_diff >= 1 the result is independent of the sort's tie order; for min_diff <= 0 ties resolve to the lowest index.
int orc_match(const uint8_t* descA, int nAall, const uint8_t* maskA, const uint8_t* descB, int nBall, const uint8_t* maskB,
              int max_hamming, int min_diff, orc_dmatch* out, int* count)
{
    *count = 0;
    std::vector<int> ia, ib;
    for (int i = 0; i < nAall; i++) if (!maskA || maskA[i]) ia.push_back(i);
    for (int i = 0; i < nBall; i++) if (!maskB || maskB[i]) ib.push_back(i);
    if (ia.empty() || ib.empty()) return 0;
    const int nA = (int)ia.size(), nB = (int)ib.size();
    const float maxDist = (float)max_hamming;
    struct Best { int idx = -1; float d0 = 0, d1 = 0; int n = 0; };
    auto direction = [&](const uint8_t* Q, const std::vector<int>& iq, const uint8_t* T, const std::vector<int>& itr,
                         std::vector<Best>& best) {
        best.assign(iq.size(), Best());
        for (size_t q = 0; q < iq.size(); q++) {
            Best b;
            for (size_t t = 0; t < itr.size(); t++) {
                float d = (float)orc_descriptor_distance(Q + 32 * (size_t)iq[q], T + 32 * (size_t)itr[t]);
                if (!(d <= maxDist)) continue;
                if (b.n == 0) { b.idx = (int)t; b.d0 = d; }
                else if (d < b.d0) { b.d1 = b.d0; b.d0 = d; b.idx = (int)t; }
                else if (b.n == 1 || d < b.d1) { b.d1 = d; }
                b.n++;
            }
            best[q] = b;
        }
    };
    std::vector<Best> fwd, bwd;
    direction(descA, ia, descB, ib, fwd);
    direction(descB, ib, descA, ia, bwd);
    std::vector<int> bestBack(nB, -1);
    for (int i = 0; i < nB; i++) {
        const Best& b = bwd[i];
        if (b.n == 0) continue;
        if (b.n > 1 && (b.d1 - b.d0) < (float)min_diff) continue;
        bestBack[i] = b.idx;
    }
    int n = 0;
    for (int i = 0; i < nA; i++) {
        const Best& b = fwd[i];
        if (b.n == 0) continue;
        if (b.n > 1 && (b.d1 - b.d0) < (float)min_diff) continue;
        if (bestBack[b.idx] == i) out[n++] = orc_dmatch{ia[i], ib[b.idx], b.d0};
    }
    *count = n;
    return n;
}

} // extern "C"

// ---- IndexedMatch -----------------------------------------------------------------------------------------------------
// ref Tracking/FeatureMatcher.cpp:192-268 (IndexedMatch) with TrackMatch (:22-57). The vocabulary lookups
// (BaseBow::QueryFeatures / BaseFeatureMatcher::QueryFeatures, BoW subsystem, out of scope) are inputs: CSR candidate lists
// in the order QueryFeatures returns them, a2b for every A feature and b2a for every B feature.
namespace {
struct TrackMatchResult { size_t MatchIdx; int HammingDistance; };

void track_match(const uint8_t* descLeft, const uint8_t* descsRight, size_t idxRight, const uint8_t* rightMask, TrackMatchResult& best,
                 TrackMatchResult& second, int maxHamming)
{
    if (!rightMask || rightMask[idxRight]) {
        int hamming_dist = orc_descriptor_distance(descLeft, descsRight + 32 * idxRight);
        if (hamming_dist < maxHamming) {
            if (hamming_dist < best.HammingDistance) { second = best; best = {idxRight, hamming_dist}; }
            else if (hamming_dist < second.HammingDistance) { second = {idxRight, hamming_dist}; }
        }
    }
}
} // namespace

extern "C" int orc_indexed_match(const uint8_t* descA, int nA, const uint8_t* maskA, const uint8_t* descB, int nB, const uint8_t* maskB,
                                 const int* a2b_off, const int* a2b, const int* b2a_off, const int* b2a, int maxHammingDist,
                                 int minHammingDifference, orc_dmatch* out)
{
    size_t cntA = 0, cntB = 0;
    for (int i = 0; i < nA; i++) cntA += (!maskA || maskA[i]);
    for (int i = 0; i < nB; i++) cntB += (!maskB || maskB[i]);
    if (cntA == 0 || cntB == 0) return 0;
    const int maxHamming = maxHammingDist + 1;
    const size_t none = (size_t)-1;
    std::vector<std::pair<size_t, size_t>> matches;
    for (size_t idxA = 0; idxA < (size_t)nA; ++idxA) {
        if (maskA && !maskA[idxA]) continue;
        TrackMatchResult best{none, maxHamming}, second{none, maxHamming};
        for (int j = a2b_off[idxA]; j < a2b_off[idxA + 1]; ++j) track_match(descA + 32 * idxA, descB, (size_t)a2b[j], maskB, best, second, maxHamming);
        if (best.HammingDistance < maxHamming &&
            (second.HammingDistance >= maxHamming || second.HammingDistance - best.HammingDistance >= minHammingDifference))
            matches.emplace_back(idxA, best.MatchIdx);
    }
    int n = 0;
    for (const auto& match : matches) {
        const size_t idxB = match.second;
        TrackMatchResult best{none, maxHamming}, second{none, maxHamming};
        for (int j = b2a_off[idxB]; j < b2a_off[idxB + 1]; ++j) track_match(descB + 32 * idxB, descA, (size_t)b2a[j], maskA, best, second, maxHamming);
        if (best.HammingDistance < maxHamming && best.MatchIdx == match.first &&
            (second.HammingDistance >= maxHamming || second.HammingDistance - best.HammingDistance >= minHammingDifference))
            out[n++] = orc_dmatch{(int)best.MatchIdx, (int)idxB, (float)best.HammingDistance};
    }
    return n;
}

// ---- RadiusMatch ------------------------------------------------------------------------------------------------------
// The reference gates candidates with a boost::geometry R*-tree (rstar<12>) built by the range constructor, i.e. boost's
// packing algorithm (boost 1.67 index/detail/rtree/pack_create.hpp, vendored by the reference): top-down, split at an
// element-count median along the longest edge of the (hint) bounding box with std::nth_element, leaves of <= 12 values.
// A box query reports values in depth-first order of that tree, so the enumeration order of the candidates -- the only thing
// RadiusMatch's result depends on beyond the candidate set -- is the order of the packed leaves. The restatement below
// reproduces that order (rank of every value); it is validated against the real boost R-tree (oracle/radius_ref.cpp).
namespace {
struct PackEntry { float c[3]; int idx; };
struct PackBox { float lo[3], hi[3]; };
struct PackCounts { size_t maxc, minc; };
constexpr size_t kRtMax = 12, kRtMin = 3;                          // rstar<12>: min = 0.3 * max

size_t packMedianCount(size_t count, const PackCounts& sc)         // pack_create.hpp calculate_median_count
{
    size_t n = count / sc.maxc, r = count % sc.maxc, median = (n / 2) * sc.maxc;
    if (r != 0) {
        if (sc.minc <= r) median = ((n + 1) / 2) * sc.maxc;
        else {
            size_t cmm = count - sc.minc;
            n = cmm / sc.maxc; r = cmm % sc.maxc;
            if (r == 0) median = ((n + 1) / 2) * sc.maxc;
            else median = (n == 0) ? r : ((n + 2) / 2) * sc.maxc;
        }
    }
    return median;
}
void packPerLevel(PackEntry* first, PackEntry* last, const PackBox& hint, size_t count, const PackCounts& sc, std::vector<int>& order);
void packPackets(PackEntry* first, PackEntry* last, const PackBox& hint, size_t count, const PackCounts& sc, const PackCounts& next, std::vector<int>& order)
{
    if (count <= sc.maxc) { packPerLevel(first, last, hint, count, next, order); return; }
    size_t medianCount = packMedianCount(count, sc);
    PackEntry* median = first + medianCount;
    float len = hint.hi[0] - hint.lo[0]; int dim = 0;               // pack_utils::biggest_edge
    for (int d = 1; d < 3; d++) { float cur = hint.hi[d] - hint.lo[d]; if (len < cur) { dim = d; len = cur; } }
    std::nth_element(first, median, last, [dim](const PackEntry& a, const PackEntry& b) { return a.c[dim] < b.c[dim]; });
    PackBox left = hint, right = hint;
    float mid = hint.lo[dim] + (hint.hi[dim] - hint.lo[dim]) / 2;
    left.hi[dim] = mid; right.lo[dim] = mid;
    packPackets(first, median, left, medianCount, sc, next, order);
    packPackets(median, last, right, count - medianCount, sc, next, order);
}
void packPerLevel(PackEntry* first, PackEntry* last, const PackBox& hint, size_t count, const PackCounts& sc, std::vector<int>& order)
{
    if (sc.maxc <= 1) { for (PackEntry* e = first; e != last; ++e) order.push_back(e->idx); return; }      // leaf: values in range order
    PackCounts next{sc.maxc / kRtMax, sc.minc / kRtMax};
    packPackets(first, last, hint, count, sc, next, order);
}
// enumeration (depth-first) order of the packed tree
void rtreeOrder(const orc_keypoint* kps, int n, std::vector<int>& order)
{
    order.clear();
    if (n <= 0) return;
    std::vector<PackEntry> e(n);
    PackBox box;
    for (int i = 0; i < n; i++) {
        e[i] = PackEntry{{kps[i].x, kps[i].y, kps[i].octave * 100.f}, i};
        for (int d = 0; d < 3; d++) {
            if (i == 0) { box.lo[d] = box.hi[d] = e[i].c[d]; }
            else { box.lo[d] = std::min(box.lo[d], e[i].c[d]); box.hi[d] = std::max(box.hi[d], e[i].c[d]); }
        }
    }
    PackCounts sc{1, 1};                                              // calculate_subtree_elements_counts
    for (size_t smax = kRtMax; smax < (size_t)n; smax *= kRtMax) sc.maxc = smax;
    sc.minc = kRtMin * (sc.maxc / kRtMax);
    packPerLevel(e.data(), e.data() + n, box, (size_t)n, sc, order);
}
} // namespace

extern "C" int orc_rtree_order(const orc_keypoint* kps, int n, int* order_out)
{
    std::vector<int> order;
    rtreeOrder(kps, n, order);
    std::copy(order.begin(), order.end(), order_out);
    return (int)order.size();
}

// ref Tracking/FeatureMatcher.cpp:294-446 + Image/KeypointSpatialIndex.cpp:89-97 (box query: same octave, |dx|,|dy| <= radius)
extern "C" int orc_radius_match(const orc_keypoint* qk, int nq, const float* qpos_override, const uint8_t* qmask, const uint8_t* qdesc,
                                const orc_keypoint* tk, int nt, const uint8_t* tmask, const uint8_t* tdesc, float radius, int maxHamming,
                                int minDiff, orc_dmatch* out)
{
    std::vector<int> order;
    rtreeOrder(tk, nt, order);
    std::vector<orc_dmatch> almost;
    for (int q = 0; q < nq; q++) {
        if (qmask && !qmask[q]) continue;
        const float px = qpos_override ? qpos_override[2 * q] : qk[q].x, py = qpos_override ? qpos_override[2 * q + 1] : qk[q].y;
        const float lox = px - radius, hix = px + radius, loy = py - radius, hiy = py + radius;
        const float loz = qk[q].octave * 100.f - 1.f, hiz = qk[q].octave * 100.f + 1.f;
        int best = maxHamming + 1, second = INT_MAX;
        orc_dmatch bm{0, -1, 0.f};
        for (int t : order) {                                                   // enumeration order of the R-tree
            const float z = tk[t].octave * 100.f;
            if (!(lox <= tk[t].x && tk[t].x <= hix && loy <= tk[t].y && tk[t].y <= hiy && loz <= z && z <= hiz)) continue;
            if (tmask && !tmask[t]) continue;
            int d = orc_descriptor_distance(qdesc + 32 * (size_t)q, tdesc + 32 * (size_t)t);
            if (d < best) { bm.train = t; bm.distance = (float)d; second = best; best = d; }
        }
        if (bm.train != -1 && (second - best) > minDiff) { bm.query = q; almost.push_back(bm); }
    }
    int n = 0;
    if (almost.size() > 1) {
        std::vector<float> bestD(nt, FLT_MAX), secondD(nt, FLT_MAX);
        for (const auto& m : almost) {
            if (m.distance < bestD[m.train]) { secondD[m.train] = bestD[m.train]; bestD[m.train] = m.distance; }
            else if (m.distance < secondD[m.train]) secondD[m.train] = m.distance;
        }
        for (const auto& m : almost) if (m.distance == bestD[m.train] && bestD[m.train] < secondD[m.train]) out[n++] = m;
    } else {
        for (const auto& m : almost) out[n++] = m;
    }
    return n;
}
