// ORB front-end for sm_100a: pyramid -> FAST-9/16 + score + NMS -> retain-best + ANMS -> orientation -> blur -> rBRIEF.
//
// Replaces OrbDetector::DetectAndCompute (ref Core/MAGESLAM/Source/Image/OpenCVModified.cpp:771-886) behind
// include/mage_b200.h. Integer/byte work, HBM/L2 bound: no tensor cores. A batch of frames is processed per launch
// (grid.y / grid.z = frame) so the 148 SMs are filled; every kernel reads its geometry from one by-value constant block.
//
// Data layout in HBM (one arena per handle):
//   pyr[f]   : levels 0..L-1 of frame f, each with a 64-byte aligned row pitch (unblurred, chained bilinear)
//   blur[f]  : same geometry, Gaussian-blurred copy (descriptors sample this one; orientation samples pyr)
//   cand[f]  : per level, packed u32 (score << 24 | y*w + x) appended with atomics by the FAST kernel
//   sel[f]   : per level, the <= n_l selected keypoints in final order, same packing
// The selection order is the canonical one of SURVEY.md section 7 (stable raster order in RetainBestFeatures, fully
// sorted ANMS prefix) -- a conforming execution of the reference's std::nth_element calls.
#include "common.cuh"

#include <cuda.h>              // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved at run time, libcuda is not linked)

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace mage {

#include "brief_base_patterns.inc"

constexpr int kMaxLevels = 16;
constexpr int kMaxCells = 2048;
constexpr int kSelThreads = 1024;        // launch bound; batches launch half of it (see the launch site)
constexpr int kSelSmemItems = 2048;      // keypoints per (frame, level) handled entirely in shared memory (34 KB: four 512-thread CTAs per SM; a level keeping more uses the global scratch)

struct LevelGeom {
    int w, h, pitch, nfeat;
    float scale;
    unsigned pyr_off;       // byte offset inside a frame's pyramid slab
    unsigned cand_off;      // entry offset inside a frame's candidate arrays
    unsigned cand_cap;
    unsigned key_off;       // entry offset inside a frame's key array (pow2-padded capacities)
    unsigned key_cap;
    unsigned sel_off;       // entry offset inside a frame's selected array
    unsigned tab_x, tab_y;  // offsets of the resize tables (destination = this level)
    int fast_tiles_x, fast_tiles_y, fast_tile_base;       // FAST tile grid over the border-inset rectangle
    int fast_x0, fast_y0;                                // staged-pixel origin of tile (0, 0); x is a multiple of 16
    int blur_tiles_x, blur_tiles_y, blur_tile_base;      // tile rows = kBlurOH (generic kernel) or 8*kBlur7Rows (7x7 fast path)
};

struct OrbGeom {
    int nlevels, width, height;
    int fast_threshold, border, patch, half_patch, use_orientation, ksize;
    int strong_response, num_cells_x, num_cells_y;
    float feature_factor, feature_strength, min_rf, max_rf;
    int umax[66];           // half_patch + 2 entries (half_patch <= 63)
    int generic_pattern;    // patch size without a pre-rotated table: runtime rotation of the cv::RNG pattern (ref :452-492)
    int gk[16];
    float gkf[8];           // first half of the float Gaussian kernel (edge .. centre), used when blur_float != 0
    int blur_float;         // 0: Q8.8 fixed point; 1 / 2: cv::GaussianBlur's float path (level ROIs are proper submatrices of the reference's packed buffer), fused / unfused
    LevelGeom lv[kMaxLevels];
};

struct OrbBuffers {
    const uint8_t* lvl0;        // level-0 pixels (user device buffer or the handle's staging copy)
    size_t lvl0_frame_stride;
    int lvl0_pitch;
    uint8_t* pyr;               // unblurred pyramid slabs (levels >= 1; level 0 only when it is the staging copy)
    uint8_t* blur;              // blurred pyramid slabs (all levels)
    size_t slab;                // bytes per frame slab
    const int* tab_ofs;         // resize tables
    const short2* tab_coef;
    uint32_t* cand;  int* cand_count;  size_t cand_stride;    // [f][cand_total], [f][kMaxLevels]
    uint32_t* kept;  uint32_t* cxy;  uint8_t* csc;            // selection scratch, same strides as cand
    unsigned long long* keys; size_t key_stride;
    uint32_t* sel;   int* sel_count;   size_t sel_stride;     // [f][sel_total], [f][kMaxLevels]
    int* status;                                                // [f]
    const int8_t* pattern;                                      // 30 x 1024 pre-rotated BRIEF table, or the 512 (x, y) points of the generic pattern
};

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ const uint8_t* level_ptr(const OrbGeom& g, const OrbBuffers& b, int f, int l, int& pitch)
{
    if (l == 0) { pitch = b.lvl0_pitch; return b.lvl0 + (size_t)f * b.lvl0_frame_stride; }
    pitch = g.lv[l].pitch;
    return b.pyr + (size_t)f * b.slab + g.lv[l].pyr_off;
}
__device__ __forceinline__ uint8_t* blur_ptr(const OrbGeom& g, const OrbBuffers& b, int f, int l)
{
    return b.blur + (size_t)f * b.slab + g.lv[l].pyr_off;
}

// ------------------------------------------------------------------------------------------------ K1: pyramid level
// ref OpenCVModified.cpp:819-842 -> cv::resize(level l-1, INTER_LINEAR) restated in SURVEY appendix A.2.
// One thread = 4 horizontally adjacent output pixels (one 32-bit store); tables hold source offsets + 11-bit coefficients.
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ OrbGeom g, const OrbBuffers b, int l)
{
    const LevelGeom& D = g.lv[l];
    const LevelGeom& S = g.lv[l - 1];
    const int f = blockIdx.z;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (dy >= D.h || dx0 >= D.w) return;
    int spitch;
    const uint8_t* src = level_ptr(g, b, f, l - 1, spitch);
    uint8_t* dst = b.pyr + (size_t)f * b.slab + D.pyr_off;
    const int sy = b.tab_ofs[D.tab_y + dy];
    const short2 cy = b.tab_coef[D.tab_y + dy];
    const int sy1 = min(sy + 1, S.h - 1);
    const uint8_t* r0 = src + (size_t)sy * spitch;
    const uint8_t* r1 = src + (size_t)sy1 * spitch;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int dx = dx0 + i;
        if (dx < D.w) {
            int sx = b.tab_ofs[D.tab_x + dx];
            short2 cx = b.tab_coef[D.tab_x + dx];
            int sx1 = min(sx + 1, S.w - 1);
            int h0 = r0[sx] * cx.x + r0[sx1] * cx.y;
            int h1 = r1[sx] * cx.x + r1[sx1] * cx.y;
            int v = ((((int)cy.x * (h0 >> 4)) >> 16) + (((int)cy.y * (h1 >> 4)) >> 16) + 2) >> 2;
            packed |= (uint32_t)(v & 0xff) << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)dy * D.pitch + dx0) = packed;      // pitch % 64 == 0: padding is writable
}

// Fast variant for scale factors <= 3 (every shipped setting): a thread owns 4 output columns and walks kResizeRows rows, so the
// column tables, tap offsets and byte selectors are set up once. Per source row and pixel PAIR it loads two aligned words, gathers
// the four taps (L_a, R_a, L_b, R_b) with one PRMT and evaluates both horizontal interpolations with IDP.2A (16-bit coefficient
// pair x byte pair) -- 8 loads / 4 PRMT / 8 IDP.2A per 4 pixels instead of 16 byte loads + 8 table loads + 16 multiplies.
constexpr int kResizeRows = 4;
__global__ void __launch_bounds__(256) k_resize4(const __grid_constant__ OrbGeom g, const OrbBuffers b, int l)
{
    const LevelGeom& D = g.lv[l];
    const LevelGeom& S = g.lv[l - 1];
    const int f = blockIdx.z;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    // programmatic dependent launch along the chain of levels: the next level's grid may become resident and set its coefficients up
    // while this one is still running; it waits below, before it touches the level this grid writes
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (dx0 >= D.w) return;
    int spitch;
    const uint8_t* src = level_ptr(g, b, f, l - 1, spitch);
    uint8_t* dst = b.pyr + (size_t)f * b.slab + D.pyr_off;
    uint32_t cf[4], sel[2];
    int o0[2], o1[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const int da = min(dx0 + 2 * p, D.w - 1), db = min(dx0 + 2 * p + 1, D.w - 1);        // columns past the width land in the padding
        const int sa = b.tab_ofs[D.tab_x + da], sb = b.tab_ofs[D.tab_x + db];
        const short2 ca = b.tab_coef[D.tab_x + da], cb = b.tab_coef[D.tab_x + db];
        cf[2 * p] = (uint32_t)(uint16_t)ca.x | ((uint32_t)(uint16_t)ca.y << 16);
        cf[2 * p + 1] = (uint32_t)(uint16_t)cb.x | ((uint32_t)(uint16_t)cb.y << 16);
        const int sa1 = min(sa + 1, S.w - 1), sb1 = min(sb + 1, S.w - 1), base = sa & ~3;
        sel[p] = (uint32_t)(sa - base) | ((uint32_t)(sa1 - base) << 4) | ((uint32_t)(sb - base) << 8) | ((uint32_t)(sb1 - base) << 12);
        o0[p] = base;
        o1[p] = (sb1 - base >= 4) ? base + 4 : base;             // the second word is only touched when a tap lies in it (never past the row)
    }
    const int dyb = blockIdx.y * (8 * kResizeRows) + threadIdx.y;
    asm volatile("griddepcontrol.wait;" ::: "memory");          // the source level is complete and visible (no-op for a plain launch)
#pragma unroll
    for (int k = 0; k < kResizeRows; k++) {
        const int dy = dyb + 8 * k;
        if (dy >= D.h) break;
        const int sy = b.tab_ofs[D.tab_y + dy];
        const short2 cy = b.tab_coef[D.tab_y + dy];
        const uint32_t cy0 = (uint32_t)cy.x << 16, cy1 = (uint32_t)cy.y << 16;        // (cy * v) >> 16 == umulhi(cy << 16, v)
        const uint8_t* r0 = src + (size_t)sy * spitch;
        const uint8_t* r1 = src + (size_t)min(sy + 1, S.h - 1) * spitch;
        uint32_t v[4];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const uint32_t x0 = __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(r0 + o0[p])), __ldg(reinterpret_cast<const uint32_t*>(r0 + o1[p])), sel[p]);
            const uint32_t x1 = __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(r1 + o0[p])), __ldg(reinterpret_cast<const uint32_t*>(r1 + o1[p])), sel[p]);
            const uint32_t h0a = __dp2a_lo(cf[2 * p], x0, 0u), h0b = __dp2a_hi(cf[2 * p + 1], x0, 0u);
            const uint32_t h1a = __dp2a_lo(cf[2 * p], x1, 0u), h1b = __dp2a_hi(cf[2 * p + 1], x1, 0u);
            v[2 * p] = (__umulhi(cy0, h0a >> 4) + __umulhi(cy1, h1a >> 4) + 2u) >> 2;
            v[2 * p + 1] = (__umulhi(cy0, h0b >> 4) + __umulhi(cy1, h1b >> 4) + 2u) >> 2;
        }
        const uint32_t packed = (v[0] & 0xffu) | ((v[1] & 0xffu) << 8) | ((v[2] & 0xffu) << 16) | (v[3] << 24);
        *reinterpret_cast<uint32_t*>(dst + (size_t)dy * D.pitch + dx0) = packed;      // pitch % 64 == 0: padding is writable
    }
}

// ------------------------------------------------------------------------------------------------ K2: FAST + score + NMS
// ref OpenCVModified.cpp:1224-1512 (FAST_t<16>), :926-1071 (cornerScore<16>), :619-639 (RunByImageBorder).
// Closed form: score = max(max_k min(d[k..k+8]), -min_k max(d[k..k+8])) - 1, corner <=> score >= thr.
//
// Only key points inside the border survive (ref :712, :619-639), so only the border-inset rectangle, dilated by the one pixel
// the 3x3 suppression looks at, is scored: the tile grid starts at the border instead of the image corner. The kernel is bound by
// instruction issue, not by HBM; what it is built around:
//   * a pixel PAIR per thread in packed 16-bit lanes; the pixel tile is staged twice as 16-bit elements (A[y][x] and
//     As[y][x] = A[y][x+1]) so that every (pixel, right neighbour) pair is one aligned 32-bit shared-memory load;
//   * the staged element is the HALF-PRECISION number 1024 + pixel (bits 0x6400 | pixel). Bit patterns of positive halves order
//     like integers, so packed integer min/max (ALU pipe) and half2 arithmetic (FMA pipe) work on the same registers: part of the
//     min/max network runs as  d = relu(a - b), max = b + d, min = a - d  (exact: every intermediate is an integer below 2048),
//     which spreads the network over both pipes instead of saturating the ALU pipe alone;
//   * work items are laid out flat over the scored rectangle of a tile, so partial tiles at the right / bottom do not idle lanes;
//   * the 3x3 suppression walks down the score tile keeping the row-wise maxima of the two previous rows in registers.
constexpr int kFastTW = 112, kFastTH = 64;                       // key-point (output) tile
constexpr int kFastPW = 128, kFastPH = 72;                       // staged pixel tile: 4-pixel apron (3 ring + 1 suppression neighbour); origin and width are multiples of 16 bytes
constexpr int kFastSW = kFastTW + 2, kFastSH = kFastTH + 2;      // score tile
constexpr int kFastNP = kFastSW / 2;                             // score pairs per row (57)
constexpr int kFastScP = 64;                                     // score row pitch in pair words (two u16 scores per word; the tail stays zero)
constexpr int kFastP16 = kFastPW;                                // 16-bit elements per staged row
constexpr int kFastAsOff = kFastPH * kFastP16;                   // As = A + kFastAsOff
constexpr size_t kFastSmemBytes = 2 * (size_t)kFastPH * kFastP16 * sizeof(uint16_t) + (size_t)kFastSH * kFastScP * sizeof(uint32_t);
static_assert(kFastSW % 2 == 0 && kFastNP <= kFastScP - 2 && kFastTW + 8 <= kFastPW && kFastTH + 8 == kFastPH, "FAST tile geometry");

__device__ __forceinline__ int min3(int a, int b, int c) { return min(min(a, b), c); }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

// half2 arithmetic on raw 32-bit registers (FMA pipe)
__device__ __forceinline__ uint32_t h2_relu_diff(uint32_t a, uint32_t b)      // relu(a - b)
{
    uint32_t d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));
    return d;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b)            // a + b
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0x3C003C00u), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b)            // a - b
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));
    return d;
}
// minimum and maximum of the same two operands: two ALU-pipe instructions, or three FMA-pipe ones
template <bool FMA> __device__ __forceinline__ void h2_minmax(uint32_t a, uint32_t b, uint32_t& mn, uint32_t& mx)
{
    if (FMA) { const uint32_t d = h2_relu_diff(a, b); mx = h2_add(b, d); mn = h2_sub(a, d); }
    else { mn = __vmins2(a, b); mx = __vmaxs2(a, b); }
}

// `base` = address of staged element (first pixel of the pair) - 1, an even element index: pair (x + DX, x + DX + 1) is A[...] when
// x + DX is even and As[x + DX - 1] when it is odd -- either way one aligned word at a compile-time offset from `base`
template <int DX, int DY>
__device__ __forceinline__ uint32_t ring_pair(const uint16_t* base)
{
    constexpr int off = DY * kFastP16 + ((DX & 1) != 0 ? DX + 1 : DX + kFastAsOff);
    return *reinterpret_cast<const uint32_t*>(base + off);
}

// NP / NX / NQ: how many of the 8 + 8 first-level (min, max) pairs and of the 8 second-level pairs run on the FMA pipe
template <int NP, int NX, int NQ>
__device__ __forceinline__ uint32_t fast_score_pair(const uint16_t* base, uint32_t thr_h2)
{
    // min/max commute with the subtraction of the centre, so the whole network runs on the RAW ring values R[0..15] and the
    // centre v enters once at the end:  score + 1 = max(v - T, S - v),  S = max_k min(R[k..k+8]),  T = min_k max(R[k..k+8]).
    // Windows k = 2j and 2j+1 share R[2j+1..2j+8]; max(min(sh, a), min(sh, b)) = min(sh, max(a, b)), hence
    //   S = max_j min(R[2j+1..2j+8], max(R[2j], R[2j+9])),  T = min_j max(R[2j+1..2j+8], min(R[2j], R[2j+9]))
    // -> 36 packed min/max per polarity.
    uint32_t R[16];
    R[0] = ring_pair<0, 3>(base);    R[1] = ring_pair<1, 3>(base);    R[2] = ring_pair<2, 2>(base);    R[3] = ring_pair<3, 1>(base);
    R[4] = ring_pair<3, 0>(base);    R[5] = ring_pair<3, -1>(base);   R[6] = ring_pair<2, -2>(base);   R[7] = ring_pair<1, -3>(base);
    R[8] = ring_pair<0, -3>(base);   R[9] = ring_pair<-1, -3>(base);  R[10] = ring_pair<-2, -2>(base); R[11] = ring_pair<-3, -1>(base);
    R[12] = ring_pair<-3, 0>(base);  R[13] = ring_pair<-3, 1>(base);  R[14] = ring_pair<-2, 2>(base);  R[15] = ring_pair<-1, 3>(base);
    uint32_t pmin[8], pmax[8], xmin[8], xmax[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < NP) h2_minmax<true>(R[(2 * j + 1) & 15], R[(2 * j + 2) & 15], pmin[j], pmax[j]);
        else h2_minmax<false>(R[(2 * j + 1) & 15], R[(2 * j + 2) & 15], pmin[j], pmax[j]);
        if (j < NX) h2_minmax<true>(R[2 * j], R[(2 * j + 9) & 15], xmin[j], xmax[j]);
        else h2_minmax<false>(R[2 * j], R[(2 * j + 9) & 15], xmin[j], xmax[j]);
    }
    uint32_t qmin[8], qmax[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < NQ) {
            qmin[j] = h2_sub(pmin[j], h2_relu_diff(pmin[j], pmin[(j + 1) & 7]));
            qmax[j] = h2_add(pmax[(j + 1) & 7], h2_relu_diff(pmax[j], pmax[(j + 1) & 7]));
        } else {
            qmin[j] = __vmins2(pmin[j], pmin[(j + 1) & 7]); qmax[j] = __vmaxs2(pmax[j], pmax[(j + 1) & 7]);
        }
    }
    uint32_t ymin[8], ymax[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        ymin[j] = __vimin3_s16x2(qmin[j], qmin[(j + 2) & 7], xmax[j]);
        ymax[j] = __vimax3_s16x2(qmax[j], qmax[(j + 2) & 7], xmin[j]);
    }
    uint32_t S = __vimax3_s16x2(ymin[0], ymin[1], ymin[2]), T = __vimin3_s16x2(ymax[0], ymax[1], ymax[2]);
    S = __vimax3_s16x2(S, ymin[3], ymin[4]); T = __vimin3_s16x2(T, ymax[3], ymax[4]);
    S = __vimax3_s16x2(S, ymin[5], ymin[6]); T = __vimin3_s16x2(T, ymax[5], ymax[6]);
    S = __vmaxs2(S, ymin[7]); T = __vmins2(T, ymax[7]);
    const uint32_t vv = ring_pair<0, 0>(base);
    // Returned BIASED: max(score + 1 - thr, 0) = max(relu(v - thr - T), relu(S - (v + thr))) -- 0 for non-corners and an
    // order-preserving positive value for corners (the suppression only compares; the emitter adds thr - 1 back). Both terms are
    // non-negative halves, so the integer maximum is their maximum; + 1024 puts the integer into the low 10 mantissa bits.
    const uint32_t a1 = h2_relu_diff(h2_sub(vv, thr_h2), T), b1 = h2_relu_diff(S, h2_add(vv, thr_h2));
    return h2_add(__vmaxs2(a1, b1), 0x64006400u) & 0x03ff03ffu;
}

// High-speed test of a pair (ref :1415-1440 tests the same compass points one pixel at a time): every 9-arc of the 16-ring contains
// at least one pixel of each opposite couple {k, k + 8}, so S <= max(R[k], R[k+8]) and T >= min(R[k], R[k+8]) for every k. With the
// couples k = 0 and k = 4:  score >= thr  =>  min(max(R0, R8), max(R4, R12)) - v > thr  or  v - max(min(R0, R8), min(R4, R12)) > thr.
// True when either pixel of the pair can still be a corner; never false for a corner.
__device__ __forceinline__ bool fast_pretest_pair(const uint16_t* base, uint32_t thr_h2)
{
    const uint32_t r0 = ring_pair<0, 3>(base), r4 = ring_pair<3, 0>(base), r8 = ring_pair<0, -3>(base), r12 = ring_pair<-3, 0>(base);
    const uint32_t vv = ring_pair<0, 0>(base);
    const uint32_t lo = __vmins2(__vmaxs2(r0, r8), __vmaxs2(r4, r12)), hi = __vmaxs2(__vmins2(r0, r8), __vmins2(r4, r12));
    const uint32_t t = h2_relu_diff(lo, h2_add(vv, thr_h2)) | h2_relu_diff(h2_sub(vv, thr_h2), hi);
    return (t & 0x7fff7fffu) != 0;
}
constexpr int kFastDenseAt = 22;       // survivors of a 32-pair chunk from which the test costs more than it saves
constexpr int kFastDenseRun = 15;      // chunks a warp stays in dense mode before it probes again

// One entry per tile of the per-frame grid, built on the host by mage_orb_create (fast_tile_table) and read by every thread of the
// CTA with two 16-byte loads: the tile's place, the part of it that is scored and the constants of the flat item walk -- about 200
// instructions of per-thread set-up otherwise. Score column sx <-> image x = px0 + 3 + sx <-> staged column 3 + sx.
struct alignas(16) FastTile {
    short l, px0, py0;                          // level, staged-pixel origin (px0 a multiple of 16)
    short xlo, xhi;                             // key-point columns of this tile inside the border: [xlo, xhi)
    signed char oy_lo, oy_hi;                   // first / last output row inside the border
    unsigned char sp_lo, npx, sy_lo, nrows;     // scored rectangle: first score pair, pairs per row, first score row, rows -- the
                                                // border-inset rectangle dilated by one pixel, clipped to 3 <= x <= w - 4 (FAST's domain)
    unsigned char dsy, dsp;                     // 256 / npx, 256 % npx
    unsigned char g_lo, g_hi;                   // first / last 16-pixel group of a staged row the rectangle touches
    unsigned int inv_npx;                       // ceil(2^32 / npx): t / npx == umulhi(t, inv_npx) for t < 256
    unsigned int pad[2];
};
static_assert(sizeof(FastTile) == 32, "FastTile is read as two uint4");
__device__ __forceinline__ FastTile fast_tile_load(const FastTile* table, int tile)
{
    union { FastTile t; uint4 q[2]; } u;
    u.q[0] = __ldg(reinterpret_cast<const uint4*>(table + tile)); u.q[1] = __ldg(reinterpret_cast<const uint4*>(table + tile) + 1);
    return u.t;
}

// 16 pixels -> 16 staged elements of A (element k = 0x6400 | pixel k) and of As (element k = 0x6400 | pixel k + 1)
__device__ __forceinline__ void fast_widen16(const uint4 v, uint32_t nxt, uint16_t* a)
{
    constexpr uint32_t H = 0x64646464u;
    uint4 o;
    o.x = __byte_perm(v.x, H, 0x4140); o.y = __byte_perm(v.x, H, 0x4342); o.z = __byte_perm(v.y, H, 0x4140); o.w = __byte_perm(v.y, H, 0x4342);
    *reinterpret_cast<uint4*>(a) = o;
    o.x = __byte_perm(v.z, H, 0x4140); o.y = __byte_perm(v.z, H, 0x4342); o.z = __byte_perm(v.w, H, 0x4140); o.w = __byte_perm(v.w, H, 0x4342);
    *reinterpret_cast<uint4*>(a + 8) = o;
    const uint32_t s0 = __funnelshift_r(v.x, v.y, 8), s1 = __funnelshift_r(v.y, v.z, 8), s2 = __funnelshift_r(v.z, v.w, 8), s3 = __funnelshift_r(v.w, nxt, 8);
    o.x = __byte_perm(s0, H, 0x4140); o.y = __byte_perm(s0, H, 0x4342); o.z = __byte_perm(s1, H, 0x4140); o.w = __byte_perm(s1, H, 0x4342);
    *reinterpret_cast<uint4*>(a + kFastAsOff) = o;
    o.x = __byte_perm(s2, H, 0x4140); o.y = __byte_perm(s2, H, 0x4342); o.z = __byte_perm(s3, H, 0x4140); o.w = __byte_perm(s3, H, 0x4342);
    *reinterpret_cast<uint4*>(a + kFastAsOff + 8) = o;
}

__device__ __forceinline__ void fast_zero_scores(uint32_t* sc)
{
    for (int i = threadIdx.x; i < kFastSH * kFastScP / 4; i += 256) reinterpret_cast<uint4*>(sc)[i] = make_uint4(0u, 0u, 0u, 0u);
}

// scores -> strict 3x3 NMS -> border cull -> append, on a pixel tile already staged (A, As) and a zeroed score tile; shared by
// both FAST kernels. Begins with the barrier that publishes the staged tile.
template <int NP, int NX, int NQ>
__device__ __forceinline__ void fast_tile_compute(const OrbGeom& g, const OrbBuffers& b, const LevelGeom& L, int f, const FastTile& T,
                                                  uint16_t* A, uint32_t* sc, uint16_t* queue, int& s_n, int& s_base)
{
    const int tid = threadIdx.x, thr = g.fast_threshold, bd = g.border, l = T.l, px0 = T.px0, py0 = T.py0;
    struct { int sp_lo, npx, sy_lo, nrows; } rg = {T.sp_lo, T.npx, T.sy_lo, T.nrows};
    if (tid == 0) s_n = 0;
    __syncthreads();
    {
        // scores: items = the pairs of the scored rectangle in raster order, a chunk of 32 per warp and step. A warp is either in
        // DENSE mode (every pair goes through the network) or in SPARSE mode: every pair takes the high-speed test first and the
        // survivors are queued per warp until 32 of them fill a full-width pass through the network -- on camera images a few
        // per cent of the pairs survive, on a corner-dense chart most do and the test would only add work. A chunk whose
        // survivors exceed kFastDenseAt switches the warp to dense mode for the next kFastDenseRun chunks, then it probes again.
        // Pairs that fail the test keep the zero the score tile was cleared to; the outputs do not depend on the mode.
        const int nit = rg.npx * rg.nrows, lane = tid & 31;
        const int nchunks = (nit + 255) >> 8;
        if (nchunks > 0) {
            const uint32_t thr_h2 = h2_sub(0x64006400u | (uint32_t)thr | ((uint32_t)thr << 16), 0x64006400u);      // (thr, thr) as halves
            const uint32_t lt = (1u << lane) - 1u;
            uint16_t* q = queue + (tid >> 5) * 64;
            int sy = (int)__umulhi((unsigned)tid, T.inv_npx), sp = tid - sy * rg.npx;
            const int dsy = T.dsy, dsp = T.dsp;
            const bool edge = bd < 5;                              // the rectangle then touches columns FAST is not defined on
            int qn = 0, dense_left = 0;
            for (int k = 0; k <= nchunks; k++) {                   // the last turn only drains the queue
                bool go = false;
                int e = 0;                                         // score-tile word of the pair: row << 6 | pair column
                if (k < nchunks) {
                    const bool valid = k * 256 + tid < nit;
                    e = ((rg.sy_lo + sy) << 6) | (rg.sp_lo + sp);
                    sp += dsp; sy += dsy;
                    if (sp >= rg.npx) { sp -= rg.npx; sy++; }
                    if (dense_left > 0) { dense_left--; go = valid; }
                    else {
                        const bool pass = valid && fast_pretest_pair(A + ((e >> 6) + 3) * kFastP16 + 2 * (e & 63) + 2, thr_h2);
                        const uint32_t m = __ballot_sync(0xffffffffu, pass);
                        const int np = __popc(m);
                        if (np >= kFastDenseAt) { dense_left = kFastDenseRun; go = valid; }
                        else if (np > 0) {
                            if (pass) q[qn + __popc(m & lt)] = (uint16_t)e;
                            qn += np;
                            __syncwarp();
                            if (qn >= 32) { qn -= 32; e = q[qn + lane]; go = true; }
                        }
                    }
                } else if (lane < qn) { e = q[lane]; go = true; }
                if (go) {
                    const int spx = e & 63;
                    uint32_t out = fast_score_pair<NP, NX, NQ>(A + ((e >> 6) + 3) * kFastP16 + 2 * spx + 2, thr_h2);
                    if (edge) {
                        const int x = px0 + 3 + 2 * spx;
                        out &= ((x >= 3 && x <= L.w - 4) ? 0x0000ffffu : 0u) | ((x + 1 >= 3 && x + 1 <= L.w - 4) ? 0xffff0000u : 0u);
                    }
                    sc[e] = out;
                }
                __syncwarp();                                      // queue reads of this turn precede the writes of the next
            }
        }
    }
    __syncthreads();
    // Strict 3x3 non-maximum suppression. Lane = 4 score columns (two pair words), warp = 8 output rows: the warp walks down 10
    // score rows; per row it forms LR = max(left, right) and H = max(LR, centre) in packed u16x2 arithmetic (the horizontal
    // neighbours of a pair are funnel shifts over the words of the adjacent lanes) and keeps them for two rows, so a pixel's
    // eight neighbours are max3(H above, H below, LR). Survivors are collected in shared memory (reusing the pixel tile) and
    // appended with ONE global atomic per CTA.
    uint32_t* list = reinterpret_cast<uint32_t*>(A);            // <= 1792 survivors (strict NMS: one per 2x2 block)
    const int lane = tid & 31, oy0 = 8 * (tid >> 5);
    const int xq = px0 + 3 + 4 * lane;                           // image column of the lane's first score column
    const int xlo = T.xlo, xhi = T.xhi;                          // key-point columns of this tile inside the border
    uint2 xmask;
    xmask.x = ((xq >= xlo && xq < xhi) ? 0x0000ffffu : 0u) | ((xq + 1 >= xlo && xq + 1 < xhi) ? 0xffff0000u : 0u);
    xmask.y = ((xq + 2 >= xlo && xq + 2 < xhi) ? 0x0000ffffu : 0u) | ((xq + 3 >= xlo && xq + 3 < xhi) ? 0xffff0000u : 0u);
    const int oy_lo = T.oy_lo, oy_hi = T.oy_hi;
    if (oy0 <= oy_hi && oy0 + 7 >= oy_lo) {                      // warp-uniform
        const uint32_t* col = sc + 2 * lane;
        const uint32_t lt = (1u << lane) - 1u, bias = (uint32_t)(thr - 1);
        uint2 cP = make_uint2(0u, 0u), hP = cP, hPP = cP, lrP = cP;
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const int r = oy0 + k;                               // score row
            const uint2 c = *reinterpret_cast<const uint2*>(col + r * kFastScP);
            const uint32_t cp = __shfl_up_sync(0xffffffffu, c.y, 1), cn = __shfl_down_sync(0xffffffffu, c.x, 1);
            // lane 0 / 31 get their own word back: they only feed columns outside the key-point range (masked below)
            const uint32_t l0 = __funnelshift_r(cp, c.x, 16), m01 = __funnelshift_r(c.x, c.y, 16), r1 = __funnelshift_r(c.y, cn, 16);
            uint2 lr, hh;
            lr.x = __vmaxu2(l0, m01); lr.y = __vmaxu2(m01, r1);
            hh.x = __vmaxu2(lr.x, c.x); hh.y = __vmaxu2(lr.y, c.y);
            const int oy = r - 2;                                // output row of the PREVIOUS score row
            if (k >= 2 && oy >= oy_lo && oy <= oy_hi) {          // warp-uniform
                const uint32_t n0 = __vimax3_u16x2(hPP.x, hh.x, lrP.x), n1 = __vimax3_u16x2(hPP.y, hh.y, lrP.y);
                // v > n  <=>  max(v, n) != n; the differences are < 256 per half, so "+ 0xff >> 8" turns each half into its nonzero flag
                const uint32_t g0 = __vmaxu2(cP.x & xmask.x, n0) ^ n0, g1 = __vmaxu2(cP.y & xmask.y, n1) ^ n1;
                // bit 0 / 16 / 1 / 17 = pixel 0 / 1 / 2 / 3 of the quad survives (at most two of them: they are not adjacent)
                const uint32_t keep = (((g0 + 0x00ff00ffu) >> 8) & 0x00010001u) | (((g1 + 0x00ff00ffu) >> 7) & 0x00020002u);
                const uint32_t b1 = __ballot_sync(0xffffffffu, keep != 0);
                if (b1 != 0) {
                    const uint32_t b2 = __ballot_sync(0xffffffffu, (keep & (keep - 1u)) != 0);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_n, __popc(b1) + __popc(b2));
                    base = __shfl_sync(0xffffffffu, base, 0) + __popc(b1 & lt) + __popc(b2 & lt);
                    const uint32_t pos = (uint32_t)((py0 + 4 + oy) * L.w + xq);
                    if (keep & 0x00000001u) list[base++] = (((cP.x & 0xffffu) + bias) << 24) | pos;
                    if (keep & 0x00010000u) list[base++] = (((cP.x >> 16) + bias) << 24) | (pos + 1);
                    if (keep & 0x00000002u) list[base++] = (((cP.y & 0xffffu) + bias) << 24) | (pos + 2);
                    if (keep & 0x00020000u) list[base++] = (((cP.y >> 16) + bias) << 24) | (pos + 3);
                }
            }
            hPP = hP; hP = hh; lrP = lr; cP = c;
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;        // uniform: s_n is shared
    if (tid == 0) s_base = atomicAdd(b.cand_count + f * kMaxLevels + l, n);
    __syncthreads();
    uint32_t* cand = b.cand + (size_t)f * b.cand_stride + L.cand_off;
    for (int i = tid; i < n; i += 256) {
        const int slot = s_base + i;
        if (slot < (int)L.cand_cap) cand[slot] = list[i];
    }
}

template <int NP, int NX, int NQ>
__global__ void __launch_bounds__(256) k_fast(const __grid_constant__ OrbGeom g, const OrbBuffers b, const FastTile* __restrict__ tiles)
{
    extern __shared__ __align__(128) uint8_t fast_smem[];
    uint16_t* A = reinterpret_cast<uint16_t*>(fast_smem);
    uint32_t* sc = reinterpret_cast<uint32_t*>(A + 2 * kFastAsOff);            // [kFastSH][kFastScP], scores as u16 pairs
    __shared__ int s_n, s_base;
    __shared__ uint16_t queue[8 * 64];                                          // sparse-mode survivors, 64 per warp
    const int f = blockIdx.y;
    const FastTile T = fast_tile_load(tiles, blockIdx.x);
    const int l = T.l, px0 = T.px0, py0 = T.py0;
    const LevelGeom& L = g.lv[l];
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);

    // stage the pixels the scored rectangle touches: 16-byte global loads (a thread = 16 pixels of one row), all issued before the
    // first use, then widened to 16-bit elements twice. Rows are inside the image by construction; a 16-pixel group left of the
    // image or running past the row pitch is loaded word by word with zero fill.
    const int g_lo = T.g_lo, g_hi = T.g_hi, sy_lo = T.sy_lo;
    const int n_rows = T.nrows > 0 && T.npx > 0 ? T.nrows + 6 : 0;
    const bool wide = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)pitch) & 15) == 0;
    constexpr int kStageIters = (kFastPH * 8 + 255) / 256;
    uint4 pv[kStageIters];
#pragma unroll
    for (int it = 0; it < kStageIters; it++) {
        const int i = it * 256 + threadIdx.x, ry = i >> 3, grp = i & 7;
        pv[it] = make_uint4(0u, 0u, 0u, 0u);
        if (ry < n_rows && grp >= g_lo && grp <= g_hi) {
            const int x = px0 + 16 * grp;
            const uint8_t* p = img + (size_t)(py0 + sy_lo + ry) * pitch + x;
            if (wide && x >= 0 && x + 16 <= pitch) pv[it] = __ldg(reinterpret_cast<const uint4*>(p));
            else {
                if (x >= 0 && x + 4 <= pitch) pv[it].x = __ldg(reinterpret_cast<const uint32_t*>(p));
                if (x + 4 >= 0 && x + 8 <= pitch) pv[it].y = __ldg(reinterpret_cast<const uint32_t*>(p + 4));
                if (x + 8 >= 0 && x + 12 <= pitch) pv[it].z = __ldg(reinterpret_cast<const uint32_t*>(p + 8));
                if (x + 12 >= 0 && x + 16 <= pitch) pv[it].w = __ldg(reinterpret_cast<const uint32_t*>(p + 12));
            }
        }
    }
    fast_zero_scores(sc);
#pragma unroll
    for (int it = 0; it < kStageIters; it++) {
        const int i = it * 256 + threadIdx.x, ry = i >> 3, grp = i & 7;
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, pv[it].x, 1);        // first pixel of the next group of the row (group 7: not needed)
        if (ry < n_rows && grp >= g_lo && grp <= g_hi) fast_widen16(pv[it], nxt, A + (sy_lo + ry) * kFastP16 + 16 * grp);
    }
    fast_tile_compute<NP, NX, NQ>(g, b, L, f, T, A, sc, queue, s_n, s_base);
}

// TMA variant: the pixel tile is fetched by ONE cp.async.bulk.tensor per CTA (3-D tensor map per level: x, y, frame; out-of-image
// bytes arrive as zeros, so there is no address or bounds logic in the kernel) into the shared memory that later holds the score
// tile, and is widened from there. The tile origin is a multiple of 16 bytes in x, as the TMA engine requires of the innermost
// coordinate (measured on B200: any other start raises an illegal-instruction fault, tools/tma_probe2.cu).
struct FastMaps { CUtensorMap m[kMaxLevels]; };
constexpr int kFastRawBytes = kFastPH * kFastPW;
static_assert(kFastRawBytes <= kFastSH * kFastScP * 4, "the raw tile is landed in the score tile's shared memory");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
                 "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, void* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

template <int NP, int NX, int NQ>
__global__ void __launch_bounds__(256) k_fast_tma(const __grid_constant__ OrbGeom g, const OrbBuffers b, const __grid_constant__ FastMaps maps,
                                                  const FastTile* __restrict__ tiles)
{
    extern __shared__ __align__(128) uint8_t fast_smem[];
    uint16_t* A = reinterpret_cast<uint16_t*>(fast_smem);
    uint32_t* sc = reinterpret_cast<uint32_t*>(A + 2 * kFastAsOff);
    uint8_t* raw = reinterpret_cast<uint8_t*>(sc);                              // [kFastPH][kFastPW] bytes, written by the TMA engine
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int s_n, s_base;
    __shared__ uint16_t queue[8 * 64];
    const int f = blockIdx.y;
    const FastTile T = fast_tile_load(tiles, blockIdx.x);
    const int l = T.l, px0 = T.px0, py0 = T.py0;
    const LevelGeom& L = g.lv[l];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, kFastRawBytes);
        tma_load_3d(raw, &maps.m[l], px0, py0, f, &bar);
    }
    const int g_lo = T.g_lo, g_hi = T.g_hi, sy_lo = T.sy_lo;
    const int n_rows = T.nrows > 0 && T.npx > 0 ? T.nrows + 6 : 0;
    __syncthreads();                                   // the barrier is initialised before anyone polls it
    mbar_wait(&bar, 0u);
    constexpr int kStageIters = (kFastPH * 8 + 255) / 256;
#pragma unroll
    for (int it = 0; it < kStageIters; it++) {
        const int i = it * 256 + threadIdx.x, ry = i >> 3, grp = i & 7;
        const bool on = ry < n_rows && grp >= g_lo && grp <= g_hi;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (on) v = *reinterpret_cast<const uint4*>(raw + (sy_lo + ry) * kFastPW + 16 * grp);
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
        if (on) fast_widen16(v, nxt, A + (sy_lo + ry) * kFastP16 + 16 * grp);
    }
    __syncthreads();                                   // the raw bytes are consumed: their memory becomes the score tile
    fast_zero_scores(sc);
    fast_tile_compute<NP, NX, NQ>(g, b, L, f, T, A, sc, queue, s_n, s_base);
}

// ------------------------------------------------------------------------------------------------ K3: selection
// ref OpenCVModified.cpp:571-617 (RetainBestFeatures) and :144-360 (AdaptiveNonMaximalSuppresion), canonical order.
// One CTA per (level, frame). Keys are unique 64-bit words, sorted descending:
//   ANMS   : r << 32 | score << 24 | (0xFFFFFF - raster)     (r desc, strength desc, raster asc)
//   raster : (0xFFFFFF - raster) << 32 | score << 24 | (0xFFFFFF - raster)
// Bitonic sort, descending, of npow2 (a power of two) keys by the whole CTA. A compare-exchange distance below 32 stays inside a warp
// (element i sits in lane i % 32 because the block size is a multiple of 32): those stages run on registers through shuffles, with no
// barrier; only the distances of 32 and more go through the array. For 1024 keys that is 21 barriers instead of 55 (the old all-in-memory
// version spent 25 000 clocks here for a one-frame call: 460 per stage).
__device__ void bitonic_sort_desc(unsigned long long* keys, int npow2)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (npow2 + nt - 1) / nt;
    if (npow2 < 64 || (nt & 31)) {                   // tiny arrays (and odd block sizes): the plain all-in-memory network
        for (int k = 2; k <= npow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npow2; i += nt) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = keys[i], c = keys[ixj];
                        const bool desc = (i & k) == 0;
                        if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        return;
    }
    // register stages j = jtop, jtop / 2, .. 1 of merge size k on the elements this thread holds (each element's partners at distances
    // below 32 are in its own warp and belong to the same round m, so an element is loaded, exchanged and stored on its own)
    auto reg_stages = [&](int k, int jtop) {
        for (int m = 0; m < per; m++) {
            const int i = tid + m * nt;
            unsigned long long a = i < npow2 ? keys[i] : 0ull;
            for (int kk = (k <= 32) ? 2 : k; kk <= k; kk <<= 1) {        // k <= 32: the whole prefix of the network at once
                for (int j = min(kk >> 1, jtop); j > 0; j >>= 1) {
                    const unsigned long long c = __shfl_xor_sync(0xffffffffu, a, j);
                    const bool take_max = ((i & kk) == 0) == ((i & j) == 0);
                    a = take_max ? (a > c ? a : c) : (a < c ? a : c);
                }
            }
            if (i < npow2) keys[i] = a;
        }
        __syncthreads();
    };
    // merge sizes 2 .. 32 entirely in registers
    reg_stages(min(32, npow2), min(16, npow2 >> 1));
    for (int k = 64; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j >= 32; j >>= 1) {
            for (int i = tid; i < npow2; i += nt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
        reg_stages(k, 16);
    }
}

// LANES: eight lanes per key point in the ANMS ring search (the latency variant, for calls of a few frames); false: one thread per key
// point (the throughput variant, for batches)
template <bool LANES>
__global__ void __launch_bounds__(kSelThreads, LANES ? 1 : 2) k_select(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int hist[256];
    __shared__ int s_vals[16];          // 0:K 1:stop 2:minX 3:maxX 4:minY 5:maxY 6:minScore 7:R 8:mcd2 9:counter
    __shared__ float s_rf;
    __shared__ int cellStart[kMaxCells + 1];
    __shared__ int cellFill[kMaxCells];

    const int l = blockIdx.x, f = blockIdx.y;
    const LevelGeom& L = g.lv[l];
    int n = b.cand_count[f * kMaxLevels + l];
    if (n > (int)L.cand_cap) { if (threadIdx.x == 0) atomicOr(b.status + f, 1); n = (int)L.cand_cap; }
    const uint32_t* cand = b.cand + (size_t)f * b.cand_stride + L.cand_off;
    uint32_t* sel = b.sel + (size_t)f * b.sel_stride + L.sel_off;
    int* selCount = b.sel_count + f * kMaxLevels + l;
    const int nKeep = L.nfeat;
    const int tid = threadIdx.x, nt = blockDim.x;
#ifdef MAGE_SELECT_TIMERS
    long long tmr[12]; int tk = 0;
#define SELT() do { __syncthreads(); if (tid == 0 && tk < 12) tmr[tk++] = clock64(); } while (0)
#else
#define SELT() do {} while (0)
#endif
    SELT();

    unsigned long long* keys;
    uint32_t *kept, *cxy;
    uint8_t* csc;
    auto bind = [&](int K) {
        if (K <= kSelSmemItems) {
            keys = reinterpret_cast<unsigned long long*>(smem_raw);
            kept = reinterpret_cast<uint32_t*>(smem_raw + kSelSmemItems * 8);
            cxy = kept + kSelSmemItems;
            csc = reinterpret_cast<uint8_t*>(cxy + kSelSmemItems);
        } else {
            keys = b.keys + (size_t)f * b.key_stride + L.key_off;
            kept = b.kept + (size_t)f * b.cand_stride + L.cand_off;
            cxy = b.cxy + (size_t)f * b.cand_stride + L.cand_off;
            csc = b.csc + (size_t)f * b.cand_stride + L.cand_off;
        }
    };

    if (n <= nKeep) {
        // no suppression: keep FAST's raster order (ref :721 branch not taken)
        bind(n);
        int np2 = 1; while (np2 < n) np2 <<= 1;
        for (int i = tid; i < np2; i += nt) {
            unsigned long long key = 0;
            if (i < n) { uint32_t c = cand[i]; uint32_t inv = 0xFFFFFFu - (c & 0xFFFFFFu); key = ((unsigned long long)inv << 32) | (c & 0xFF000000u) | inv; }
            keys[i] = key;
        }
        __syncthreads();
        bitonic_sort_desc(keys, np2);
        for (int i = tid; i < n; i += nt) {
            uint32_t lo = (uint32_t)keys[i];
            sel[i] = (lo & 0xFF000000u) | (0xFFFFFFu - (lo & 0xFFFFFFu));
        }
        if (tid == 0) *selCount = n;
        return;
    }

    // ---- RetainBestFeatures: histogram of scores, whole-bin cut
    SELT();
    for (int i = tid; i < 256; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) atomicAdd(&hist[cand[i] >> 24], 1);
    __syncthreads();
    SELT();
    // suffix sums cum[i] = sum_{j >= i} hist[j] by the first 256 threads (Hillis-Steele in shared memory), then the two
    // thresholds of RetainBestFeatures as max-reductions -- replaces two serial 256-bin walks by one thread
    __shared__ int cum[256];
    if (tid < 256) cum[tid] = hist[tid];
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        int v = 0;
        if (tid < 256 && tid + o < 256) v = cum[tid + o];
        __syncthreads();
        if (tid < 256) cum[tid] += v;
        __syncthreads();
    }
    if (tid == 0) { s_vals[10] = -1; s_vals[11] = -1; s_vals[12] = 256; }
    __syncthreads();
    const int minThreshold = g.fast_threshold;
    if (tid < 256 && tid >= minThreshold && cum[tid] >= nKeep) atomicMax(&s_vals[10], tid);       // first loop of ref :592-600
    __syncthreads();
    {
        const int minNumThreshold = s_vals[10] >= 0 ? s_vals[10] : minThreshold;
        const int maxNum = (int)__fmul_rn((float)nKeep, g.feature_factor);
        const int lo = max((int)__fmul_rn((float)minNumThreshold, g.feature_strength), minThreshold);
        if (tid < 256 && tid >= lo && cum[tid] >= maxNum) atomicMax(&s_vals[11], tid);                // second loop of ref :605-613
        __syncthreads();
        const int stop_ = s_vals[11] >= 0 ? s_vals[11] : lo;
        if (tid < 256 && tid >= stop_ && hist[tid] > 0) atomicMin(&s_vals[12], tid);
        __syncthreads();
        if (tid == 0) {
            s_vals[0] = stop_ < 256 ? cum[stop_] : 0; s_vals[1] = stop_; s_vals[6] = min(s_vals[12], 255);
            s_vals[2] = INT_MAX; s_vals[3] = INT_MIN; s_vals[4] = INT_MAX; s_vals[5] = INT_MIN; s_vals[9] = 0;
        }
    }
    __syncthreads();
    const int K = s_vals[0], stop = s_vals[1];
    if (K < nKeep) {
        // RetainBestFeatures kept fewer than n_l (feature_strength > 1 lifts the cut above the n_l-th score, or feature_factor < 1):
        // AdaptiveNonMaximalSuppresion returns early (ref :181-184) and the K survivors stay in RetainBest's (canonical: raster) order
        bind(n);
        int np2 = 1; while (np2 < n) np2 <<= 1;
        for (int i = tid; i < np2; i += nt) {
            unsigned long long key = 0;
            if (i < n) {
                const uint32_t c = cand[i];
                if ((int)(c >> 24) >= stop) { const uint32_t inv = 0xFFFFFFu - (c & 0xFFFFFFu); key = ((unsigned long long)inv << 32) | (c & 0xFF000000u) | inv; }
            }
            keys[i] = key;
        }
        __syncthreads();
        bitonic_sort_desc(keys, np2);
        for (int i = tid; i < K; i += nt) {
            const uint32_t lo = (uint32_t)keys[i];
            sel[i] = (lo & 0xFF000000u) | (0xFFFFFFu - (lo & 0xFFFFFFu));
        }
        if (tid == 0) *selCount = K;
        return;
    }
    SELT();
    bind(K);
    {
        int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
        for (int i = tid; i < n; i += nt) {
            uint32_t c = cand[i];
            if ((int)(c >> 24) >= stop) {
                int j = atomicAdd(&s_vals[9], 1);
                kept[j] = c;
                int pos = c & 0xFFFFFF, y = pos / L.w, x = pos - y * L.w;
                mnx = min(mnx, x); mxx = max(mxx, x); mny = min(mny, y); mxy = max(mxy, y);
            }
        }
        if (mnx != INT_MAX) { atomicMin(&s_vals[2], mnx); atomicMax(&s_vals[3], mxx); atomicMin(&s_vals[4], mny); atomicMax(&s_vals[5], mxy); }
    }
    SELT();
    const int numX = g.num_cells_x, numY = g.num_cells_y, nCells = numX * numY;
    for (int i = tid; i < nCells; i += nt) { cellStart[i] = 0; cellFill[i] = 0; }
    __syncthreads();
    const int minX = s_vals[2], maxX = s_vals[3], minY = s_vals[4], maxY = s_vals[5];
    if (tid == 0) {
        // ref :207-214 robustness factor (float32, no contraction), :259 globalMaxR2, :261-266 minCellDelta2
        const float thrf = (float)g.fast_threshold;
        float hi = __fsub_rn((float)g.strong_response, thrf);
        float val = fminf(hi, fmaxf(0.0f, __fsub_rn((float)s_vals[6], thrf)));
        float range = fmaxf(0.0f, __fsub_rn(g.max_rf, g.min_rf));
        s_rf = __fsub_rn(g.max_rf, __fmul_rn(__fdiv_rn(val, (float)(g.strong_response - g.fast_threshold)), range));
        s_vals[7] = (int)(((double)(maxX - minX)) * ((double)(maxY - minY)) / (double)nKeep);
        int dx = max((maxX - minX) / numX, 1), dy = max((maxY - minY) / numY, 1);
        s_vals[8] = min(dx, dy) * min(dx, dy);
    }
    // cell histogram; (cell, x | y << 16) of every kept key point is parked in the key array (free until the radii are written) so that the
    // scatter pass below does not repeat the three integer divisions
    for (int j = tid; j < K; j += nt) {
        int pos = kept[j] & 0xFFFFFF, y = pos / L.w, x = pos - y * L.w;
        int cell = ((y - minY) * numY / (maxY + 1 - minY)) * numX + (x - minX) * numX / (maxX + 1 - minX);
        atomicAdd(&cellStart[cell], 1);
        keys[j] = ((unsigned long long)(uint32_t)cell << 32) | (uint32_t)x | ((uint32_t)y << 16);
    }
    __syncthreads();
    {   // exclusive scan over the <= 2048 cells by the whole CTA: a few cells per thread, warp scan, scan of the warp totals
        __shared__ int wsum[32];
        const int per = (nCells + nt - 1) / nt, beg = tid * per, end = min(beg + per, nCells), lane = tid & 31, warp = tid >> 5;
        int sum = 0;
        for (int i = beg; i < end; i++) sum += cellStart[i];
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < (nt >> 5) ? wsum[lane] : 0;
            int wi = w;
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
            wsum[lane] = wi - w;                              // exclusive warp offsets
            if (lane == 31) cellStart[nCells] = wi;
        }
        __syncthreads();
        int run = wsum[warp] + incl - sum;
        for (int i = beg; i < end; i++) { const int c = cellStart[i]; cellStart[i] = run; run += c; }
    }
    __syncthreads();
    for (int j = tid; j < K; j += nt) {
        const unsigned long long cz = keys[j];
        const int cell = (int)(cz >> 32);
        const int slot = cellStart[cell] + atomicAdd(&cellFill[cell], 1);
        cxy[slot] = (uint32_t)cz;
        csc[slot] = (uint8_t)(kept[j] >> 24);
    }
    __syncthreads();
    SELT();
    // ---- ANMS radii: literal ring search of ref :268-326, one thread per keypoint
    const int R = s_vals[7], mcd2 = s_vals[8];
    const float rf = s_rf;
    int np2 = 1; while (np2 < K) np2 <<= 1;
    // how far the ring search of this level can reach: rings d with (d - 1)^2 mcd2 < R
    int dmax = 1;
    while (dmax * dmax * mcd2 < R && dmax < numX + numY) dmax++;
    if (LANES && dmax >= 5) {
        // Eight lanes per key point (levels whose searches reach five rings or more: the small ones; below that the plain loop is faster): ring d of the cell grid (one cell for d = 0, else the 8 d cells of the square's border) is dealt out
        // over the lanes, the minimum squared distance to a stronger key point is reduced over the group after every ring -- the reference
        // tests its stopping rule once per ring too and only ever takes minima inside one, so the radius is the same whatever the order of
        // the cells. On the small levels a key point walks up to 17 x 17 mostly empty cells: one thread per key point is a chain of
        // 36 000 clocks there (in-kernel timers, -DMAGE_SELECT_TIMERS), the critical path of a one-frame call.
        const int gl = tid & 7, ngrp = nt >> 3;
        for (int base = 0; base < np2; base += ngrp) {
            const int slot = base + (tid >> 3);
            const bool live = slot < K;
            int x = 0, y = 0, sc = 0, cx = 0, cy = 0;
            if (live) {
                const uint32_t me = cxy[slot];
                x = me & 0xFFFF; y = me >> 16; sc = csc[slot];
                cx = (x - minX) * numX / (maxX + 1 - minX); cy = (y - minY) * numY / (maxY + 1 - minY);
            }
            const float s = __fadd_rn(__fmul_rn((float)sc, rf), 0.002f);           // strength >= 0 always (FAST scores)
            int minR2 = R;
            bool go = live;
            for (int d = 0; ; d++) {
                go = go && max(0, d - 1) * max(0, d - 1) * mcd2 < minR2 && d <= numX + numY;      // beyond numX + numY there are no cells
                if (!__any_sync(0xffffffffu, go)) break;
                if (go) {
                    // work items of ring d: the two full rows at cy -+ d, each ONE range of slots (the cells of a grid row are neighbours
                    // in the sorted arrays), then the two end cells of every row in between
                    const int items = d == 0 ? 1 : 4 * d;
                    const int x0 = max(cx - d, 0), x1 = min(cx + d, numX - 1);
                    for (int i = gl; i < items; i += 8) {
                        int c0 = 0, c1 = 0;
                        if (i < 2) {
                            const int cYY = i ? cy + d : cy - d;
                            if (cYY >= 0 && cYY < numY) { c0 = cellStart[cYY * numX + x0]; c1 = cellStart[cYY * numX + x1 + 1]; }
                        } else {
                            const int k = i - 2, cYY = cy - d + 1 + (k >> 1), cXX = (k & 1) ? cx + d : cx - d;
                            if (cXX >= 0 && cXX < numX && cYY >= 0 && cYY < numY) { c0 = cellStart[cYY * numX + cXX]; c1 = cellStart[cYY * numX + cXX + 1]; }
                        }
                        for (int o = c0; o < c1; o++) {
                            if ((float)csc[o] > s) {
                                const uint32_t ot = cxy[o];
                                const int ddx = x - (int)(ot & 0xFFFF), ddy = y - (int)(ot >> 16);
                                minR2 = min(minR2, ddx * ddx + ddy * ddy);
                            }
                        }
                    }
                }
                minR2 = min(minR2, __shfl_xor_sync(0xffffffffu, minR2, 1));
                minR2 = min(minR2, __shfl_xor_sync(0xffffffffu, minR2, 2));
                minR2 = min(minR2, __shfl_xor_sync(0xffffffffu, minR2, 4));
            }
            if (gl == 0 && slot < np2) {
                unsigned long long key = 0;
                if (live) {
                    const uint32_t inv = 0xFFFFFFu - (uint32_t)(y * L.w + x);
                    key = ((unsigned long long)(uint32_t)minR2 << 32) | ((uint32_t)sc << 24) | inv;
                }
                keys[slot] = key;
            }
        }
    } else {
    for (int slot = tid; slot < np2; slot += nt) {
        unsigned long long key = 0;
        if (slot < K) {
            const uint32_t me = cxy[slot];
            const int x = me & 0xFFFF, y = me >> 16, sc = csc[slot];
            const int cx = (x - minX) * numX / (maxX + 1 - minX), cy = (y - minY) * numY / (maxY + 1 - minY);
            const float s = __fadd_rn(__fmul_rn((float)sc, rf), 0.002f);       // strength >= 0 always (FAST scores)
            int minR2 = R;
            for (int d = 0; max(0, d - 1) * max(0, d - 1) * mcd2 < minR2; d++) {
                if (d > numX + numY) break;                                     // no cells out there
                // ring d: the full rows at cy -+ d are ONE range of slots each (the cells of a grid row are neighbours in the sorted
                // arrays), the rows in between contribute their two end cells
                auto scan = [&](int c0, int c1) {
                    for (int o = c0; o < c1; o++) {
                        if ((float)csc[o] > s) {
                            const uint32_t ot = cxy[o];
                            const int ddx = x - (int)(ot & 0xFFFF), ddy = y - (int)(ot >> 16);
                            minR2 = min(minR2, ddx * ddx + ddy * ddy);
                        }
                    }
                };
                const int x0 = max(cx - d, 0), x1 = min(cx + d, numX - 1);
                if (cy - d >= 0) scan(cellStart[(cy - d) * numX + x0], cellStart[(cy - d) * numX + x1 + 1]);
                if (d > 0 && cy + d < numY) scan(cellStart[(cy + d) * numX + x0], cellStart[(cy + d) * numX + x1 + 1]);
                for (int cYY = max(cy - d + 1, 0); cYY <= min(cy + d - 1, numY - 1); cYY++) {
                    if (cx - d >= 0) scan(cellStart[cYY * numX + cx - d], cellStart[cYY * numX + cx - d + 1]);
                    if (cx + d < numX) scan(cellStart[cYY * numX + cx + d], cellStart[cYY * numX + cx + d + 1]);
                }
            }
            uint32_t inv = 0xFFFFFFu - (uint32_t)(y * L.w + x);
            key = ((unsigned long long)(uint32_t)minR2 << 32) | ((uint32_t)sc << 24) | inv;
        }
        keys[slot] = key;
    }
    }
    __syncthreads();
    SELT();
    bitonic_sort_desc(keys, np2);
    SELT();
    for (int i = tid; i < nKeep; i += nt) {
        uint32_t lo = (uint32_t)keys[i];
        sel[i] = (lo & 0xFF000000u) | (0xFFFFFFu - (lo & 0xFFFFFFu));
    }
    if (tid == 0) *selCount = nKeep;
#ifdef MAGE_SELECT_TIMERS
    SELT();
    if (tid == 0 && f == 0) printf("[k_select] level %d n %d K %d nKeep %d clk: start %lld hist %lld cut %lld compact %lld cells %lld rings %lld sort %lld out %lld\n", l, n, K, nKeep,
                                   tmr[1] - tmr[0], tmr[2] - tmr[1], tmr[3] - tmr[2], tmr[4] - tmr[3], tmr[5] - tmr[4], tmr[6] - tmr[5], tmr[7] - tmr[6], tmr[8] - tmr[7]);
#endif
}

// ------------------------------------------------------------------------------------------------ K5: Gaussian blur
// ref OpenCVModified.cpp:853-865 -> cv::GaussianBlur(k x k, sigma 2, REFLECT_101), bit-exact Q8.8 path.
constexpr int kBlurOW = 128, kBlurOH = 32, kBlurMaxR = 7;
constexpr int kBlurIW = kBlurOW + 2 * kBlurMaxR + 2;    // 144
constexpr int kBlurIH = kBlurOH + 2 * kBlurMaxR;        // 46

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    __shared__ uint8_t in[kBlurIH * kBlurIW];
    __shared__ uint16_t mid[kBlurIH * kBlurOW];
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_base) l++;
    const LevelGeom& L = g.lv[l];
    const int t = blockIdx.x - L.blur_tile_base;
    const int x0 = (t % L.blur_tiles_x) * kBlurOW, y0 = (t / L.blur_tiles_x) * kBlurOH;
    const int r = g.ksize / 2, ks = g.ksize;
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);
    const int iw = kBlurOW + 2 * r, ih = kBlurOH + 2 * r;
    for (int i = threadIdx.x; i < ih * iw; i += blockDim.x) {
        int ry = i / iw, rx = i - ry * iw;
        int y = reflect101(min(y0 + ry - r, L.h + r), L.h), x = reflect101(min(x0 + rx - r, L.w + r), L.w);
        in[ry * kBlurIW + rx] = img[(size_t)y * pitch + x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ih * kBlurOW; i += blockDim.x) {
        int ry = i / kBlurOW, ox = i - ry * kBlurOW;
        unsigned acc = 0;
        for (int k = 0; k < ks; k++) acc += (unsigned)g.gk[k] * in[ry * kBlurIW + ox + k];
        mid[ry * kBlurOW + ox] = (uint16_t)acc;
    }
    __syncthreads();
    uint8_t* out = blur_ptr(g, b, f, l);
    for (int i = threadIdx.x; i < kBlurOH * (kBlurOW / 4); i += blockDim.x) {
        int oy = i / (kBlurOW / 4), ox = (i - oy * (kBlurOW / 4)) * 4;
        int y = y0 + oy, x = x0 + ox;
        if (y >= L.h || x >= L.w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned acc = 0;
            for (int k = 0; k < ks; k++) acc += (unsigned)g.gk[k] * mid[(oy + k) * kBlurOW + ox + j];
            packed |= ((acc + 32768u) >> 16) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(out + (size_t)y * L.pitch + x) = packed;
    }
}

// 7x7 fast path (the reference default GaussianKernelSize and every BASELINE config). The Q8.8 convolution is exact
// integer arithmetic (sum of products <= 2^24, one rounding at the end), so the horizontal taps run as two IDP.4A per pixel on
// funnel-shifted words and the vertical taps as IMADs over a register column; no shared memory, 32-bit coalesced loads/stores.
constexpr int kBlur7Rows = 16;      // output rows per thread (22 source rows incl. the +-3 halo)

__device__ __noinline__ uint32_t load_word_reflect_slow(const uint8_t* row, int x, int w)
{
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v |= (uint32_t)row[reflect101(x + i, w)] << (8 * i);
    return v;
}
__device__ __forceinline__ uint32_t load_word_reflect(const uint8_t* row, int x, int w)
{
    if (x >= 0 && x + 4 <= w) return *reinterpret_cast<const uint32_t*>(row + x);
    return load_word_reflect_slow(row, x, w);
}

__global__ void __launch_bounds__(256) k_blur7(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_base) l++;
    const LevelGeom& L = g.lv[l];
    const int t = blockIdx.x - L.blur_tile_base;
    const int x = (t % L.blur_tiles_x) * 128 + (threadIdx.x & 31) * 4;
    const int y0 = (t / L.blur_tiles_x) * (8 * kBlur7Rows) + (threadIdx.x >> 5) * kBlur7Rows;
    if (x < 4 || x + 8 > L.w || y0 >= L.h) return;       // edge columns (needing REFLECT_101 in x) are done by k_blur7_edges
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);
    const uint32_t K0 = (uint32_t)g.gk[0] | ((uint32_t)g.gk[1] << 8) | ((uint32_t)g.gk[2] << 16) | ((uint32_t)g.gk[3] << 24);
    const uint32_t K1 = (uint32_t)g.gk[4] | ((uint32_t)g.gk[5] << 8) | ((uint32_t)g.gk[6] << 16);
    const unsigned k0 = g.gk[0], k1 = g.gk[1], k2 = g.gk[2], k3 = g.gk[3];      // symmetric kernel: k[6-i] == k[i]
    int T[kBlur7Rows + 6][4];
#pragma unroll
    for (int i = 0; i < kBlur7Rows + 6; i++) {
        // branch-free REFLECT_101 of the row index (levels are >= 8 rows, the overshoot is at most 3 after the clamp), so that
        // all 66 loads of the thread can be issued back to back
        int sy = min(y0 + i - 3, L.h + 2);
        sy = sy < 0 ? -sy : sy;
        sy = sy >= L.h ? 2 * L.h - 2 - sy : sy;
        const uint8_t* row = img + (size_t)sy * pitch;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(row + x - 4), w1 = *reinterpret_cast<const uint32_t*>(row + x),
                       w2 = *reinterpret_cast<const uint32_t*>(row + x + 4);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t v0 = (j == 3) ? w1 : __funnelshift_r(w0, w1, 8 * (1 + j));     // bytes x-3+j .. x+j
            const uint32_t v1 = (j == 3) ? w2 : __funnelshift_r(w1, w2, 8 * (1 + j));     // bytes x+1+j .. x+4+j
            T[i][j] = (int)__dp4a(v1, K1, __dp4a(v0, K0, 0u));
        }
    }
    uint8_t* out = blur_ptr(g, b, f, l);
#pragma unroll
    for (int r = 0; r < kBlur7Rows; r++) {
        const int y = y0 + r;
        if (y < L.h) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned acc = 32768u + k0 * (unsigned)(T[r][j] + T[r + 6][j]) + k1 * (unsigned)(T[r + 1][j] + T[r + 5][j]) +
                                     k2 * (unsigned)(T[r + 2][j] + T[r + 4][j]) + k3 * (unsigned)T[r + 3][j];
                packed |= (acc >> 16) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(out + (size_t)y * L.pitch + x) = packed;
        }
    }
}

// Edge columns of the 7x7 blur: x in [0, 4) and [xr, w) with xr = the first column not covered by an interior thread.
// One thread per output pixel, plain 49-tap evaluation with REFLECT_101 (about 1.5 % of the pixels).
__global__ void __launch_bounds__(256) k_blur7_edges(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    const int f = blockIdx.z, l = blockIdx.y;
    if (l >= g.nlevels) return;
    const LevelGeom& L = g.lv[l];
    // interior threads cover x in [4, xr) where xr = 4 * floor((w - 8) / 4) + 4 ... i.e. the largest multiple of 4 with x + 8 <= w, plus 4
    const int last = ((L.w - 8) / 4) * 4;                       // last interior thread column (>= 4 when w >= 12)
    const int xr = (L.w >= 12) ? last + 4 : 0;                  // right strip starts here; whole row when the level is tiny
    const int nleft = (L.w >= 12) ? 4 : 0, nright = L.w - xr, ncols = nleft + nright;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols * L.h) return;
    const int y = i / ncols, c = i - y * ncols;
    const int x = (c < nleft) ? c : xr + (c - nleft);
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);
    int xs[7];
#pragma unroll
    for (int k = 0; k < 7; k++) xs[k] = reflect101(x + k - 3, L.w);          // column taps once per pixel
    unsigned acc = 32768u;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        int sy = y + j - 3;                                                   // branch-free row reflect (|overshoot| <= 3 < h)
        sy = sy < 0 ? -sy : sy;
        sy = sy >= L.h ? 2 * L.h - 2 - sy : sy;
        const uint8_t* row = img + (size_t)sy * pitch;
        unsigned t = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) t += (unsigned)g.gk[k] * row[xs[k]];
        acc += (unsigned)g.gk[j] * t;
    }
    blur_ptr(g, b, f, l)[(size_t)y * L.pitch + x] = (uint8_t)(acc >> 16);
}


// ------------------------------------------------------------------------------------------------ K5b: blur, float path
// cv::GaussianBlur of a SUBMATRIX source -- the reference blurs level ROIs of its packed pyramid buffer in place (ref :853-865),
// which OpenCV does not route to the fixed-point path above but through sepFilter2D with CV_32F kernels:
//   row    s = K[0]*x[0]; s = fma(K[k], x[k], s), k = 1 .. ksize-1          (uchar -> float, left to right)
//   column s = Kc*r[c];   s = fma(K[c+k], r[c+k] + r[c-k], s), k = 1 .. ksize/2, then cvRound + saturate (float -> uchar)
// evaluated with fused multiply-adds exactly like the stock OpenCV 4.13 build the parity tests compare with.
// Every operation is an explicitly rounded intrinsic, so the result is bit-identical.
__device__ __forceinline__ float u8f(uint32_t w, int byte)         // byte -> float without the quarter-rate I2F: 2^23 + x as bits, minus 2^23
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)byte)), 8388608.f);
}
__device__ __forceinline__ uint32_t f2u8(float s)
{
    return (uint32_t)min(max(__float2int_rn(s), 0), 255);
}

// FUSED = true: one FMA per multiply-add (an AVX2/FMA build of OpenCV, what the cv2 4.13 wheel runs); false: product and sum rounded
// separately (an SSE2-baseline build such as the reference's MSVC x64 one). Selected by mage_orb_set_blur_mode.
template <bool FUSED> __device__ __forceinline__ float blur_mad(float k, float v, float s)
{
    return FUSED ? __fmaf_rn(k, v, s) : __fadd_rn(s, __fmul_rn(k, v));
}

template <bool FUSED>
__global__ void __launch_bounds__(256) k_blurf(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    __shared__ uint8_t in[kBlurIH * kBlurIW];
    __shared__ float mid[kBlurIH * kBlurOW];
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_base) l++;
    const LevelGeom& L = g.lv[l];
    const int t = blockIdx.x - L.blur_tile_base;
    const int x0 = (t % L.blur_tiles_x) * kBlurOW, y0 = (t / L.blur_tiles_x) * kBlurOH;
    const int r = g.ksize / 2, ks = g.ksize;
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);
    const int iw = kBlurOW + 2 * r, ih = kBlurOH + 2 * r;
    for (int i = threadIdx.x; i < ih * iw; i += blockDim.x) {
        int ry = i / iw, rx = i - ry * iw;
        int y = reflect101(min(y0 + ry - r, L.h + r), L.h), x = reflect101(min(x0 + rx - r, L.w + r), L.w);
        in[ry * kBlurIW + rx] = img[(size_t)y * pitch + x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ih * kBlurOW; i += blockDim.x) {
        int ry = i / kBlurOW, ox = i - ry * kBlurOW;
        float s = __fmul_rn(g.gkf[0], (float)in[ry * kBlurIW + ox]);
        for (int k = 1; k < ks; k++) s = blur_mad<FUSED>(g.gkf[k <= r ? k : 2 * r - k], (float)in[ry * kBlurIW + ox + k], s);
        mid[ry * kBlurOW + ox] = s;
    }
    __syncthreads();
    uint8_t* out = blur_ptr(g, b, f, l);
    for (int i = threadIdx.x; i < kBlurOH * (kBlurOW / 4); i += blockDim.x) {
        int oy = i / (kBlurOW / 4), ox = (i - oy * (kBlurOW / 4)) * 4;
        int y = y0 + oy, x = x0 + ox;
        if (y >= L.h || x >= L.w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = __fmul_rn(g.gkf[r], mid[(oy + r) * kBlurOW + ox + j]);
            for (int k = 1; k <= r; k++) s = blur_mad<FUSED>(g.gkf[r - k], __fadd_rn(mid[(oy + r + k) * kBlurOW + ox + j], mid[(oy + r - k) * kBlurOW + ox + j]), s);
            packed |= f2u8(s) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(out + (size_t)y * L.pitch + x) = packed;
    }
}

// 7x7 float path, same work split as k_blur7: a thread owns 4 columns x kBlur7Rows rows, the row pass of its 22 source rows stays
// in registers (88 floats), no shared memory. Interior columns only; k_blur7f_edges does the REFLECT_101 columns.
template <bool FUSED>
__global__ void __launch_bounds__(256) k_blur7f(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_base) l++;
    const LevelGeom& L = g.lv[l];
    const int t = blockIdx.x - L.blur_tile_base;
    const int x = (t % L.blur_tiles_x) * 128 + (threadIdx.x & 31) * 4;
    const int y0 = (t / L.blur_tiles_x) * (8 * kBlur7Rows) + (threadIdx.x >> 5) * kBlur7Rows;
    if (x < 4 || x + 8 > L.w || y0 >= L.h) return;
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);
    const float k0 = g.gkf[0], k1 = g.gkf[1], k2 = g.gkf[2], k3 = g.gkf[3];
    float T[kBlur7Rows + 6][4];
#pragma unroll
    for (int i = 0; i < kBlur7Rows + 6; i++) {
        int sy = min(y0 + i - 3, L.h + 2);
        sy = sy < 0 ? -sy : sy;
        sy = sy >= L.h ? 2 * L.h - 2 - sy : sy;
        const uint8_t* row = img + (size_t)sy * pitch;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(row + x - 4), w1 = *reinterpret_cast<const uint32_t*>(row + x),
                       w2 = *reinterpret_cast<const uint32_t*>(row + x + 4);
        float v[10];                                           // pixels x-3 .. x+6
        v[0] = u8f(w0, 1); v[1] = u8f(w0, 2); v[2] = u8f(w0, 3);
        v[3] = u8f(w1, 0); v[4] = u8f(w1, 1); v[5] = u8f(w1, 2); v[6] = u8f(w1, 3);
        v[7] = u8f(w2, 0); v[8] = u8f(w2, 1); v[9] = u8f(w2, 2);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = __fmul_rn(k0, v[j]);
            s = blur_mad<FUSED>(k1, v[j + 1], s); s = blur_mad<FUSED>(k2, v[j + 2], s); s = blur_mad<FUSED>(k3, v[j + 3], s);
            s = blur_mad<FUSED>(k2, v[j + 4], s); s = blur_mad<FUSED>(k1, v[j + 5], s); s = blur_mad<FUSED>(k0, v[j + 6], s);
            T[i][j] = s;
        }
    }
    uint8_t* out = blur_ptr(g, b, f, l);
#pragma unroll
    for (int r = 0; r < kBlur7Rows; r++) {
        const int y = y0 + r;
        if (y < L.h) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float s = __fmul_rn(k3, T[r + 3][j]);
                s = blur_mad<FUSED>(k2, __fadd_rn(T[r + 4][j], T[r + 2][j]), s);
                s = blur_mad<FUSED>(k1, __fadd_rn(T[r + 5][j], T[r + 1][j]), s);
                s = blur_mad<FUSED>(k0, __fadd_rn(T[r + 6][j], T[r][j]), s);
                packed |= f2u8(s) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(out + (size_t)y * L.pitch + x) = packed;
        }
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(256) k_blur7f_edges(const __grid_constant__ OrbGeom g, const OrbBuffers b)
{
    const int f = blockIdx.z, l = blockIdx.y;
    if (l >= g.nlevels) return;
    const LevelGeom& L = g.lv[l];
    const int last = ((L.w - 8) / 4) * 4;
    const int xr = (L.w >= 12) ? last + 4 : 0;
    const int nleft = (L.w >= 12) ? 4 : 0, nright = L.w - xr, ncols = nleft + nright;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols * L.h) return;
    const int y = i / ncols, c = i - y * ncols;
    const int x = (c < nleft) ? c : xr + (c - nleft);
    int pitch;
    const uint8_t* img = level_ptr(g, b, f, l, pitch);
    int xs[7];
#pragma unroll
    for (int k = 0; k < 7; k++) xs[k] = reflect101(x + k - 3, L.w);
    const float kk[4] = {g.gkf[0], g.gkf[1], g.gkf[2], g.gkf[3]};
    float R[7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        int sy = y + j - 3;
        sy = sy < 0 ? -sy : sy;
        sy = sy >= L.h ? 2 * L.h - 2 - sy : sy;
        const uint8_t* row = img + (size_t)sy * pitch;
        float s = __fmul_rn(kk[0], (float)row[xs[0]]);
#pragma unroll
        for (int k = 1; k < 7; k++) s = blur_mad<FUSED>(kk[k <= 3 ? k : 6 - k], (float)row[xs[k]], s);
        R[j] = s;
    }
    float s = __fmul_rn(kk[3], R[3]);
    s = blur_mad<FUSED>(kk[2], __fadd_rn(R[4], R[2]), s);
    s = blur_mad<FUSED>(kk[1], __fadd_rn(R[5], R[1]), s);
    s = blur_mad<FUSED>(kk[0], __fadd_rn(R[6], R[0]), s);
    blur_ptr(g, b, f, l)[(size_t)y * L.pitch + x] = (uint8_t)f2u8(s);
}

// ------------------------------------------------------------------------------------------------ K4+K7: orientation + rBRIEF
// ref OpenCVModified.cpp:399-437 (ICAngles), :756-760 (rescale), :502-549 (ComputeOrbDescriptorsPrerotated),
// cv::fastAtan2 restated in SURVEY appendix A.1 (float32, no FMA contraction).
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// lanes span u = -hp..hp (hp <= 15 -> at most 31 columns); rows v are unrolled when hp is a compile-time constant
template <int HP>
__device__ __forceinline__ void orient_moments(const uint8_t* center, int pitch, int lane, const int* umax, int& m01, int& m10, int hp_rt = 0)
{
    const int hp = HP ? HP : hp_rt;
    const int u = lane - hp, au = u < 0 ? -u : u;
    if (HP && lane > 2 * hp) return;
    if (HP) {
        int val[2 * (HP ? HP : 1) + 1];
#pragma unroll
        for (int i = 0; i < 2 * HP + 1; i++) {
            const int v = i - HP, av = v < 0 ? -v : v;
            val[i] = (au <= umax[av]) ? (int)center[v * pitch + u] : 0;
        }
        int rowsum = 0;                                          // m10 = u * (column sum): one multiply per lane instead of one per row
#pragma unroll
        for (int i = 0; i < 2 * HP + 1; i++) { rowsum += val[i]; m01 += (i - HP) * val[i]; }
        m10 += u * rowsum;
    } else {
        for (int uu = lane; uu <= 2 * hp; uu += 32) {                // half patches above 15 need more than one column per lane
            const int u2 = uu - hp, au2 = u2 < 0 ? -u2 : u2;
            for (int v = -hp; v <= hp; v++) {
                const int av = v < 0 ? -v : v;
                if (au2 <= umax[av]) { const int x = center[v * pitch + u2]; m10 += u2 * x; m01 += v * x; }
            }
        }
    }
}

__global__ void __launch_bounds__(256, 5) k_orient_describe(const __grid_constant__ OrbGeom g, const OrbBuffers b,
                                                         mage_keypoint* __restrict__ out_kps, uint8_t* __restrict__ out_desc,
                                                         int* __restrict__ out_counts, int capacity)
{
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // output slot of this warp
    const int* selCount = b.sel_count + f * kMaxLevels;
    // level of output slot i: one coalesced load of the per-level counts + a warp prefix sum (kMaxLevels = 16 <= warp)
    const int cnt = lane < g.nlevels ? selCount[lane] : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < kMaxLevels; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, kMaxLevels - 1);
    const uint32_t lm = __ballot_sync(0xffffffffu, lane < g.nlevels && i < incl);
    const int l = lm ? __ffs(lm) - 1 : -1;
    const int j = i - __shfl_sync(0xffffffffu, incl - cnt, l < 0 ? 0 : l);
    if (i == 0 && lane == 0) out_counts[f] = min(total, capacity);
    if (l < 0 || i >= capacity) return;
    const LevelGeom& L = g.lv[l];
    const uint32_t c = b.sel[(size_t)f * b.sel_stride + L.sel_off + j];
    const int pos = c & 0xFFFFFF, score = c >> 24;
    const int y = pos / L.w, x = pos - y * L.w;

    float angle = 0.f;
    if (g.use_orientation) {
        int pitch;
        const uint8_t* img = level_ptr(g, b, f, l, pitch);
        const uint8_t* center = img + (size_t)y * pitch + x;
        int m01 = 0, m10 = 0;
        if (g.half_patch == 15) orient_moments<15>(center, pitch, lane, g.umax, m01, m10);
        else if (g.half_patch == 7) orient_moments<7>(center, pitch, lane, g.umax, m01, m10);
        else orient_moments<0>(center, pitch, lane, g.umax, m01, m10, g.half_patch);
        m01 = warp_reduce_sum(m01);
        m10 = warp_reduce_sum(m10);
        angle = fast_atan2_deg((float)m01, (float)m10);
    }
    const float ptx = __fmul_rn((float)x, L.scale), pty = __fmul_rn((float)y, L.scale);
    if (lane == 0) {
        mage_keypoint kp;
        kp.x = ptx; kp.y = pty;
        kp.size = __fmul_rn((float)g.patch, L.scale);
        kp.angle = angle; kp.response = (float)score; kp.octave = l; kp.class_id = -1;
        out_kps[(size_t)f * capacity + i] = kp;
    }
    // descriptor: lane = byte index; 8 tests of 4 int8 coordinates each = 32 contiguous pattern bytes per lane
    const float inv = __fdiv_rn(1.f, L.scale);
    const int bin = __float2int_rn(__fdiv_rn(angle, 12.f)) % 30;
    const int cyy = __float2int_rn(__fmul_rn(pty, inv)), cxx = __float2int_rn(__fmul_rn(ptx, inv));
    const uint8_t* base = (g.ksize > 1) ? blur_ptr(g, b, f, l) : nullptr;
    int pitch = L.pitch;
    if (!base) base = level_ptr(g, b, f, l, pitch);
    const uint8_t* ctr = base + (size_t)cyy * pitch + cxx;
    if (g.generic_pattern) {
        // ref :452-492 ComputeOrbDescriptors: a = (float)cos(angle), b = (float)sin(angle) in double precision on the float angle,
        // x = px*a - py*b, y = px*b + py*a in float32 with separately rounded products, sampled at (cvRound(y), cvRound(x))
        const float rad = __fmul_rn(angle, (float)(3.1415926535897932384626433832795 / 180.0f));
        const float ca = (float)cos((double)rad), sa = (float)sin((double)rad);
        const int4* gp = reinterpret_cast<const int4*>(b.pattern + lane * 32);
        const int4 qa = __ldg(gp), qb = __ldg(gp + 1);
        const int q8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        int gval = 0;
#pragma unroll
        for (int bit = 0; bit < 8; bit++) {
            const int w = q8[bit];
            const float x0 = (float)(signed char)(w & 0xff), y0 = (float)(signed char)((w >> 8) & 0xff);
            const float x1 = (float)(signed char)((w >> 16) & 0xff), y1 = (float)(signed char)((w >> 24) & 0xff);
            const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sa))), iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sa), __fmul_rn(y0, ca)));
            const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sa))), iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sa), __fmul_rn(y1, ca)));
            const int t0 = ctr[iy0 * pitch + ix0], t1 = ctr[iy1 * pitch + ix1];
            gval |= (t0 < t1) << bit;
        }
        out_desc[((size_t)f * capacity + i) * 32 + lane] = (uint8_t)gval;
        return;
    }
    const int4* pp = reinterpret_cast<const int4*>(b.pattern + bin * 1024 + lane * 32);
    const int4 pa = __ldg(pp), pb = __ldg(pp + 1);
    const int w8[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
        int w = w8[bit];
        int x0 = (int)(signed char)(w & 0xff), y0 = (int)(signed char)((w >> 8) & 0xff);
        int x1 = (int)(signed char)((w >> 16) & 0xff), y1 = (int)(signed char)((w >> 24) & 0xff);
        int t0 = ctr[y0 * pitch + x0], t1 = ctr[y1 * pitch + x1];
        val |= (t0 < t1) << bit;
    }
    out_desc[((size_t)f * capacity + i) * 32 + lane] = (uint8_t)val;
}

// raster-order tap used by the stage parity tests
__global__ void k_sort_candidates(const __grid_constant__ OrbGeom g, const OrbBuffers b, int f, int l, uint32_t* out, int cap, int* count)
{
    const LevelGeom& L = g.lv[l];
    int n = min(b.cand_count[f * kMaxLevels + l], (int)L.cand_cap);
    const uint32_t* cand = b.cand + (size_t)f * b.cand_stride + L.cand_off;
    // rank by counting (debug path, O(n^2/threads))
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t me = cand[i] & 0xFFFFFF;
        int rank = 0;
        for (int k = 0; k < n; k++) rank += (cand[k] & 0xFFFFFF) < me;
        if (rank < cap) out[rank] = cand[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *count = n;
}

} // namespace mage

// ====================================================================================================================
// Host side: handle, geometry, launches
// ====================================================================================================================
using namespace mage;

struct mage_orb_s {
    mage_orb_params params;
    OrbGeom g;
    OrbBuffers b;
    DeviceArena arena;
    int max_batch = 0;
    size_t sel_total = 0, cand_total = 0, key_total = 0;
    int fast_tiles = 0, blur_tiles = 0;
    size_t off_stage_kps = 0, off_stage_desc = 0, off_stage_counts = 0, off_lvl0 = 0;
    uint8_t* h_back = nullptr; size_t h_back_bytes = 0;      // pinned landing buffer of the host path's read-back (allocated at the first call)
    int last_n = 0;
    int stage_cap = 0;          // staging capacity per frame = max(nfeatures, sum of per-level budgets)
    cudaStream_t own_stream = nullptr;
    cudaStream_t aux_stream = nullptr;      // the blur runs here, concurrently with FAST + selection (it only feeds the descriptors)
    // TMA path of FAST: one 3-D tensor map (x, y, frame) per level over the handle's pyramid slabs; level 0 is re-encoded per call
    // when the pixels come from a caller's device buffer. Unavailable (=> k_fast) if the driver entry point cannot be resolved.
    FastMaps maps;
    bool tma_ok = false, maps_lvl0_foreign = false;
    // host path: the kernel chain of a call (memset + 12 launches + the blur fork/join) is captured once per (frames, capacity) and
    // replayed as one CUDA graph -- per-frame calls are launch-latency bound
    struct ChainGraph { int n, cap; cudaGraphExec_t exec; };
    std::vector<ChainGraph> graphs;
    bool graphs_off = false;
    int sm_count = 0;
    int fast_variant = 0;
    const FastTile* d_fast_tiles = nullptr;      // per-frame tile table of k_fast (in the arena)
    void* encode_fn = nullptr;
    cudaEvent_t ev_pyr = nullptr, ev_blur = nullptr;
};

namespace {

typedef void (*FastKernel)(const OrbGeom, const OrbBuffers, const FastTile*);
typedef void (*FastTmaKernel)(const OrbGeom, const OrbBuffers, const FastMaps, const FastTile*);
struct FastVariant { FastKernel k; FastTmaKernel kt; };
#define MAGE_FV(np, nx, nq) {k_fast<np, nx, nq>, k_fast_tma<np, nx, nq>}
const FastVariant kFastVariants[] = {MAGE_FV(8, 8, 0), MAGE_FV(0, 0, 0), MAGE_FV(4, 4, 0), MAGE_FV(6, 6, 0), MAGE_FV(8, 8, 4), MAGE_FV(8, 8, 8), MAGE_FV(8, 4, 0), MAGE_FV(4, 8, 0)};
#undef MAGE_FV

inline int cvRoundF(float v) { return (int)lrintf(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeilF(float v) { int i = (int)v; return i + (i < v); }

const int* gauss_kernel_q8(int ksize)
{
    // Q8.8 kernels of cv::GaussianBlur's bit-exact path for sigma = 2 (OpenCV 4.x), SURVEY appendix A.3
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

// getGaussianKernel(ksize, 2, CV_32F) of OpenCV 4.13, first half (edge .. centre): the kernels of cv::GaussianBlur's float path
const float* gauss_kernel_f32(int ksize)
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

// cv::resize INTER_LINEAR coefficient tables for one axis
void resize_axis(int ssize, int dsize, int* ofs, short2* coef)
{
    double scale = (double)ssize / dsize;
    for (int d = 0; d < dsize; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(fx);
        fx -= s;
        if (s < 0) { s = 0; fx = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; fx = 0.f; }
        ofs[d] = s;
        coef[d] = make_short2((short)cvRoundF((1.f - fx) * 2048.f), (short)cvRoundF(fx * 2048.f));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D byte tensor (x < w, y < h, frame < frames) with row stride `pitch` and frame stride `fstride`; box = one FAST pixel tile
bool encode_level_map(void* fn, CUtensorMap* out, const void* base, int w, int h, int frames, size_t pitch, size_t fstride)
{
    if (!fn || ((uintptr_t)base & 15) || (pitch & 15) || (fstride & 15)) return false;
    const cuuint64_t dim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    const cuuint64_t stride[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {(cuuint32_t)kFastPW, (cuuint32_t)kFastPH, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return reinterpret_cast<EncodeTiledFn>(fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int select_smem_bytes() { return kSelSmemItems * (8 + 4 + 4 + 1); }

// The tile table of k_fast: for every tile of every level (in launch order) its origin, the part of it that is scored -- the
// border-inset rectangle dilated by one pixel and clipped to FAST's domain 3 <= x <= w - 4, in pair granularity -- and the key-point
// rows / columns it may emit (ref :619-639 RunByImageBorder).
std::vector<FastTile> fast_tile_table(const OrbGeom& g)
{
    std::vector<FastTile> out;
    const int bd = g.border;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        for (int ty = 0; ty < L.fast_tiles_y; ty++)
            for (int tx = 0; tx < L.fast_tiles_x; tx++) {
                FastTile t;
                memset(&t, 0, sizeof(t));
                const int px0 = L.fast_x0 + kFastTW * tx, py0 = L.fast_y0 + kFastTH * ty;
                const int vx_lo = std::max(3, bd - 1), vx_hi = std::min(L.w - 4, L.w - bd), vy_lo = std::max(3, bd - 1), vy_hi = std::min(L.h - 4, L.h - bd);
                const int sx_lo = std::max(0, vx_lo - (px0 + 3)), sx_hi = std::min(kFastSW - 1, vx_hi - (px0 + 3));
                const int sy_lo = std::max(0, vy_lo - (py0 + 3)), sy_hi = std::min(kFastSH - 1, vy_hi - (py0 + 3));
                int sp_lo = sx_lo >> 1, npx = sx_hi >= sx_lo ? (sx_hi >> 1) - sp_lo + 1 : 0, nrows = std::max(sy_hi - sy_lo + 1, 0);
                if (npx <= 0 || nrows <= 0) { npx = 0; nrows = 0; sp_lo = 0; }
                t.l = (short)l; t.px0 = (short)px0; t.py0 = (short)py0;
                t.xlo = (short)std::max(bd, px0 + 4); t.xhi = (short)std::min(L.w - bd, px0 + 4 + kFastTW);
                t.oy_lo = (signed char)std::max(0, bd - (py0 + 4)); t.oy_hi = (signed char)std::min(kFastTH - 1, L.h - bd - 1 - (py0 + 4));
                t.sp_lo = (unsigned char)sp_lo; t.npx = (unsigned char)npx; t.sy_lo = (unsigned char)(nrows ? sy_lo : 0); t.nrows = (unsigned char)nrows;
                if (npx > 0) {
                    t.dsy = (unsigned char)(256 / npx); t.dsp = (unsigned char)(256 % npx);
                    t.g_lo = (unsigned char)((2 * sp_lo) >> 4); t.g_hi = (unsigned char)((2 * (sp_lo + npx - 1) + 7) >> 4);
                    t.inv_npx = (unsigned int)((0x100000000ull + (unsigned)npx - 1) / (unsigned)npx);
                    for (unsigned q = 0; q < 256; q++)                                   // the reciprocal is exact over the range it is used on
                        if ((unsigned)(((unsigned long long)q * t.inv_npx) >> 32) != q / (unsigned)npx) { t.inv_npx = 0; break; }
                }
                out.push_back(t);
            }
    }
    return out;
}

} // namespace

extern "C" int mage_orb_create(const mage_orb_params* p, int width, int height, int max_batch, mage_orb_t* out)
{
    MAGE_REQUIRE(p && out, MAGE_ERR_INVALID, "mage_orb_create: null argument");
    MAGE_REQUIRE(width >= 16 && height >= 16 && max_batch >= 1, MAGE_ERR_INVALID, "mage_orb_create: bad size %dx%d batch %d", width, height, max_batch);
    MAGE_REQUIRE((size_t)width * height < (1u << 24), MAGE_ERR_UNSUPPORTED, "image larger than 2^24 pixels");
    MAGE_REQUIRE(p->patch_size >= 2, MAGE_ERR_INVALID, "patch_size must be >= 2 (CV_Assert in the reference)");
    MAGE_REQUIRE(p->patch_size <= 127, MAGE_ERR_UNSUPPORTED, "patch_size %u: at most 127", p->patch_size);
    MAGE_REQUIRE(p->nlevels >= 1 && p->nlevels <= (unsigned)kMaxLevels, MAGE_ERR_UNSUPPORTED, "nlevels must be 1..%d", kMaxLevels);
    MAGE_REQUIRE(p->gaussian_kernel_size <= 1 || gauss_kernel_q8((int)p->gaussian_kernel_size), MAGE_ERR_UNSUPPORTED,
                 "gaussian_kernel_size %u unsupported (odd 3..15)", p->gaussian_kernel_size);
    MAGE_REQUIRE(p->num_cells_x >= 1 && p->num_cells_y >= 1 && p->num_cells_x * p->num_cells_y <= kMaxCells, MAGE_ERR_UNSUPPORTED,
                 "num_cells_x*num_cells_y must be 1..%d", kMaxCells);
    MAGE_REQUIRE(p->fast_threshold >= 1 && p->fast_threshold <= 255, MAGE_ERR_INVALID, "fast_threshold must be 1..255 (reference asserts > 0)");
    MAGE_REQUIRE(p->scale_factor > 1.0f || p->nlevels == 1, MAGE_ERR_INVALID, "scale_factor must be > 1");
    MAGE_REQUIRE(p->nfeatures >= 1, MAGE_ERR_INVALID, "nfeatures must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the ORB path has no CPU fallback"); return MAGE_ERR_CUDA; }

    mage_orb_s* h = new mage_orb_s();
    h->params = *p;
    h->max_batch = max_batch;
    OrbGeom& g = h->g;
    memset(&g, 0, sizeof(g));
    g.nlevels = (int)p->nlevels; g.width = width; g.height = height;
    g.fast_threshold = (int)p->fast_threshold; g.patch = (int)p->patch_size; g.half_patch = g.patch / 2;
    g.use_orientation = p->use_orientation ? 1 : 0;
    g.generic_pattern = (g.patch != 31 && g.patch != 15) ? 1 : 0;              // ref :866-885
    g.border = g.use_orientation ? cvCeilF(g.half_patch * std::sqrt(2.0f)) : g.half_patch;        // ref :712
    g.ksize = p->gaussian_kernel_size > 1 ? (int)p->gaussian_kernel_size : 1;
    g.strong_response = p->strong_response; g.num_cells_x = p->num_cells_x; g.num_cells_y = p->num_cells_y;
    g.feature_factor = p->feature_factor; g.feature_strength = p->feature_strength; g.min_rf = p->min_robust_factor; g.max_rf = p->max_robust_factor;
    if (g.ksize > 1) {
        const int* k = gauss_kernel_q8(g.ksize); for (int i = 0; i < g.ksize; i++) g.gk[i] = k[i];
        const float* kf = gauss_kernel_f32(g.ksize); for (int i = 0; i <= g.ksize / 2; i++) g.gkf[i] = kf[i];
        // ref :792,:812,:860: the blur source is imagePyramid(layerInfo[level]) -- a proper submatrix of the (cols+15)&-16 wide packed
        // buffer unless there is a single level that fills it; cv::GaussianBlur only takes its fixed-point path for non-submatrix sources
        g.blur_float = (g.nlevels > 1 || (width & 15) != 0) ? 1 : 0;
    }
    {   // ref :673-689 umax
        int hp = g.half_patch, v, v0, vmax = cvFloorF(hp * std::sqrt(2.f) / 2 + 1), vmin = cvCeilF(hp * std::sqrt(2.f) / 2);
        for (v = 0; v <= vmax; ++v) g.umax[v] = (int)lrint(std::sqrt((double)hp * hp - v * v));
        for (v = hp, v0 = 0; v >= vmin; --v) { while (g.umax[v0] == g.umax[v0 + 1]) ++v0; g.umax[v] = v0; ++v0; }
    }
    // level sizes (ref :795-799), per-level budget (ref :660-670)
    {
        int nfeatures = (int)p->nfeatures;
        float factor = 1.0f / p->scale_factor;
        float ndesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)g.nlevels));
        int sum = 0;
        for (int l = 0; l < g.nlevels; l++) {
            float scale = (float)std::pow((double)p->scale_factor, (double)l);
            g.lv[l].scale = scale;
            g.lv[l].w = cvRoundF(width / scale);
            g.lv[l].h = cvRoundF(height / scale);
            if (l < g.nlevels - 1) { g.lv[l].nfeat = cvRoundF(ndesired); sum += g.lv[l].nfeat; ndesired *= factor; }
            else g.lv[l].nfeat = std::max(nfeatures - sum, 0);
            if (g.lv[l].w < 8 || g.lv[l].h < 8) { delete h; set_error("level %d is %dx%d: too small", l, g.lv[l].w, g.lv[l].h); return MAGE_ERR_UNSUPPORTED; }
        }
    }
    // offsets
    size_t pyr = 0, cand = 0, key = 0, sel = 0, tab = 0;
    int ftiles = 0, btiles = 0;
    for (int l = 0; l < g.nlevels; l++) {
        LevelGeom& L = g.lv[l];
        L.pitch = (int)align_up((size_t)L.w, 64);
        L.pyr_off = (unsigned)pyr; pyr += align_up((size_t)L.pitch * L.h, 256);
        L.cand_cap = (unsigned)(((L.w + 1) / 2) * ((L.h + 1) / 2));          // strict 3x3 NMS: at most one per 2x2 block
        L.cand_off = (unsigned)cand; cand += align_up(L.cand_cap, 64);
        unsigned kc = 1; while (kc < L.cand_cap) kc <<= 1;
        L.key_cap = kc; L.key_off = (unsigned)key; key += kc;
        L.sel_off = (unsigned)sel; sel += align_up((size_t)std::max(L.nfeat, 1), 32);
        L.tab_x = (unsigned)tab; tab += L.w; L.tab_y = (unsigned)tab; tab += L.h;
        // key points live in [border, w - border) x [border, h - border) (ref :619-639, :706-711); the first tile's 4-pixel apron
        // starts at border - 4 rounded down to a multiple of 16 bytes (vector loads / the TMA engine's innermost coordinate)
        L.fast_x0 = (int)(((g.border - 4 + 1024) & ~15) - 1024); L.fast_y0 = g.border - 4;
        const bool has_kp = L.w > 2 * g.border && L.h > 2 * g.border;
        L.fast_tiles_x = has_kp ? div_up(L.w - g.border - (L.fast_x0 + 4), kFastTW) : 0;
        L.fast_tiles_y = has_kp ? div_up(L.h - 2 * g.border, kFastTH) : 0;
        L.fast_tile_base = ftiles;
        ftiles += L.fast_tiles_x * L.fast_tiles_y;
        L.blur_tiles_x = div_up(L.w, kBlurOW); L.blur_tiles_y = div_up(L.h, g.ksize == 7 ? 8 * kBlur7Rows : kBlurOH); L.blur_tile_base = btiles;
        btiles += L.blur_tiles_x * L.blur_tiles_y;
    }
    h->fast_tiles = ftiles; h->blur_tiles = btiles;
    h->cand_total = cand; h->key_total = key; h->sel_total = sel;

    DeviceArena& A = h->arena;
    const size_t B = (size_t)max_batch;
    size_t o_pyr = A.reserve(pyr * B), o_blur = A.reserve(pyr * B);
    size_t o_tabo = A.reserve(tab * sizeof(int)), o_tabc = A.reserve(tab * sizeof(short2));
    size_t o_cand = A.reserve(cand * 4 * B), o_kept = A.reserve(cand * 4 * B), o_cxy = A.reserve(cand * 4 * B), o_csc = A.reserve(cand * B);
    size_t o_keys = A.reserve(key * 8 * B);
    size_t o_sel = A.reserve(sel * 4 * B);
    size_t o_cnt = A.reserve(sizeof(int) * kMaxLevels * B * 2 + sizeof(int) * B);     // cand_count, sel_count, status
    size_t o_pat = A.reserve(30 * 1024);
    size_t o_ftiles = A.reserve(sizeof(FastTile) * (size_t)std::max(ftiles, 1));
    int sum_nl = 0;
    for (int l = 0; l < g.nlevels; l++) sum_nl += g.lv[l].nfeat;
    const int cap = std::max((int)p->nfeatures, sum_nl);
    h->stage_cap = cap;
    h->off_stage_kps = A.reserve(sizeof(mage_keypoint) * cap * B);
    h->off_stage_desc = A.reserve((size_t)32 * cap * B);
    h->off_stage_counts = A.reserve(sizeof(int) * B);
    if (A.commit() != cudaSuccess) { set_error("cudaMalloc of %zu bytes failed", A.used); delete h; return MAGE_ERR_CUDA; }
    cudaMemset(A.base, 0, A.size);

    OrbBuffers& b = h->b;
    b.pyr = A.at<uint8_t>(o_pyr); b.blur = A.at<uint8_t>(o_blur); b.slab = pyr;
    b.lvl0 = b.pyr; b.lvl0_frame_stride = pyr; b.lvl0_pitch = g.lv[0].pitch;
    b.tab_ofs = A.at<int>(o_tabo); b.tab_coef = A.at<short2>(o_tabc);
    b.cand = A.at<uint32_t>(o_cand); b.kept = A.at<uint32_t>(o_kept); b.cxy = A.at<uint32_t>(o_cxy); b.csc = A.at<uint8_t>(o_csc);
    b.cand_stride = cand;
    b.keys = A.at<unsigned long long>(o_keys); b.key_stride = key;
    b.sel = A.at<uint32_t>(o_sel); b.sel_stride = sel;
    b.cand_count = A.at<int>(o_cnt); b.sel_count = b.cand_count + kMaxLevels * B; b.status = b.sel_count + kMaxLevels * B;
    b.pattern = A.at<int8_t>(o_pat);
    h->d_fast_tiles = A.at<FastTile>(o_ftiles);

    // tables
    std::vector<int> tofs(tab); std::vector<short2> tcoef(tab);
    for (int l = 1; l < g.nlevels; l++) {
        resize_axis(g.lv[l - 1].w, g.lv[l].w, &tofs[g.lv[l].tab_x], &tcoef[g.lv[l].tab_x]);
        resize_axis(g.lv[l - 1].h, g.lv[l].h, &tofs[g.lv[l].tab_y], &tcoef[g.lv[l].tab_y]);
    }
    std::vector<int8_t> pat(30 * 1024);
    if (g.generic_pattern) {
        // ref :551-560 MakeRandomPattern: cv::RNG(0x34985739) -- multiply-with-carry, next() = (unsigned)(state = (unsigned)state *
        // 4164903690 + (state >> 32)); uniform(a, b) = next() % (b - a) + a -- x then y for each of the 512 points
        uint64_t state = 0x34985739u;
        auto next = [&]() { state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32); return (uint32_t)state; };
        const int lo = -g.patch / 2, hi = g.patch / 2 + 1;
        for (int i = 0; i < 1024; i++) pat[i] = (int8_t)(int)(next() % (uint32_t)(hi - lo) + lo);
    } else
    {   // rows 1..29 = row 0 rotated by 12r degrees, float32, round-half-even
        const signed char* base = g.patch == 31 ? kBriefBase31 : kBriefBase15;
        for (int r = 0; r < 30; r++) {
            double ang = (double)(12 * r) * M_PI / 180.0;
            float c = (float)std::cos(ang), s = (float)std::sin(ang);
            for (int i = 0; i < 512; i++) {
                float x = base[2 * i], y = base[2 * i + 1];
                volatile float xc = x * c, ys = y * s, xs = x * s, yc = y * c;      // products rounded separately (no contraction)
                pat[r * 1024 + 2 * i] = (int8_t)lrintf(xc - ys);
                pat[r * 1024 + 2 * i + 1] = (int8_t)lrintf(xs + yc);
            }
        }
    }
    cudaError_t e = cudaMemcpy((void*)b.tab_ofs, tofs.data(), tab * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void*)b.tab_coef, tcoef.data(), tab * sizeof(short2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void*)b.pattern, pat.data(), pat.size(), cudaMemcpyHostToDevice);
    {
        const std::vector<FastTile> ft = fast_tile_table(g);
        if (e == cudaSuccess && !ft.empty()) e = cudaMemcpy((void*)h->d_fast_tiles, ft.data(), ft.size() * sizeof(FastTile), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_select<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, select_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_select<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, select_smem_bytes());
    {   // which share of the min/max network runs on the FMA pipe: tuning switch, every variant is bit-identical
        const char* env = getenv("MAGE_FAST_VARIANT");
        const int v = env ? atoi(env) : 0;
        h->fast_variant = v >= 0 && v < (int)(sizeof(kFastVariants) / sizeof(kFastVariants[0])) ? v : 0;
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)kFastVariants[h->fast_variant].k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmemBytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pyr, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_blur, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("mage_orb_create: %s", cudaGetErrorString(e)); A.release(); delete h; return MAGE_ERR_CUDA; }
    {   // TMA path of FAST: the default (MAGE_FAST_TMA=0 selects the register-staged kernel). One bulk tensor load per CTA brings the
        // tile's box in (zero fill outside the image by the tensor map); on B200 the two kernels are within 1.5 % of each other
        // (0.422 vs 0.416 ms per 128 frames on the chart scene, 0.293 vs 0.289 ms on camera-like frames): the kernel is ALU-bound
        const char* env = getenv("MAGE_FAST_TMA");
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        if ((!env || atoi(env) != 0) && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess && fn) {
            h->encode_fn = fn;
            bool ok = cudaFuncSetAttribute((const void*)kFastVariants[h->fast_variant].kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmemBytes) == cudaSuccess;
            for (int l = 0; l < g.nlevels && ok; l++)
                ok = encode_level_map(fn, &h->maps.m[l], b.pyr + g.lv[l].pyr_off, g.lv[l].w, g.lv[l].h, max_batch, (size_t)g.lv[l].pitch, pyr);
            h->tma_ok = ok;
        }
        cudaGetLastError();
    }
    *out = h;
    return MAGE_OK;
}

extern "C" void mage_orb_destroy(mage_orb_t h)
{
    if (!h) return;
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_pyr) cudaEventDestroy(h->ev_pyr);
    if (h->ev_blur) cudaEventDestroy(h->ev_blur);
    for (auto& gph : h->graphs) cudaGraphExecDestroy(gph.exec);
    if (h->h_back) cudaFreeHost(h->h_back);
    h->arena.release();
    delete h;
}

// Which arithmetic cv::GaussianBlur runs depends on the OpenCV build behind the reference; the default is what
// OpenCV 4.13 does on the reference's call. Captured launch graphs hold the geometry by value, so they are dropped.
extern "C" int mage_orb_set_blur_mode(mage_orb_t h, int mode)
{
    MAGE_REQUIRE(h, MAGE_ERR_INVALID, "null handle");
    MAGE_REQUIRE(mode >= MAGE_BLUR_AUTO && mode <= MAGE_BLUR_FIXED, MAGE_ERR_INVALID, "blur mode %d: 0..3", mode);
    OrbGeom& g = h->g;
    if (g.ksize > 1) {
        const bool submatrix = g.nlevels > 1 || (g.width & 15) != 0;
        g.blur_float = mode == MAGE_BLUR_FLOAT_FUSED ? 1 : mode == MAGE_BLUR_FLOAT_UNFUSED ? 2 : mode == MAGE_BLUR_FIXED ? 0 : (submatrix ? 1 : 0);
    }
    for (auto& gph : h->graphs) cudaGraphExecDestroy(gph.exec);
    h->graphs.clear();
    return MAGE_OK;
}

extern "C" int mage_orb_level_info(mage_orb_t h, int* widths, int* heights, float* scales, int* nfeat)
{
    MAGE_REQUIRE(h, MAGE_ERR_INVALID, "null handle");
    for (int l = 0; l < h->g.nlevels; l++) {
        if (widths) widths[l] = h->g.lv[l].w;
        if (heights) heights[l] = h->g.lv[l].h;
        if (scales) scales[l] = h->g.lv[l].scale;
        if (nfeat) nfeat[l] = h->g.lv[l].nfeat;
    }
    return MAGE_OK;
}

// The kernel chain for n frames whose level-0 pixels are described by bufs.lvl0*. Asynchronous on s.
static int orb_launch(mage_orb_t h, const OrbBuffers& bufs, int n, mage_keypoint* d_kps, uint8_t* d_desc, int capacity, int* d_counts, cudaStream_t s)
{
    const OrbGeom& g = h->g;
    MAGE_CUDA_TRY(cudaMemsetAsync(bufs.cand_count, 0, sizeof(int) * kMaxLevels * h->max_batch * 2 + sizeof(int) * h->max_batch, s));
    {
        ProfScope ps(PROF_RESIZE, s);       // the L-1 chained launches are timed as one group
        bool prev_fast = false;
        for (int l = 1; l < g.nlevels; l++) {
            dim3 block(32, 8);
            // a pair of adjacent output pixels must find its four taps inside two aligned words: source step <= 3 pixels
            const bool fast = (double)g.lv[l - 1].w / g.lv[l].w <= 3.0 && g.lv[l - 1].w >= 8;
            if (fast) {
                // levels 2.. follow a k_resize4 launch: programmatic dependent launch overlaps their launch latency and set-up with the
                // tail of the level before (seven dependent launches of a few microseconds each are a third of a one-frame call)
                static const bool pdl_on = !(getenv("MAGE_ORB_PDL") && atoi(getenv("MAGE_ORB_PDL")) == 0);
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(div_up(g.lv[l].w, 128), div_up(g.lv[l].h, 8 * kResizeRows), n); cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
                cudaLaunchAttribute at{};
                at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
                const bool dep = pdl_on && prev_fast && l >= 2;
                cfg.attrs = dep ? &at : nullptr; cfg.numAttrs = dep ? 1 : 0;
                MAGE_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_resize4, g, bufs, l));
            }
            else k_resize<<<dim3(div_up(g.lv[l].w, 128), div_up(g.lv[l].h, 8), n), block, 0, s>>>(g, bufs, l);
            prev_fast = fast;
        }
    }
    // fork: the blurred pyramid is only read by the descriptor stage, so it is produced on a second stream while FAST and the
    // selection run on the caller's stream (when per-kernel timing is on, everything stays on one stream to keep the events clean)
    const bool fork = g.ksize > 1 && !prof_enabled();
    cudaStream_t sb = fork ? h->aux_stream : s;
    if (fork) { MAGE_CUDA_TRY(cudaEventRecord(h->ev_pyr, s)); MAGE_CUDA_TRY(cudaStreamWaitEvent(sb, h->ev_pyr, 0)); }
    if (g.ksize == 7) {
        ProfScope ps(PROF_BLUR, sb);
        if (g.blur_float == 1) {
            k_blur7f<true><<<dim3(h->blur_tiles, n), 256, 0, sb>>>(g, bufs);
            k_blur7f_edges<true><<<dim3(div_up(15 * g.lv[0].h, 256), g.nlevels, n), 256, 0, sb>>>(g, bufs);
        } else if (g.blur_float == 2) {
            k_blur7f<false><<<dim3(h->blur_tiles, n), 256, 0, sb>>>(g, bufs);
            k_blur7f_edges<false><<<dim3(div_up(15 * g.lv[0].h, 256), g.nlevels, n), 256, 0, sb>>>(g, bufs);
        } else {
            k_blur7<<<dim3(h->blur_tiles, n), 256, 0, sb>>>(g, bufs);
            k_blur7_edges<<<dim3(div_up(15 * g.lv[0].h, 256), g.nlevels, n), 256, 0, sb>>>(g, bufs);
        }
    }
    else if (g.ksize > 1) {
        ProfScope ps(PROF_BLUR, sb);
        if (g.blur_float == 1) k_blurf<true><<<dim3(h->blur_tiles, n), 256, 0, sb>>>(g, bufs);
        else if (g.blur_float == 2) k_blurf<false><<<dim3(h->blur_tiles, n), 256, 0, sb>>>(g, bufs);
        else k_blur<<<dim3(h->blur_tiles, n), 256, 0, sb>>>(g, bufs);
    }
    if (fork) MAGE_CUDA_TRY(cudaEventRecord(h->ev_blur, sb));
    {
        ProfScope ps(PROF_FAST, s);
        bool tma = h->tma_ok;
        if (tma && bufs.lvl0 != h->b.lvl0)          // level 0 lives in the caller's device buffer: its own map (falls back if it is not 16-byte regular)
            tma = encode_level_map(h->encode_fn, &h->maps.m[0], bufs.lvl0, g.lv[0].w, g.lv[0].h, n, (size_t)bufs.lvl0_pitch, bufs.lvl0_frame_stride);
        else if (tma && h->maps_lvl0_foreign)
            tma = encode_level_map(h->encode_fn, &h->maps.m[0], h->b.pyr + g.lv[0].pyr_off, g.lv[0].w, g.lv[0].h, h->max_batch, (size_t)g.lv[0].pitch, h->b.slab);
        h->maps_lvl0_foreign = bufs.lvl0 != h->b.lvl0;
        if (h->fast_tiles > 0) {         // no tile: every level is narrower than twice the border, nothing can be a key point
            if (tma) kFastVariants[h->fast_variant].kt<<<dim3(h->fast_tiles, n), 256, kFastSmemBytes, s>>>(g, bufs, h->maps, h->d_fast_tiles);
            else kFastVariants[h->fast_variant].k<<<dim3(h->fast_tiles, n), 256, kFastSmemBytes, s>>>(g, bufs, h->d_fast_tiles);
        }
    }
    // one CTA per (level, frame): a chain of dependent phases. With a few frames the call's latency is this kernel's (58 us for one 640 x 480
    // frame with 512 threads, 46 us with 1024); with a batch four 512-thread CTAs per SM give the better throughput
    {
        static const int lanes_env = getenv("MAGE_SELECT_LANES") ? atoi(getenv("MAGE_SELECT_LANES")) : -1;      // -1: by batch size
        const bool lanes = lanes_env >= 0 ? lanes_env != 0 : n <= 4;
        ProfScope ps(PROF_SELECT, s);
        if (lanes) k_select<true><<<dim3(g.nlevels, n), n <= 4 ? kSelThreads : kSelThreads / 2, select_smem_bytes(), s>>>(g, bufs);
        else k_select<false><<<dim3(g.nlevels, n), n <= 4 ? kSelThreads : kSelThreads / 2, select_smem_bytes(), s>>>(g, bufs);
    }
    if (fork) MAGE_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_blur, 0));
    { ProfScope ps(PROF_ORIENT_DESCRIBE, s); k_orient_describe<<<dim3(div_up(capacity, 8), n), 256, 0, s>>>(g, bufs, d_kps, d_desc, d_counts, capacity); }
    MAGE_CUDA_TRY(cudaGetLastError());
    h->last_n = n;
    return MAGE_OK;
}

extern "C" int mage_orb_extract_device(mage_orb_t h, const uint8_t* d_images, int n, int width, int height, int stride,
                                       size_t frame_stride, mage_keypoint* d_kps, uint8_t* d_desc, int capacity,
                                       int* d_counts, void* stream)
{
    MAGE_REQUIRE(h && d_images && d_kps && d_desc && d_counts, MAGE_ERR_INVALID, "mage_orb_extract_device: null argument");
    MAGE_REQUIRE(n >= 1 && n <= h->max_batch, MAGE_ERR_INVALID, "batch %d exceeds max_batch %d", n, h->max_batch);
    MAGE_REQUIRE(width == h->g.width && height == h->g.height, MAGE_ERR_INVALID, "image %dx%d does not match the handle (%dx%d)", width, height, h->g.width, h->g.height);
    MAGE_REQUIRE(stride >= width && stride % 4 == 0 && ((uintptr_t)d_images % 16) == 0 && frame_stride % 4 == 0, MAGE_ERR_INVALID,
                 "device images need stride %% 4 == 0 and a 16-byte aligned base");
    MAGE_REQUIRE(capacity >= 1, MAGE_ERR_INVALID, "capacity must be >= 1");
    OrbBuffers bufs = h->b;
    bufs.lvl0 = d_images; bufs.lvl0_frame_stride = frame_stride; bufs.lvl0_pitch = stride;
    return orb_launch(h, bufs, n, d_kps, d_desc, capacity, d_counts, (cudaStream_t)stream);
}

extern "C" int mage_orb_detect_and_compute_batch(mage_orb_t h, const uint8_t* images, int n, int width, int height, int stride,
                                                 size_t frame_stride, mage_keypoint* kps, uint8_t* desc, int capacity,
                                                 int* counts, void* stream)
{
    MAGE_REQUIRE(h && images && kps && desc && counts, MAGE_ERR_INVALID, "mage_orb_detect_and_compute: null argument");
    MAGE_REQUIRE(n >= 1 && n <= h->max_batch, MAGE_ERR_INVALID, "batch %d exceeds max_batch %d", n, h->max_batch);
    MAGE_REQUIRE(width == h->g.width && height == h->g.height && stride >= width, MAGE_ERR_INVALID,
                 "image %dx%d (stride %d) does not match the handle (%dx%d)", width, height, stride, h->g.width, h->g.height);
    MAGE_REQUIRE(capacity >= 1, MAGE_ERR_INVALID, "capacity must be >= 1");
    cudaStream_t s = stream ? (cudaStream_t)stream : h->own_stream;
    const OrbGeom& g = h->g;
    const int cap = std::min(capacity, h->stage_cap);     // at most sum(n_l) <= stage_cap keypoints exist per frame
    // level 0 = copy of the source image (ref :838) straight into the pyramid slab
    for (int f = 0; f < n; f++)
        MAGE_CUDA_TRY(cudaMemcpy2DAsync(h->b.pyr + (size_t)f * h->b.slab + g.lv[0].pyr_off, g.lv[0].pitch, images + (size_t)f * frame_stride,
                                        stride, width, height, cudaMemcpyHostToDevice, s));
    mage_keypoint* d_kps = h->arena.at<mage_keypoint>(h->off_stage_kps);
    uint8_t* d_desc = h->arena.at<uint8_t>(h->off_stage_desc);
    int* d_counts = h->arena.at<int>(h->off_stage_counts);
    int rc = MAGE_OK;
    cudaGraphExec_t exec = nullptr;
    if (!h->graphs_off && !prof_enabled()) {
        for (auto& gph : h->graphs) if (gph.n == n && gph.cap == cap) exec = gph.exec;
        if (!exec) {
            static const bool env_off = getenv("MAGE_ORB_GRAPH") && atoi(getenv("MAGE_ORB_GRAPH")) == 0;
            cudaGraph_t graph = nullptr;
            if (env_off || cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) h->graphs_off = true;
            else {
                rc = orb_launch(h, h->b, n, d_kps, d_desc, cap, d_counts, s);
                const cudaError_t ce = cudaStreamEndCapture(s, &graph);
                if (rc != MAGE_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { exec = nullptr; h->graphs_off = true; }
                if (graph) cudaGraphDestroy(graph);
                if (exec) h->graphs.push_back({n, cap, exec});
                cudaGetLastError();
                rc = MAGE_OK;
            }
        }
    }
    if (exec) { MAGE_CUDA_TRY(cudaGraphLaunch(exec, s)); h->last_n = n; }
    else rc = orb_launch(h, h->b, n, d_kps, d_desc, cap, d_counts, s);
    if (rc != MAGE_OK) return rc;
    // read-back. The staged key points, descriptors and counts are neighbours in the arena: when the call fills most of that span (a
    // one-frame detector always does) it comes back as ONE copy into a pinned landing buffer and is handed out from there -- four copies
    // into the caller's pageable arrays are four staged, synchronous transfers (about 8 us each on a 200 us call)
    const size_t span = h->off_stage_counts + sizeof(int) * (size_t)h->max_batch - h->off_stage_kps, span_al = align_up(span, 256);
    if ((size_t)n * cap * (sizeof(mage_keypoint) + 32) * 2 >= span) {
        if (h->h_back_bytes < span_al + sizeof(int) * (size_t)h->max_batch) {
            if (h->h_back) cudaFreeHost(h->h_back);
            h->h_back = nullptr; h->h_back_bytes = 0;
            if (cudaHostAlloc(reinterpret_cast<void**>(&h->h_back), span_al + sizeof(int) * (size_t)h->max_batch, cudaHostAllocDefault) == cudaSuccess)
                h->h_back_bytes = span_al + sizeof(int) * (size_t)h->max_batch;
            else cudaGetLastError();
        }
    }
    if (h->h_back_bytes && (size_t)n * cap * (sizeof(mage_keypoint) + 32) * 2 >= span) {
        MAGE_CUDA_TRY(cudaMemcpyAsync(h->h_back, h->arena.base + h->off_stage_kps, span, cudaMemcpyDeviceToHost, s));
        MAGE_CUDA_TRY(cudaMemcpyAsync(h->h_back + span_al, h->b.status, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
        MAGE_CUDA_TRY(cudaStreamSynchronize(s));
        const uint8_t* bk = h->h_back;
        const uint8_t* bd = h->h_back + (h->off_stage_desc - h->off_stage_kps);
        for (int f = 0; f < n; f++) {
            memcpy(kps + (size_t)capacity * f, bk + sizeof(mage_keypoint) * (size_t)cap * f, sizeof(mage_keypoint) * (size_t)cap);
            memcpy(desc + (size_t)32 * capacity * f, bd + (size_t)32 * cap * f, (size_t)32 * cap);
        }
        memcpy(counts, h->h_back + (h->off_stage_counts - h->off_stage_kps), sizeof(int) * n);
        const int* st = reinterpret_cast<const int*>(h->h_back + span_al);
        for (int f = 0; f < n; f++) MAGE_REQUIRE(st[f] == 0, MAGE_ERR_OVERFLOW, "frame %d: candidate list overflow", f);
        return MAGE_OK;
    }
    if (cap == capacity) {
        MAGE_CUDA_TRY(cudaMemcpyAsync(kps, d_kps, sizeof(mage_keypoint) * (size_t)cap * n, cudaMemcpyDeviceToHost, s));
        MAGE_CUDA_TRY(cudaMemcpyAsync(desc, d_desc, (size_t)32 * cap * n, cudaMemcpyDeviceToHost, s));
    } else {
        MAGE_CUDA_TRY(cudaMemcpy2DAsync(kps, sizeof(mage_keypoint) * (size_t)capacity, d_kps, sizeof(mage_keypoint) * (size_t)cap,
                                        sizeof(mage_keypoint) * (size_t)cap, n, cudaMemcpyDeviceToHost, s));
        MAGE_CUDA_TRY(cudaMemcpy2DAsync(desc, (size_t)32 * capacity, d_desc, (size_t)32 * cap, (size_t)32 * cap, n, cudaMemcpyDeviceToHost, s));
    }
    MAGE_CUDA_TRY(cudaMemcpyAsync(counts, d_counts, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    std::vector<int> status(n);
    MAGE_CUDA_TRY(cudaMemcpyAsync(status.data(), h->b.status, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    MAGE_CUDA_TRY(cudaStreamSynchronize(s));
    for (int f = 0; f < n; f++) MAGE_REQUIRE(status[f] == 0, MAGE_ERR_OVERFLOW, "frame %d: candidate list overflow", f);
    return MAGE_OK;
}

extern "C" int mage_orb_detect_and_compute(mage_orb_t h, const uint8_t* image, int width, int height, int stride,
                                           mage_keypoint* kps, uint8_t* desc, int capacity, int* count, void* stream)
{
    return mage_orb_detect_and_compute_batch(h, image, 1, width, height, stride, (size_t)stride * height, kps, desc, capacity, count, stream);
}

extern "C" int mage_orb_debug_get_level(mage_orb_t h, int frame, int level, int blurred, uint8_t* out)
{
    MAGE_REQUIRE(h && out && frame >= 0 && frame < h->max_batch && level >= 0 && level < h->g.nlevels, MAGE_ERR_INVALID, "bad debug request");
    const LevelGeom& L = h->g.lv[level];
    const uint8_t* src = (blurred ? h->b.blur : h->b.pyr) + (size_t)frame * h->b.slab + L.pyr_off;
    MAGE_CUDA_TRY(cudaDeviceSynchronize());
    MAGE_CUDA_TRY(cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return MAGE_OK;
}

extern "C" int mage_orb_debug_get_candidates(mage_orb_t h, int frame, int level, uint32_t* out, int capacity, int* count)
{
    MAGE_REQUIRE(h && out && count && frame >= 0 && frame < h->max_batch && level >= 0 && level < h->g.nlevels, MAGE_ERR_INVALID, "bad debug request");
    uint32_t* d_out = nullptr; int* d_cnt = nullptr;
    MAGE_CUDA_TRY(cudaDeviceSynchronize());
    MAGE_CUDA_TRY(cudaMalloc(&d_out, sizeof(uint32_t) * (size_t)capacity + sizeof(int)));
    d_cnt = reinterpret_cast<int*>(d_out + capacity);
    k_sort_candidates<<<64, 256>>>(h->g, h->b, frame, level, d_out, capacity, d_cnt);
    cudaError_t e = cudaMemcpy(out, d_out, sizeof(uint32_t) * (size_t)capacity, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(count, d_cnt, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    MAGE_CUDA_TRY(e);
    return MAGE_OK;
}
