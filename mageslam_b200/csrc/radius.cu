// RadiusMatch for sm_100a: position-gated best/second-best Hamming matching of query keypoints against the keypoints of one
// image -- the per-frame "match against the local map" of steady-state tracking.
//
// Replaces RadiusMatch (ref Core/MAGESLAM/Source/Tracking/FeatureMatcher.cpp:294-446, both overloads) and the query side of
// KeypointSpatialIndex (ref Image/KeypointSpatialIndex.cpp:26-58, :89-97) behind include/mage_b200.h.
//
// The reference enumerates candidates through a boost::geometry R*-tree (rstar<12>, built by the range constructor = boost's
// packing algorithm) and its result depends on that enumeration order: the best candidate is the FIRST one with the smallest
// distance and "second best" is the running minimum seen BEFORE it (ref :425-436). A box query reports values in depth-first
// order of the packed tree, so instead of building a tree the host computes each target's RANK in that order -- by restating
// the packing recursion (element-count median split along the longest edge of the hint box with std::nth_element, leaves of
// <= 12 values; boost 1.67 index/detail/rtree/pack_create.hpp) -- and the kernel scans the targets by brute force:
//     candidates  = same octave (z = octave*100, query range +-1) and |dx|, |dy| <= radius (closed float box)
//     (d*, r*)    = lexicographic minimum of (distance, rank) over candidates with distance <= maxHamming
//     second      = min(maxHamming + 1, min distance over candidates with rank < r*)
//     accept      <=> a best exists and second - d* > minHammingDifference
// followed by the per-target uniqueness filter of the multi-query overload (ref :344-372).
// Integer/byte work: 2000 x 2000 box tests + ~16 Hamming distances per query; one warp per query.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace mage {

struct RadiusIndexDev {
    const float *x, *y;         // [n]
    const int *octave, *rank;   // [n]
    int n;
};

constexpr unsigned kNone = 0xFFFFFFFFu;

// one warp per query: almost[q] = (distance << 20 | train) or kNone; per-target best/second distance via integer atomics
__global__ void __launch_bounds__(256) k_radius_best(RadiusIndexDev ix, const mage_keypoint* __restrict__ qk, int nq, const float* __restrict__ qpos,
                                                     const uint8_t* __restrict__ qmask, const uint32_t* __restrict__ qdesc,
                                                     const uint8_t* __restrict__ tmask, const uint32_t* __restrict__ tdesc, float radius,
                                                     int maxHamming, int minDiff, unsigned* __restrict__ almost, unsigned* __restrict__ tbest,
                                                     unsigned* __restrict__ tsecond)
{
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    if (qmask && !qmask[q]) { if (lane == 0) almost[q] = kNone; return; }
    const float px = qpos ? qpos[2 * q] : qk[q].x, py = qpos ? qpos[2 * q + 1] : qk[q].y;
    const float lox = __fsub_rn(px, radius), hix = __fadd_rn(px, radius), loy = __fsub_rn(py, radius), hiy = __fadd_rn(py, radius);
    const int oct = qk[q].octave;
    uint32_t qw[8];
#pragma unroll
    for (int w = 0; w < 8; w++) qw[w] = qdesc[(size_t)q * 8 + w];
    // pass 1: lexicographic min of (distance, rank); key = distance << 20 | rank, train index carried alongside
    unsigned bestKey = kNone; int bestT = -1;
    for (int t = lane; t < ix.n; t += 32) {
        const float tx = ix.x[t], ty = ix.y[t];
        if (ix.octave[t] != oct || !(lox <= tx && tx <= hix && loy <= ty && ty <= hiy)) continue;
        if (tmask && !tmask[t]) continue;
        int d = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) d += __popc(qw[w] ^ tdesc[(size_t)t * 8 + w]);
        if (d > maxHamming) continue;                         // bestHammingDistance starts at maxHamming + 1, strict <
        const unsigned key = ((unsigned)d << 20) | (unsigned)ix.rank[t];
        if (key < bestKey) { bestKey = key; bestT = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned ok = __shfl_xor_sync(0xffffffffu, bestKey, o);
        const int ot = __shfl_xor_sync(0xffffffffu, bestT, o);
        if (ok < bestKey) { bestKey = ok; bestT = ot; }
    }
    if (bestT < 0) { if (lane == 0) almost[q] = kNone; return; }
    const int bestD = (int)(bestKey >> 20), bestRank = (int)(bestKey & 0xFFFFF);
    // pass 2: running minimum before the best in enumeration order
    int second = maxHamming + 1;
    for (int t = lane; t < ix.n; t += 32) {
        if (ix.rank[t] >= bestRank) continue;
        const float tx = ix.x[t], ty = ix.y[t];
        if (ix.octave[t] != oct || !(lox <= tx && tx <= hix && loy <= ty && ty <= hiy)) continue;
        if (tmask && !tmask[t]) continue;
        int d = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) d += __popc(qw[w] ^ tdesc[(size_t)t * 8 + w]);
        second = min(second, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) second = min(second, __shfl_xor_sync(0xffffffffu, second, o));
    if (lane == 0) {
        // before the first record the reference's "second" is INT_MAX; after it, the previous record (<= maxHamming + 1)
        const bool accept = (second - bestD) > minDiff;
        if (accept) {
            almost[q] = ((unsigned)bestD << 20) | (unsigned)bestT;
            const unsigned old = atomicMin(&tbest[bestT], (unsigned)bestD);
            atomicMin(&tsecond[bestT], max(old, (unsigned)bestD));
        } else almost[q] = kNone;
    }
}

// ref :344-372: keep a match iff its distance is the unique minimum among the accepted matches of its target; ascending query order
__global__ void __launch_bounds__(256) k_radius_emit(const unsigned* __restrict__ almost, int nq, const unsigned* __restrict__ tbest,
                                                     const unsigned* __restrict__ tsecond, mage_dmatch* __restrict__ out, int* __restrict__ count)
{
    __shared__ int warp_sum[8];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int q0 = 0; q0 < nq; q0 += blockDim.x) {
        const int q = q0 + threadIdx.x;
        bool ok = false; unsigned key = kNone;
        if (q < nq) {
            key = almost[q];
            if (key != kNone) { const unsigned t = key & 0xFFFFF, d = key >> 20; ok = d == tbest[t] && tbest[t] < tsecond[t]; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_sum[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += warp_sum[w];
        off += __popc(m & ((1u << lane) - 1));
        if (ok) { mage_dmatch dm; dm.query_idx = q; dm.train_idx = (int)(key & 0xFFFFF); dm.distance = (float)(key >> 20); out[off] = dm; }
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 8; w++) tot += warp_sum[w]; base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

} // namespace mage

using namespace mage;

// ---------------------------------------------------------------------------------------------------------------------
// Host: enumeration rank of every keypoint in the packed R*-tree (boost 1.67 pack_create.hpp restated)
namespace {

struct Entry { float c[3]; int idx; };
struct Box3 { float lo[3], hi[3]; };
constexpr size_t kMaxElements = 12, kMinElements = 3;              // rstar<12>, default min = 30 %

struct Packer {
    std::vector<int>& order;
    static size_t median_count(size_t count, size_t maxc, size_t minc)
    {
        size_t n = count / maxc, r = count % maxc, m = (n / 2) * maxc;
        if (r == 0) return m;
        if (minc <= r) return ((n + 1) / 2) * maxc;
        const size_t rest = count - minc;
        n = rest / maxc; r = rest % maxc;
        if (r == 0) return ((n + 1) / 2) * maxc;
        return n == 0 ? r : ((n + 2) / 2) * maxc;
    }
    // one tree level: split [first, last) into packets of <= maxc values, each packet becomes a child subtree
    void level(Entry* first, Entry* last, const Box3& hint, size_t maxc, size_t minc)
    {
        if (maxc <= 1) { for (Entry* e = first; e != last; ++e) order.push_back(e->idx); return; }        // leaf
        packets(first, last, hint, maxc, minc);
    }
    void packets(Entry* first, Entry* last, const Box3& hint, size_t maxc, size_t minc)
    {
        const size_t count = (size_t)(last - first);
        if (count <= maxc) { level(first, last, hint, maxc / kMaxElements, minc / kMaxElements); return; }
        const size_t mc = median_count(count, maxc, minc);
        int dim = 0; float len = hint.hi[0] - hint.lo[0];
        for (int d = 1; d < 3; d++) { const float cur = hint.hi[d] - hint.lo[d]; if (len < cur) { dim = d; len = cur; } }
        std::nth_element(first, first + mc, last, [dim](const Entry& a, const Entry& b) { return a.c[dim] < b.c[dim]; });
        Box3 left = hint, right = hint;
        const float mid = hint.lo[dim] + (hint.hi[dim] - hint.lo[dim]) / 2;
        left.hi[dim] = mid; right.lo[dim] = mid;
        packets(first, first + mc, left, maxc, minc);
        packets(first + mc, last, right, maxc, minc);
    }
};

void packed_rtree_rank(const mage_keypoint* kps, int n, std::vector<int>& rank)
{
    rank.assign(n, 0);
    if (n <= 0) return;
    std::vector<Entry> e(n);
    Box3 box;
    for (int i = 0; i < n; i++) {
        e[i].c[0] = kps[i].x; e[i].c[1] = kps[i].y; e[i].c[2] = kps[i].octave * 100.f; e[i].idx = i;      // octaveSpacing = 100
        for (int d = 0; d < 3; d++) {
            if (i == 0) box.lo[d] = box.hi[d] = e[i].c[d];
            else { box.lo[d] = std::min(box.lo[d], e[i].c[d]); box.hi[d] = std::max(box.hi[d], e[i].c[d]); }
        }
    }
    size_t maxc = 1;
    for (size_t smax = kMaxElements; smax < (size_t)n; smax *= kMaxElements) maxc = smax;
    const size_t minc = kMinElements * (maxc / kMaxElements);
    std::vector<int> order;
    order.reserve(n);
    Packer pk{order};
    pk.level(e.data(), e.data() + n, box, maxc, minc);
    for (int r = 0; r < n; r++) rank[order[r]] = r;
}

} // namespace

struct mage_spatial_index_s {
    int n = 0;
    DevStage mem;                   // coordinates, octaves and ranks of the indexed key points (from the process-wide scratch pool: the index is
                                    // rebuilt for every analysed frame, so its set-up cost is per-frame latency)
    RadiusIndexDev dev{};
    std::vector<int> rank;
};

extern "C" int mage_spatial_index_create(const mage_keypoint* keypoints, int n, mage_spatial_index_s** out)
{
    MAGE_REQUIRE(out && n >= 0 && (keypoints || n == 0), MAGE_ERR_INVALID, "mage_spatial_index_create: bad argument");
    MAGE_REQUIRE(n < (1 << 20), MAGE_ERR_UNSUPPORTED, "at most 2^20 keypoints per index");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: RadiusMatch has no CPU fallback"); return MAGE_ERR_CUDA; }
    mage_spatial_index_s* ix = new mage_spatial_index_s();
    ix->n = n;
    packed_rtree_rank(keypoints, n, ix->rank);
    // one pooled buffer, one packed upload
    const size_t cnt = (size_t)std::max(n, 1);
    size_t used = 0;
    auto take = [&](size_t bytes) { used = align_up(used, 16); const size_t o = used; used += bytes; return o; };
    const size_t ox = take(4 * cnt), oy = take(4 * cnt), oo = take(4 * cnt), orank = take(4 * cnt);
    std::vector<uint8_t> host(used);
    float* x = reinterpret_cast<float*>(host.data() + ox); float* y = reinterpret_cast<float*>(host.data() + oy);
    int* oc = reinterpret_cast<int*>(host.data() + oo);
    for (int i = 0; i < n; i++) { x[i] = keypoints[i].x; y[i] = keypoints[i].y; oc[i] = keypoints[i].octave; }
    if (n) memcpy(host.data() + orank, ix->rank.data(), 4 * (size_t)n);
    ix->mem = dev_stage_acquire(used);
    cudaError_t e = ix->mem.p ? cudaSuccess : cudaErrorMemoryAllocation;
    if (e == cudaSuccess && n) e = cudaMemcpy(ix->mem.p, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("mage_spatial_index_create: %s", cudaGetErrorString(e)); dev_stage_release(ix->mem); delete ix; return MAGE_ERR_CUDA; }
    uint8_t* base = ix->mem.p;
    ix->dev.x = reinterpret_cast<float*>(base + ox); ix->dev.y = reinterpret_cast<float*>(base + oy);
    ix->dev.octave = reinterpret_cast<int*>(base + oo); ix->dev.rank = reinterpret_cast<int*>(base + orank); ix->dev.n = n;
    *out = ix;
    return MAGE_OK;
}

extern "C" void mage_spatial_index_destroy(mage_spatial_index_s* ix)
{
    if (!ix) return;
    dev_stage_release(ix->mem);                              // every match call synchronises before it returns, nothing is in flight
    delete ix;
}

// enumeration rank of every indexed keypoint (inspection tap for the parity tests)
extern "C" int mage_spatial_index_rank(mage_spatial_index_s* ix, int* rank_out)
{
    MAGE_REQUIRE(ix && rank_out, MAGE_ERR_INVALID, "null argument");
    std::copy(ix->rank.begin(), ix->rank.end(), rank_out);
    return MAGE_OK;
}

extern "C" int mage_radius_match(mage_spatial_index_s* ix, const mage_keypoint* query_kps, int nq, const float* query_pos_override,
                                 const uint8_t* query_mask, const uint8_t* query_desc, const uint8_t* target_mask, const uint8_t* target_desc,
                                 float radius, int max_hamming, int min_hamming_diff, mage_dmatch* out, int* count, void* cuda_stream)
{
    MAGE_REQUIRE(ix && count && nq >= 0 && (nq == 0 || (query_kps && query_desc && out)), MAGE_ERR_INVALID, "mage_radius_match: bad argument");
    *count = 0;
    if (nq == 0 || ix->n == 0) return MAGE_OK;
    MAGE_REQUIRE(target_desc, MAGE_ERR_INVALID, "mage_radius_match: null target descriptors");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : thread_stream();      // (a null result is the default stream: still correct)
    const size_t nT = (size_t)ix->n, nQ = (size_t)nq;
    // scratch layout: query kps | query pos | query mask | query desc | target mask | target desc | almost | tbest | tsecond | out | count
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_qk = take(sizeof(mage_keypoint) * nQ), o_qp = take(8 * nQ), o_qm = take(nQ), o_qd = take(32 * nQ), o_tm = take(nT), o_td = take(32 * nT);
    const size_t o_al = take(4 * nQ), o_tb = take(4 * nT), o_ts = take(4 * nT), o_out = take(sizeof(mage_dmatch) * nQ), o_cnt = take(4);
    // the scratch comes from the process-wide pool (an index lives for one frame: a scratch of its own would be allocated every frame)
    DevStage ds = dev_stage_acquire(off);
    MAGE_REQUIRE(ds.p, MAGE_ERR_CUDA, "mage_radius_match: no device memory for %zu bytes of scratch", off);
    uint8_t* S = ds.p;
    // the inputs are packed into one pinned staging buffer laid out like the scratch (query kps .. target desc are neighbours) and go up in
    // ONE copy, the two target tables are cleared by one memset, matches + count come back in one copy: up to six uploads and two
    // read-backs between the device and the caller's pageable arrays were 60 us of a 107 us call
    PinnedStage st = stage_acquire(off);
    if (!st.p) { dev_stage_release(ds); MAGE_REQUIRE(false, MAGE_ERR_CUDA, "mage_radius_match: no pinned staging memory"); }
    memcpy(st.p + o_qk, query_kps, sizeof(mage_keypoint) * nQ);
    if (query_pos_override) memcpy(st.p + o_qp, query_pos_override, 8 * nQ);
    if (query_mask) memcpy(st.p + o_qm, query_mask, nQ);
    memcpy(st.p + o_qd, query_desc, 32 * nQ);
    if (target_mask) memcpy(st.p + o_tm, target_mask, nT);
    memcpy(st.p + o_td, target_desc, 32 * nT);
    cudaError_t e = cudaMemcpyAsync(S + o_qk, st.p + o_qk, o_td + 32 * nT - o_qk, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(S + o_tb, 0xFF, o_ts + 4 * nT - o_tb, s);
    if (e != cudaSuccess) { cudaStreamSynchronize(s); stage_release(st); dev_stage_release(ds); MAGE_CUDA_TRY(e); }
    k_radius_best<<<div_up(nq, 8), 256, 0, s>>>(ix->dev, reinterpret_cast<const mage_keypoint*>(S + o_qk), nq,
                                                query_pos_override ? reinterpret_cast<const float*>(S + o_qp) : nullptr,
                                                query_mask ? S + o_qm : nullptr, reinterpret_cast<const uint32_t*>(S + o_qd),
                                                target_mask ? S + o_tm : nullptr, reinterpret_cast<const uint32_t*>(S + o_td), radius, max_hamming,
                                                min_hamming_diff, reinterpret_cast<unsigned*>(S + o_al), reinterpret_cast<unsigned*>(S + o_tb),
                                                reinterpret_cast<unsigned*>(S + o_ts));
    k_radius_emit<<<1, 256, 0, s>>>(reinterpret_cast<const unsigned*>(S + o_al), nq, reinterpret_cast<const unsigned*>(S + o_tb),
                                    reinterpret_cast<const unsigned*>(S + o_ts), reinterpret_cast<mage_dmatch*>(S + o_out), reinterpret_cast<int*>(S + o_cnt));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(st.p + o_out, S + o_out, o_cnt + 4 - o_out, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) {
        *count = *reinterpret_cast<const int*>(st.p + o_cnt);
        memcpy(out, st.p + o_out, sizeof(mage_dmatch) * (size_t)std::max(0, std::min(*count, nq)));
    }
    if (e != cudaSuccess) cudaStreamSynchronize(s);          // nothing of this call may still be running when the buffers go back
    stage_release(st);
    dev_stage_release(ds);
    MAGE_CUDA_TRY(e);
    return MAGE_OK;
}
