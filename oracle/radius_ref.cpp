/*
 * oracle/radius_ref.cpp -- TEST INFRASTRUCTURE ONLY.
 * RadiusMatch oracle on top of the REAL boost R-tree the reference uses (vendored headers under
 * /root/reference/Dependencies/boost, boost 1.67): the index is built exactly like KeypointSpatialIndex
 * (ref Core/MAGESLAM/Source/Image/KeypointSpatialIndex.cpp:26-58: rstar<12>, points (x, y, octave*100), range constructor =
 * packing algorithm) and queried exactly like KeypointSpatialIndex::Query (:89-97). The matching loops restate
 * ref Tracking/FeatureMatcher.cpp:294-446 (they need only the enumeration order of the query results, which is the
 * implementation-defined part). Built by `make -C oracle ref` into oracle/_ref/libradius_ref.so; never linked into the product.
 */
#include <boost/geometry/algorithms/equals.hpp>
#include <boost/geometry/strategies/strategies.hpp>
#include <boost/geometry/index/rtree.hpp>

#include <cstdint>
#include <cstring>
#include <limits>
#include <tuple>
#include <vector>

namespace {
using indexType = boost::geometry::index::rstar<12>;
using rtree_point = boost::geometry::model::point<float, 3, boost::geometry::cs::cartesian>;
using rtree_box = boost::geometry::model::box<rtree_point>;
using rtree_value = std::tuple<rtree_point, size_t>;
using rtree = boost::geometry::index::rtree<rtree_value, indexType>;
constexpr float octaveSpacing = 100, octaveQueryRange = 1;       // ref KeypointSpatialIndex.h:34-35

struct KP { float x, y, size, angle, response; int32_t octave, class_id; };
struct DM { int32_t query, train; float distance; };

struct Index { rtree index; Index(const std::vector<rtree_value>& v) : index(v.begin(), v.end()) {} };

int descriptorDistance(const uint8_t* a, const uint8_t* b)
{
    int r = 0;
    for (int i = 0; i < 8; i++) { uint32_t x, y; memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4); r += __builtin_popcount(x ^ y); }
    return r;
}
void query(const Index& ix, float cx, float cy, int octave, float radius, std::vector<size_t>& out)
{
    rtree_box box{rtree_point{cx - radius, cy - radius, octave * octaveSpacing - octaveQueryRange},
                  rtree_point{cx + radius, cy + radius, octave * octaveSpacing + octaveQueryRange}};
    struct Out {
        std::vector<size_t>* r;
        Out& operator=(const rtree_value& v) { r->push_back(std::get<1>(v)); return *this; }
        Out& operator*() { return *this; }
        Out& operator++() { return *this; }
    } o{&out};
    ix.index.query(boost::geometry::index::intersects(box), o);
}
}

extern "C" {

void* rmref_index_create(const KP* kps, int n)
{
    std::vector<rtree_value> v;
    v.reserve(n);
    for (int i = 0; i < n; i++) v.emplace_back(rtree_point(kps[i].x, kps[i].y, kps[i].octave * octaveSpacing), (size_t)i);
    return new Index(v);
}
void rmref_index_destroy(void* h) { delete static_cast<Index*>(h); }

int rmref_query(void* h, float cx, float cy, int octave, float radius, int* out, int cap)
{
    std::vector<size_t> r;
    query(*static_cast<Index*>(h), cx, cy, octave, radius, r);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = (int)r[i];
    return (int)r.size();
}

// ref FeatureMatcher.cpp:294-376 (multi query) + :384-446 (single query)
int rmref_radius_match(void* h, const KP* qk, int nq, const float* qpos_override /*nullable, 2 per query*/, const uint8_t* qmask,
                       const uint8_t* qdesc, int nt, const uint8_t* tmask, const uint8_t* tdesc, float radius, int maxHamming,
                       int minDiff, DM* out)
{
    const Index& ix = *static_cast<Index*>(h);
    std::vector<DM> almost;
    std::vector<size_t> res;
    for (int q = 0; q < nq; q++) {
        if (qmask && !qmask[q]) continue;
        float px = qpos_override ? qpos_override[2 * q] : qk[q].x, py = qpos_override ? qpos_override[2 * q + 1] : qk[q].y;
        int best = maxHamming + 1, second = std::numeric_limits<int>::max();
        DM bm{0, -1, 0.f};
        res.clear();
        query(ix, px, py, qk[q].octave, radius, res);
        for (size_t t : res) {
            if (tmask == nullptr || tmask[t]) {
                int d = descriptorDistance(qdesc + 32 * (size_t)q, tdesc + 32 * t);
                if (d < best) { bm.train = (int)t; bm.distance = (float)d; second = best; best = d; }
            }
        }
        if (bm.train != -1 && (second - best) > minDiff) { bm.query = q; almost.push_back(bm); }
    }
    int n = 0;
    if (almost.size() > 1) {
        std::vector<float> bestD(nt, std::numeric_limits<float>::max()), secondD(nt, std::numeric_limits<float>::max());
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

} // extern "C"
